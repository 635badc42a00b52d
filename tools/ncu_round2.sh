set -x
python tools/ncu_step.py > gpurun_out/r2_ncu_step_plain.log 2>&1 || exit 1
K=$(cat gpurun_out/step_k.txt)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k "regex:^(igemm_kernel|igemm_pair_kernel|wgrad_kernel|wgrad_pair_kernel|slab_kernel|slab3_kernel|wslab_kernel|first_fwd_kernel|first_wgrad_kernel)$" -s $((2*K)) -c $K --csv --log-file gpurun_out/step_metrics.csv python tools/ncu_step.py > gpurun_out/r2_ncu_step.log 2>&1
echo "ncu step rc $?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_short.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_ncu_bench.log 2>&1
echo "ncu launches rc $?"
