# Final round-2 measurement run (one GPU): tests, ncu tables, bench lines.  Outputs under gpurun_out/.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_final_pytest.log
tail -2 gpurun_out/r2_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_final_smoke.log 2>&1; tail -1 gpurun_out/r2_final_smoke.log
bash tools/ncu_round2.sh > gpurun_out/r2_final_ncu.log 2>&1
python bench.py --layers-out gpurun_out/r2_final_layers.json > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2>/dev/null
for m in unet segnet fcdensenet; do python bench.py --model $m --steps 10 --no-cpu-baseline --layers-out gpurun_out/r2_final_layers_$m.json > gpurun_out/r2_final_bench_$m.json 2>/dev/null; done
python bench.py --workload infer --steps 20 > gpurun_out/r2_final_bench_infer.json 2>/dev/null
python bench.py --model fcdensenet --res 384x1248 --batch 8 --steps 4 --no-cpu-baseline > gpurun_out/r2_final_bench_fcdensenet_full.json 2>/dev/null
echo done
