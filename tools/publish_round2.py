"""Copy the outputs of tools/final_round2.sh from gpurun_out/ into profiles/ (tracked) and rebuild the derived tables:
r2_ncu_step_tc_launches.md + ncu_traffic.json (tools/ncu_merge.py), r2_launches_summary.md, r2_results.md."""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def line(path):
    return json.loads([l for l in open(path) if l.startswith("{")][-1])


copies = {"r2_final_bench.json": "r2_bench_n1.json", "r2_final_layers.json": "r2_layers_n1.json",
          "r2_final_bench_ref.json": "r2_bench_reference_arm.json", "r2_final_bench_unet.json": "r2_bench_unet.json",
          "r2_final_bench_segnet.json": "r2_bench_segnet.json", "r2_final_bench_fcdensenet.json": "r2_bench_fcdensenet.json",
          "r2_final_bench_fcdensenet_full.json": "r2_bench_fcdensenet_384x1248.json", "r2_final_bench_infer.json": "r2_bench_infer.json",
          "r2_final_pytest.log": "r2_gpu_tests.log", "r2_launches.csv": "r2_launches.csv"}
for src, dst in copies.items():
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))

# per-launch ncu table of one step + traffic json
md = os.path.join(P, "r2_ncu_step_tc_launches.md")
head = open(md).read().split("| # |")[0]
open(md, "w").write(head)
subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_merge.py"), os.path.join(G, "step_metrics.csv"),
                       os.path.join(G, "step_calls.json"), md, os.path.join(P, "ncu_traffic.json")])

# launch list summary: one step = the launches between two consecutive first_fwd_kernel launches
rows = list(csv.reader(open(os.path.join(P, "r2_launches.csv"))))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ls = [(r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", ""), float(r[vi].replace(",", "")) / 1e6)
      for r in rows[start + 1:] if len(r) == len(hdr)]
firsts = [i for i, (k, _) in enumerate(ls) if k.startswith("first_fwd_kernel")]
a, b = firsts[3], firsts[4]
step = ls[a:b]
agg = {}
for k, t in step:
    e = agg.setdefault(k, [0, 0.0])
    e[0] += 1
    e[1] += t
tot = sum(t for _, t in step)
out = ["# Round 2 — ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary` (B200, N=1)", "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv` (cold-cache, serialised: compare SHARES, not absolutes).",
       f"Raw list: `profiles/r2_launches.csv`.  The table is ONE training step (the launches between the 4th and 5th `first_fwd_kernel`: {len(step)} launches, {tot:.2f} ms serialised under ncu).", "",
       "| kernel | launches / step | ms | share |", "|---|---:|---:|---:|"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {n} | {t:.3f} | {100 * t / tot:.1f}% |")
open(os.path.join(P, "r2_launches_summary.md"), "w").write("\n".join(out) + "\n")

# results
b = line(os.path.join(P, "r2_bench_n1.json"))
sec = b.get("secondary") or {}
res = {"FCN_IPS": f"{b['value']:.0f}", "FCN_MS": f"{b['ms_per_step']:.2f}", "FCN_E2E": f"{b['e2e']['value']:.0f}",
       "GEMM_MS": f"{b['conv_gemms']['ms_per_step']:.2f}", "GEMM_TF": f"{b['conv_gemms']['tflops'] / 1e3:.2f}",
       "GEMM_FRAC": f"{100 * b['conv_gemms']['frac_of_burst_peak']:.0f}",
       "INF_IPS": f"{sec['fcn_infer_384x1248_b16']['images_per_s']:.0f}" if sec else "?",
       "LAT": f"{sec['fcn_infer_160x576_b1_latency_ms']['p50']:.2f} ({sec['fcn_infer_160x576_b1_latency_ms']['cuda_graph']['p50']:.2f} from a CUDA graph)" if sec else "?"}
for m, key in (("unet", "UNET"), ("segnet", "SEGNET"), ("fcdensenet", "DN")):
    p = os.path.join(P, f"r2_bench_{m}.json")
    if os.path.exists(p):
        d = line(p)
        res[key + "_IPS"] = f"{d['value']:.0f}"
        res[key + "_MS"] = f"{d['ms_per_step']:.1f}"
for n in (2, 4, 8):
    p = os.path.join(P, f"r2_bench_n{n}.json")
    if os.path.exists(p):
        d = line(p)
        res[f"N{n}_IPS"] = f"{d['value']:.0f}"
        res[f"N{n}_MS"] = f"{d['ms_per_step']:.2f}"
json.dump(res, open(os.path.join(P, "r2_results.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
print("roofline", json.dumps(b["roofline"])[:600])
