"""Merge an ncu --csv metrics capture of one training step's tensor-core launches (tools/ncu_step.py) with the ordered
call list it wrote (gpurun_out/step_calls.json):
    python tools/ncu_merge.py gpurun_out/step_metrics.csv gpurun_out/step_calls.json profiles/r2_ncu_step_tc_launches.md profiles/ncu_traffic.json
Writes the per-launch table (markdown) and the per-call-shape DRAM traffic / tensor-pipe table bench.py reads."""
import csv, json, sys

csv_path, calls_path, md_path, json_path = sys.argv[1:5]
rows = list(csv.reader(open(csv_path)))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
launches = {}
for r in rows[start + 1:]:
    if len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    L = launches.setdefault(int(d["ID"]), {"kernel": d["Kernel Name"]})
    try:
        L[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        pass
calls = json.load(open(calls_path))
ids = sorted(launches)
assert len(ids) == len(calls), (len(ids), len(calls))
M = {"t": "gpu__time_duration.sum", "r": "dram__bytes_read.sum", "w": "dram__bytes_write.sum",
     "tp": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l2": "lts__t_sector_hit_rate.pct"}
out = ["| # | C-ABI call [int args] | kernel | time us | tensor pipe active % | DRAM read MB | DRAM write MB | L2 hit % |",
       "|---:|---|---|---:|---:|---:|---:|---:|"]
table, tot_t, tot_tp = {}, 0.0, 0.0
for i, (k, call) in enumerate(zip(ids, calls)):
    L = launches[k]
    name = L["kernel"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    t, tp = L[M["t"]] / 1e3, L.get(M["tp"], 0.0)
    out.append(f"| {i} | `{call}` | `{name}` | {t:.1f} | {tp:.1f} | {L[M['r']] / 1e6:.1f} | {L[M['w']] / 1e6:.1f} | {L.get(M['l2'], 0.0):.1f} |")
    tot_t += t
    tot_tp += t * tp
    e = table.setdefault(call, {"dram_bytes": 0.0, "tensor_pipe_active_pct": 0.0, "n": 0, "kernel": name, "source": md_path})
    e["dram_bytes"] += L[M["r"]] + L[M["w"]]
    e["tensor_pipe_active_pct"] += tp
    e["n"] += 1
for e in table.values():
    e["dram_bytes"] /= e["n"]
    e["tensor_pipe_active_pct"] /= e["n"]
out.append("")
out.append(f"Sum of the {len(ids)} launches: {tot_t / 1e3:.3f} ms; time-weighted tensor-pipe activity {tot_tp / tot_t:.1f} %.")
open(md_path, "a").write("\n".join(out) + "\n")
json.dump(table, open(json_path, "w"), indent=1)
print(out[-1])
