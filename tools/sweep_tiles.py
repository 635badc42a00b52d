"""Sweep BLOCK_N / split-K overrides per layer shape and op (run on the GPU box)."""
import os, sys, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
layers = sys.argv[1:] or ["conv1_2", "conv2_2", "conv3_2", "conv4_2", "conv5_2", "conv6", "conv7"]
res = []
for bn in (0, 64, 128, 256):
    for sp in (0, 1, 2, 4, 8):
        env = dict(os.environ, SEGK_FORCE_BN=str(bn), SEGK_FORCE_KSPLIT=str(sp), SEGK_FORCE_WSPLIT=str(sp))
        if bn == 0 and sp != 0:
            continue
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prof_layers.py"), "--time"] + layers,
                           env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        for line in r.stdout.splitlines():
            f = line.split()
            if len(f) >= 4 and f[1] in ("fwd", "dgrad", "wgrad"):
                res.append({"bn": bn, "split": sp, "layer": f[0], "op": f[1], "ms": float(f[2])})
        if r.returncode != 0:
            print("FAILED bn", bn, "split", sp, r.stdout[-500:])
best = {}
for r in res:
    k = (r["layer"], r["op"])
    if k not in best or r["ms"] < best[k]["ms"]:
        best[k] = r
base = {(r["layer"], r["op"]): r for r in res if r["bn"] == 0 and r["split"] == 0}
for k in sorted(best):
    b = best[k]
    print("%-8s %-6s default %.4f ms | best %.4f ms (bn %d split %d)" % (k[0], k[1], base[k]["ms"], b["ms"], b["bn"], b["split"]))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)
