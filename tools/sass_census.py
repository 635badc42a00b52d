"""SASS opcode census of libsegk.so -> profiles/r2_sass_census.md (cuobjdump -sass; tensor-core / TMA / multimem opcodes per kernel)."""
import os, re, subprocess, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "semanticsegmentation_tensorflow_b200", "libsegk.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
kernels, cur = [], None
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        d = next(it)
        d = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", d)
        if d.endswith(")"):                      # drop the argument list (the last balanced parenthesis group)
            depth, i = 0, len(d) - 1
            while i >= 0:
                depth += d[i] == ")"
                depth -= d[i] == "("
                if depth == 0:
                    break
                i -= 1
            d = d[:i]
        d = d.replace("(int)", "").replace("(bool)", "")
        cur = {"name": d, "n": 0, "ops": collections.Counter()}
        kernels.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P[T\d]+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        cur["n"] += 1
        cur["ops"][m.group(1)] += 1
tot = collections.Counter()
for k in kernels:
    tot.update(k["ops"])
def fam(c, prefix):
    return sum(v for o, v in c.items() if o.startswith(prefix))
def exact(c, pat):
    return sum(v for o, v in c.items() if re.fullmatch(pat, o))
lines = ["# SASS opcode census of `libsegk.so` (round 2)", "",
         f"`python tools/sass_census.py` = `cuobjdump -sass semanticsegmentation_tensorflow_b200/libsegk.so` of the committed build (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a`); {len(kernels)} kernels in total.",
         f"Library-wide counts: **{fam(tot, 'UTCHMMA')} `UTCHMMA`** (tcgen05.mma; {fam(tot, 'UTCHMMA.2CTA')} of them `UTCHMMA.2CTA` = `cta_group::2`, M = 256 across a CTA pair), "
         f"**{fam(tot, 'UTMALDG')} `UTMALDG`** ({exact(tot, r'UTMALDG\.3D.*')} `.3D` + {exact(tot, r'UTMALDG\.4D.*')} `.4D` TMA loads, {sum(v for o, v in tot.items() if o.startswith('UTMALDG') and '2CTA' in o)} of them `.2CTA`: completion on the pair leader's mbarrier), "
         f"**{fam(tot, 'UTMASTG')} `UTMASTG.4D`** (TMA stores), **{fam(tot, 'LDTM')} `LDTM.x32`** (tcgen05.ld), {fam(tot, 'UTCBAR')} `UTCBAR` (tcgen05.commit; {sum(v for o, v in tot.items() if o.startswith('UTCBAR') and 'MULTICAST' in o)} `.2CTA.MULTICAST`), "
         f"{fam(tot, 'UTCATOMSWS')} `UTCATOMSWS` (TMEM alloc / dealloc), {fam(tot, 'SYNCS.PHASECHK')} `SYNCS.PHASECHK` + {fam(tot, 'SYNCS.ARRIVE')} `SYNCS.ARRIVE` (mbarrier), "
         f"**{fam(tot, 'LDGMC')} `LDGMC.E.ADD.F32x4`** (multimem.ld_reduce in the NVSwitch exchange kernels), {fam(tot, 'STG.E.ENL2.256')} `STG.E.ENL2.256` (256-bit dW stores), "
         f"{sum(fam(k['ops'], 'REDG') + fam(k['ops'], 'ATOMG') for k in kernels if fam(k['ops'], 'UTCHMMA'))} `REDG` / `ATOMG` in the tensor-core kernels (no atomics on the data path).  "
         "No cuBLAS / cuDNN / CUTLASS symbol is linked (`ldd`: libstdc++, libm, libc; the CUDA runtime is linked statically).", "",
         "Kernels that contain tensor-core / TMA / multimem instructions:", "",
         "| kernel | SASS instrs | UTCHMMA | of which .2CTA | UTMALDG | UTMASTG | LDTM | LDGMC |", "|---|---:|---:|---:|---:|---:|---:|---:|"]
rows = [k for k in kernels if fam(k["ops"], "UTCHMMA") or fam(k["ops"], "UTMALDG") or fam(k["ops"], "LDGMC")]
rows.sort(key=lambda k: (-fam(k["ops"], "UTCHMMA"), k["name"]))
for k in rows:
    c = k["ops"]
    lines.append(f"| `{k['name']}` | {k['n']} | {fam(c, 'UTCHMMA')} | {fam(c, 'UTCHMMA.2CTA')} | {fam(c, 'UTMALDG')} | {fam(c, 'UTMASTG')} | {fam(c, 'LDTM')} | {fam(c, 'LDGMC')} |")
open(os.path.join(ROOT, "profiles", "r2_sass_census.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:4]))
