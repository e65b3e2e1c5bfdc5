"""Static SASS instruction counts of the hot kernels in lib/libife_cuda.so (cuobjdump -sass), as a markdown table.
Usage: python profiles/sass_evidence.py > profiles/r2_sass_evidence.md"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "image-feature-extraction_b200", "lib", "libife_cuda.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kernels, cur = {}, None
for line in txt.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1); kernels[cur] = []; continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", line)
    if m and cur: kernels[cur].append(m.group(1))
names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
cols = [("UTMALDG (TMA load)", r"\bUTMALDG"), ("UTMASTG (TMA store)", r"\bUTMASTG"), ("UTMAPF (TMA L2 prefetch)", r"\bUTMAPF"),
        ("SYNCS (mbarrier)", r"\bSYNCS"), ("ELECT", r"\bELECT"), ("LDGSTS (cp.async)", r"\bLDGSTS"), ("STG.E.128", r"\bSTG\.E\.128"),
        ("LDS.128", r"\bLDS\.128"), ("FP64 (DFMA+DMUL+DADD)", r"\b(DFMA|DMUL|DADD)\b"), ("MATCH", r"\bMATCH")]
print("# SASS evidence (closing build, `cuobjdump -sass lib/libife_cuda.so`, `profiles/sass_evidence.py`): static instruction counts per kernel\n")
print("| kernel | instructions | " + " | ".join(c for c, _ in cols) + " |")
print("|---|---|" + "---|" * len(cols))
rows = []
for mangled, name in zip(kernels, names):
    short = re.sub(r"\(.*", "", name).replace("void ", "")
    short = re.sub(r"\((int|bool)\)", "", short)
    if not re.search(r"iir_tma_kernel|features_march4_kernel<0, (false|true), true|rs_scatter_kernel$", short): continue
    ins = kernels[mangled]
    rows.append((short, len(ins), [sum(1 for i in ins if re.search(p, i)) for _, p in cols]))
for short, n, c in sorted(rows):
    print("| `%s` | %d | " % (short, n) + " | ".join(str(x) for x in c) + " |")
print("""
`iir_tma_kernel<AXIS, INMODE, DIVIDE, FMA, MINB, MASKMODE>`: AXIS 0 = z, 1 = y, 2 = x; INMODE 0 = two float fields (or one field in two
stacks of rows), 1 = image + uint8 certainty, 2 = image + float certainty; MASKMODE 1 / 2 = uint8 / float output mask in the divide.
No `HMMA` / `UTC*MMA` anywhere: the path has no dense contraction.
`features_march4_kernel<0, false, true, 1>` = 8 feature volumes out (16-byte `cp.async` staging, `STG.E.128` stores); `<0, true, true, 2>` = histograms only.""")
