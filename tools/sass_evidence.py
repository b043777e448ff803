"""SASS evidence per tensor-core kernel: counts of the tcgen05 / TMA / TMEM mnemonics in libb200spk.so (cuobjdump -sass),
written to profiles/r02_sass_evidence.md.  UTCHMMA = tcgen05.mma (kind::f16), UTMALDG / UTMASTG = TMA tensor load / store
(the .IM2COL form is the im2col load), LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, LDGSTS = cp.async."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "3d-speaker_b200", "b200spk", "libb200spk.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "LDGSTS", "UCGABAR", "FFMA2", "SYNCS"]
cur, counts, im2col = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("void ", "").replace("spk::", "").replace("__nv_bfloat16", "bf16")
        cur = cur.split("(")[0] if not cur.startswith("(") else cur
        counts.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    for p in pats:
        if re.search(r"\b" + p + r"\b|\b" + p + r"\.", line):
            counts[cur][p] += 1
    if "UTMALDG" in line and "IM2COL" in line:
        im2col[cur] += 1
with open(os.path.join(ROOT, "profiles", "r02_sass_evidence.md"), "w") as f:
    f.write("# SASS mnemonics per kernel of libb200spk.so (`cuobjdump -sass`, sm_100a)\n\n" + __doc__.split("\n", 1)[1].strip() + "\n\n")
    f.write("| kernel | " + " | ".join(pats) + " | UTMALDG.IM2COL |\n|---|" + "---|" * (len(pats) + 1) + "\n")
    for k, c in counts.items():
        if not any(c[p] for p in pats[:7]):
            continue
        f.write("| `%s` | " % k[:100] + " | ".join(str(c[p]) for p in pats) + " | %d |\n" % im2col[k])
print(open(os.path.join(ROOT, "profiles", "r02_sass_evidence.md")).read()[:3000])
