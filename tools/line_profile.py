"""Per-source-line share of warp-stall samples and executed instructions of one `ncu --set full --import-source on`
capture:  python tools/line_profile.py report.ncu-rep source.cu out.md [min_pct]"""
import csv, io, os, subprocess, sys
rep, src_path, out_path = sys.argv[1:4]
min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = next(r for r in rows if "# Samples" in r)
ismp, iex = h.index("# Samples"), h.index("Instructions Executed")
cur, agg, kernel = None, {}, ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if len(r) == 2 and r[0] == "Function Name":
        kernel = r[1]; continue
    if len(r) <= ismp or not r[0].isdigit():
        continue
    k = (cur, int(r[0]))
    a = agg.get(k, (0, 0))
    num = lambda v: int(v) if v.strip().isdigit() else 0
    agg[k] = (a[0] + num(r[ismp]), a[1] + num(r[iex]))
tot, totex = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
src = open(src_path).read().split("\n")
name = os.path.basename(src_path)
with open(out_path, "w") as f:
    f.write("# Where `%s` spends its issue slots\n\n`%s`, capture `%s` (ncu --set full, warp-state sampling): %d samples, %d warp "
            "instructions.  Lines with at least %.1f %% of either; lines of inlined headers (packed f32x2 intrinsics) are summed per header line.\n\n"
            % (kernel.split("(")[0], name, os.path.basename(rep), tot, totex, min_pct))
    f.write("| line | samples % | instructions % | source |\n|---|---|---|---|\n")
    for k in sorted(agg, key=lambda k: (k[0] != name, k[1])):
        s, e = agg[k]
        if 100.0 * s / tot >= min_pct or 100.0 * e / totex >= min_pct:
            code = src[k[1] - 1].strip()[:110].replace("|", "\\|") if k[0] == name and k[1] <= len(src) else "(inlined from %s)" % k[0]
            f.write("| %s:%d | %.1f | %.1f | `%s` |\n" % (k[0], k[1], 100.0 * s / tot, 100.0 * e / totex, code))
print(open(out_path).read()[:1500])
