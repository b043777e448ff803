"""DRAM traffic of one CAM++ forward call from an ncu pass (bench.py's roofline.traffic).

On the GPU box (tools/profile_round.sh does this):
    python tools/run_forward.py --segments 2048 --iters 1 > gpurun_out/plain.log 2>&1 && \
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/r02_traffic.csv python tools/run_forward.py --segments 2048 --iters 1
Here:
    python tools/measure_traffic.py gpurun_out/r02_traffic.csv 2048
writes profiles/r02_traffic.json (+ .md): bytes per segment over every kernel of the forward call (fbank and the
one-off weight conversions excluded), tagged with the hash of csrc/ so bench.py can tell a stale profile."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main(path, segments, tag="r02_traffic", title="CAM++ forward call (%d x 1.5 s segments, bf16)", model_args=""):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    iid = hdr.index("ID")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "usecond": 1e3, "nsecond": 1.0, "msecond": 1e6}
    per = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])     # launches, read, write, ns
    seen = set()
    for r in rows[1:]:
        name = r[ik].split("(")[0].replace("void ", "").replace("spk::", "").replace("<unnamed>::", "")
        if "f32_to_bf16" in name or "fbank" in name or "cmn_kernel" in name or "at::" in name or "elementwise" in name:
            continue
        v = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
        e = per[name]
        if (r[iid], name) not in seen:
            seen.add((r[iid], name))
            e[0] += 1
        if r[im].startswith("dram__bytes_read"):
            e[1] += v
        elif r[im].startswith("dram__bytes_write"):
            e[2] += v
        elif r[im].startswith("gpu__time_duration"):
            e[3] += v
    tot_r = sum(e[1] for e in per.values())
    tot_w = sum(e[2] for e in per.values())
    tot_t = sum(e[3] for e in per.values())
    out = {"segments": segments, "dram_bytes_read": tot_r, "dram_bytes_write": tot_w,
           "dram_bytes_per_segment": (tot_r + tot_w) / segments, "kernel_time_ms_serialised": tot_t / 1e6,
           "launches": sum(e[0] for e in per.values()), "source_hash": bench.source_hash(),
           "command": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                      "python tools/run_forward.py %s--segments %d --iters 1" % (model_args, segments),
           "per_kernel": {k: {"launches": e[0], "read_MB": e[1] / 1e6, "write_MB": e[2] / 1e6, "ms": e[3] / 1e6}
                          for k, e in sorted(per.items(), key=lambda kv: -(kv[1][1] + kv[1][2]))}}
    json.dump(out, open(os.path.join(ROOT, "profiles", tag + ".json"), "w"), indent=1)
    with open(os.path.join(ROOT, "profiles", tag + ".md"), "w") as f:
        f.write("# DRAM traffic and device time per kernel of one %s, ncu dram__bytes + gpu__time_duration\n\n" % (title % segments))
        f.write("`%s`\n\nTotal %.2f GB read + %.2f GB written = **%.2f MB per segment**, %d launches, %.2f ms of serialised kernel time; csrc hash %s.\n\n"
                % (out["command"], tot_r / 1e9, tot_w / 1e9, out["dram_bytes_per_segment"] / 1e6, out["launches"], tot_t / 1e6, out["source_hash"]))
        f.write("| kernel | launches | read MB | write MB | ms | GB/s |\n|---|---|---|---|---|---|\n")
        for k, e in out["per_kernel"].items():
            f.write("| `%s` | %d | %.1f | %.1f | %.3f | %.0f |\n" % (k[:90], e["launches"], e["read_MB"], e["write_MB"], e["ms"],
                                                                   (e["read_MB"] + e["write_MB"]) / max(e["ms"], 1e-9)))
    print(json.dumps({k: v for k, v in out.items() if k != "per_kernel"}))


if __name__ == "__main__":
    # python tools/measure_traffic.py csv segments [tag title model_args]
    if len(sys.argv) > 3:
        main(sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else "")
    else:
        main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2048)
