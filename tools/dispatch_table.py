"""Which kernel family takes which conv of which network (SPK_TRACE_DISPATCH=1): prints a count table per model / T."""
import collections, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("campplus", 1.5, 8), ("campplus", 3.0, 8), ("campplus", 10.0, 4), ("eres", 3.0, 8), ("eres_w24", 3.0, 8), ("eres", 1.5, 8), ("ecapa", 10.0, 8), ("ecapa", 1.5, 8)]
for model, sec, n in CASES:
    env = dict(os.environ, SPK_TRACE_DISPATCH="1", SPK_GRAPH_MAX_BATCH="0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_forward.py"), "--model", model, "--seconds", str(sec), "--segments", str(n), "--iters", "1"],
                         capture_output=True, text=True, env=env)
    c = collections.Counter(l.split()[2] for l in out.stderr.splitlines() if l.startswith("[spk dispatch]"))
    odd = sorted(set(l for l in out.stderr.splitlines() if l.startswith("[spk dispatch]") and l.split()[2] in ("tc2_gather", "slab2", "simt")))
    print("%-9s %4.1f s: %s" % (model, sec, dict(c)))
    for l in odd:
        print("      ", l[15:])
