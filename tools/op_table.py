"""Op list of a network next to an ncu launch list: python tools/op_table.py eres 298 [traffic.csv]
Builds the program on the CPU mock (no GPU needed) and prints one line per op with its GEMM shape per segment."""
import csv, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-speaker_b200"), os.path.join(ROOT, "tests")]
import b200spk
from b200spk import _lib, campplus, ecapa_tdnn, eres2netv2
from test_program_cpu import MockModel

name, T = sys.argv[1], int(sys.argv[2])
if name == "eres":
    mod, eng = b200spk.ERes2NetV2(precision="bf16"), eres2netv2._Engine
elif name == "eres_w24":
    mod, eng = b200spk.ERes2NetV2(baseWidth=24, scale=4, expansion=4, precision="bf16"), eres2netv2._Engine
elif name == "ecapa":
    mod, eng = b200spk.ECAPA_TDNN(80, channels=[1024, 1024, 1024, 1024, 3072], precision="bf16"), ecapa_tdnn._Engine
else:
    mod, eng = b200spk.CAMPPlus(embedding_size=512, precision="bf16"), campplus._Engine
mm = MockModel(_lib.PREC_BF16)
e = eng(mod.eval(), mm)
e.compile(T)
prog = mm.programs[T]
kinds = {v: k for k, v in vars(_lib).items() if k.startswith("OP_")}
tot = 0
for i, op in enumerate(prog.ops):
    M = op.Ho * op.Wo
    fl = 2.0 * M * op.KH * op.KW * op.Cin * op.Cout if op.kind in (_lib.OP_CONV, _lib.OP_CAM_LOCAL) else 0
    tot += fl
    print("%3d %-14s in %dx%dx%d k%dx%d s%d -> %dx%dx%d  M=%d K=%d N=%d  %.1f MFLOP  pro=%d res=%d ph=%d" % (
        i, kinds.get(op.kind, "?"), op.H, op.W, op.Cin, op.KH, op.KW, op.sh, op.Ho, op.Wo, op.Cout, M, op.KH * op.KW * op.Cin, op.Cout,
        fl / 1e6, op.pro_scale >= 0, op.res_buf >= 0, op.phase))
print("total GFLOP per segment (as executed, padded):", tot / 1e9)
