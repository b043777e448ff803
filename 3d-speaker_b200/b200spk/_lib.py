"""ctypes binding of libb200spk.so (include/b200spk.h).  No torch extension, no JIT.

The library is built in-tree by ``3d-speaker_b200/build.py``.  Loading fails loudly when the
.so is missing; every compute call fails loudly (``SpkError``) when there is no sm_100 device.
There is no CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200spk.so")

SPK_OK = 0
ERR_NAMES = {-1: "SPK_ERR_INVALID", -2: "SPK_ERR_CUDA", -3: "SPK_ERR_UNSUPPORTED",
             -4: "SPK_ERR_WORKSPACE", -5: "SPK_ERR_NO_DEVICE", -6: "SPK_ERR_KERNEL"}
PREC_F32, PREC_BF16 = 0, 1
DT_F32, DT_BF16 = 0, 1
OP_STEM, OP_CONV, OP_CAM_GATE, OP_STATS_POOL, OP_AFF_BLEND, OP_CAM_LOCAL, OP_SE_SCALE, OP_ASP_POOL, OP_STEM_BLOCK = 1, 2, 3, 4, 5, 6, 7, 8, 9
ACT_NONE, ACT_RELU, ACT_CLAMP20, ACT_SILU, ACT_TANH = 0, 1, 2, 3, 4


class SpkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERR_NAMES.get(code, "SPK_ERR"), code, msg))
        self.code = code


class SpkBuf(C.Structure):
    _fields_ = [("elems", C.c_int64), ("dtype", C.c_int32), ("reserved", C.c_int32)]


class SpkOp(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("in_buf", C.c_int32), ("in_ld", C.c_int32), ("in_choff", C.c_int32),
        ("out_buf", C.c_int32), ("out_ld", C.c_int32), ("out_choff", C.c_int32),
        ("res_buf", C.c_int32), ("res_ld", C.c_int32), ("res_choff", C.c_int32),
        ("gate_buf", C.c_int32), ("gate_win", C.c_int32),
        ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32),
        ("Ho", C.c_int32), ("Wo", C.c_int32), ("Cout", C.c_int32),
        ("KH", C.c_int32), ("KW", C.c_int32), ("sh", C.c_int32), ("sw", C.c_int32),
        ("ph", C.c_int32), ("pw", C.c_int32), ("dh", C.c_int32), ("dw", C.c_int32),
        ("w", C.c_int32),
        ("pro_scale", C.c_int32), ("pro_shift", C.c_int32), ("pro_relu", C.c_int32),
        ("epi_scale", C.c_int32), ("epi_shift", C.c_int32), ("act", C.c_int32),
        ("aux", C.c_int32 * 4), ("iaux", C.c_int32 * 4), ("faux", C.c_float * 2),
        ("phase", C.c_int32), ("reserved", C.c_int32),
    ]


def make_op(kind, **kw):
    op = SpkOp()
    op.kind = kind
    for f in ("res_buf", "gate_buf", "w", "pro_scale", "pro_shift", "epi_scale", "epi_shift"):
        setattr(op, f, -1)
    op.aux[:] = [-1, -1, -1, -1]
    op.KH = op.KW = op.sh = op.sw = op.dh = op.dw = 1
    for k, v in kw.items():
        if k in ("aux", "iaux"):
            arr = getattr(op, k)
            for i, x in enumerate(v):
                arr[i] = int(x)
        elif k == "faux":
            for i, x in enumerate(v):
                op.faux[i] = float(x)
        else:
            if not hasattr(op, k):
                raise AttributeError(k)
            setattr(op, k, int(v))
    return op


_SIGS = {
    "spk_abi_version": (C.c_int, []),
    "spk_last_error": (C.c_char_p, []),
    "spk_device_check": (C.c_int, [C.c_int]),
    "spk_launch_count": (C.c_int64, []),
    "spk_add_launches": (None, [C.c_int64]),
    "spk_fbank_num_frames": (C.c_int64, [C.c_int64]),
    "spk_fbank_set_tables": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "spk_fbank_set_repair": (C.c_int, [C.c_float, C.c_int]),
    "spk_fbank_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int,
                                C.c_void_p]),
    "spk_fbank_i16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int,
                                C.c_void_p]),
    "spk_fbank_windows": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "spk_segment_mean": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "spk_fbank_host_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int]),
    "spk_ahc_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64]),
    "spk_ahc": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "spk_model_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "spk_model_destroy": (C.c_int, [C.c_void_p]),
    "spk_model_add_param": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64]),
    "spk_model_set_program": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(SpkBuf), C.c_int32, C.POINTER(SpkOp),
                                        C.c_int32]),
    "spk_model_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64]),
    "spk_model_forward": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "spk_model_read_buffer": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_void_p,
                                        C.c_void_p, C.c_int64, C.c_void_p]),
    "spk_affinity_laplacian": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                         C.c_int64, C.c_void_p]),
    "spk_affinity_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64]),
    "spk_eig_smallest": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_void_p]),
    "spk_eig_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int32]),
    "spk_lanczos_extend": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "spk_lanczos_ritz": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                   C.c_void_p]),
    "spk_lanczos_max_dim": (C.c_int32, [C.c_int64, C.c_int32]),
    "spk_kmeans": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_float,
                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "spk_kmeans_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32]),
    "spk_cosine_pairs": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                   C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    """The loaded library (loads on first use; raises if the .so has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libb200spk.so not built: run `python 3d-speaker_b200/build.py` "
                              "(there is no fallback implementation)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def exported_names():
    return sorted(_SIGS)


def last_error():
    return (lib().spk_last_error() or b"").decode("utf-8", "replace")


def check(code):
    """Raise SpkError for a negative return code; pass non-negative values through."""
    if code < 0:
        raise SpkError(int(code), last_error())
    return code


def current_stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
