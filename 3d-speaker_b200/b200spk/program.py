"""Host-side compiler from an eval-mode network (a state_dict) to the library's fused-op list.

Folds BatchNorm (running stats) into per-channel scale/shift, repacks convolution weights to
``[Cout][KH][KW][Cin]`` (channels-last implicit-GEMM order), uploads parameters once through
``spk_model_add_param`` and registers one op list per frame count through
``spk_model_set_program``.  Activations are channels-last buffers in a caller-owned workspace.
"""
import collections
import ctypes as C
import os

import torch

from . import _lib

BN_EPS = 1e-5


def fold_bn(sd, prefix, affine=True, eps=BN_EPS):
    """Eval-mode BatchNorm as y = x*scale + shift."""
    var = sd[prefix + ".running_var"].double()
    mean = sd[prefix + ".running_mean"].double()
    scale = 1.0 / torch.sqrt(var + eps)
    if affine:
        scale = scale * sd[prefix + ".weight"].double()
    shift = -mean * scale
    if affine:
        shift = shift + sd[prefix + ".bias"].double()
    return scale.float(), shift.float()


def pack_conv2d(w):
    """[Cout, Cin, KH, KW] -> [Cout, KH, KW, Cin]."""
    return w.permute(0, 2, 3, 1).contiguous()


def pack_conv1d(w):
    """[Cout, Cin, K] -> [Cout, 1, K, Cin]."""
    return w.permute(0, 2, 1).contiguous()


class Model:
    """Owns a ``spk_model_t`` handle, its uploaded parameters and per-T programs."""

    def __init__(self, precision, device):
        self.precision = precision
        self.device = torch.device(device)
        self.act_dtype = _lib.DT_BF16 if precision == _lib.PREC_BF16 else _lib.DT_F32
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().spk_model_create(C.byref(h), precision))
        self.handle = h
        self.programs = {}
        self._ws = {}
        # CUDA-graph replay of small-batch forwards (the reference call sites batch 64 windows: ~125 launches of
        # a few microseconds each, where host launch cost and inter-kernel gaps are most of the call)
        self.graph_max_batch = int(os.environ.get("SPK_GRAPH_MAX_BATCH", "128"))
        self._graphs = collections.OrderedDict()

    def close(self):
        if self.handle is not None and self.handle.value:
            _lib.lib().spk_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def param(self, t):
        t = t.detach().to(torch.float32).contiguous().cpu()
        with torch.cuda.device(self.device):
            return int(_lib.check(_lib.lib().spk_model_add_param(self.handle, C.c_void_p(t.data_ptr()), t.numel())))

    def set_program(self, T, prog):
        bufs = (_lib.SpkBuf * len(prog.bufs))(*prog.bufs)
        ops = (_lib.SpkOp * len(prog.ops))(*prog.ops)
        _lib.check(_lib.lib().spk_model_set_program(self.handle, int(T), bufs, len(prog.bufs), ops, len(prog.ops)))
        self.programs[int(T)] = prog

    def workspace(self, T, chunk, fine):
        key = (int(T), int(chunk), int(fine))
        ws = self._ws.get(key)
        if ws is None:
            n = int(_lib.check(_lib.lib().spk_model_workspace_bytes(self.handle, int(T), int(chunk), int(fine))))
            self._ws = {}            # keep one workspace alive (captured graphs hold their own reference)
            ws = torch.empty(max(n, 16), dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def forward(self, T, feats, emb_dim, chunk, fine=0):
        """feats [B,T,F] f32 contiguous CUDA -> emb [B,E] f32.  ``chunk``/``fine``: coarse and fine
        sub-batch sizes (see spk_model_forward)."""
        B = feats.shape[0]
        emb = torch.empty((B, emb_dim), dtype=torch.float32, device=feats.device)
        if B == 0:          # an empty batch is a valid call (the reference returns [0, E]); nothing to launch
            return emb
        chunk = max(1, min(int(chunk), max(B, 1)))
        fine = chunk if fine <= 0 else max(1, min(int(fine), chunk))
        self._last_chunk = (chunk, fine)
        if B <= self.graph_max_batch and not torch.cuda.is_current_stream_capturing():
            return self._forward_graph(T, feats, emb_dim, chunk, fine)
        ws = self.workspace(T, chunk, fine)
        self._last_ws = ws
        self._launch(T, feats, emb, ws, chunk, fine)
        return emb

    def _launch(self, T, feats, emb, ws, chunk, fine):
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().spk_model_forward(self.handle, int(T), C.c_void_p(feats.data_ptr()), feats.shape[0],
                                                    C.c_void_p(emb.data_ptr()), C.c_void_p(ws.data_ptr()),
                                                    ws.numel(), chunk, fine, _lib.current_stream_ptr()))

    def _forward_graph(self, T, feats, emb_dim, chunk, fine):
        """One captured graph per (T, batch): static input / output / workspace buffers, the forward's launches
        (programmatic-dependent-launch edges included) replayed with one call."""
        key = (int(T), tuple(feats.shape), int(emb_dim), chunk, fine)
        ent = self._graphs.get(key)
        if ent is None:
            ws = self.workspace(T, chunk, fine)
            sfeats = torch.empty_like(feats)
            semb = torch.empty((feats.shape[0], emb_dim), dtype=torch.float32, device=feats.device)
            sfeats.copy_(feats)
            self._launch(T, sfeats, semb, ws, chunk, fine)          # eager once: attribute / tensor-map caches, warm-up
            torch.cuda.current_stream(self.device).synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.lib().spk_launch_count()
            with torch.cuda.graph(graph):
                self._launch(T, sfeats, semb, ws, chunk, fine)
            ent = (graph, sfeats, semb, ws, int(_lib.lib().spk_launch_count() - n0))
            self._graphs[key] = ent
            while len(self._graphs) > 8:
                self._graphs.popitem(last=False)
        else:
            self._graphs.move_to_end(key)
        graph, sfeats, semb, ws, n_launch = ent
        self._last_ws = ws              # read_buffer looks at the workspace the last forward really used
        sfeats.copy_(feats)
        graph.replay()
        _lib.lib().spk_add_launches(n_launch)       # the replay ran them; keep the library's launch counter truthful
        return semb.clone()

    def read_buffer(self, T, name, n_segments):
        """Widened copy of a named workspace buffer as the last sub-batch of the last forward
        left it (per-layer parity checks)."""
        chunk, fine = self._last_chunk
        prog = self.programs[int(T)]
        bid = prog.names[name]
        n = prog.bufs[bid].elems * n_segments
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        ws = self._last_ws
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().spk_model_read_buffer(self.handle, int(T), bid, chunk, fine, C.c_void_p(ws.data_ptr()),
                                                        C.c_void_p(out.data_ptr()), n, _lib.current_stream_ptr()))
        return out


class Program:
    """Op list + buffer table under construction."""

    def __init__(self, feat_elems, emb_elems):
        self.bufs = [_lib.SpkBuf(int(feat_elems), _lib.DT_F32, 0), _lib.SpkBuf(int(emb_elems), _lib.DT_F32, 0)]
        self.names = {"feats": 0, "emb": 1}
        self.ops = []
        self.phase = 0           # phase given to ops added from now on

    def buf(self, name, elems, dtype):
        self.bufs.append(_lib.SpkBuf(int(elems), int(dtype), 0))
        self.names[name] = len(self.bufs) - 1
        return len(self.bufs) - 1

    def op(self, kind, **kw):
        kw.setdefault("phase", self.phase)
        self.ops.append(_lib.make_op(kind, **kw))


def conv_out(n, k, s, p, d=1):
    return (n + 2 * p - d * (k - 1) - 1) // s + 1


class EngineBase:
    """Shared host-side plumbing of a compiled network: parameter cache (uploaded once, shared by
    every per-T program) and lazy per-T compilation."""

    def __init__(self, module, model):
        self.m = module
        self.model = model
        self.sd = {k: v.detach().float().cpu() for k, v in module.state_dict().items()}
        self._pcache = {}
        self._compiled = set()

    def close(self):
        self.model.close()

    def _p(self, key, fn):
        if key not in self._pcache:
            self._pcache[key] = self.model.param(fn())
        return self._pcache[key]

    def _bn(self, prefix, affine=True):
        s = self._p(("bn_s", prefix), lambda: fold_bn(self.sd, prefix, affine)[0])
        b = self._p(("bn_b", prefix), lambda: fold_bn(self.sd, prefix, affine)[1])
        return s, b

    def _w2d(self, key):
        return self._p(("w", key), lambda: pack_conv2d(self.sd[key]))

    def _w1d(self, key):
        return self._p(("w", key), lambda: pack_conv1d(self.sd[key]))

    def _raw(self, key):
        return self._p(("raw", key), lambda: self.sd[key].reshape(-1))

    def default_chunks(self, T):
        raise NotImplementedError

    def compile(self, T):
        raise NotImplementedError

    def run(self, feats, emb_dim, chunk=None):
        T = feats.shape[1]
        if T not in self._compiled:
            self.compile(T)
            self._compiled.add(T)
        coarse, fine = self.default_chunks(T)
        if chunk:
            coarse, fine = (chunk if isinstance(chunk, (tuple, list)) else (chunk, min(fine, chunk)))
        return self.model.forward(T, feats, emb_dim, coarse, fine)


class EngineModule(torch.nn.Module):
    """nn.Module side of a drop-in model: holds the reference-layout parameters, owns the compiled
    engine and invalidates it when weights or device change."""
    engine_cls = None

    def _init_engine(self, precision, chunk):
        self.precision = precision
        self.chunk = chunk
        self._engine = None
        self._engine_key = None

    def invalidate(self):
        """Drop the compiled engine (call after editing weights in place; load_state_dict and
        .to()/.cuda() do it automatically)."""
        if getattr(self, "_engine", None) is not None:
            self._engine.close()
        self._engine = None
        self._engine_key = None

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate()
        return out

    def _get_engine(self, device):
        key = (str(device), self.precision)
        if self._engine is None or self._engine_key != key:
            self.invalidate()
            prec = _lib.PREC_BF16 if self.precision in ("bf16", "bfloat16") else _lib.PREC_F32
            self._engine = self.engine_cls(self, Model(prec, device))
            self._engine_key = key
        return self._engine

    def _run(self, x, feat_dim, emb_dim):
        assert not self.training, "b200spk models are inference engines: call .eval()"
        assert x.dim() == 3 and x.shape[2] == feat_dim
        if not x.is_cuda:
            raise RuntimeError("b200spk models need CUDA tensors (no CPU fallback); move the model and "
                               "features to a B200 device")
        x = x.to(torch.float32).contiguous()
        return self._get_engine(x.device).run(x, emb_dim, self.chunk)
