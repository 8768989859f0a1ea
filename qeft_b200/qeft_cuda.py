"""Python face of the C ABI with the operator signatures of the reference's ``qeft_cuda`` module.

Reference boundary: ``qeft/kernel/qeft_cuda.cpp:10-27`` binds ``gemm_4bit`` (gemm/gemm_cuda.cu:929),
``gemv_4bit`` (gemv/gemv_cuda.cu:358) and ``gemv_4bit_qeft`` (gemv/gemv_cuda_qeft.cu:392).  The three
functions below keep those names, argument orders and meanings, allocate the output with the torch
caching allocator like the reference (`torch::empty`) and raise ``RuntimeError`` on bad input.  Unlike
the reference they launch on torch's *current* stream of the tensors' device, so they are
multi-device-safe and CUDA-graph-capturable.

The remaining functions are the fused B200 entry points the module layer uses (one launch per
projection group, outliers/bias/o_proj gather inside the kernel).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib

_PDL_DEFAULT = True


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


class _NoGuard:
    __slots__ = ()

    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def _on(t: torch.Tensor):
    """Device guard for the launch: a no-op when the tensor's device is already current (the common case; entering
    ``torch.cuda.device`` costs a few microseconds of host time per call, a third of an eager decode GEMV's)."""
    idx = t.device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(t.device)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("qeft_b200 has no CPU path: all operands must be CUDA tensors")


def _f16(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dtype != torch.float16:
        raise RuntimeError(f"expected scalar type Half for {what} but found {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _flags(pdl: Optional[bool]) -> int:
    return _lib.F_PDL if (_PDL_DEFAULT if pdl is None else pdl) else 0


# ----------------------------------------------------------------------------------------------
# reference signatures
# ----------------------------------------------------------------------------------------------
def gemv_4bit_qeft(in_feats, kernel, scaling_factors, zeros, oweight, m, n, k, group_size):
    """Reference ``gemv_4bit_qeft``: ``oweight`` is the row-pair interleaved ``[N/2, 2r]`` tensor."""
    return gemv_w4(in_feats, kernel, scaling_factors, zeros, oweight, m, n, k, group_size,
                   ow_layout=_lib.OW_INTERLEAVED)


def gemv_4bit(in_feats, kernel, scaling_factors, zeros, m, n, k, group_size):
    """Reference ``gemv_4bit`` (no outlier columns)."""
    return gemv_w4(in_feats, kernel, scaling_factors, zeros, None, m, n, k, group_size, ow_layout=_lib.OW_NONE)


def gemm_4bit(in_feats, kernel, scales, zeros):
    """Reference ``gemm_4bit``: all K int4 columns, no outlier term, no bias."""
    return gemm_w4(in_feats, kernel, scales, zeros, None, None)


# ----------------------------------------------------------------------------------------------
# fused entry points
# ----------------------------------------------------------------------------------------------
def gemv_w4(x, qweight, scales, scaled_zeros, oweight, m, n, k, group_size, *, ow_layout, bias=None,
            x_gather=None, out=None, pdl=None):
    _need_cuda(x, qweight, scales, scaled_zeros, oweight, bias, x_gather)
    x = _f16(x, "in_feats")
    scales = _f16(scales, "scaling_factors")
    scaled_zeros = _f16(scaled_zeros, "zeros")
    if m < 1 or m > 8 or x.numel() != m * k:
        raise RuntimeError("Unsupported batch size for gemv kernel.")
    r = 0
    if oweight is not None:
        oweight = _f16(oweight, "oweight")
        r = oweight.shape[1] // 2 if ow_layout == _lib.OW_INTERLEAVED else oweight.shape[1]
    if out is None:
        out = torch.empty(x.shape[:-1] + (n,), dtype=x.dtype, device=x.device)
    with _on(x):
        st = _lib.load().qeft_gemv_w4(_ptr(x), _ptr(qweight), _ptr(scales), _ptr(scaled_zeros), _ptr(oweight),
                                      ow_layout, _ptr(bias), _ptr(x_gather), _ptr(out), m, n, k, r, group_size,
                                      _flags(pdl), _stream(x))
    _lib.check(st, "qeft_gemv_w4")
    return out


def gemv_w4_multi(x, parts: Sequence[dict], m, k, r, group_size, *, ow_layout, x_gather=None, pdl=None):
    """One launch for several projections that share ``x``.  ``parts``: dicts with qweight, scales,
    scaled_zeros, oweight, bias, N (+ optional preallocated ``y``).  Returns the list of outputs."""
    _need_cuda(x)
    x = _f16(x, "in_feats")
    arr = (_lib.GemvPart * len(parts))()
    outs = []
    for i, p in enumerate(parts):
        y = p.get("y")
        if y is None:
            y = torch.empty(x.shape[:-1] + (p["N"],), dtype=x.dtype, device=x.device)
        outs.append(y)
        arr[i] = _lib.GemvPart(_ptr(p["qweight"]), _ptr(p["scales"]), _ptr(p["scaled_zeros"]),
                               _ptr(p.get("oweight")), _ptr(p.get("bias")), _ptr(y), p["N"])
    with _on(x):
        st = _lib.load().qeft_gemv_w4_multi(_ptr(x), arr, len(parts), ow_layout, _ptr(x_gather), m, k, r, group_size,
                                            _flags(pdl), _stream(x))
    _lib.check(st, "qeft_gemv_w4_multi")
    return outs


def gemv_w4_multi_gather(x, parts: Sequence[dict], m, k, r, group_size, gather, *, ow_layout, x_gather=None, pdl=None):
    """Column-sharded decode: like :func:`gemv_w4_multi`, but every part's result is stored by the kernel into
    every rank's gathered buffer (``gather``: a prepared ``_lib.Gather``; see include/qeft_b200.h).  Returns None."""
    _need_cuda(x)
    x = _f16(x, "in_feats")
    arr = (_lib.GemvPart * len(parts))()
    for i, p in enumerate(parts):
        arr[i] = _lib.GemvPart(_ptr(p["qweight"]), _ptr(p["scales"]), _ptr(p["scaled_zeros"]),
                               _ptr(p.get("oweight")), _ptr(p.get("bias")), None, p["N"])
    with _on(x):
        st = _lib.load().qeft_gemv_w4_multi_gather(_ptr(x), arr, len(parts), ow_layout, _ptr(x_gather), m, k, r,
                                                   group_size, _flags(pdl), C.byref(gather), _stream(x))
    _lib.check(st, "qeft_gemv_w4_multi_gather")


def gather_wait(counter_ptr: int, epoch: torch.Tensor, nranks: int):
    """Order the current stream after the arrival of every rank's slice of a fused-gather launch (include/qeft_b200.h)."""
    with _on(epoch):
        st = _lib.load().qeft_gather_wait(counter_ptr, _ptr(epoch), nranks, _stream(epoch))
    _lib.check(st, "qeft_gather_wait")


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float16:
        return _lib.DT_F16
    if t.dtype == torch.bfloat16:
        return _lib.DT_BF16
    raise RuntimeError(f"expected scalar type Half or BFloat16 but found {t.dtype}")


def _check_packed(what, x_dtype, qweight, scales, scaled_zeros, oweight, bias, K, group_size):
    """The C ABI sees pointers only: check here what it cannot (the advisor's round-1 findings): the scale tables must
    have the shape the kernel will index with ``group_size``, be fp16 like the checkpoint, and the dense outlier block
    must have the activations' dtype (the kernels copy its bits unconverted)."""
    N = qweight.shape[0] * 4
    if qweight.dim() != 2 or qweight.shape[1] != K or qweight.dtype != torch.int16 or not qweight.is_contiguous():
        raise RuntimeError(f"{what}: qweight must be a contiguous int16 [N / 4, {K}] tensor (pack_intweight), found "
                           f"{qweight.dtype} {tuple(qweight.shape)}")
    G = K if group_size in (-1, K) else group_size
    if G <= 0 or K % G != 0:
        raise RuntimeError(f"{what}: group_size {group_size} does not divide K = {K}")
    for nm, t in (("scales", scales), ("scaled_zeros", scaled_zeros)):
        if t.dtype != torch.float16:
            raise RuntimeError(f"{what}: expected scalar type Half for {nm} but found {t.dtype}")
        if tuple(t.shape) != (K // G, N) or not t.is_contiguous():
            raise RuntimeError(f"{what}: {nm} must be a contiguous [{K // G}, {N}] tensor for group_size {G}, "
                               f"found {tuple(t.shape)}")
    if bias is not None and bias.dtype != torch.float16:
        raise RuntimeError(f"{what}: expected scalar type Half for bias but found {bias.dtype}")
    if oweight is not None and oweight.dtype != x_dtype:
        raise RuntimeError(f"{what}: oweight is {oweight.dtype} but the activations are {x_dtype}; "
                           "cast the outlier columns to the activations' dtype")


def _check_out(what, out, shape_numel, dtype, device):
    """A caller-provided result tensor is written through its raw pointer: it must be exactly what the kernel assumes."""
    if out.dtype != dtype or out.numel() != shape_numel or not out.is_contiguous() or out.device != device:
        raise RuntimeError(f"{what}: out must be a contiguous {dtype} tensor of {shape_numel} elements on {device}, found "
                           f"{out.dtype} {tuple(out.shape)} on {out.device}")


def gemm_w4(x, qweight, scales, scaled_zeros, oweight, bias, *, group_size=128, out=None, pdl=None):
    """``y = x . Wdense^T (+ bias)``; ``oweight`` plain ``[N, r]`` fp16/bf16 or None (r = 0)."""
    _need_cuda(x, qweight, scales, scaled_zeros, oweight, bias)
    dt = _dt(x)
    x = x if x.is_contiguous() else x.contiguous()
    K = x.shape[-1]
    M = x.numel() // K
    N = qweight.shape[0] * 4
    r = 0 if oweight is None else oweight.shape[1]
    _check_packed("gemm_w4", x.dtype, qweight, scales, scaled_zeros, oweight, bias, K, group_size)
    if oweight is not None and not oweight.is_contiguous():
        oweight = oweight.contiguous()
    if out is None:
        out = torch.empty(x.shape[:-1] + (N,), dtype=x.dtype, device=x.device)
    else:
        _check_out("gemm_w4", out, M * N, x.dtype, x.device)
    with _on(x):
        st = _lib.load().qeft_gemm_w4(_ptr(x), _ptr(qweight), _ptr(scales), _ptr(scaled_zeros), _ptr(oweight),
                                      _ptr(bias), _ptr(out), M, N, K, r, group_size, dt, _flags(pdl), _stream(x))
    _lib.check(st, "qeft_gemm_w4")
    return out



def gemm_w4_gather(x, qweight, scales, scaled_zeros, oweight, bias, gather, *, group_size=128, pdl=None):
    """Column-sharded prefill: :func:`gemm_w4` on this rank's row slab, every output tile stored by the kernel into
    every rank's gathered ``[M, y_ld]`` buffer (``gather``: a prepared ``_lib.Gather``; include/qeft_b200.h)."""
    _need_cuda(x, qweight, scales, scaled_zeros, oweight, bias)
    dt = _dt(x)
    if not x.is_contiguous():
        raise RuntimeError("gemm_w4_gather: x must be contiguous (it is a gathered buffer)")
    K = x.shape[-1]
    M = x.numel() // K
    N = qweight.shape[0] * 4
    r = 0 if oweight is None else oweight.shape[1]
    _check_packed("gemm_w4_gather", x.dtype, qweight, scales, scaled_zeros, oweight, bias, K, group_size)
    with _on(x):
        st = _lib.load().qeft_gemm_w4_gather(_ptr(x), _ptr(qweight), _ptr(scales), _ptr(scaled_zeros), _ptr(oweight),
                                             _ptr(bias), M, N, K, r, group_size, dt, _flags(pdl), C.byref(gather),
                                             _stream(x))
    _lib.check(st, "qeft_gemm_w4_gather")

def gemm_w4_dx(dy, qweight, scales, scaled_zeros, oweight, K, *, group_size=128, out=None, pdl=None):
    """``dx[M, K] = dy[M, N] . Wdense``."""
    _need_cuda(dy, qweight, scales, scaled_zeros, oweight)
    dt = _dt(dy)
    dy = dy if dy.is_contiguous() else dy.contiguous()
    N = dy.shape[-1]
    M = dy.numel() // N
    r = 0 if oweight is None else oweight.shape[1]
    if qweight.shape[0] * 4 != N:
        raise RuntimeError(f"gemm_w4_dx: dy has {N} features but qweight packs {qweight.shape[0] * 4} rows")
    _check_packed("gemm_w4_dx", dy.dtype, qweight, scales, scaled_zeros, oweight, None, K, group_size)
    if oweight is not None and not oweight.is_contiguous():
        oweight = oweight.contiguous()               # (the kernel's TMA map assumes rows of r elements)
    if out is None:
        out = torch.empty(dy.shape[:-1] + (K,), dtype=dy.dtype, device=dy.device)
    else:
        _check_out("gemm_w4_dx", out, M * K, dy.dtype, dy.device)
    with _on(dy):
        st = _lib.load().qeft_gemm_w4_dx(_ptr(dy), _ptr(qweight), _ptr(scales), _ptr(scaled_zeros), _ptr(oweight),
                                         _ptr(out), M, N, K, r, group_size, dt, _flags(pdl), _stream(dy))
    _lib.check(st, "qeft_gemm_w4_dx")
    return out


def gemm_w4_dx_plan(M: int, N: int, K: int, sm_count: int = 148):
    """Host-only: ``(splits, whole_tiles, ctas)`` of the dX launch for this shape (see ``qeft_gemm_w4_dx_plan``)."""
    import ctypes
    sp, wt, ct = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    st = _lib.load().qeft_gemm_w4_dx_plan(int(M), int(N), int(K), int(sm_count), ctypes.byref(sp), ctypes.byref(wt),
                                          ctypes.byref(ct))
    _lib.check(st, "qeft_gemm_w4_dx_plan")
    return sp.value, wt.value, ct.value


def dow(dy, x, r, *, out=None, accumulate=False, pdl=None):
    """``dow[N, r] (fp32) (+)= dy^T . x[:, K-r:]``."""
    _need_cuda(dy, x)
    dt = _dt(dy)
    if x.dtype != dy.dtype:
        raise RuntimeError("dy and x must have the same dtype")
    dy = dy if dy.is_contiguous() else dy.contiguous()
    x = x if x.is_contiguous() else x.contiguous()
    N, K = dy.shape[-1], x.shape[-1]
    M = dy.numel() // N
    if out is None:
        out = torch.empty((N, r), dtype=torch.float32, device=dy.device)
        accumulate = False
    with _on(dy):
        st = _lib.load().qeft_dow(_ptr(dy), _ptr(x), _ptr(out), M, N, K, r, dt, int(accumulate), _flags(pdl),
                                  _stream(dy))
    _lib.check(st, "qeft_dow")
    return out


def pack_w4(intweight: torch.Tensor) -> torch.Tensor:
    """Device packer: int32 ``[N, K]`` -> int16 ``[N/4, K]`` (bit-exact with the reference's pack_intweight)."""
    _need_cuda(intweight)
    q = intweight.to(torch.int32).contiguous()
    N, K = q.shape
    out = torch.empty((N // 4, K), dtype=torch.int16, device=q.device)
    with _on(q):
        st = _lib.load().qeft_pack_w4(_ptr(q), _ptr(out), N, K, _stream(q))
    _lib.check(st, "qeft_pack_w4")
    return out


def unpack_w4(qweight: torch.Tensor) -> torch.Tensor:
    _need_cuda(qweight)
    qweight = qweight.contiguous()
    Nq, K = qweight.shape
    out = torch.empty((Nq * 4, K), dtype=torch.int32, device=qweight.device)
    with _on(qweight):
        st = _lib.load().qeft_unpack_w4(_ptr(qweight), _ptr(out), Nq * 4, K, _stream(qweight))
    _lib.check(st, "qeft_unpack_w4")
    return out


def dequant_w4(qweight, scales, scaled_zeros, oweight=None, group_size=128, dtype=torch.float16):
    """Dense ``[N, K]`` weight the packed layer stands for (debug / checks; never on the hot path)."""
    _need_cuda(qweight, scales, scaled_zeros, oweight)
    Nq, K = qweight.shape
    N = Nq * 4
    r = 0 if oweight is None else oweight.shape[1]
    out = torch.empty((N, K), dtype=dtype, device=qweight.device)
    dt = _lib.DT_F16 if dtype == torch.float16 else _lib.DT_BF16
    with _on(qweight):
        st = _lib.load().qeft_dequant_w4(_ptr(qweight), _ptr(scales), _ptr(scaled_zeros), _ptr(oweight), _ptr(out),
                                         N, K, r, group_size, dt, _stream(qweight))
    _lib.check(st, "qeft_dequant_w4")
    return out


def interleave_oweight(oweight: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``pack_oweight`` on the device; accepts the fp32 master copy used during fine-tuning."""
    _need_cuda(oweight)
    ow = oweight.detach()
    if ow.dtype not in (torch.float16, torch.float32):
        raise RuntimeError(f"oweight must be fp16 or fp32, found {ow.dtype}")
    ow = ow if ow.is_contiguous() else ow.contiguous()
    N, r = ow.shape
    if out is None:
        out = torch.empty((N // 2, 2 * r), dtype=torch.float16, device=ow.device)
    with _on(ow):
        st = _lib.load().qeft_interleave_oweight(_ptr(ow), _ptr(out), N, r, int(ow.dtype == torch.float32), _stream(ow))
    _lib.check(st, "qeft_interleave_oweight")
    return out


def launch_count() -> int:
    return _lib.launch_count()


# ----------------------------------------------------------------------------------------------
# decode programs: a chain of dependent decode GEMVs as one persistent cooperative launch
# ----------------------------------------------------------------------------------------------
class DecodeProgram:
    """A list of decode stages compiled into one persistent-kernel program (include/qeft_b200.h,
    ``qeft_decode_program_*``; kernel ``csrc/decode_w4.cu``).

    ``stages``: dicts with ``x`` (fp16 ``[m, K]``), ``parts`` (dicts with qweight, scales, scaled_zeros, oweight (plain
    ``[N, r]``), optional bias, ``y`` (fp16 ``[m, N]``), N), ``K``, ``r``, ``G`` and optionally ``x_gather`` (int32 ``[K]``),
    ``norm_weight`` / ``norm_eps`` (RMSNorm on the way in), ``epilogue`` ("swiglu" | "residual") and ``residual``.
    Replaces the reference's per-projection launches of ``gemv_4bit_qeft`` (qeft/qlinear.py:251-263).  The tensors are
    referenced, not copied: the program keeps them alive."""

    _EPI = {None: _lib.EPI_NONE, "none": _lib.EPI_NONE, "swiglu": _lib.EPI_SWIGLU, "residual": _lib.EPI_RESIDUAL}

    def __init__(self, stages: Sequence[dict], m: int = 1):
        if not stages:
            raise RuntimeError("DecodeProgram: no stages")
        self.m = m
        self._keep = []
        arr = (_lib.DecodeStage * len(stages))()
        dev = None
        for i, st in enumerate(stages):
            x = st["x"]
            _need_cuda(x)
            if x.dtype != torch.float16 or not x.is_contiguous():
                raise RuntimeError("expected contiguous scalar type Half for in_feats")
            dev = x.device if dev is None else dev
            d = arr[i]
            d.nparts = len(st["parts"])
            if d.nparts > _lib.GEMV_MAX_PARTS:
                raise RuntimeError("DecodeProgram: at most 4 projections per stage")
            for j, p in enumerate(st["parts"]):
                for k in ("qweight", "scales", "scaled_zeros"):
                    _need_cuda(p[k])
                ow = p.get("oweight")
                if ow is not None and (ow.dtype != torch.float16 or not ow.is_contiguous()):
                    raise RuntimeError("expected contiguous scalar type Half for oweight (plain [N, r] layout)")
                d.parts[j] = _lib.GemvPart(_ptr(p["qweight"]), _ptr(p["scales"]), _ptr(p["scaled_zeros"]), _ptr(ow),
                                           _ptr(p.get("bias")), _ptr(p.get("y")), p["N"])
                self._keep.append(p)
            d.K, d.r, d.G = st["K"], st["r"], st["G"]
            d.x = _ptr(x)
            d.x_gather = _ptr(st.get("x_gather"))
            d.norm_weight = _ptr(st.get("norm_weight"))
            d.norm_eps = float(st.get("norm_eps", 0.0))
            d.epilogue = self._EPI[st.get("epilogue")]
            d.residual = _ptr(st.get("residual"))
            self._keep.append(st)
        self.device = dev
        self.nstages = len(stages)
        handle = C.c_void_p()
        with _on(stages[0]["x"]):
            st = _lib.load().qeft_decode_program_create(arr, len(stages), m, C.byref(handle))
        _lib.check(st, "qeft_decode_program_create")
        self._h = handle

    def set_ranks(self, nranks: int, rank: int, barrier_ptrs: Sequence[int]):
        """Column-sharded program: ``barrier_ptrs[p]`` = address of rank p's barrier word (uint32, zero, peer-mapped)."""
        arr = (C.c_void_p * nranks)(*[C.c_void_p(int(a)) for a in barrier_ptrs])
        _lib.check(_lib.load().qeft_decode_program_set_ranks(self._h, nranks, rank, arr), "qeft_decode_program_set_ranks")
        self.nranks, self.rank = nranks, rank

    def shard(self, stage: int, part: int, y_full_ptrs: Sequence[int]):
        """The part's output is this rank's slice of a gathered row: ``y_full_ptrs[p]`` = base of rank p's copy."""
        arr = (C.c_void_p * len(y_full_ptrs))(*[C.c_void_p(int(a)) for a in y_full_ptrs])
        _lib.check(_lib.load().qeft_decode_program_shard(self._h, stage, part, arr), "qeft_decode_program_shard")

    def run(self, begin: int = 0, end: Optional[int] = None):
        end = self.nstages if end is None else end
        with (_NO_GUARD if self.device.index in (None, torch.cuda.current_device()) else torch.cuda.device(self.device)):
            st = _lib.load().qeft_decode_program_run(self._h, begin, end, 0, torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(st, "qeft_decode_program_run")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().qeft_decode_program_destroy(h)
            except Exception:
                pass
            self._h = None
