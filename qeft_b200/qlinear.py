"""Packed 4-bit linear layer with dense, trainable fp16 outlier columns (QEFT), B200-native.

Mirrors the module API of the reference's ``qeft/qlinear.py`` -- ``QuantLinear`` (:123-330), the autograd
functions ``QuantMatMulQEFT`` / ``QuantMatMul`` (:13-68) and the packers ``pack_intweight`` /
``pack_oweight`` (:70-121) -- with the same constructor arguments, buffer names, shapes and dtypes, so
reference checkpoints load unchanged.  What differs is below the API:

* every forward is ONE kernel launch on the current stream (outlier columns, bias and the o_proj
  ``index_select`` are fused into the GEMV / GEMM; the reference launches 2-4);
* the backward implements the math the reference intended:
  ``dX = dY . Wdense`` and ``dOW = dY^T . X[:, K-r:]`` (fp32);
* ``oweight_interleaved`` is refreshed from ``oweight`` after fine-tuning (`refresh_oweight_interleaved`),
  which the reference forgets (utils/modelutils.py:185-198).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib, qeft_cuda
from .reorder import sparse_to_dense_ids

__all__ = ["QuantLinear", "QuantMatMul", "QuantMatMulQEFT", "pack_intweight", "pack_oweight", "unpack_intweight"]


def _nvtx_wrapped(fn, label):
    """``QEFT_NVTX=1``: an NVTX range around every forward, named after the dispatch target and the layer (the ranges
    the reference carries commented out at qlinear.py:270,276,301,312,329).  Off by default: ``forward`` stays the bound
    method itself."""
    def forward(x):
        torch.cuda.nvtx.range_push(label)
        try:
            return fn(x)
        finally:
            torch.cuda.nvtx.range_pop()
    forward.__name__ = fn.__name__
    return forward


# --------------------------------------------------------------------------------------------------
# packers (offline path; torch index arithmetic, run wherever the tensor lives)
# --------------------------------------------------------------------------------------------------
def _tile_positions(device):
    """(int16 index e, nibble i) of weight (row-in-tile j, column-in-tile kk), both ``[4, 64]``."""
    j = torch.arange(4, device=device).view(4, 1)
    kk = torch.arange(64, device=device).view(1, 64)
    half, k32 = kk // 32, kk % 32
    nib = (k32 % 2) * 4 + k32 // 8          # slot inside the 32-bit word
    word = (k32 % 8) // 2
    lin = 64 * j + 32 * half + 8 * word + nib
    return lin // 4, lin % 4


def pack_intweight(unpacked_qweight: torch.Tensor, interleave: int = 4, kstride: int = 64) -> torch.Tensor:
    """int ``[N, K]`` (0..15) -> int16 ``[N/4, K]``; same bytes as the reference's ``pack_intweight``."""
    if interleave != 4 or kstride != 64:
        raise ValueError("the packed layout is defined for interleave=4, kstride=64")
    q = unpacked_qweight
    N, K = q.shape
    if q.is_cuda:
        return qeft_cuda.pack_w4(q)
    e, i = _tile_positions(q.device)
    # source (j, kk) of nibble i of int16 number e: invert the position map once
    src = torch.empty((64, 4), dtype=torch.long, device=q.device)
    src[e.reshape(-1), i.reshape(-1)] = torch.arange(256, device=q.device)
    tiles = q.to(torch.int64).view(N // 4, 4, K // 64, 64).permute(0, 2, 1, 3).reshape(N // 4, K // 64, 256)
    nib = tiles[:, :, src.reshape(-1)].view(N // 4, K // 64, 64, 4)
    # OR, not add, and no masking: an out-of-range value spills into its neighbours exactly like the
    # reference's `a | b << 4 | c << 8 | d << 12` on unclamped ints (qlinear.py:109-114)
    out = nib[..., 0] | (nib[..., 1] << 4) | (nib[..., 2] << 8) | (nib[..., 3] << 12)
    out = out & 0xFFFF
    out = torch.where(out >= 0x8000, out - 0x10000, out)
    return out.view(N // 4, K).to(torch.int16).contiguous()


def unpack_intweight(qweight: torch.Tensor) -> torch.Tensor:
    """Inverse of :func:`pack_intweight` -> int32 ``[N, K]``."""
    if qweight.is_cuda:
        return qeft_cuda.unpack_w4(qweight)
    Nq, K = qweight.shape
    e, i = _tile_positions(qweight.device)
    tiles = (qweight.to(torch.int64) & 0xFFFF).view(Nq, K // 64, 64)
    vals = tiles[:, :, e.reshape(-1)].view(Nq, K // 64, 4, 64)
    vals = (vals >> (4 * i).view(1, 1, 4, 64)) & 0xF
    return vals.permute(0, 2, 1, 3).reshape(Nq * 4, K).to(torch.int32)


def pack_oweight(oweight: torch.Tensor, interleave: int = 4) -> torch.Tensor:
    """fp16 ``[N, r]`` -> ``[N/2, 2r]``: rows n and n+4 of every 8-row block interleaved element-wise per 32 columns."""
    if interleave != 4:
        raise ValueError("interleave must be 4")
    if oweight.is_cuda and oweight.dtype in (torch.float16, torch.float32):
        return qeft_cuda.interleave_oweight(oweight)
    N, r = oweight.shape
    v = oweight.reshape(N // 8, 2, 4, r // 32, 32)            # [block, s, j, c, t]
    return v.permute(0, 2, 3, 4, 1).reshape(N // 2, 2 * r).contiguous()


# --------------------------------------------------------------------------------------------------
# autograd
# --------------------------------------------------------------------------------------------------
class QuantMatMulQEFT(torch.autograd.Function):
    """``y = x . Wdense^T + bias`` with the outlier columns trainable (reference: qlinear.py:13-44).

    ``group_size`` is an extra trailing argument (the reference's kernels hard-code 128, gemm_cuda.cu:954): the
    scales of a layer packed with another group size are indexed with it in forward, dX and nowhere else.
    The launches are made WITHOUT programmatic dependent launch: ``oweight.to(fp16)`` runs as the kernel right
    before the GEMM and the GEMM's dequant warps read ``oweight`` without waiting for the previous grid."""

    @staticmethod
    def forward(ctx, x, oweight, qweight, scales, scaled_zeros, n_out, bias, name, group_size=128):
        dtype = scales.dtype
        xh = x.to(dtype)
        ow_h = oweight.to(dtype)
        y = qeft_cuda.gemm_w4(xh, qweight, scales, scaled_zeros, ow_h, bias, group_size=group_size, pdl=False)
        ctx.n_out = n_out
        ctx.group_size = group_size
        ctx.in_dtype = x.dtype
        ctx.ow_dtype = oweight.dtype
        # only the r outlier activations are needed for dOW: keep a compact copy, not a view of x
        x_out = xh[..., xh.shape[-1] - n_out:].contiguous() if ctx.needs_input_grad[1] else None
        ctx.save_for_backward(x_out, ow_h, qweight, scales, scaled_zeros)
        ctx.K = xh.shape[-1]
        return y

    @staticmethod
    def backward(ctx, grad_output):
        x_out, ow_h, qweight, scales, scaled_zeros = ctx.saved_tensors
        dy = grad_output.to(scales.dtype)
        grad_x = grad_ow = None
        if ctx.needs_input_grad[0]:
            grad_x = qeft_cuda.gemm_w4_dx(dy, qweight, scales, scaled_zeros, ow_h, ctx.K, group_size=ctx.group_size,
                                          pdl=False).to(ctx.in_dtype)
        if ctx.needs_input_grad[1]:
            grad_ow = qeft_cuda.dow(dy, x_out, ctx.n_out, pdl=False).to(ctx.ow_dtype)
        return grad_x, grad_ow, None, None, None, None, None, None, None


class QuantMatMul(torch.autograd.Function):
    """No outlier columns (reference: qlinear.py:46-68)."""

    @staticmethod
    def forward(ctx, x, qweight, scales, scaled_zeros, n_out, bias, name, group_size=128):
        dtype = scales.dtype
        y = qeft_cuda.gemm_w4(x.to(dtype), qweight, scales, scaled_zeros, None, bias, group_size=group_size, pdl=False)
        ctx.save_for_backward(qweight, scales, scaled_zeros)
        ctx.K = x.shape[-1]
        ctx.group_size = group_size
        ctx.in_dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, grad_output):
        qweight, scales, scaled_zeros = ctx.saved_tensors
        grad_x = None
        if ctx.needs_input_grad[0]:
            grad_x = qeft_cuda.gemm_w4_dx(grad_output.to(scales.dtype), qweight, scales, scaled_zeros, None,
                                          ctx.K, group_size=ctx.group_size, pdl=False).to(ctx.in_dtype)
        return grad_x, None, None, None, None, None, None, None


# --------------------------------------------------------------------------------------------------
# module
# --------------------------------------------------------------------------------------------------
class QuantLinear(nn.Module):
    """Same constructor and buffers as the reference ``QuantLinear`` (qlinear.py:125-178)."""

    GEMV_MAX_ROWS = 8   # reference: `seq_len < 8` -> GEMV (qlinear.py:252)

    def __init__(self, bits, infeatures, outfeatures, bias, dtype, outlierfeatures, group_size, reorder, name):
        super().__init__()
        assert bits in [4], "Only 4 bits is supported."
        assert dtype == torch.float16, "Only fp16 is supported."
        self.bits = bits
        self.infeatures = infeatures
        self.outfeatures = outfeatures
        self.outlierfeatures = outlierfeatures
        self.group_size = group_size if group_size != -1 else infeatures
        self.interleave = 4
        assert infeatures % self.group_size == 0
        assert outfeatures % (32 // bits) == 0
        numgroup = infeatures // self.group_size

        self.register_buffer("qweight", torch.empty((outfeatures // self.interleave, infeatures), dtype=torch.int16))
        self.register_buffer("scales", torch.empty((numgroup, outfeatures), dtype=dtype))
        self.register_buffer("scaled_zeros", torch.empty((numgroup, outfeatures), dtype=dtype))
        if bias:
            self.register_buffer("bias", torch.empty((outfeatures,), dtype=torch.float16))
        else:
            self.bias = None
        if outlierfeatures > 0:
            self.register_buffer("oweight", torch.empty((outfeatures, outlierfeatures), dtype=dtype))
            self.register_buffer("oweight_interleaved", torch.empty((outfeatures // 2, outlierfeatures * 2), dtype=dtype))
            self.register_buffer("outlieridx", torch.zeros((outlierfeatures,), dtype=torch.int))
        self.faster = True
        self.dtype = dtype
        self.name = name
        self.reorder = reorder
        self.training = False
        self.gemv = self.gemm = self.matmul = None

    # ---- offline ---------------------------------------------------------------------------------
    @torch.no_grad()
    def pack(self, linear, scales, zeros, outlieridx: torch.Tensor, sym: bool = False):
        """Quantise ``linear.weight`` with per-group ``scales``/``zeros`` ``[N, K/G]`` and fill the buffers.

        Same arithmetic as the reference (qlinear.py:180-215): round without clamp, outlier columns carry
        the zero point of their group, fp16 ``scales`` and ``scaled_zeros = -(zeros*scales)``, the last
        ``outlierfeatures`` columns of the (already OGR-reordered) weight become ``oweight``.  With ``sym``
        the zero points are shifted by 8 in place, as the reference does.
        """
        self.sym = sym
        if sym:
            zeros += 2 ** (self.bits - 1)
        if linear.bias is not None:
            self.bias = linear.bias.detach().to(self.dtype)
        K, G, r = self.infeatures, self.group_size, self.outlierfeatures
        rep = 1 if G == K else G
        w = linear.weight.data
        sz = zeros * scales
        q = torch.round((w + torch.repeat_interleave(sz, rep, dim=1)) / torch.repeat_interleave(scales, rep, dim=1))
        q = q.to(torch.int32)
        if r > 0:
            cols = torch.arange(K - r, K, device=q.device)
            q[:, K - r:] = zeros[:, cols // G].to(torch.int32)
        self.qweight = pack_intweight(q, interleave=4, kstride=64)
        self.scales = scales.t().contiguous().to(self.dtype)
        self.scaled_zeros = -sz.t().contiguous().to(self.dtype)
        if r > 0:
            self.oweight = w[:, K - r:].clone()
            self.oweight_interleaved = pack_oweight(self.oweight, interleave=4)
            self.outlieridx = outlieridx

    # ---- load time -------------------------------------------------------------------------------
    def set_kernel(self, training=False):
        """Bind the kernels for this layer's format (reference: qlinear.py:217-237)."""
        _lib.load()   # fail here, loudly, if the CUDA library is missing
        self.training = training
        r = self.outlierfeatures
        if r > 0:
            # The tensor-core path consumes the dense block in 64-column slabs.  For r % 64 != 0 the slab is
            # completed at the FRONT with the dequantised int4 columns K-r_pad..K-r-1, so the result is the
            # same as treating exactly r columns as outliers (the reference zero-pads, qlinear.py:221-222,
            # which only works for r % 64 == 0).  `oweight` itself keeps its checkpoint shape [N, r].
            pad = (-r) % 64
            self._ow_pad = pad
            if pad:
                K = self.infeatures
                dense = qeft_cuda.dequant_w4(self.qweight, self.scales, self.scaled_zeros, None, self.group_size)
                front = dense[:, K - r - pad:K - r]
                self.register_buffer("oweight_gemm", torch.cat([front, self.oweight.to(self.dtype)], dim=1).contiguous(),
                                     persistent=False)
                if training:
                    raise NotImplementedError("fine-tuning needs outlierfeatures % 64 == 0")
            self.gemv = qeft_cuda.gemv_4bit_qeft
            self.gemm = qeft_cuda.gemm_w4
            self.forward = self.forward_outlier
            if "o_proj" in self.name or "out_proj" in self.name:
                ids = sparse_to_dense_ids(self.outlieridx, self.infeatures)
                self.register_buffer("reorder_ids", ids)
                self.register_buffer("reorder_ids32", ids.to(torch.int32), persistent=False)
                self.forward = self.forward_outlier_out_proj
            if training:
                self.matmul = QuantMatMulQEFT.apply
        else:
            self.gemv = qeft_cuda.gemv_4bit
            self.gemm = qeft_cuda.gemm_w4
            self.forward = self.forward_normal
            if training:
                self.matmul = QuantMatMul.apply
        if os.environ.get("QEFT_NVTX", "0") == "1":
            self.forward = _nvtx_wrapped(self.forward, f"QuantLinear.{self.forward.__name__}:{self.name}")

    def set_for_wct(self):
        """Freeze the packed weight, make the outlier columns an fp32 trainable parameter (qlinear.py:239-242)."""
        self.qweight = nn.Parameter(self.qweight, requires_grad=False)
        if self.outlierfeatures > 0:
            self.oweight = nn.Parameter(self.oweight.to(dtype=torch.float), requires_grad=True)

    @torch.no_grad()
    def refresh_oweight_interleaved(self):
        """Re-derive the GEMV copy of the outlier columns after ``oweight`` changed (fine-tuning / WCT load)."""
        if self.outlierfeatures > 0:
            r = self.outlierfeatures
            src = self.oweight.detach()[:, -r:]
            self.oweight_interleaved = pack_oweight(src.contiguous(), interleave=4).to(self.dtype)

    # ---- forwards --------------------------------------------------------------------------------
    def _oweight_plain(self, dtype=None):
        """The dense outlier block in the activations' dtype (the GEMM kernels take fp16 or bf16 operands and do not
        convert: an fp16 block under bf16 activations would be read as bf16 bits)."""
        dtype = self.dtype if dtype is None else dtype
        ow = self.oweight_gemm if getattr(self, "_ow_pad", 0) else self.oweight
        return ow if ow.dtype == dtype else ow.to(dtype)

    def _decode(self, x, seq_len, x_gather=None):
        r = self.outlierfeatures
        return qeft_cuda.gemv_w4(x, self.qweight, self.scales, self.scaled_zeros,
                                 self.oweight_interleaved if r > 0 else None, seq_len, self.outfeatures,
                                 self.infeatures, self.group_size,
                                 ow_layout=_lib.OW_INTERLEAVED if r > 0 else _lib.OW_NONE,
                                 bias=self.bias, x_gather=x_gather)

    def _train_path(self):
        return self.training and self.matmul is not None

    def forward_outlier(self, x):
        if self._train_path():
            return self.matmul(x, self.oweight, self.qweight, self.scales, self.scaled_zeros,
                               self.outlierfeatures, self.bias, self.name, self.group_size)
        seq_len = x.numel() // x.shape[-1]
        if seq_len < self.GEMV_MAX_ROWS:
            return self._decode(x, seq_len)
        return self.gemm(x, self.qweight, self.scales, self.scaled_zeros, self._oweight_plain(x.dtype), self.bias,
                         group_size=self.group_size)

    def forward_outlier_out_proj(self, x):
        """o_proj: the input arrives in model order; the outlier channels are gathered to the back
        (reference: qlinear.py:273-304).  Decode fuses the gather into the GEMV's x load."""
        seq_len = x.numel() // x.shape[-1]
        if not self._train_path() and seq_len < self.GEMV_MAX_ROWS:
            return self._decode(x, seq_len, x_gather=self.reorder_ids32)
        inputs = torch.index_select(x, -1, self.reorder_ids)
        if self._train_path():
            return self.matmul(inputs, self.oweight, self.qweight, self.scales, self.scaled_zeros,
                               self.outlierfeatures, self.bias, self.name, self.group_size)
        return self.gemm(inputs, self.qweight, self.scales, self.scaled_zeros, self._oweight_plain(x.dtype), self.bias,
                         group_size=self.group_size)

    def forward_normal(self, x):
        if self._train_path():
            return self.matmul(x, self.qweight, self.scales, self.scaled_zeros, self.outlierfeatures, self.bias,
                               self.name, self.group_size)
        seq_len = x.numel() // x.shape[-1]
        if seq_len < self.GEMV_MAX_ROWS:
            return self._decode(x, seq_len)
        return self.gemm(x, self.qweight, self.scales, self.scaled_zeros, None, self.bias, group_size=self.group_size)
