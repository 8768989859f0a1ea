"""CPU restatement (numpy) of the reference arithmetic for the packed QuantLinear path.

Test infrastructure only -- see ``oracle/__init__.py``.  Every function cites the
reference ``file:line`` (relative to ``/root/reference``) it restates.  Nothing here
is copied from the reference: the packers are written from the closed-form index
map of the layout, not from the reference's chain of
reshape/transpose calls, and are checked against the reference's own outputs in
``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "tile_index_map", "pack_intweight", "unpack_intweight", "pack_oweight",
    "unpack_oweight", "sparse_to_dense_ids", "dequant_weight", "dense_weight",
    "forward", "backward", "quantize_for_pack", "gemv_algorithmic_bytes",
    "synth_layer", "cpu_dequant_matmul",
]

INTERLEAVE = 4      # qeft/qlinear.py:135  (self.interleave = 4)
KSTRIDE = 64        # qeft/qlinear.py:203  (pack_intweight(..., kstride=64))


def tile_index_map():
    """Position of weight (j, kk) of a 4-row x 64-column tile inside its 128-byte image.

    Restates the three permutations of ``pack_intweight`` (qeft/qlinear.py:88-108) in
    closed form.  Returns ``(e, i)``, both ``[4, 64]`` int arrays: weight ``(j, kk)`` of the
    tile lives in nibble ``i`` (bits ``4i..4i+3``) of int16 number ``e`` of the tile.
    """
    j = np.arange(4)[:, None]
    kk = np.arange(64)[None, :]
    half = kk // 32                    # which 16-byte "thread chunk" of the row
    k32 = kk % 32
    q8 = k32 // 8                      # 0..3: the 8-wide quarter (ldmatrix-style interleave)
    g = (k32 % 8) // 2                 # 32-bit word inside the chunk
    odd = k32 % 2
    n = odd * 4 + q8                   # nibble slot inside the word ([0,2,4,6,1,3,5,7] order)
    p = 32 * half + 8 * g + n          # nibble index inside the row's 64 nibbles
    lin = 64 * j + p                   # nibble index inside the tile (row-interleave of 4)
    return lin // 4, lin % 4


def pack_intweight(q):
    """int weights ``q[N, K]`` (values 0..15) -> int16 ``[N/4, K]`` (qeft/qlinear.py:81-121)."""
    q = np.asarray(q)
    N, K = q.shape
    assert N % INTERLEAVE == 0 and K % KSTRIDE == 0
    e, i = tile_index_map()
    tiles = q.reshape(N // 4, 4, K // 64, 64).transpose(0, 2, 1, 3).astype(np.int64)  # [b, T, j, kk]
    out = np.zeros((N // 4, K // 64, 64), dtype=np.int64)
    for jj in range(4):
        for ii in range(4):
            # no clamp, like the reference: an out-of-range value spills into its neighbours.
            # For a fixed (row, nibble slot) the 16 target int16 numbers are distinct.
            sel = np.nonzero(i[jj] == ii)[0]
            out[:, :, e[jj][sel]] |= tiles[:, :, jj, sel] << (4 * ii)
    return out.reshape(N // 4, K).astype(np.uint16).view(np.int16)


def unpack_intweight(qweight):
    """Inverse of :func:`pack_intweight`: int16 ``[N/4, K]`` -> int32 ``[N, K]`` in 0..15."""
    qw = np.ascontiguousarray(qweight).view(np.uint16).astype(np.int32)
    Nq, K = qw.shape
    e, i = tile_index_map()
    tiles = qw.reshape(Nq, K // 64, 64)
    out = np.empty((Nq, K // 64, 4, 64), dtype=np.int32)
    for jj in range(4):
        out[:, :, jj, :] = (tiles[:, :, e[jj]] >> (4 * i[jj])) & 0xF
    return out.transpose(0, 2, 1, 3).reshape(Nq * 4, K)


def pack_oweight(ow):
    """fp16 ``[N, r]`` -> row-pair-interleaved ``[N/2, 2r]`` (qeft/qlinear.py:70-79).

    ``out[(n//8)*4 + n%4, 64*(j//32) + 2*(j%32) + (n%8)//4] = ow[n, j]``.
    """
    ow = np.asarray(ow)
    N, r = ow.shape
    assert N % 8 == 0 and r % 32 == 0
    v = ow.reshape(N // 8, 2, 4, r // 32, 32)          # [blk, s, j4, c, t]
    return np.ascontiguousarray(v.transpose(0, 2, 3, 4, 1)).reshape(N // 2, 2 * r)


def unpack_oweight(owi):
    """Inverse of :func:`pack_oweight`."""
    owi = np.asarray(owi)
    Nh, r2 = owi.shape
    r = r2 // 2
    v = owi.reshape(Nh // 4, 4, r // 32, 32, 2)        # [blk, j4, c, t, s]
    return np.ascontiguousarray(v.transpose(0, 4, 1, 2, 3)).reshape(Nh * 2, r)


def sparse_to_dense_ids(ids, length):
    """``[all k not in ids, ascending] ++ ids`` (qeft/reorder.py:6-12)."""
    ids = np.asarray(ids, dtype=np.int64)
    assert len(ids) < length
    mask = np.ones(length, dtype=bool)
    mask[ids] = False
    return np.concatenate([np.arange(length, dtype=np.int64)[mask], ids])


def dequant_weight(qweight, scales, scaled_zeros, group_size=128):
    """fp16 ``W[N, K] = fma.rn.f16(q, s, sz)``.

    Follows ``dequantize_s4_to_fp16x2`` (kernel/quantization_new/dequantize.cuh:14-77: exact
    0..15 as fp16) and the ``__hfma2(w, scale, scaled_zero)`` of the GEMV/GEMM kernels
    (gemv/gemv_cuda.cu:149-159, gemm/gemm_cuda.cu:728-743).  ``q*s + sz`` is exact in
    float64 (4-bit x 11-bit product plus an fp16), so one cast gives the single rounding.
    """
    q = unpack_intweight(qweight).astype(np.float64)
    N, K = q.shape
    G = K if group_size in (-1, K) else group_size
    s = np.asarray(scales, dtype=np.float16).astype(np.float64)        # [K/G, N]
    z = np.asarray(scaled_zeros, dtype=np.float16).astype(np.float64)
    gi = np.arange(K) // G
    return (q * s.T[:, gi] + z.T[:, gi]).astype(np.float16)


def dense_weight(qweight, scales, scaled_zeros, oweight=None, group_size=128, semantics="gemv"):
    """The dense fp32 ``[N, K]`` matrix the packed layer stands for.

    ``semantics="gemv"``: the last r columns are *replaced* by ``oweight``
    (gemv/gemv_cuda_qeft.cu:168-176).  ``semantics="gemm"``: the int4 residual of those
    columns is kept and ``oweight`` is added (qeft/qlinear.py:264-266).
    """
    W = dequant_weight(qweight, scales, scaled_zeros, group_size).astype(np.float32)
    if oweight is not None:
        ow = np.asarray(oweight).astype(np.float32)
        r = ow.shape[1]
        if semantics == "gemv":
            W[:, W.shape[1] - r:] = ow
        elif semantics == "gemm":
            W[:, W.shape[1] - r:] += ow
        else:
            raise ValueError(semantics)
    return W


def forward(x, qweight, scales, scaled_zeros, oweight=None, bias=None, group_size=128,
            semantics="gemv", reorder_ids=None, acc=np.float64):
    """``y = x . W^T (+ bias)`` with wide accumulation, fp16 result.

    Restates ``QuantLinear.forward_outlier[_out_proj]`` (qeft/qlinear.py:244-304): optional
    ``index_select(x, -1, reorder_ids)`` for o_proj (:275), then the GEMV (:253-263) or
    GEMM + outlier ``F.linear`` (:264-266) arithmetic, then ``+ bias`` (:268).
    """
    x = np.asarray(x)
    if reorder_ids is not None:
        x = np.take(x, np.asarray(reorder_ids), axis=-1)
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1]).astype(acc)
    W = dense_weight(qweight, scales, scaled_zeros, oweight, group_size, semantics).astype(acc)
    y = x2 @ W.T
    if bias is not None:
        y = y + np.asarray(bias).astype(acc)
    return y.astype(np.float16).reshape(*lead, -1)


def backward(dy, x, qweight, scales, scaled_zeros, oweight, group_size=128, acc=np.float64):
    """Gradients of ``y = x . W_dense^T`` w.r.t. ``x`` and the trainable outlier columns.

    The reference's ``QuantMatMulQEFT.backward`` (qeft/qlinear.py:28-44) is not usable
    (SURVEY.md section 0); this is the math BASELINE.json's north_star defines:
    ``dX = dY . W_dense`` and ``dOW[N, r] = dY^T . X[:, K-r:]``, identical to autograd
    through ``F.linear(x, W_dense)``.  Returns ``(dx fp16 [.., K], dow fp32 [N, r])``.
    """
    dy = np.asarray(dy)
    x = np.asarray(x)
    lead = x.shape[:-1]
    dy2 = dy.reshape(-1, dy.shape[-1]).astype(acc)
    x2 = x.reshape(-1, x.shape[-1]).astype(acc)
    W = dense_weight(qweight, scales, scaled_zeros, oweight, group_size, "gemv").astype(acc)
    dx = (dy2 @ W).astype(np.float16).reshape(*lead, -1)
    dow = None
    if oweight is not None:
        r = np.asarray(oweight).shape[1]
        dow = (dy2.T @ x2[:, x2.shape[1] - r:]).astype(np.float32)
    return dx, dow


def quantize_for_pack(weight, scales, zeros, n_out, group_size, sym=False):
    """Restates ``QuantLinear.pack`` (qeft/qlinear.py:180-215) up to the packer calls.

    ``weight`` fp [N, K]; ``scales``/``zeros`` [N, K/G].  Returns a dict with the int
    matrix, fp16 ``scales``/``scaled_zeros`` ``[K/G, N]`` and ``oweight``.  Like the
    reference: no clamp (:197), the outlier columns take ``zeros[:, k // G]`` (:200-202),
    ``sym`` shifts the zero points by 8 (:184-185; the reference does it in place, here
    the caller's array is left alone).
    """
    import torch  # torch.round / division in the weight dtype, like the reference

    w = torch.as_tensor(np.asarray(weight))
    s = torch.as_tensor(np.asarray(scales))
    z = torch.as_tensor(np.asarray(zeros)).clone()
    if sym:
        z = z + 8
    N, K = w.shape
    G = K if group_size in (-1, K) else group_size
    rep = 1 if G == K else G
    sz = z * s
    q = torch.round((w + torch.repeat_interleave(sz, rep, dim=1)) /
                    torch.repeat_interleave(s, rep, dim=1)).to(torch.int32)
    if n_out > 0:
        cols = torch.arange(K - n_out, K)
        q[:, K - n_out:] = z[:, cols // G].to(torch.int32)
    out = {
        "intweight": q.numpy(),
        "scales": s.t().contiguous().to(torch.float16).numpy(),
        "scaled_zeros": (-sz.t().contiguous().to(torch.float16)).numpy(),
    }
    if n_out > 0:
        out["oweight"] = w[:, K - n_out:].clone().numpy()
    return out


def gemv_algorithmic_bytes(N, K, m=1, r=128, G=128, bias=False):
    """bytes one decode GEMV call must move (dead outlier int4 not counted)."""
    b = N * (K - r) // 2 + 4 * N * ((K - r) // G) + 2 * N * r + 2 * K * m + 2 * N * m
    return b + (2 * N if bias else 0)


def synth_layer(N, K, r=128, G=128, seed=0, bias=False, o_proj=False):
    """Synthetic packed layer per SURVEY.md 8(d) (seeded; all numpy, packed by the oracle packer)."""
    rng = np.random.default_rng(seed)
    ng = K // G
    q = rng.integers(0, 16, size=(N, K), dtype=np.int32)
    zero = rng.integers(0, 16, size=(ng, N), dtype=np.int32)
    scale = rng.uniform(0.002, 0.012, size=(ng, N)).astype(np.float16)
    if r > 0:
        q[:, K - r:] = zero[(np.arange(K - r, K) // G), :].T
    sz = (-(zero.astype(np.float32) * scale.astype(np.float32))).astype(np.float16)
    layer = {
        "N": N, "K": K, "r": r, "G": G,
        "intweight": q,
        "qweight": pack_intweight(q),
        "scales": scale,
        "scaled_zeros": sz,
    }
    if r > 0:
        ow = (rng.standard_normal((N, r)) * 0.02).astype(np.float16)
        layer["oweight"] = ow
        layer["oweight_interleaved"] = pack_oweight(ow)
        if o_proj:
            idx = np.sort(rng.choice(K, size=r, replace=False)).astype(np.int32)
        else:
            idx = np.arange(K - r, K, dtype=np.int32)
        layer["outlieridx"] = idx
    if bias:
        layer["bias"] = (rng.standard_normal(N) * 0.02).astype(np.float16)
    return layer


def _unpack_torch(qweight_t):
    """Same map as :func:`unpack_intweight`, as torch reshapes/permutes (multi-threaded, for the timed baseline)."""
    import torch

    Nq, K = qweight_t.shape
    v = (qweight_t.to(torch.int32) & 0xFFFF).view(Nq, K // 64, 64, 1)
    nib = (v >> torch.tensor([0, 4, 8, 12], dtype=torch.int32).view(1, 1, 1, 4)) & 0xF      # [b, T, e, i]
    # linear nibble index 4e+i = 64 j + 32 half + 8 g + 4 odd + q8 ;  column = 32 half + 8 q8 + 2 g + odd
    nib = nib.reshape(Nq, K // 64, 4, 2, 4, 2, 4)                                           # [b, T, j, half, g, odd, q8]
    nib = nib.permute(0, 2, 1, 3, 6, 4, 5)                                                  # [b, j, T, half, q8, g, odd]
    return nib.reshape(Nq * 4, K)


def cpu_dequant_matmul(x, layer, threads=None, cached_weight=None):
    """The "reference torch dequant+matmul on CPU" of BASELINE.json config 1, in torch.

    unpack -> ``w = fp16(q*s+sz)`` -> fp32 matmul + outlier matmul.  Used by ``bench.py`` as
    the timed CPU baseline (``cpu_baseline.kind == "port"``).  Returns ``(y, W_fp32)``.
    """
    import torch

    if threads:
        torch.set_num_threads(threads)
    xt = torch.as_tensor(np.asarray(x)).float()
    r = layer["r"]
    K = layer["K"]
    if cached_weight is None:
        q = _unpack_torch(torch.as_tensor(layer["qweight"]))                        # int32 [N, K]
        G = layer["G"]
        N = q.shape[0]
        s = torch.as_tensor(layer["scales"]).t().float()                            # [N, K/G]
        z = torch.as_tensor(layer["scaled_zeros"]).t().float()
        # q*s is exact in fp32 (4-bit x 11-bit); the sum is rounded once to fp32 and again to fp16.  Double
        # rounding can differ from the single-rounding fma by one fp16 ulp in rare ties; this function is the
        # TIMED baseline, the checker is `forward` (float64, single rounding).
        W = (q.view(N, K // G, G).float() * s.unsqueeze(-1) + z.unsqueeze(-1)).half().float().view(N, K)
    else:
        W = cached_weight
    xm = xt.reshape(-1, K)
    if r > 0:
        y = xm[:, :K - r] @ W[:, :K - r].t() + xm[:, K - r:] @ torch.as_tensor(layer["oweight"]).float().t()
    else:
        y = xm @ W.t()
    if "bias" in layer:
        y = y + torch.as_tensor(layer["bias"]).float()
    return y.half().reshape(*xt.shape[:-1], -1), W
