"""Single-GPU checks of the prefill stack and of the gather entry point with one rank (peer = self)."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def test_gemm_gather_single_rank_writes_pitched_columns():
    """nranks = 1, y_ld > N: the epilogue writes this rank's column block of a wider gathered buffer and signals."""
    from qeft_b200 import _lib, qeft_cuda
    N, K, r, M, W = 256, 512, 128, 200, 768
    L = oracle.synth_layer(N, K, r=r, G=128, seed=11, bias=True)
    d = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()  # noqa: E731
    x = np.random.default_rng(3).standard_normal((M, K)).astype(np.float16)
    full = torch.full((M, W), 7.0, dtype=torch.float16, device="cuda")
    flags = torch.zeros(2, dtype=torch.int32, device="cuda")
    epoch = torch.ones(1, dtype=torch.int32, device="cuda")
    g = _lib.Gather()
    g.nranks, g.y_ld = 1, W
    g.y_peer[0][0] = full.data_ptr() + 2 * 256           # columns 256 .. 511
    g.done_peer[0] = flags.data_ptr()
    g.wait_flag = None
    g.epoch = epoch.data_ptr()
    T = {k: d(L[k]) for k in ("qweight", "scales", "scaled_zeros", "oweight", "bias")}
    qeft_cuda.gemm_w4_gather(d(x), T["qweight"], T["scales"], T["scaled_zeros"], T["oweight"], T["bias"], g)
    want = qeft_cuda.gemm_w4(d(x), T["qweight"], T["scales"], T["scaled_zeros"], T["oweight"], T["bias"])
    torch.cuda.synchronize()
    assert torch.equal(full[:, 256:512], want)
    assert torch.all(full[:, :256] == 7.0) and torch.all(full[:, 512:] == 7.0)
    assert flags.tolist() == [65536, 0]                  # the CTAs' shares add up to QEFT_ARRIVALS_PER_LAUNCH
    ref = oracle.forward(x, L["qweight"], L["scales"], L["scaled_zeros"], L["oweight"], L["bias"])
    err = np.max(np.abs(want.cpu().numpy().astype(np.float64) - ref.astype(np.float64))) / np.max(np.abs(ref.astype(np.float64)))
    assert err <= 1e-3
    # a second launch that waits on the first one's counter (epoch 1 x 1 rank reached) runs through
    g2 = _lib.Gather()
    g2.nranks, g2.y_ld = 1, W
    g2.y_peer[0][0] = full.data_ptr()
    flags2 = torch.zeros(2, dtype=torch.int32, device="cuda")
    g2.done_peer[0] = flags2.data_ptr()
    g2.wait_flag = flags.data_ptr()
    g2.epoch = epoch.data_ptr()
    qeft_cuda.gemm_w4_gather(d(x), T["qweight"], T["scales"], T["scaled_zeros"], T["oweight"], T["bias"], g2)
    torch.cuda.synchronize()
    assert torch.equal(full[:, :256], want) and flags2.tolist() == [65536, 0]


def test_prefill_stack_single_gpu_matches_per_layer_gemm():
    from qeft_b200 import qeft_cuda
    from qeft_b200.prefill import PackedPrefillStack
    st = PackedPrefillStack((256, 512, 3, 128), M=160, seed=2, fast_synth=False)
    y = st.step()
    t = st.blocks[2]["down"]
    want = qeft_cuda.gemm_w4(st.x_f, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], None)
    torch.cuda.synchronize()
    assert torch.equal(y, want)
    assert st.launches_per_step() == 21 and st.flops_per_step() > 0
