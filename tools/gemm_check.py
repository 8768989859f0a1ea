import sys, torch
sys.path.insert(0, '/root/repo')
from qeft_b200 import qeft_cuda
from qeft_b200.synth import synth_tensors
for (N,K,M) in [(4096,4096,2048),(256,512,128),(11008,4096,2048)]:
    t = synth_tensors(N, K, seed=3)
    x = torch.randn((M, K), device="cuda").half()
    f = lambda xx: qeft_cuda.gemm_w4(xx, t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"], None, pdl=False)
    y1 = f(x); y2 = f(x); torch.cuda.synchronize()
    print(N,K,M,"deterministic:", torch.equal(y1,y2), "ndiff", (y1!=y2).sum().item())
    w = qeft_cuda.dequant_w4(t["qweight"], t["scales"], t["scaled_zeros"], t["oweight"])
    ref = (x.float() @ w.float().t())
    err = (y1.float()-ref).abs().max().item()/ref.abs().max().item()
    print(" max rel err vs torch fp32 matmul on dequantised weight:", err)
    y3 = f((x*2).half()); torch.cuda.synchronize()
    d = (y3 != (y1.float()*2).half())
    print(" linearity ndiff", d.sum().item(), "of", d.numel())
    if d.any():
        idx = d.nonzero()[:5]
        for i,j in idx.tolist(): print("   ", i,j, y1[i,j].item(), y3[i,j].item())
