import sys, os, copy, torch
sys.path.insert(0, '/root/repo')
from sparsefactorization_b200 import generators
from sparsefactorization_b200._psfnet import MLPBlock
torch.manual_seed(0)
n, B, n_W, L = 8192, 4, 13, 14
fs = torch.nn.ModuleList([MLPBlock([32, "GELU"], 32, L) for _ in range(n_W)]).cuda()
g = MLPBlock([32, "GELU"], 32, 8).cuda()
for scale in (1.0, 1.41):
    x = torch.randn(B * n, 32, device="cuda") * scale
    W, V0 = generators.fused_generators(fs, x, n, 1, g)
    fs64 = copy.deepcopy(fs).double()
    W64 = torch.stack([f(x.double()) for f in fs64], 0).permute(0, 2, 1)
    W32 = torch.stack([f(x) for f in fs], 0).permute(0, 2, 1)
    ulp = torch.pow(2.0, torch.floor(torch.log2(W64.abs().clamp_min(1e-30))) - 23)
    for name, t in (("ours", W), ("torch32", W32)):
        e = (t.double() - W64) * torch.sign(W64) / ulp
        m = W64.abs() > 1e-3
        print(scale, name, "bias ulp", float(e[m].mean()), "rms ulp", float(e[m].pow(2).mean().sqrt()), "rel bias", float(((t.double()-W64)*torch.sign(W64))[m].mean()/W64.abs()[m].mean()))
