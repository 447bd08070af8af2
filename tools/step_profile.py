#!/usr/bin/env python
"""torch.profiler breakdown of one PSFNet training step at the bench configuration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from bench import psf_kwargs, N_VEC
from sparsefactorization_b200 import data, dp
from sparsefactorization_b200.psf import PSFNet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 132
dev = torch.device("cuda:0")
torch.manual_seed(42)
net = PSFNet(**psf_kwargs(True)).to(dev)
opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)
X, Y = data.adding(B, N_VEC, seed=1, device=dev)
loss_fn = torch.nn.MSELoss()
for _ in range(3):
    dp.train_step(net, None, opt, loss_fn, X, Y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        dp.train_step(net, None, opt, loss_fn, X, Y)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
