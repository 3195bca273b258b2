import torch, numpy as np, sys
sys.path.insert(0, '.')  # run from the repository root
from dodonaphy_b200 import engine
S, B = 2000, 148
xs = torch.as_tensor(np.random.default_rng(7).normal(0, 0.05, (B, S, 5))).cuda()
pdm = engine.pdm_forward(xs, -1.0, "torch")
for _ in range(2):
    engine.nj_hard(pdm)
engine.nj_soft(pdm, 1e-8)  # warm-up (the first call allocates the workspace)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); engine.nj_hard(pdm); b.record(); torch.cuda.synchronize()
print("nj_hard B=148 ms", a.elapsed_time(b))
a.record(); engine.nj_soft(pdm, 1e-8); b.record(); torch.cuda.synchronize()
print("nj_soft B=148 ms", a.elapsed_time(b))
