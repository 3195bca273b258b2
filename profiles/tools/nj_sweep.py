import torch, numpy as np, sys
sys.path.insert(0, '.')  # run from the repository root
from dodonaphy_b200 import engine
S, B = int(sys.argv[1]), int(sys.argv[2])
xs = torch.as_tensor(np.random.default_rng(3).normal(0, 0.05, (B, S, 5))).cuda()
pdm = engine.pdm_forward(xs, -1.0, "torch")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print(S, B, "soft ms %.3f hard ms %.3f" % (t(lambda: engine.nj_soft(pdm, 1e-8)), t(lambda: engine.nj_hard(pdm))))
