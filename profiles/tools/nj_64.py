import torch, numpy as np, sys
sys.path.insert(0, '.')  # run from the repository root
from dodonaphy_b200 import engine, synth
prob = synth.make_problem(64, 100, 128, D=5, seed=1002)
xs = torch.as_tensor(prob["xs"]).cuda()
pdm = engine.pdm_forward(xs, -1.0, "torch")
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print("soft ms", t(lambda: engine.nj_soft(pdm, 1e-8)), "hard ms", t(lambda: engine.nj_hard(pdm)))
