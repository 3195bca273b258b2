"""Time soft / hard NJ for S taxa x B trees; soft NJ with and without the lean small-tree kernel.
   python profiles/tools/nj_sweep.py S B"""
import os, sys
import numpy as np, torch
sys.path.insert(0, '.')  # run from the repository root
from dodonaphy_b200 import engine
S, B = int(sys.argv[1]), int(sys.argv[2])
xs = torch.as_tensor(np.random.default_rng(3).normal(0, 0.05, (B, S, 5))).cuda()
pdm = engine.pdm_forward(xs, -1.0, "torch")
pdm_np = engine.pdm_forward(xs, -1.0, "numpy")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
lean = t(lambda: engine.nj_soft(pdm, 1e-8))
p1, b1, f1 = engine.nj_soft(pdm, 1e-8, return_flags=True)
os.environ["DODO_NJ_NO_LEAN"] = "1"
gen = t(lambda: engine.nj_soft(pdm, 1e-8))
p0, b0, f0 = engine.nj_soft(pdm, 1e-8, return_flags=True)
hard_gen = t(lambda: engine.nj_hard(pdm_np))
hp0, hb0 = engine.nj_hard(pdm_np)
del os.environ["DODO_NJ_NO_LEAN"]
hp1, hb1 = engine.nj_hard(pdm_np)
same_hard = bool(torch.equal(hp0, hp1)) and bool(torch.equal(hb0, hb1))
same = bool(torch.equal(p0, p1)), float((b0 - b1).abs().max()), bool(torch.equal(f0, f1))
print(S, B, "soft ms %.3f (general %.3f) hard ms %.3f (general %.3f, identical %s)" % (lean, gen, t(lambda: engine.nj_hard(pdm_np)), hard_gen, same_hard), "lean==general:", same,
      "W", os.environ.get("DODO_NJ_LEAN_W", "auto"))
