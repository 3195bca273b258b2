"""One soft + one hard NJ launch at a given size (for ncu): python profiles/tools/nj_one.py S B"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")  # run from the repository root
from dodonaphy_b200 import engine

S, B = int(sys.argv[1]), int(sys.argv[2])
xs = torch.as_tensor(np.random.default_rng(3).normal(0, 0.05, (B, S, 5))).cuda()
pdm = engine.pdm_forward(xs, -1.0, "torch")
for _ in range(2):
    engine.nj_soft(pdm, 1e-8)
    engine.nj_hard(pdm)
torch.cuda.synchronize()
