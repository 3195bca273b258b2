"""Time the batched blens Jacobian + log-det at the C2 shape: python profiles/tools/jac_c2.py"""
import sys

import torch

sys.path.insert(0, ".")  # run from the repository root
from dodonaphy_b200 import synth
from dodonaphy_b200.engine import SubstModel
from dodonaphy_b200.pipeline import HotPath

S, L, B, D = 64, 10000, 128, 5
prob = synth.make_problem(S, 2000, B, D=D, seed=1002)
hp = HotPath(prob["masks"], prob["weights"], SubstModel.gtr(prob["rates"], prob["freqs"]),
             rates=SubstModel.discrete_gamma(0.5, 4), mix=True, tau=1e-8)
x = torch.as_tensor(prob["xs"]).cuda()


def t(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


print("jacobian only ms %.3f" % t(lambda: hp.blens_jacobian(x, logdet=False)))
print("jacobian + logdet ms %.3f" % t(lambda: hp.blens_jacobian(x, logdet=True)))
