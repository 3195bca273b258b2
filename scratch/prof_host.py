import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from dodonaphy_b200 import synth, engine
from dodonaphy_b200.engine import SubstModel
from dodonaphy_b200.pipeline import HotPath
S,L,B,D,K=64,10000,128,5,4
prob = synth.make_problem(S,L,B,D=D,seed=1002)
hp = HotPath(prob["masks"], prob["weights"], SubstModel.gtr(prob["rates"], prob["freqs"]), rates=SubstModel.discrete_gamma(0.5,K), mix=True)
x = torch.as_tensor(prob["xs"]).cuda()
xp = torch.as_tensor(prob["xs"]).pin_memory(); llp = torch.empty(B,dtype=torch.float64).pin_memory(); gxp = torch.empty((B,S,D),dtype=torch.float64).pin_memory()
for _ in range(3): hp.value_and_grad(x)
torch.cuda.synchronize()
def t(fn, n=10):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn(); torch.cuda.synchronize()
    return (time.perf_counter()-t0)/n*1e3
def cpu_only(fn, n=10):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    t1=time.perf_counter(); torch.cuda.synchronize()
    return (t1-t0)/n*1e3
print("step sync each (ms):", t(lambda: hp.value_and_grad(x)))
print("step host-enqueue time (ms):", cpu_only(lambda: hp.value_and_grad(x)))
print("e2e sync each (ms):", t(lambda: hp.value_and_grad_host(xp, llp, gxp)))
pdm = engine.pdm_forward(x, -1.0)
print("pdm fwd:", t(lambda: engine.pdm_forward(x, -1.0)))
peel, blens, flags = engine.nj_soft(pdm, 1e-8, return_flags=True)
print("nj soft:", t(lambda: engine.nj_soft(pdm, 1e-8, return_flags=True)))
print("prune:", t(lambda: hp.prune.loglik(peel, blens, hp.model, rates=hp.rates, mix=True, grad=True)))
ll, g = hp.prune.loglik(peel, blens, hp.model, rates=hp.rates, mix=True, grad=True)
print("nj bwd:", t(lambda: engine.nj_soft_backward(peel, flags, g)))
gp = engine.nj_soft_backward(peel, flags, g)
print("pdm bwd:", t(lambda: engine.pdm_backward(x, gp, -1.0)))
print("h2d:", t(lambda: xp.to('cuda', non_blocking=True)))
print("d2h:", t(lambda: (llp.copy_(ll, non_blocking=True), gxp.copy_(x, non_blocking=True))))
