"""The learnable-GTR pruning launch alone at the C2 size (64 taxa x 10k patterns x 128 samples x 4 categories):
dodo_prune_model (per-thread W) next to the explicit route (d lnL / d P per edge + dodo_model_chain) and the
branch-length-only launch.   python profiles/tools/gmodel_kernel.py [reps]   (ncu: -k regex:prune_kernel -s 2 -c 1)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dodonaphy_b200 import synth
from dodonaphy_b200.engine import PruneEngine, SubstModel
from dodonaphy_b200.pipeline import HotPath

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
only_new = "--only-new" in sys.argv
prob = synth.make_problem(64, 10000, 128, D=5, seed=1002)
m = SubstModel.gtr(prob["rates"], prob["freqs"])
rates = SubstModel.discrete_gamma(0.5, 4)
hp = HotPath(prob["masks"], prob["weights"], m, rates=rates, mix=True, tau=1e-8)
x = torch.tensor(prob["xs"]).cuda()
dec = hp.decode(x)
peel, blens = dec["peel"], dec["blens"]
eng = hp.prune

def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

t_new = timed(lambda: eng.loglik_model_grad(peel, blens, m, rates=rates, mix=True))
if only_new:
    print(f"dodo_prune_model {t_new:.3f} ms"); sys.exit(0)
t_old = timed(lambda: eng.loglik_model_grad(peel, blens, m, rates=rates, mix=True, via_mats=True))
t_bl = timed(lambda: eng.loglik(peel, blens, m, rates=rates, mix=True, grad=True))
print(f"blens only {t_bl:.3f} ms | dodo_prune_model {t_new:.3f} ms ({t_new / t_bl:.2f}x) | GMATS + dodo_model_chain {t_old:.3f} ms ({t_old / t_bl:.2f}x)")
