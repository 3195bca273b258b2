"""C2 with a LEARNABLE GTR model (the reference's --model GTR VI always learns rates and frequencies,
phylomodel.py:69-74): one step = lnL of 128 samples + gradients to the locations, the curvature leaf and the
unconstrained rate / frequency leaves, through BaseModel.log_likelihood_batch.  Timed next to the fixed-model
step of the headline.   python profiles/tools/gtr_learn.py [B]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dodonaphy_b200 import synth
from dodonaphy_b200.base_model import BaseModel
from dodonaphy_b200.engine import SubstModel
from dodonaphy_b200.pipeline import HotPath

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
prob = synth.make_problem(64, 10000, B, D=5, seed=1002)
tips = [torch.tensor(np.stack([(m >> j) & 1 for j in range(4)]).astype(np.float64)) for m in prob["masks"]]
w = torch.tensor(prob["weights"])
def timed(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
x = torch.tensor(prob["xs"]).cuda()
KS = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 4]
for K in KS:
    m = BaseModel("vi", tips, w, 5, soft_temp=1e-8, model_name="GTR", gamma_shape=0.5 if K > 1 else None, n_rate_categories=K)
    m.phylomodel.sub_rates = torch.tensor(prob["rates"]); m.phylomodel.freqs = torch.tensor(prob["freqs"])
    xl = x.clone().requires_grad_(True)
    def learn_step():
        ll, peel, blens = m.log_likelihood_batch(xl)
        ll.sum().backward()
    hp = HotPath(prob["masks"], prob["weights"], SubstModel.gtr(prob["rates"], prob["freqs"]),
                 rates=None if K == 1 else SubstModel.discrete_gamma(0.5, K), mix=K > 1, tau=1e-8)
    t_fixed = timed(lambda: hp.value_and_grad(x))
    t_learn = timed(learn_step)
    print(f"K={K} B={B}: fixed-model step {t_fixed:.2f} ms, learnable-GTR step {t_learn:.2f} ms ({t_learn / t_fixed:.2f}x)", flush=True)
