"""Host-side profile of the learnable-GTR step at C2 (where the wall time between the kernels goes):
python profiles/tools/gtr_host_profile.py"""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dodonaphy_b200 import synth
from dodonaphy_b200.base_model import BaseModel

prob = synth.make_problem(64, 10000, 128, D=5, seed=1002)
tips = [torch.tensor(np.stack([(m >> j) & 1 for j in range(4)]).astype(np.float64)) for m in prob["masks"]]
m = BaseModel("vi", tips, torch.tensor(prob["weights"]), 5, soft_temp=1e-8, model_name="GTR", gamma_shape=0.5, n_rate_categories=4)
m.phylomodel.sub_rates = torch.tensor(prob["rates"]); m.phylomodel.freqs = torch.tensor(prob["freqs"])
xl = torch.tensor(prob["xs"]).cuda().requires_grad_(True)

def step():
    ll, _, _ = m.log_likelihood_batch(xl)
    ll.sum().backward()

for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize()
print("step ms", (time.perf_counter() - t0) / 20 * 1e3)
# forward / backward split
tf = tb = 0.0
for _ in range(20):
    torch.cuda.synchronize(); a = time.perf_counter()
    ll, _, _ = m.log_likelihood_batch(xl)
    s = ll.sum()
    torch.cuda.synchronize(); b = time.perf_counter()
    s.backward()
    torch.cuda.synchronize(); c = time.perf_counter()
    tf += b - a; tb += c - b
print("forward ms", tf / 20 * 1e3, "backward ms", tb / 20 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(18)
