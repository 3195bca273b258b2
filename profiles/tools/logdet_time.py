"""gram_logdet_kernel at the C2 size (128 samples x 126 branch lengths): time against the number of columns C --
the slope is the Gram phase, the intercept the elimination.   python profiles/tools/logdet_time.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dodonaphy_b200 import engine

rng = np.random.default_rng(5)
for B, E, C in ((128, 126, 126), (128, 126, 320), (128, 126, 640), (16, 126, 320), (128, 62, 160), (64, 190, 480)):
    J = torch.tensor(rng.normal(size=(B, E, C))).cuda()
    engine.gram_logdet(J); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): out = engine.gram_logdet(J)
    e1.record(); torch.cuda.synchronize()
    want = 0.5 * np.linalg.slogdet(J[0].cpu().numpy() @ J[0].cpu().numpy().T)[1]
    print(f"B={B} E={E} C={C}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us   (sample 0: {out[0].item():.12g} vs LAPACK {want:.12g})", flush=True)
