"""Hard NJ of a few trees per GPU (a shard of configs[2]: 64 chains over 8 GPUs = 8 trees of 200 taxa): the lean
kernel with one CTA per tree against one thread-block cluster per tree (DODO_NJ_CLUSTER forces the cluster size;
unset = the library's own choice).  Results must be identical bit for bit.   python profiles/tools/nj_cluster.py [S]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dodonaphy_b200 import engine

S = int(sys.argv[1]) if len(sys.argv) > 1 else 200
if len(sys.argv) > 2:
    os.environ["DODO_NJ_LEAN_W"] = sys.argv[2]  # warps per CTA (8 or 16)
rng = np.random.default_rng(11)
for B in (8, 16, 32, 64):
    xs = torch.as_tensor(rng.normal(0, 0.05, (B, S, 3))).cuda()
    pdm = engine.pdm_forward(xs, -1.0, "numpy")
    ref = None
    line = [f"S={S} B={B:3d}:"]
    for cl in ("1", "2", "4", "8", None):
        if cl is None:
            os.environ.pop("DODO_NJ_CLUSTER", None)
        else:
            os.environ["DODO_NJ_CLUSTER"] = cl
        for _ in range(2):
            peel, blens = engine.nj_hard(pdm)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            peel, blens = engine.nj_hard(pdm)
        b.record(); torch.cuda.synchronize()
        if ref is None:
            ref = (peel.clone(), blens.clone())
        same = torch.equal(peel, ref[0]) and torch.equal(blens, ref[1])
        line.append(f"CL={cl or 'auto'} {a.elapsed_time(b) / 5:.3f} ms{'' if same else ' MISMATCH'}")
    print("  ".join(line), flush=True)
