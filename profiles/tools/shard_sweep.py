#!/usr/bin/env python
"""Per-rank workloads of the strong-scaling legs, emulated on ONE GPU: the shard a rank would hold at
N = 1, 2, 4, 8 is run alone and timed stage by stage (CUDA events), which predicts the strong-scaling
efficiency of the kernels (collectives excluded) without spending multi-GPU time.

    python profiles/tools/shard_sweep.py [c2] [c3] [c5]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dodonaphy_b200 import engine, synth  # noqa: E402
from dodonaphy_b200.engine import PruneEngine, SubstModel  # noqa: E402
from dodonaphy_b200.pipeline import GraphedStep, HotPath  # noqa: E402


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    which = sys.argv[1:] or ["c2", "c3", "c5"]
    NS = tuple(int(v) for v in os.environ.get("SHARDS", "1,2,4,8").split(","))  # which N to emulate
    out = {}
    if "c2" in which:
        prob = synth.make_problem(64, 10000, 128, D=5, seed=1002)
        hp = HotPath(prob["masks"], prob["weights"], SubstModel.gtr(prob["rates"], prob["freqs"]),
                     rates=SubstModel.discrete_gamma(0.5, 4), mix=True, tau=1e-8)
        rows = []
        for n in NS:
            x = torch.as_tensor(prob["xs"][:128 // n]).cuda()
            g = GraphedStep(hp, x, grad=True)
            pdm = engine.pdm_forward(x, -1.0, "torch")
            peel, blens, _ = engine.nj_soft(pdm, 1e-8, return_flags=True)
            rows.append(dict(n=n, samples=128 // n, step_graph_ms=timed(lambda: g.run(x)),
                             nj_soft_ms=timed(lambda: engine.nj_soft(pdm, 1e-8)),
                             prune_ms=timed(lambda: hp.prune.loglik(peel, blens, hp.model, rates=hp.rates, mix=True, grad=True))))
        for r in rows:
            r["efficiency_vs_n1"] = rows[0]["step_graph_ms"] / (r["n"] * r["step_graph_ms"])
        out["c2_128_total"] = rows
    if "c3" in which:
        prob = synth.make_problem(200, 20000, 64, D=5, seed=1003)
        masks = np.tile(prob["masks"], (1, 3))[:, :50000]
        weights = np.tile(prob["weights"], 3)[:50000]
        hp = HotPath(masks, weights, SubstModel.jc69())
        rows = []
        for n in NS:
            x = torch.as_tensor(prob["xs"][:64 // n]).cuda()
            g = GraphedStep(hp, x, grad=False)
            pdm = engine.pdm_forward(x, -1.0, "numpy")
            peel, blens = engine.nj_hard(pdm)
            rows.append(dict(n=n, chains=64 // n, sweep_graph_ms=timed(lambda: g.run(x)),
                             nj_hard_ms=timed(lambda: engine.nj_hard(pdm)),
                             prune_ms=timed(lambda: hp.prune.loglik(peel, blens, hp.model, grad=False))))
        for r in rows:
            r["efficiency_vs_n1"] = rows[0]["sweep_graph_ms"] / (r["n"] * r["sweep_graph_ms"])
        out["c3_64_chains"] = rows
    if "c5" in which:
        prob = synth.make_problem(500, 20000, 1, D=5, seed=1005)
        L = 1000000
        masks = np.tile(prob["masks"], (1, 50))[:, :L]
        weights = np.tile(prob["weights"], 50)[:L]
        peel, blens = engine.nj_hard(engine.pdm_forward(torch.as_tensor(prob["xs"][0]).cuda(), -1.0, "numpy"))
        rows = []
        for n in NS:
            eng = PruneEngine(np.ascontiguousarray(masks[:, :L // n]), weights[:L // n])
            rows.append(dict(n=n, patterns=L // n, eval_ms=timed(lambda: eng.loglik(peel, blens, SubstModel.jc69(), grad=True), n=5)))
            del eng
            torch.cuda.empty_cache()
        for r in rows:
            r["efficiency_vs_n1"] = rows[0]["eval_ms"] / (r["n"] * r["eval_ms"])
        out["c5_1M_patterns"] = rows
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
