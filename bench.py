#!/usr/bin/env python
"""bench.py -- headline benchmark of the dodonaphy_b200 hot path (contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): VI on a synthetic 64-taxon x 10 000-site-pattern alignment,
GTR + Gamma4, 128 samples per step and per GPU, embedding dimension 5.  One *step* = one pass of the
whole hot path over the batch:

    locations -> pairwise hyperbolic distance -> soft NJ -> P(t) + pruning lnL -> d lnL/d blens
              -> soft-NJ backward -> distance backward -> d lnL / d locations

metric  = log-lik+grad evaluations per second expressed as taxa x site-patterns x samples / s.
value   = device-timed (CUDA events, max over ranks), inputs resident in HBM.
e2e     = the same through the public API with HOST buffers: pinned H2D of the step's locations and
          D2H of lnL + gradients inside the timed region.
roofline= algorithmic bytes of the pruning fwd+bwd kernel (V(5S-8)+2S per pattern and
          sample, V = 32K) / that kernel's duration measured with CUDA events around the launch.
cpu_baseline / --impl reference = the reference's own modules from oracle/_ref (compiled Cython + its Python
          package: Chyp_np.get_pdm, Cpeeler.nj_np, PhyloModel.get_transition_mats, phylo.calculate_treelikelihood,
          torch autograd; kind "reference") on this box's host cores, on a bounded sample of the same workload;
          the torch port of those lines (kind "port") only where oracle/_ref is absent.

N > 1 (torchrun, one rank per GPU): samples are sharded, every rank holds its own 128 samples (weak
scaling); the only exchange is an all_gather of the per-sample lnL (the IWAE logsumexp needs all of
them).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "log-lik+grad evals/sec (taxa x sites x samples / s)"
UNIT = "taxa*sites*samples/s"
WORKLOADS = {
    # name: S, L, B per GPU, D, K, gamma alpha, seed
    "c2": dict(S=64, L=10000, B=128, D=5, K=4, alpha=0.5, seed=1002, tau=1e-8,
               desc="VI, synthetic 64 taxa x 10k site patterns, GTR+G4, 128 samples/step, d5 (BASELINE.json configs[1])"),
    "tiny": dict(S=16, L=500, B=8, D=3, K=4, alpha=0.5, seed=1, tau=1e-8, desc="debug"),
}
# the other BASELINE.json configurations: parity-test shapes, timed by `bench.py --side NAME` into
# profiles/ (NOT the headline line; they print their own small JSON)
SIDE = {
    "c1": dict(S=27, L=934, B=4, D=3, desc="DS1-sized (27 taxa x 934 patterns, JC69): MCMC sweep of 4 chains and VI step of 8 samples, launch-by-launch vs one CUDA graph (configs[0])"),
    "c3": dict(S=200, L=50000, B=64, D=5, desc="MCMC, 64 chains, 200 taxa x 50k patterns, JC69, forward only (configs[2])"),
    "c4": dict(S=2000, B=256, D=5, desc="pairwise distance + batched NJ, 2000 taxa x 256 embeddings (configs[3])"),
    "c5": dict(S=500, L=1000000, B=1, D=5, desc="long alignment, 500 taxa x 1M patterns, fwd+grad (configs[4]; one GPU holds L/world)"),
    "gtr": dict(S=64, L=10000, B=128, D=5, desc="configs[1] with the GTR rates and frequencies LEARNT (the reference's --model GTR always learns them, phylomodel.py:69-74): lnL + gradients to locations, curvature, rate and frequency leaves"),
    "jac": dict(S=64, L=10000, B=128, D=5, desc="configs[1]: the change-of-variables term of every sample (vi.py:376-387): d blens / d locations (126 x 320 per sample) + log sqrt det(J J^T)"),
}


def config_dict(cfg, world):
    """The `config` object of the JSON line -- identical for both arms (the reference arm's bounded sample is
    described in its cpu_baseline.sample, not here)."""
    return {"workload": cfg["desc"], "taxa": cfg["S"], "site_patterns": cfg["L"], "samples_per_step_per_gpu": cfg["B"],
            "global_samples_per_step": cfg["B"] * world, "rate_categories": cfg["K"], "model": "GTR+G4",
            "embedding_dim": cfg["D"], "soft_nj_tau": cfg["tau"],
            "sharding": "by sample; all_gather of per-sample lnL" if world > 1 else "single GPU",
            "l2": "pruning scratch streams ~20 GB/step through HBM (>> 126 MB L2); no flush needed"}


def quiet_nccl():
    """Rank 0 prints ONE JSON line on stdout.  NCCL writes its version banner to stdout at the VERSION and WARN levels
    and honours NCCL_DEBUG_FILE only above VERSION: so VERSION becomes WARN, and whatever NCCL logs goes to a file."""
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    if os.environ.get("NCCL_DEBUG"):
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/dodonaphy_b200_nccl.%h.%p.log")


def kernel_source_sha():
    """sha256 of the pruning kernel's source: profiles/roofline_traffic.json records the one its ncu capture was
    taken with, so a stale DRAM-traffic figure is never quoted for a changed kernel."""
    import hashlib

    with open(os.path.join(ROOT, "dodonaphy_b200", "csrc", "prune.cu"), "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def algorithmic_bytes_fwd_bwd(S, K):
    """bytes per (pattern, sample) of a fwd+bwd evaluation that materialises every
    partial once: V(5S-8) + 2S with V = 32 K."""
    V = 32 * K
    return V * (5 * S - 8) + 2 * S


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return None
        top = max(sm)
        loaded = [v for v in sm if v >= 0.5 * top]
        return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU reference arm (the only place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------
def cpu_reference(prob, cfg, n_trees, repeats=1, threads=None):
    """Reference CPU arithmetic for n_trees samples of the workload; returns (seconds per pass,
    description).  Topology stage: the reference's soft NJ needs 209 s and 18.8 GB per 64-taxon tree
    (SURVEY.md Appendix B), so its hard NJ (compiled Cpeeler.nj_np when available) stands in."""
    import torch

    from oracle import hyp, nj, phylo, refload

    torch.set_num_threads(threads or os.cpu_count() or 1)
    ref = refload.load()
    rates = torch.tensor(phylo.discrete_gamma_rates(cfg["alpha"], cfg["K"]))
    tips = [torch.tensor(phylo.mask_to_partial(m)) for m in prob["masks"]]
    w = torch.tensor(prob["weights"])
    sub = torch.tensor(prob["rates"])
    fr = torch.tensor(prob["freqs"])
    # The REAL reference where its package travelled (oracle/_ref: compiled Cython modules + its Python modules):
    # Chyp_np.get_pdm -> Cpeeler.nj_np -> PhyloModel("GTR").get_transition_mats with the K category rates as the
    # reference's own category axis -> phylo.calculate_treelikelihood -> torch autograd.  The reference has no gamma
    # MIXTURE: with a category axis its broadcasting sums the per-category lnL (phylo.py:137-139) -- the same
    # arithmetic as the mixture minus one per-site logsumexp.  Otherwise the torch port of those lines.
    real = bool(ref.compiled and ref.python)
    kind = "reference" if real else "port"
    if real:
        pm = ref.phylomodel.PhyloModel("GTR")
        pm._freqs = pm.freqs_transform.inv(fr).clone().detach()
        pm._sub_rates = pm.sub_rates_transform.inv(sub).clone().detach()
        L = int(prob["masks"].shape[1])
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        for b in range(n_trees):
            x = prob["xs"][b % prob["xs"].shape[0]]
            if ref.compiled:
                pdm = np.ascontiguousarray(ref.Chyp_np.get_pdm(x))
                peel, blens = ref.Cpeeler.nj_np(pdm)
                peel = np.asarray(peel)
            else:
                peel, blens = nj.nj_hard(hyp.get_pdm_np(x))
            bt = torch.tensor(blens, requires_grad=True)
            if real:
                parts = list(tips) + [torch.zeros((1, 4, L)) for _ in range(cfg["S"] - 1)]  # BaseModel.__init__, base_model.py:86-101
                mats = pm.get_transition_mats(bt[:, None] * rates[None, :])
                ll = ref.phylo.calculate_treelikelihood(parts, w, peel, mats, pm.freqs)
            else:
                mats = phylo.gtr_p_t_torch(bt[:, None] * rates[None, :], sub, fr)
                ll = phylo.treelikelihood_torch(tips, w, peel, mats, fr, gamma_mix=cfg["K"] > 1)
            ll.backward()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    if real:
        how = ("the reference's own modules (oracle/_ref): Chyp_np.get_pdm, Cpeeler.nj_np, PhyloModel.get_transition_mats with the "
               f"{cfg['K']} gamma rates as its category axis, phylo.calculate_treelikelihood, torch autograd (per-category lnL summed: "
               "the reference has no gamma mixture)")
    else:
        how = (f"pdm + hard NJ ({'reference Cython modules (oracle/_ref)' if ref.compiled else 'oracle port'}) + GTR P(t) + pruning "
               "lnL + autograd gradient (torch float64 port of phylo.py:132-139 / phylomodel.py:96-114)")
    desc = (f"{n_trees} of {cfg['B']} samples: {how}; hard NJ stands in for the reference's soft NJ (209 s and 18.8 GB per "
            f"{cfg['S']}-taxon tree); {torch.get_num_threads()} threads")
    return best, desc, kind, torch.get_num_threads()


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dodonaphy_b200 import synth

    prob = synth.make_problem(cfg["S"], cfg["L"], cfg["B"], D=cfg["D"], seed=cfg["seed"])
    n_trees = args.cpu_trees
    for _ in range(args.warmup):
        cpu_reference(prob, cfg, max(1, n_trees // 4))
    times = []
    desc = kind = None
    cores = 1
    for _ in range(args.steps):
        dt, desc, kind, cores = cpu_reference(prob, cfg, n_trees)
        times.append(dt)
    sec = sum(times) / len(times)
    value = cfg["S"] * cfg["L"] * n_trees / sec
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(cfg, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def run_ours(args, cfg):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: dodonaphy_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        quiet_nccl()
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from dodonaphy_b200 import _lib, synth
    from dodonaphy_b200.engine import SubstModel
    from dodonaphy_b200.pipeline import HotPath

    S, L, B, D, K = cfg["S"], cfg["L"], cfg["B"], cfg["D"], cfg["K"]
    prob = synth.make_problem(S, L, B * world, D=D, seed=cfg["seed"])
    xs = prob["xs"][rank * B:(rank + 1) * B]
    model = SubstModel.gtr(prob["rates"], prob["freqs"])
    rates = SubstModel.discrete_gamma(cfg["alpha"], K)
    hp = HotPath(prob["masks"], prob["weights"], model, rates=rates, mix=K > 1, tau=cfg["tau"])
    if args.grid:
        hp.prune.grid_hint = args.grid
    if args.threads:
        hp.prune.threads_hint = args.threads
    lib = _lib.load()
    x_dev = torch.as_tensor(xs).cuda()
    x_pin = torch.as_tensor(xs).pin_memory()
    ll_pin = torch.empty(B, dtype=torch.float64).pin_memory()
    gx_pin = torch.empty((B, S, D), dtype=torch.float64).pin_memory()
    gathered = torch.empty(B * world, dtype=torch.float64, device="cuda") if world > 1 else None

    from dodonaphy_b200.pipeline import GraphedStep

    def step_direct():  # kernel by kernel: used for the per-kernel CUDA events and the launch count
        return hp.value_and_grad(x_dev)

    graph = GraphedStep(hp, x_dev, grad=True)  # the step as the product runs it: captured once, replayed

    def step():
        r = graph.run(x_dev)
        if world > 1:
            dist.all_gather_into_tensor(gathered, r["lnL"])
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_direct()
        r = step()
    barrier()
    assert torch.isfinite(r["lnL"]).all(), "non-finite lnL in warm-up"
    assert int((r["flags"] & 12).sum().item()) == 0, "soft-NJ exactness guard tripped"

    # ---- device-timed region: exactly K steps
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b_ in evs:  # instantiate the handles
        a.record()
        b_.record()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.2)
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for i in range(args.steps):
        step()
    t_end.record()
    barrier()
    ms_total = t_start.elapsed_time(t_end)
    # the same step launched kernel by kernel (outside the timed region): CUDA events around the pruning kernel for
    # the roofline line, and the number of launches one step consists of (a replay launches the same kernels)
    launches0 = _lib.launch_count()
    for i in range(args.steps):
        lib.dodo_set_timing_events(evs[i][0].cuda_event, evs[i][1].cuda_event)
        step_direct()
    torch.cuda.synchronize()
    lib.dodo_set_timing_events(None, None)
    launches = _lib.launch_count() - launches0
    kern_ms = [a.elapsed_time(b_) for a, b_ in evs]

    # ---- end-to-end region: host buffers in, host buffers out, every step
    for _ in range(2):
        hp.value_and_grad_host(x_pin, ll_pin, gx_pin)
        torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        rr = hp.value_and_grad_host(x_pin, ll_pin, gx_pin)
        if world > 1:
            dist.all_gather_into_tensor(gathered, rr["lnL"])
        torch.cuda.current_stream().synchronize()  # the caller reads ll_pin / gx_pin now
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None  # sampled over both timed regions

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = t.tolist()
    ms_step = ms_total / args.steps
    units = S * L * B * world
    value = units / (ms_step * 1e-3)
    e2e_value = units / (ms_e2e / args.steps * 1e-3)

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = algorithmic_bytes_fwd_bwd(S, K) * L * B
        k_ms = sum(kern_ms) / len(kern_ms)
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        traffic = traffic_note = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                tj = json.load(fh)
            if tj.get("workload") == args.workload and tj.get("prune_cu_sha256") == kernel_source_sha():
                traffic = tj.get("dram_bytes_per_launch")
                traffic_note = tj.get("source")
            else:
                traffic_note = "profiles/roofline_traffic.json was captured for another kernel source or workload: not quoted"
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": config_dict(cfg, world),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(x_pin.numel() * 8),
                        "d2h_bytes_per_step": int((ll_pin.numel() + gx_pin.numel()) * 8), "ms_per_step": ms_e2e / args.steps},
                "notes": ("K = 4 discrete-gamma MIXTURE (lnL = sum_p w_p log mean_k f_pk) does not exist in the reference (no rate "
                          "heterogeneity): parity of that mode is pinned to the restatement only; K = 1 and the reference's summed "
                          "category axis are pinned to the reference itself at this size (tests/golden/prune_c2.npz, chain_c2.npz)"),
                "gpu_launches": int(launches),
                "launch_mode": "one CUDA graph replay per step (pipeline.GraphedStep); gpu_launches counts the kernels of the K steps, counted from the same steps launched directly",
                "roofline": {"bound": "hbm", "kernel": "dodo::prune_kernel<1,true,false,1,0> (MODE=grad, resident tables, no rescaling, 1 pattern/thread, generic P: fused P(t)+pruning forward + branch-length gradient)", "achieved": achieved,
                             "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                             "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms,
                             "kernel_share_of_step": k_ms / ms_step, "traffic": traffic, "traffic_source": traffic_note},
                "clocks": clocks}
        if world == 1 and not args.no_cpu:
            from dodonaphy_b200 import synth as _s  # noqa: F401

            dt, desc, kind, cores = cpu_reference(prob, cfg, args.cpu_trees)
            line["cpu_baseline"] = {"value": S * L * args.cpu_trees / dt, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": desc, "seconds": dt}
            n1 = max(1, args.cpu_trees // 8)  # the same arithmetic on one thread (SURVEY.md 8d asks for both)
            dt1, _, _, _ = cpu_reference(prob, cfg, n1, threads=1)
            line["cpu_baseline"]["single_thread"] = {"value": S * L * n1 / dt1, "samples": n1, "seconds": dt1}
    strong = None
    if not args.no_strong:
        strong = strong_legs(args, rank, world, dist)
    if rank == 0:
        if strong is not None:
            line["strong"] = strong
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# Strong-scaling legs: the multi-GPU configurations BASELINE.json names, total work FIXED as N grows,
# measured in the same job as the headline.  Each leg reports the time of the whole problem on ONE rank
# (rank 0 alone, the others idle at the barrier) and sharded over all N ranks, device-timed (CUDA
# events, barrier + synchronize on both sides, max over ranks).
# ---------------------------------------------------------------------------------------------
def _timed_all_ranks(fn, steps, warmup, world, dist, active=True):
    """ms per call of ``fn`` over ``steps`` calls; ranks with ``active=False`` only take part in the
    barriers (used for the 1-rank reference time inside an N-rank job)."""
    import torch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if active:
        for _ in range(warmup):
            fn()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    if active:
        for _ in range(steps):
            fn()
    b.record()
    barrier()
    t = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def strong_legs(args, rank, world, dist, which=("c2", "c3", "c5")):
    import torch

    from dodonaphy_b200 import dist as ddist, engine, synth
    from dodonaphy_b200.engine import PruneEngine, SubstModel
    from dodonaphy_b200.pipeline import GraphedStep, HotPath

    steps, warmup = max(3, min(args.steps, 10)), 3
    legs = {}

    def finish(name, desc, units, ms1, msn, extra):
        legs[name] = {"config": desc, "n_gpus": world, "ms_1gpu": ms1, "ms": msn, "value": units / (msn * 1e-3), "unit": UNIT,
                      "speedup": ms1 / msn, "efficiency": ms1 / (world * msn), **extra}

    if "c2" in which:
        # (i) the headline workload with 128 samples IN TOTAL (vi.py:328-333 shards by sample)
        cfg = WORKLOADS["c2"]
        S, L, B, D, K = cfg["S"], cfg["L"], cfg["B"], cfg["D"], cfg["K"]
        prob = synth.make_problem(S, L, B, D=D, seed=cfg["seed"])
        hp = HotPath(prob["masks"], prob["weights"], SubstModel.gtr(prob["rates"], prob["freqs"]),
                     rates=SubstModel.discrete_gamma(cfg["alpha"], K), mix=True, tau=cfg["tau"])
        lo, hi = ddist.shard_bounds(B, rank, world)
        x_all, x_loc = torch.as_tensor(prob["xs"]).cuda(), torch.as_tensor(prob["xs"][lo:hi]).cuda()
        gathered = torch.empty(B, dtype=torch.float64, device="cuda")
        ms1 = None
        if rank == 0 or world == 1:
            g1 = GraphedStep(hp, x_all, grad=True)
        ms1 = _timed_all_ranks((lambda: g1.run(x_all)) if rank == 0 else None, steps, warmup, world, dist, active=rank == 0)
        if world > 1:
            gn = GraphedStep(hp, x_loc, grad=True)
            even = B % world == 0

            def step_n():
                r = gn.run(x_loc)
                if even:
                    dist.all_gather_into_tensor(gathered, r["lnL"])
                else:
                    ddist.gather_samples(r["lnL"], B)

            msn = _timed_all_ranks(step_n, steps, warmup, world, dist)
        else:
            msn = ms1
        finish("c2_128_samples_total", cfg["desc"] + " -- 128 samples in total, sharded by sample", S * L * B, ms1, msn,
               {"samples_per_gpu": hi - lo, "exchange": "all_gather of per-sample lnL", "graph": True})
        del hp, x_all, x_loc
        torch.cuda.empty_cache()

    if "c3" in which:
        # (ii) configs[2]: 64 chains in total, 200 taxa x 50k patterns, JC69, forward only (mcmc.py:136-146)
        cfg = SIDE["c3"]
        S, L, B, D = cfg["S"], cfg["L"], cfg["B"], cfg["D"]
        prob = synth.make_problem(S, 20000, B, D=D, seed=1003)
        reps = -(-L // prob["masks"].shape[1])
        masks = np.tile(prob["masks"], (1, reps))[:, :L]
        weights = np.tile(prob["weights"], reps)[:L]
        hp = HotPath(masks, weights, SubstModel.jc69())
        lo, hi = ddist.shard_bounds(B, rank, world)
        x_all, x_loc = torch.as_tensor(prob["xs"]).cuda(), torch.as_tensor(prob["xs"][lo:hi]).cuda()
        gathered = torch.empty(B, dtype=torch.float64, device="cuda")
        if rank == 0 or world == 1:
            g1 = GraphedStep(hp, x_all, grad=False)
        ms1 = _timed_all_ranks((lambda: g1.run(x_all)) if rank == 0 else None, steps, warmup, world, dist, active=rank == 0)
        if world > 1:
            gn = GraphedStep(hp, x_loc, grad=False)

            def step_n():
                r = gn.run(x_loc)
                if B % world == 0:
                    dist.all_gather_into_tensor(gathered, r["lnL"])
                else:
                    ddist.gather_samples(r["lnL"], B)

            msn = _timed_all_ranks(step_n, steps, warmup, world, dist)
        else:
            msn = ms1
        finish("c3_64_chains_total", cfg["desc"] + " -- sharded by chain", S * L * B, ms1, msn,
               {"chains_per_gpu": hi - lo, "exchange": "all_gather of per-chain lnL", "graph": True})
        del hp, x_all, x_loc
        torch.cuda.empty_cache()

    if "c5" in which:
        # (iii) configs[4]: 500 taxa x 1M patterns, lnL + gradient, site blocks + ONE all_reduce of 1 + 2S-2 doubles
        cfg = SIDE["c5"]
        S, L, D = cfg["S"], cfg["L"], cfg["D"]
        prob = synth.make_problem(S, 20000, 1, D=D, seed=1005)
        reps = -(-L // prob["masks"].shape[1])
        masks = np.tile(prob["masks"], (1, reps))[:, :L]
        weights = np.tile(prob["weights"], reps)[:L]
        peel, blens = engine.nj_hard(engine.pdm_forward(torch.as_tensor(prob["xs"][0]).cuda(), -1.0, "numpy"))
        ms1 = None
        if rank == 0 or world == 1:
            full = ddist.SiteShardedStep(PruneEngine(masks, weights), SubstModel.jc69(), peel, blens, grad=True, reduce=False)
        ms1 = _timed_all_ranks((lambda: full.run()) if rank == 0 else None, steps, warmup, world, dist, active=rank == 0)
        extra = {"exchange": "one all_reduce of 1 + 2S-2 = 999 doubles per evaluation (NCCL), captured in the step's CUDA graph"}
        if world > 1:
            ll_full = full.lnl.clone() if rank == 0 else None
            if rank == 0:
                del full
            torch.cuda.empty_cache()
            lo, hi = ddist.shard_bounds(L, rank, world)
            part = ddist.SiteShardedStep(PruneEngine(np.ascontiguousarray(masks[:, lo:hi]), weights[lo:hi]), SubstModel.jc69(),
                                         peel, blens, grad=True)
            msn = _timed_all_ranks(lambda: part.run(), steps, warmup, world, dist)
            extra.update(patterns_per_gpu=hi - lo, graph=part.graph is not None)
            if rank == 0:
                extra["lnL_rel_diff_vs_1gpu"] = abs(float(part.lnl[0]) - float(ll_full[0])) / abs(float(ll_full[0]))
        else:
            msn = ms1
            extra.update(patterns_per_gpu=L, graph=full.graph is not None)
        finish("c5_1M_site_patterns", cfg["desc"] + " -- site blocks", S * L, ms1, msn, extra)
    return legs


def run_side(args):
    """Timings of the non-headline configurations (device time, CUDA events)."""
    import torch

    from dodonaphy_b200 import dist as ddist, engine, synth
    from dodonaphy_b200.engine import PruneEngine, SubstModel
    from dodonaphy_b200.pipeline import HotPath

    cfg = SIDE[args.side]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        import torch.distributed as dist

        quiet_nccl()
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))

    def timed(fn, n):
        for _ in range(2):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    out = {"side": args.side, "config": cfg["desc"], "n_gpus": world}
    S, B, D = cfg["S"], cfg["B"], cfg["D"]
    rng = np.random.default_rng(7)
    if args.side == "c1":
        from dodonaphy_b200.pipeline import GraphedStep

        prob = synth.make_problem(S, cfg["L"], 8, D=D, seed=1001)
        hp = HotPath(prob["masks"], prob["weights"], SubstModel.jc69())
        xs = torch.as_tensor(prob["xs"]).cuda()
        chains, samples = xs[:B].contiguous(), xs
        g_mcmc, g_vi = GraphedStep(hp, chains, grad=False), GraphedStep(hp, samples, grad=True)
        out["mcmc_sweep_ms"] = timed(lambda: hp.value(chains), 200)
        out["mcmc_sweep_graph_ms"] = timed(lambda: g_mcmc.run(chains), 200)
        out["vi_step_ms"] = timed(lambda: hp.value_and_grad(samples), 200)
        out["vi_step_graph_ms"] = timed(lambda: g_vi.run(samples), 200)
    elif args.side in ("gtr", "jac"):
        c2 = WORKLOADS["c2"]
        prob = synth.make_problem(S, cfg["L"], B, D=D, seed=c2["seed"])
        grates = SubstModel.discrete_gamma(c2["alpha"], c2["K"])
        hp = HotPath(prob["masks"], prob["weights"], SubstModel.gtr(prob["rates"], prob["freqs"]), rates=grates, mix=True, tau=c2["tau"])
        x = torch.as_tensor(prob["xs"]).cuda()
        out["fixed_model_step_ms"] = timed(lambda: hp.value_and_grad(x), 10)
        if args.side == "gtr":
            from dodonaphy_b200.base_model import BaseModel

            tips = [torch.tensor(np.stack([(m >> j) & 1 for j in range(4)]).astype(np.float64)) for m in prob["masks"]]
            m = BaseModel("vi", tips, torch.tensor(prob["weights"]), D, soft_temp=c2["tau"], model_name="GTR", gamma_shape=c2["alpha"],
                          n_rate_categories=c2["K"])
            m.phylomodel.sub_rates = torch.tensor(prob["rates"])
            m.phylomodel.freqs = torch.tensor(prob["freqs"])
            xl = x.clone().requires_grad_(True)

            def learn_step():
                ll, _, _ = m.log_likelihood_batch(xl)
                ll.sum().backward()

            out["learnable_gtr_step_ms"] = timed(learn_step, 10)
            out["device_part_ms"] = timed(lambda: m._hotpath().value_and_grad_model(x), 10)
            out["ratio_to_fixed_model"] = out["learnable_gtr_step_ms"] / out["fixed_model_step_ms"]
        else:
            dec = hp.decode_checked(x)
            out["jacobian_ms"] = timed(lambda: hp.blens_jacobian(x, dec["peel"], dec["flags"], logdet=False), 10)
            out["jacobian_logdet_ms"] = timed(lambda: hp.blens_jacobian(x, dec["peel"], dec["flags"], logdet=True), 10)
            out["jacobian_by_replays_ms"] = timed(lambda: hp.blens_jacobian(x, dec["peel"], dec["flags"], logdet=False, replay=True), 3)
    elif args.side == "c4":
        lo, hi = ddist.shard_bounds(B, rank, world)
        xs = torch.as_tensor(rng.normal(0, 0.05, (B, S, D))[lo:hi]).cuda()
        pdm = engine.pdm_forward(xs, -1.0, "torch")
        out["pdm_ms"] = timed(lambda: engine.pdm_forward(xs, -1.0, "torch"), 5)
        out["nj_hard_ms"] = timed(lambda: engine.nj_hard(pdm), 1)
        out["nj_soft_ms"] = timed(lambda: engine.nj_soft(pdm, 1e-8), 1)
        out["trees_per_gpu"] = hi - lo
        out["nj_hard_GBps_algorithmic"] = (hi - lo) * 4 * (S ** 3 - S) / 3 / (out["nj_hard_ms"] * 1e-3) / 1e9
    else:
        L = cfg["L"]
        prob = synth.make_problem(S, min(L, 20000), B * (world if args.side == "c3" else 1), D=D, seed=1003)
        reps = -(-L // prob["masks"].shape[1])
        masks = np.tile(prob["masks"], (1, reps))[:, :L]  # columns repeated to reach the full length
        weights = np.tile(prob["weights"], reps)[:L]
        if args.side == "c3":
            hp = HotPath(masks, weights, SubstModel.jc69())
            lo, hi = ddist.shard_bounds(B * world, rank, world)
            xs = torch.as_tensor(prob["xs"][lo:hi]).cuda()
            ms = timed(lambda: hp.value(xs), 3)
            out.update(ms_per_sweep=ms, value=S * L * B * world / (ms * 1e-3), unit=UNIT)
        else:
            lo, hi = ddist.shard_bounds(L, rank, world)
            eng = PruneEngine(np.ascontiguousarray(masks[:, lo:hi]), weights[lo:hi])
            peel, blens = engine.nj_hard(engine.pdm_forward(torch.as_tensor(prob["xs"][0]).cuda(), -1.0, "numpy"))

            def ev():
                ll, g = eng.loglik(peel, blens, SubstModel.jc69(), grad=True)
                if world > 1:
                    ddist.reduce_site_blocks(ll, g)

            ms = timed(ev, 3)
            out.update(ms_per_eval=ms, value=S * L / (ms * 1e-3), unit=UNIT, patterns_per_gpu=hi - lo)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-trees", type=int, default=128, help="samples in the bounded CPU-baseline pass")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling legs (configs[1] with 128 samples in total, configs[2], configs[4])")
    ap.add_argument("--grid", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--side", default=None, choices=sorted(SIDE), help="time a non-headline configuration instead")
    args = ap.parse_args()
    if args.side:
        return run_side(args)
    cfg = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
