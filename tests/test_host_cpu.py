"""CPU tests (no GPU): host-side logic of the package, the C-ABI library's exported symbols, the
no-CPU-fallback contract, and the multi-process sharding logic on the gloo backend (world size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ C ABI
def test_library_builds_and_exports_every_declared_symbol():
    from dodonaphy_b200 import _lib, build

    lib_path = build.build()
    handle = ctypes.CDLL(lib_path)
    header = open(os.path.join(ROOT, "include", "dodonaphy_b200.h")).read()
    declared = set(re.findall(r"\b(dodo_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in the header"
    for name in declared:
        assert hasattr(handle, name), f"{name} is declared in include/dodonaphy_b200.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert handle.dodo_version() >= 100


def test_plan_and_argument_validation_without_gpu():
    """dodo_prune_plan is pure host arithmetic: sizes, tiling, error codes."""
    from dodonaphy_b200 import _lib

    lib = _lib.load()
    plan = _lib.PrunePlan()
    assert lib.dodo_prune_plan(128, 64, 10000, 4, _lib.PRUNE_GRAD | _lib.PRUNE_MIX, 444, 0, ctypes.byref(plan)) == 0
    assert (plan.threads, plan.tile_patterns, plan.grid) == (256, 256, 444)
    assert plan.n_tiles == 40 and plan.n_items == 128 * 40
    assert plan.ws_bytes > plan.off_scratch > plan.off_pmat > 0
    assert lib.dodo_prune_plan(1, 1, 10, 1, 0, 1, 0, ctypes.byref(plan)) == _lib.DODO_EINVAL
    assert b"S>=2" in lib.dodo_last_error()
    assert lib.dodo_prune_plan(1, 8, 10, 9, 0, 1, 0, ctypes.byref(plan)) == _lib.DODO_EINVAL
    with pytest.raises(ValueError):
        _lib.check(lib.dodo_prune_plan(1, 8, 10, 1, _lib.PRUNE_GRAD | _lib.PRUNE_GMATS, 1, 0, ctypes.byref(plan)))
    assert lib.dodo_nj_workspace_bytes(3, 10) == 3 * (2 * 32 * 32 + 48) * 8  # stride 10 + 16 spare slots -> 32
    assert lib.dodo_pdm_workspace_bytes(2, 5, 3) == 2 * 5 * 5 * 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a box without CUDA")
def test_no_cpu_fallback():
    from dodonaphy_b200 import engine

    with pytest.raises(RuntimeError, match="no CPU path"):
        engine.PruneEngine(np.ones((4, 8), dtype=np.uint8), np.ones(8))
    with pytest.raises(RuntimeError):
        engine.pdm_forward(torch.zeros(4, 3, dtype=torch.float64))


def test_product_code_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dodonaphy_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports oracle/"


# ------------------------------------------------------------------ host helpers
def test_pack_tips_and_validate_peel():
    from dodonaphy_b200.engine import pack_tip_partials, validate_peel
    from oracle.phylo import mask_to_partial

    masks = np.array([[1, 2, 4, 8, 15, 5], [3, 9, 6, 12, 14, 7]], dtype=np.uint8)
    assert np.array_equal(pack_tip_partials([mask_to_partial(m) for m in masks]), masks)
    assert np.array_equal(pack_tip_partials([torch.tensor(mask_to_partial(m)) for m in masks]), masks)
    with pytest.raises(ValueError):
        pack_tip_partials([np.full((4, 3), 0.5)])
    validate_peel(np.array([[0, 1, 4], [2, 3, 5], [4, 5, 6]]), 4)
    for bad in ([[0, 1, 4], [2, 3, 5], [4, 4, 6]], [[0, 1, 4], [2, 3, 4], [4, 5, 6]], [[0, 1, 9], [2, 3, 5], [4, 5, 6]]):
        with pytest.raises(ValueError):
            validate_peel(np.array(bad), 4)
    with pytest.raises(ValueError):
        validate_peel(np.zeros((2, 3)), 4)


def test_subst_model_matches_oracle():
    from dodonaphy_b200.engine import SubstModel
    from oracle import phylo as ophylo

    rng = np.random.default_rng(0)
    rates, freqs = rng.dirichlet(np.ones(6)), rng.dirichlet(np.ones(4))
    m = SubstModel.gtr(rates, freqs)
    np.testing.assert_allclose(m.Q, ophylo.gtr_q(rates, freqs), rtol=1e-14)
    t = 0.37
    P = (m.U * np.exp(m.lam * t)) @ m.Uinv
    np.testing.assert_allclose(P, ophylo.gtr_p_t(np.array([t]), rates, freqs)[0], rtol=1e-12)
    np.testing.assert_allclose(SubstModel.discrete_gamma(0.5, 4), ophylo.discrete_gamma_rates(0.5, 4), rtol=1e-14)
    jc = SubstModel.jc69()
    np.testing.assert_allclose(jc.Q.sum(axis=1), 0.0, atol=1e-15)
    with pytest.raises(ValueError):
        SubstModel.gtr(np.ones(5), freqs)


def test_phylomodel_surface():
    """test/test_phylo.py:276-300 against the host-side PhyloModel mirror."""
    from dodonaphy_b200.phylomodel import PhyloModel

    b = torch.tensor([0.1], dtype=torch.double)
    pm = PhyloModel("GTR", freqs=torch.full([4], 0.25, dtype=torch.double), sub_rates=torch.full([6], 1 / 6, dtype=torch.double))
    assert torch.allclose(pm.JC69_p_t(b)[:, 0], pm.GTR_p_t(b))
    for blens_in, size in (([[1.0, 1.0]], [1, 2, 4, 4]), ([[1.0]], [1, 1, 4, 4]), ([[1.0], [1.0]], [2, 1, 4, 4])):
        assert list(PhyloModel("GTR").GTR_p_t(torch.tensor(blens_in, dtype=torch.double)).shape) == size
    with pytest.raises(ValueError):
        PhyloModel("HKY").get_transition_mats(b)
    assert pm._sub_rates.shape == (5,) and pm._freqs.shape == (3,) and pm._freqs.requires_grad
    assert PhyloModel("JC69") == "JC69"


def test_compress_alignment_matches_oracle():
    from dodonaphy_b200 import phylo
    from oracle import phylo as ophylo

    class _NS:
        def labels(self):
            return ["a", "b", "c"]

    class _Aln:
        taxon_namespace = _NS()

        def sequences(self):
            return ["ACGTNACGT-", "ACGTRACGTY", "AAGTNAAGT?"]

    parts, w = phylo.compress_alignment(_Aln())
    tips, w_ref, _ = ophylo.compress_alignment(_Aln().sequences())
    assert w.tolist() == w_ref.tolist()
    for p, t in zip(parts, tips):
        assert np.array_equal(p.numpy(), t)


def test_newick_round_trip_and_reference_convention():
    """test/test_tree.py:19-43: peel/blens of the reference's 6-taxon example."""
    from dodonaphy_b200 import newick

    data = ("(bird:6.000000e-02,((chimp:5.000000e-02,dog:2.000000e-02):17.000000e-02,"
            "(crocodile:4.000000e-02,emu:3.000000e-02):18.000000e-02):19.000000e-02,kangaroo:1.000000e-02);")
    labels = ["bird", "chimp", "dog", "crocodile", "emu", "kangaroo"]
    peel, blens = newick.newick_to_peel(data, labels)
    assert peel.tolist() == [[1, 2, 6], [3, 4, 7], [6, 7, 8], [0, 8, 9], [5, 9, 10]]
    np.testing.assert_allclose(blens, [0.06, 0.05, 0.02, 0.04, 0.03, 0.01, 0.17, 0.18, 0.19, 0.0])
    again, b2 = newick.newick_to_peel(newick.peel_to_newick(peel, blens, labels), labels)
    assert again.shape == peel.shape and sorted(b2.tolist()) == sorted(blens.tolist())


def test_tree_writer_matches_reference_text(tmp_path):
    """tree.tree_to_newick (dodonaphy/tree.py:207-267), strings produced by the real reference."""
    import json

    from dodonaphy_b200 import tree

    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "tree_newick.json")))
    assert len(cases) >= 8
    for c in cases:
        name_id = {f"T{i + 1}": i for i in range(c["S"])}
        assert tree.tree_to_newick(name_id, np.array(c["peel"]), np.array(c["blens"]), c["rooted"]) == c["newick"]
    assert tree.tree_to_newick({}, [], torch.tensor([]), True) == ";"  # test/test_tree.py:9-16
    with pytest.raises(TypeError):
        tree.tree_to_newick([], [], [])
    c = cases[-2]
    name_id = {f"T{i + 1}": i for i in range(c["S"])}
    peels, blens = np.array([c["peel"]] * 3), np.array([c["blens"]] * 3)
    tree.save_trees(str(tmp_path), "samples", peels, blens, 10, np.array([-1.0, -2.0, -3.0]), np.zeros(3), name_id, lnQ=np.ones(3))
    lines = open(tmp_path / "samples.t").read().splitlines()
    assert len(lines) == 3 and lines[1].startswith("\ttree STATE_11 [&lnL=-2.000000, &lnPr=0.000000, &lnQ=1.000000] = [&U] (")


def test_parent_id_vector():
    from dodonaphy_b200.phylo import get_parent_id_vector

    peel = np.array([[0, 1, 4], [2, 4, 5], [3, 5, 6]])
    assert get_parent_id_vector(peel, rooted=False) == [4, 4, 5, 5, 5]
    assert get_parent_id_vector(peel, rooted=True) == [4, 4, 5, 6, 5, 6]


def test_shard_bounds_cover_everything():
    from dodonaphy_b200.dist import shard_bounds

    for n in (1, 7, 128, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_synthetic_generator_is_seeded_and_consistent():
    from dodonaphy_b200 import synth
    from dodonaphy_b200.engine import validate_peel

    a = synth.make_problem(12, 50, 3, D=3, seed=4)
    b = synth.make_problem(12, 50, 3, D=3, seed=4)
    assert np.array_equal(a["masks"], b["masks"]) and np.array_equal(a["xs"], b["xs"])
    validate_peel(a["guide_peel"], 12)
    assert a["masks"].dtype == np.uint8 and set(np.unique(a["masks"])) <= {1, 2, 4, 8, 15}
    assert (a["guide_blens"] > 0).all() and a["weights"].min() >= 1


def test_gamma_dirichlet_prior_mrbayes_known_answers():
    """test/test_phylo.py:183-211: 17.34844 and 13.56829 (+-1e-5), and the batched form."""
    from dodonaphy_b200 import newick, priors

    labels = [str(i) for i in range(1, 7)]
    t0 = "(6:2.000000e-02,((5:2.000000e-02,2:2.000000e-02):2.000000e-02,(4:2.000000e-02,3:2.000000e-02):2.000000e-02):2.000000e-02,1:2.000000e-02);"
    t1 = "((2:1.179257e-01,(6:1.047047e-03,3:1.426334e-03):7.713732e-02):7.894008e-02,(4:2.848264e-03,5:2.711671e-03):1.879317e-03,1:4.419083e-03);"
    b0 = torch.tensor(newick.newick_to_peel(t0, labels)[1])
    b1 = torch.tensor(newick.newick_to_peel(t1, labels)[1])
    assert priors.compute_prior_gamma_dir(b0, aT=torch.ones(1), bT=torch.full((1,), 0.1), a=torch.ones(1), c=torch.ones(1)).item() == pytest.approx(17.34844, abs=1e-5)
    assert priors.compute_prior_gamma_dir(b1).item() == pytest.approx(13.56829, abs=1e-5)
    both = priors.compute_prior_gamma_dir(torch.stack([b0, b1]))
    assert both[0].item() == pytest.approx(17.34844, abs=1e-5) and both[1].item() == pytest.approx(13.56829, abs=1e-5)
    assert priors.compute_prior_exponential(torch.full((3, 10), 0.1), rate=10.0).shape == (3,)


# ------------------------------------------------------------------ world_size 2 over gloo
_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from dodonaphy_b200 import dist as ddist, synth
from oracle import hyp, nj, phylo

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
prob = synth.make_problem(9, 101, 5, D=3, seed=21)
tips_full = [phylo.mask_to_partial(m) for m in prob["masks"]]
peels, blens = zip(*[nj.nj_hard(hyp.get_pdm_np(x)) for x in prob["xs"]])
full = np.array([phylo.compute_LL_np(tips_full, prob["weights"], p, b) for p, b in zip(peels, blens)])

# (1) sample sharding: each rank evaluates its slice, all_gather reassembles the batch in order
lo, hi = ddist.shard_bounds(5, rank, 2)
mine = ddist.shard_samples(torch.arange(5))
assert mine.tolist() == list(range(lo, hi))
local = torch.tensor(full[lo:hi])
got = ddist.gather_samples(local, 5)
assert np.array_equal(got.numpy(), full), (got, full)

# (2) site sharding: per-block lnL (+ gradient vector) summed with one all_reduce
masks_blk, w_blk = ddist.shard_sites(prob["masks"], prob["weights"])
tips_blk = [phylo.mask_to_partial(m) for m in masks_blk]
part = torch.tensor([phylo.compute_LL_np(tips_blk, w_blk, p, b) for p, b in zip(peels, blens)], dtype=torch.float64)
gpart = torch.full((5, 16), float(rank + 1), dtype=torch.float64)
tot, gtot = ddist.reduce_site_blocks(part, gpart)
np.testing.assert_allclose(tot.numpy(), full, rtol=1e-12)
assert torch.equal(gtot, torch.full((5, 16), 3.0, dtype=torch.float64))
tot_only, none = ddist.reduce_site_blocks(part)
assert none is None
np.testing.assert_allclose(tot_only.numpy(), full, rtol=1e-12)
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_sharding_world_size_2_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{out}"
        assert f"rank {r} ok" in out


def test_chart_changes_against_reference_fixture():
    """Chyp_torch chart changes (Chyp_torch.pyx:39-117,212-362,425-457) vs outputs of the reference.
    The reference accumulates the two log-Jacobians, real2ball_LADJ and hyper_to_poincare in float32
    buffers (``torch.zeros(1)`` / ``torch.zeros(dim)``), hence the looser tolerance on those."""
    import torch

    from conftest import load_golden
    from dodonaphy_b200 import Chyp_torch as C

    g = load_golden("charts.npz")
    x, mu = torch.tensor(g["x"]), torch.tensor(g["mu"])
    np.testing.assert_allclose(C.t02p(x).numpy(), g["t02p"], rtol=0, atol=5e-8)
    p, jac = C.t02p(x, mu, get_jacobian=True)
    np.testing.assert_allclose(p.numpy(), g["t02p_mu"], rtol=0, atol=5e-8)
    np.testing.assert_allclose(jac.numpy(), g["t02p_mu_jac"], rtol=1e-5)
    pr = torch.tensor(g["t02p_mu"])
    q, qjac = C.p2t0(pr, mu, get_jacobian=True)
    np.testing.assert_allclose(q.numpy(), g["p2t0_mu"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(qjac.numpy(), g["p2t0_mu_jac"], rtol=1e-5)
    np.testing.assert_allclose(C.p2t0(pr).numpy(), g["p2t0"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(C.p2t0(C.t02p(x)).numpy(), g["x"], rtol=0, atol=1e-12)  # inverse charts at the origin
    np.testing.assert_allclose(C.real2ball(x, 2.0).numpy(), g["real2ball"], rtol=1e-14)
    np.testing.assert_allclose(C.ball2real(pr).numpy(), g["ball2real"], rtol=1e-14)
    np.testing.assert_allclose(C.real2ball_LADJ(x, 2.0).numpy(), g["ladj"], rtol=1e-6)
    muh, xh = C.project_up(mu), C.project_up(x)
    np.testing.assert_allclose(C.tangent_to_hyper(muh, x).numpy(), g["tangent_to_hyper"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(C.parallel_transport(xh - muh, muh, xh.roll(-1, 0)).numpy(), g["transport"], rtol=1e-13, atol=1e-15)
    v = C.exp_map_inverse(xh, muh)
    np.testing.assert_allclose(v.numpy(), g["exp_inverse"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(C.exponential_map(muh, v).numpy(), g["exp_map"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(C.hyper_to_poincare_jacobian(xh).numpy(), g["h2p_jac"], rtol=1e-13)
    # NumPy twins
    from dodonaphy_b200 import Chyp_np as N

    xn, mun = g["x"], g["mu"]
    xhn, muhn = N.project_up_2d(xn), N.project_up_2d(mun)
    np.testing.assert_allclose(N.tangent_to_hyper(muhn, xn), g["np_tangent_to_hyper"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(N.unwrap_2d(xhn), g["np_unwrap"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(N.exp_map_inverse(xhn, muhn), g["np_exp_inverse"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(N.parallel_transport(xhn - muhn, muhn, np.roll(xhn, -1, 0)), g["np_transport"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(N.hyper_to_poincare(xhn), g["np_h2p"], rtol=1e-14)
    np.testing.assert_allclose(N.hyper_to_poincare_jacobian(xhn), g["np_h2p_jac"], rtol=1e-13)
    np.testing.assert_array_equal(N.project_up_jacobian(xn), g["np_project_up_jacobian"])


# ------------------------------------------------------------------ drop-in by module alias
def test_dropin_binds_the_reference_callers_to_this_package():
    """dropin.install(): the reference's own vi.py / chain.py / hmap.py / base_model.py (unmodified, from
    oracle/_ref) import and see THIS package's hot-path modules under the reference's names; uninstall()
    restores the reference's own.  (The numerical comparison of both bindings runs on the GPU:
    tests/test_dropin_gpu.py.)"""
    import importlib

    from oracle import refload

    r = refload.load()
    if not (r.compiled and r.python):
        pytest.skip("the reference package is not present in oracle/_ref (built only where /root/reference exists)")
    import dodonaphy_b200
    from dodonaphy_b200 import dropin

    try:
        dropin.install()
        vi = importlib.import_module("dodonaphy.vi")
        chain = importlib.import_module("dodonaphy.chain")
        hmap = importlib.import_module("dodonaphy.hmap")
        base = importlib.import_module("dodonaphy.base_model")
        mcmc = importlib.import_module("dodonaphy.mcmc")
        assert vi.Chyp_torch is dodonaphy_b200.Chyp_torch and vi.peeler is dodonaphy_b200.peeler
        assert chain.Cphylo is dodonaphy_b200.Cphylo and chain.Cpeeler is dodonaphy_b200.Cpeeler
        assert chain.Chyp_np is dodonaphy_b200.Chyp_np and hmap.Chyp_torch is dodonaphy_b200.Chyp_torch
        assert base.calculate_treelikelihood is dodonaphy_b200.phylo.calculate_treelikelihood
        assert mcmc.Cphylo is dodonaphy_b200.Cphylo
        # every attribute the reference's callers take from the aliased modules exists here
        import re

        for caller in ("vi", "chain", "mcmc", "base_model", "hmap", "laplace", "utils"):
            src = open(os.path.join(refload.REF_PY_DIR, caller + ".py")).read()
            for mod, attr in set(re.findall(r"\b(Chyp_torch|Chyp_np|Cphylo|Cpeeler|peeler|soft|phylo)\.([A-Za-z_][A-Za-z_0-9]*)\(", src)):
                assert hasattr(getattr(dodonaphy_b200, mod), attr), f"{caller}.py uses {mod}.{attr}"
        # constructing the reference's classes needs no GPU (tips are uploaded lazily, on first use)
        tips = [torch.tensor(np.eye(4)[:, [0, 1, 2, 3, 0]]) for _ in range(5)]
        v = vi.DodonaphyVI(tips, torch.ones(5, dtype=torch.int64), 2, embedder="up", connector="nj", soft_temp=1e-8, path_write=None)
        assert v.S == 5 and v.L == 5
    finally:
        dropin.uninstall()
    vi = importlib.import_module("dodonaphy.vi")
    assert vi.Chyp_torch is r.Chyp_torch


def test_gamma_dirichlet_prior_numpy_twin_matches_compiled_reference():
    """Cphylo.compute_prior_gamma_dir_np (Cphylo.pyx:53-114; chain.py:129) incl. its epsilon floors."""
    from dodonaphy_b200 import Cphylo
    from oracle import refload

    r = refload.load()
    rng = np.random.default_rng(0)
    for n in (3, 8, 27):
        b = rng.random(2 * n - 2) * 0.1
        b[1] = 0.0
        if r.compiled:
            assert Cphylo.compute_prior_gamma_dir_np(b) == pytest.approx(float(r.Cphylo.compute_prior_gamma_dir_np(b)), rel=1e-13)
            assert Cphylo.compute_prior_gamma_dir_np(b, 2.0, 0.3, 1.5, 0.7) == pytest.approx(
                float(r.Cphylo.compute_prior_gamma_dir_np(b, 2.0, 0.3, 1.5, 0.7)), rel=1e-13)
