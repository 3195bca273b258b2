"""The reference's OWN caller classes running on this library, unchanged (north_star:
"so the existing `dodo --infer vi|mcmc` drivers run unchanged").

``oracle/_ref/dodonaphy`` holds the unmodified reference package (its compiled Cython modules plus its
pure-Python modules, placed there by oracle/build_ref.py; git-ignored, travels with the snapshot).
Every test runs the same reference code twice in this process:

  1. as it is -- its own Chyp_torch / Chyp_np / peeler / Cpeeler / phylo / Cphylo / soft (CPU), then
  2. after ``dodonaphy_b200.dropin.install()`` -- the same classes, the hot-path modules now resolving
     to the CUDA implementations behind the C ABI,

and compares.  Bars: peels bit-exact, values and gradients 1e-9 relative.
"""
import importlib
import sys
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _TN:
    def __init__(self, labels):
        self._l = labels

    def labels(self):
        return self._l


def _ref_partials(masks):
    from oracle.phylo import mask_to_partial

    return [torch.tensor(mask_to_partial(m)) for m in masks]


@pytest.fixture
def reference():
    """Yields load(aliased: bool) -> namespace of the reference's caller modules, and restores the
    un-aliased package afterwards."""
    from oracle import refload

    r = refload.load()
    if not (r.compiled and r.python):
        pytest.fail("oracle/_ref lacks the reference package: run `python oracle/build_ref.py` where /root/reference exists")
    from dodonaphy_b200 import dropin

    def load(aliased):
        (dropin.install if aliased else dropin.uninstall)()
        if not aliased:  # callers imported earlier hold whatever was bound then
            for name in dropin._CALLERS:
                if name not in dropin.ALIASES:
                    sys.modules.pop(f"dodonaphy.{name}", None)
        ns = {name: importlib.import_module(f"dodonaphy.{name}") for name in ("vi", "chain", "hmap", "base_model")}
        import dodonaphy_b200

        for name in dropin.ALIASES:
            bound = getattr(ns["base_model"], name, None) or getattr(ns["vi"], name, None) or getattr(ns["chain"], name, None)
            if bound is not None:
                assert (bound is getattr(dodonaphy_b200, name)) == aliased, name
        return ns

    torch.set_default_dtype(torch.float64)
    try:
        yield load
    finally:
        dropin.uninstall()


def _vi_once(vim, prob, embedder, model_name, seed):
    S, D = prob["S"], prob["D"]
    torch.manual_seed(17)  # the GTR model draws its initial rates / frequencies from the global generator
    vi = vim.DodonaphyVI(_ref_partials(prob["masks"]), torch.tensor(prob["weights"]), D, embedder=embedder, connector="nj",
                         soft_temp=1e-8, path_write=None, model_name=model_name)
    vi.set_params_optim({"leaf_mu": torch.tensor(prob["x"]).clone(), "leaf_sigma": torch.full((S, D), -9.0)})
    rec, orig = [], vi.rsample_tree

    def wrapped(*a, **k):
        rec.append(orig(*a, **k))
        return rec[-1]

    vi.rsample_tree = wrapped
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        elbo = vi.elbo_siwae(importance=1)
    elbo.backward()
    t = rec[0]
    out = dict(peel=np.asarray(t["peel"]), blens=t["blens"].detach().numpy().copy(), ln_p=float(t["ln_p"]),
               ln_prior=float(t["ln_prior"]), log_q=float(t["logQ"]), jac=float(t["jacobian"]),
               g_mu=vi.params_optim["leaf_mu"].grad.numpy().copy(), g_sigma=vi.params_optim["leaf_sigma"].grad.numpy().copy(),
               g_curv=vi._curvature.grad.numpy().copy() if vi._curvature.grad is not None else np.zeros(1),
               elbo_less_jac=elbo.item() - (0.0 if t["jacobian"] == -torch.inf else float(t["jacobian"])))
    if model_name == "GTR":
        out["g_rates"] = vi.phylomodel._sub_rates.grad.numpy().copy()
        out["g_freqs"] = vi.phylomodel._freqs.grad.numpy().copy()
    return out


@pytest.mark.parametrize("embedder,model_name", [("up", "JC69"), ("wrap", "JC69"), ("up", "GTR")])
def test_reference_DodonaphyVI_runs_unchanged(reference, embedder, model_name):
    """vi.py:161-202,315-491 -- the reference's DodonaphyVI.elbo_siwae + backward: sample, connect
    (get_pdm -> nj_torch), its autograd Jacobian of the branch lengths (2S-2 backward passes through OUR
    autograd functions), compute_LL, prior; gradients back to leaf_mu / leaf_sigma / curvature / GTR leaves.
    The log|det J J^T| term is rounding noise in the reference itself (two rows of J coincide, DESIGN.md
    section 4), so it is taken out of the compared ELBO; with one importance sample it does not enter
    the gradient."""
    from dodonaphy_b200 import synth

    prob = synth.make_problem(9, 80, 1, D=3, seed=77)
    want = _vi_once(reference(False)["vi"], prob, embedder, model_name, seed=5)
    got = _vi_once(reference(True)["vi"], prob, embedder, model_name, seed=5)
    assert np.array_equal(got["peel"], want["peel"])
    for key in ("ln_p", "ln_prior", "log_q", "elbo_less_jac"):
        assert got[key] == pytest.approx(want[key], rel=1e-9), key
    np.testing.assert_allclose(got["blens"], want["blens"], rtol=1e-9, atol=1e-15)
    for key in ("g_mu", "g_sigma", "g_curv") + (("g_rates", "g_freqs") if model_name == "GTR" else ()):
        np.testing.assert_allclose(got[key], want[key], rtol=1e-7, atol=1e-9 * np.abs(want[key]).max(), err_msg=key)


def _chain_run(chm, masks, weights, x0, n_steps, prior, connector="nj"):
    S, D = x0.shape
    partials = _ref_partials(masks)
    ch = chm.Chain(partials, weights, D, leaf_x=x0.copy(), connector=connector, embedder="up", step_scale=1e-4,
                   prior=prior, mcmc_alg="MH", converge_length=None)
    ch.rng = np.random.default_rng(9)  # the reference leaves its generators unseeded (chain.py:93,329)
    np.random.seed(4)
    ch.set_probability(x0.copy())
    trace = [dict(ln_p=float(ch.ln_p), ln_prior=float(ch.ln_prior), peel=np.array(ch.peel), blens=np.array(ch.blens),
                  x=np.array(ch.leaf_x))]
    for _ in range(n_steps):
        ch.evolve()
        trace.append(dict(ln_p=float(ch.ln_p), ln_prior=float(ch.ln_prior), peel=np.array(ch.peel),
                          blens=np.array(ch.blens), x=np.array(ch.leaf_x)))
    return trace, ch.accepted


def _splits(peel, S):
    """Unrooted tree as its set of tip bipartitions (invariant to the order of the exactly-tied last joins)."""
    below = {i: frozenset([i]) for i in range(S)}
    for l, r, p in peel:
        below[int(p)] = below[int(l)] | below[int(r)]
    full = frozenset(range(S))
    return {min(s, full - s, key=lambda t: (len(t), sorted(t))) for s in below.values() if 1 < len(s) < S - 1}


def test_reference_Chain_runs_unchanged(reference):
    """chain.py:104-258,403-479 -- the reference's Chain on DS1 (configs[0]: up/nj, JC69, --prior
    gammadir): set_probability + Metropolis steps (sample_leaf_np -> connect -> Chyp_np.get_pdm ->
    Cpeeler.nj_np -> Cphylo.compute_LL_np -> Cphylo.compute_prior_gamma_dir_np -> accept/reject), the same
    seeded proposal stream both ways.  The trajectories must coincide: same acceptances, lnL and prior
    to 1e-9 at every step.  Distances agree to rounding only (host BLAS / libm vs device), and NJ's last
    joins are exact ties of Q that follow the last bit: the decoded tree is compared
    as an unrooted tree with its branch lengths, and the rate of bit-identical peels is printed."""
    from conftest import load_golden

    d = load_golden("ds1.npz")
    rng = np.random.default_rng(3)
    x0 = rng.normal(0.0, 0.05, (27, 3))
    want, acc_w = _chain_run(reference(False)["chain"], d["masks"], d["weights"], x0, 12, "gammadir")
    got, acc_g = _chain_run(reference(True)["chain"], d["masks"], d["weights"], x0, 12, "gammadir")
    assert acc_g == acc_w
    same_peel = 0
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a["x"], b["x"])
        assert a["ln_p"] == pytest.approx(b["ln_p"], rel=1e-9)
        assert a["ln_prior"] == pytest.approx(b["ln_prior"], rel=1e-9)
        assert _splits(a["peel"], 27) == _splits(b["peel"], 27)
        assert np.sum(a["blens"]) == pytest.approx(np.sum(b["blens"]), rel=1e-9)
        same_peel += int(np.array_equal(a["peel"], b["peel"]))
        if np.array_equal(a["peel"], b["peel"]):
            np.testing.assert_allclose(a["blens"], b["blens"], rtol=1e-9, atol=1e-15)
    print(f"\n[dropin] Chain on DS1: {same_peel}/{len(want)} peels bit-identical to the reference's (the rest differ only in the tied last joins)")


def test_reference_HMAP_runs_unchanged(reference):
    """hmap.py:185-192,310-358 -- the reference's HMAP.compute_loss (MAP objective: lnL + tree prior) and its
    gradient to the tip locations and the curvature leaf."""
    from dodonaphy_b200 import synth

    prob = synth.make_problem(10, 90, 1, D=3, seed=31)

    def once(hm):
        m = hm.HMAP(_ref_partials(prob["masks"]), torch.tensor(prob["weights"]), dim=3, soft_temp=1e-8, loss_fn="likelihood",
                    path_write=None, prior="gammadir", embedder="up", connector="nj")
        m.params_optim["leaf_loc"] = torch.tensor(prob["xs"][0], requires_grad=True)
        loss = m.compute_loss()
        loss.backward()
        gc = m.params_optim["curvature"].grad
        return dict(loss=loss.item(), ln_p=float(m.ln_p), ln_prior=float(m.ln_prior), peel=np.asarray(m.peel),
                    g=m.params_optim["leaf_loc"].grad.numpy().copy(), g_curv=float(gc))

    want = once(reference(False)["hmap"])
    got = once(reference(True)["hmap"])
    assert np.array_equal(got["peel"], want["peel"])
    assert got["loss"] == pytest.approx(want["loss"], rel=1e-9)
    assert got["ln_p"] == pytest.approx(want["ln_p"], rel=1e-9)
    assert got["ln_prior"] == pytest.approx(want["ln_prior"], rel=1e-9)
    assert got["g_curv"] == pytest.approx(want["g_curv"], rel=1e-7)
    np.testing.assert_allclose(got["g"], want["g"], rtol=1e-7, atol=1e-9 * np.abs(want["g"]).max())


def test_hard_nj_peel_match_rate_from_locations():
    """north_star: "NJ topologies and peel orders bit-exact".  On IDENTICAL distance matrices the hard-NJ
    kernel is bit-identical to Cpeeler.nj_np (tests/test_configs_gpu.py, 200 / 500 / 2000 taxa).  From
    LOCATIONS the distances themselves come from different libraries (the reference: OpenBLAS dsyrk + the
    host's arccosh; here: an FMA chain + CUDA's acosh) and agree to rounding, not bit for bit; NJ's Q has
    exact mathematical ties once <= 4 nodes remain, so the order of the LAST joins follows the last bit.
    This test measures that against the compiled reference (Chyp_np.get_pdm -> Cpeeler.nj_np) over the 64
    chains of configs[2] (200 taxa) and 16 DS1-sized embeddings: every tree must be the same unrooted tree
    with the same branch lengths (1e-9) and every peel must agree in all joins before the tied ones; the
    rate of fully bit-identical peels and distance matrices is printed."""
    from dodonaphy_b200 import engine, synth
    from oracle import refload

    r = refload.load()
    assert r.compiled, "oracle/_ref holds no compiled reference modules"
    for name, S, D, xs in (("configs[2] chains", 200, 5, synth.make_problem(200, 100, 64, D=5, seed=1003)["xs"]),
                           ("DS1-sized", 27, 3, np.random.default_rng(8).normal(0.0, 0.05, (16, 27, 3)))):
        pdm = engine.pdm_forward(torch.as_tensor(xs).cuda(), -1.0, "numpy")
        peel, blens = engine.nj_hard(pdm)
        pdm, peel, blens = pdm.cpu().numpy(), peel.cpu().numpy(), blens.cpu().numpy()
        same_peel = same_pdm = 0
        first_diff = []
        for b in range(xs.shape[0]):
            ref_pdm = np.ascontiguousarray(r.Chyp_np.get_pdm(xs[b]))
            ref_peel, ref_blens = r.Cpeeler.nj_np(ref_pdm)
            ref_peel = np.asarray(ref_peel)
            off = ~np.eye(S, dtype=bool)
            np.testing.assert_allclose(pdm[b][off], ref_pdm[off], rtol=1e-11, atol=1e-15)
            same_pdm += int(np.array_equal(pdm[b][off], ref_pdm[off]))
            assert _splits(peel[b], S) == _splits(ref_peel, S)
            assert blens[b].sum() == pytest.approx(ref_blens.sum(), rel=1e-9)
            rows = np.nonzero((peel[b] != ref_peel).any(axis=1))[0]
            if len(rows) == 0:
                same_peel += 1
                np.testing.assert_allclose(blens[b], ref_blens, rtol=1e-9, atol=1e-15)
            else:
                first_diff.append(int(rows[0]))
                assert rows[0] >= S - 4, "peels differ before the exactly-tied last joins"
        print(f"\n[match rate] {name}: {same_peel}/{xs.shape[0]} peels and {same_pdm}/{xs.shape[0]} distance matrices "
              f"bit-identical to the reference's; first differing join (of {S - 1}): {sorted(set(first_diff))}")
