"""GPU edge cases: tile boundaries, tiny and ragged inputs, every IUPAC code, degenerate weights,
limits and error paths -- each checked against the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.as_tensor(a).cuda()


def _random_tree(S, rng):
    from oracle import hyp, nj

    return nj.nj_hard(hyp.get_pdm_np(rng.normal(0, 0.1, (S, 3))))


def _oracle(masks, weights, peel, blens, rates, model, mix, prob_rates=None, prob_freqs=None):
    from oracle import phylo

    bt = torch.tensor(blens, requires_grad=True)
    t = bt[:, None] * torch.tensor(rates)[None, :]
    if model == "JC69":
        mats, fr = phylo.jc69_p_t_torch(t), torch.full([4], 0.25, dtype=torch.float64)
    else:
        mats, fr = phylo.gtr_p_t_torch(t, torch.tensor(prob_rates), torch.tensor(prob_freqs)), torch.tensor(prob_freqs)
    ll = phylo.treelikelihood_torch([torch.tensor(phylo.mask_to_partial(m)) for m in masks], torch.tensor(weights), peel, mats, fr, gamma_mix=mix)
    ll.backward()
    return ll.item(), bt.grad.numpy()


@pytest.mark.parametrize("L", [1, 2, 31, 32, 33, 255, 256, 257, 513])
def test_pattern_tile_boundaries(L):
    from dodonaphy_b200.engine import PruneEngine, SubstModel

    rng = np.random.default_rng(L)
    S = 9
    peel, blens = _random_tree(S, rng)
    masks = (1 << rng.integers(0, 4, size=(S, L))).astype(np.uint8)
    weights = rng.integers(1, 5, size=L)
    eng = PruneEngine(masks, weights)
    ll, g = eng.loglik(_dev(peel), _dev(blens), SubstModel.jc69(), grad=True)
    ref, gref = _oracle(masks, weights, peel, blens, np.ones(1), "JC69", False)
    assert ll.item() == pytest.approx(ref, rel=1e-9)
    np.testing.assert_allclose(g[0].cpu().numpy(), gref, rtol=1e-9, atol=1e-9 * np.abs(gref).max())


def test_every_iupac_code_value_and_gradient():
    """Masks 1..15 (two- and three-state codes take the column-summing path), GTR, K=3, both
    category semantics, small- and large-table kernel variants."""
    from dodonaphy_b200.engine import PruneEngine, SubstModel

    rng = np.random.default_rng(2)
    for S in (6, 110):
        L = 300
        peel, blens = _random_tree(S, rng)
        masks = rng.integers(1, 16, size=(S, L)).astype(np.uint8)
        weights = rng.integers(0, 4, size=L)  # zero weights included
        rates6, freqs = rng.dirichlet(np.ones(6)), rng.dirichlet(np.ones(4))
        cat = SubstModel.discrete_gamma(0.7, 3)
        eng = PruneEngine(masks, weights)
        m = SubstModel.gtr(rates6, freqs)
        for mix in (False, True):
            ll, g = eng.loglik(_dev(peel), _dev(blens), m, rates=cat, mix=mix, grad=True)
            ref, gref = _oracle(masks, weights, peel, blens, cat, "GTR", mix, rates6, freqs)
            assert ll.item() == pytest.approx(ref, rel=1e-9)
            np.testing.assert_allclose(g[0].cpu().numpy(), gref, rtol=1e-9, atol=1e-9 * np.abs(gref).max())


def test_two_and_three_taxa_and_eight_categories():
    from dodonaphy_b200.engine import PruneEngine, SubstModel

    rng = np.random.default_rng(3)
    for S in (2, 3):
        peel, blens = _random_tree(S, rng)
        masks = (1 << rng.integers(0, 4, size=(S, 40))).astype(np.uint8)
        weights = np.ones(40, dtype=np.int64)
        cat = SubstModel.discrete_gamma(0.3, 8)
        ll, g = PruneEngine(masks, weights).loglik(_dev(peel), _dev(blens), SubstModel.jc69(), rates=cat, mix=True, grad=True)
        ref, gref = _oracle(masks, weights, peel, blens, cat, "JC69", True)
        assert ll.item() == pytest.approx(ref, rel=1e-9)
        np.testing.assert_allclose(g[0].cpu().numpy(), gref, rtol=1e-9, atol=1e-9 * np.abs(gref).max())


def test_all_gap_columns_and_zero_branches():
    """A column of N/-/? only contributes log(1) = 0; zero-length branches give P = I."""
    from dodonaphy_b200.engine import PruneEngine, SubstModel

    rng = np.random.default_rng(4)
    S, L = 7, 64
    peel, blens = _random_tree(S, rng)
    masks = (1 << rng.integers(0, 4, size=(S, L))).astype(np.uint8)
    masks[:, :5] = 15
    weights = np.ones(L, dtype=np.int64)
    eng = PruneEngine(masks, weights)
    ll, _ = eng.loglik(_dev(peel), _dev(blens), SubstModel.jc69())
    ll_without, _ = PruneEngine(masks[:, 5:], weights[5:]).loglik(_dev(peel), _dev(blens), SubstModel.jc69())
    assert ll.item() == pytest.approx(ll_without.item(), rel=1e-12)
    z = blens.copy()
    z[peel[-1][1]] = 0.0  # the reference's fake-root edge of an unrooted tree (tree.py:112-121)
    ref, _ = _oracle(masks, weights, peel, z, np.ones(1), "JC69", False)
    got, _ = eng.loglik(_dev(peel), _dev(z), SubstModel.jc69())
    assert got.item() == pytest.approx(ref, rel=1e-9)


def test_limits_and_errors():
    from dodonaphy_b200 import engine
    from dodonaphy_b200.engine import PruneEngine, SubstModel

    eng = PruneEngine(np.ones((4, 10), dtype=np.uint8), np.ones(10))
    peel = _dev(np.array([[0, 1, 4], [2, 3, 5], [4, 5, 6]]))
    with pytest.raises(ValueError):
        eng.loglik(peel, _dev(np.full(6, 0.1)), SubstModel.jc69(), rates=np.ones(9))
    with pytest.raises(TypeError):
        PruneEngine(np.ones((4, 10), dtype=np.float64), np.ones(10))
    big = PruneEngine(np.ones((5000, 4), dtype=np.uint8), np.ones(4))
    with pytest.raises(NotImplementedError):
        big.loglik(_dev(np.zeros((4999, 3), dtype=np.int64)), _dev(np.full(9998, 0.1)), SubstModel.jc69(), grad=True)
    with pytest.raises(ValueError):
        engine.nj_soft(_dev(np.zeros((4, 4))), 0.0)
    with pytest.raises(ValueError):
        engine.nj_hard(_dev(np.zeros((1, 1))))
    with pytest.raises(ValueError):
        engine.pdm_forward(_dev(np.zeros((3, 16))), -1.0)


def test_degenerate_distances_stay_memory_safe():
    """NaN / all-equal distances must not crash the device (flags report the failure)."""
    from dodonaphy_b200 import engine
    from dodonaphy_b200.engine import validate_peel

    d = np.full((8, 8), np.nan)
    peel, blens = engine.nj_hard(_dev(d))
    assert peel.shape == (7, 3)
    p2, b2, flags = engine.nj_soft(_dev(d), 1e-8, return_flags=True)
    assert int((flags & 8).sum().item()) > 0
    same = np.ones((40, 40)) - np.eye(40)
    p3, b3, f3 = engine.nj_soft(_dev(same), 1e-8, return_flags=True)  # 780 exact ties -> general path
    validate_peel(p3.cpu().numpy(), 40)
    from oracle import nj

    p_ref, b_ref = nj.nj_soft(same, 1e-8)
    assert np.array_equal(p3.cpu().numpy(), p_ref)
    torch.cuda.synchronize()


def test_pdm_shapes():
    from dodonaphy_b200 import engine
    from oracle import hyp

    rng = np.random.default_rng(6)
    for S, D in ((1, 3), (2, 1), (5, 15), (33, 2)):
        x = rng.normal(0, 0.2, (S, D))
        np.testing.assert_allclose(engine.pdm_forward(_dev(x), -1.0).cpu().numpy(), hyp.get_pdm_torch(x), rtol=1e-10, atol=1e-13)


def _shaped_tree(S, shape, rng):
    """peel (S-1,3) + blens for a perfectly balanced tree (S/2 cherries, S a power of two) or a
    caterpillar (one cherry)."""
    peel, nxt = [], S
    if shape == "balanced":
        level = list(range(S))
        while len(level) > 1:
            up = []
            for i in range(0, len(level), 2):
                peel.append((level[i], level[i + 1], nxt))
                up.append(nxt)
                nxt += 1
            level = up
    else:
        cur = 0
        for t in range(1, S):
            peel.append((cur, t, nxt) if t % 2 else (t, cur, nxt))
            cur = nxt
            nxt += 1
    return np.asarray(peel, dtype=np.int64), rng.uniform(0.01, 0.3, 2 * S - 2)


@pytest.mark.parametrize("shape", ["balanced", "caterpillar"])
@pytest.mark.parametrize("model", ["JC69", "GTR"])
def test_cherry_tables_and_recomputed_cherries(shape, model):
    """64 taxa x 4 categories: a balanced tree has 32 cherries, more than the shared memory left
    over can tabulate, so tabulated, recomputed and stored cherries all occur; 2 % fully ambiguous
    cells plus a few two-state codes force the per-warp fallback of the table lookup.  Value,
    branch-length gradient, d lnL/d P and the rescaling variant against the oracle."""
    from dodonaphy_b200 import synth
    from dodonaphy_b200.engine import PruneEngine, SubstModel

    rng = np.random.default_rng(11)
    S, L, K = 64, 700, 4
    peel, blens = _shaped_tree(S, shape, rng)
    masks = (1 << rng.integers(0, 4, size=(S, L))).astype(np.uint8)
    masks[rng.random((S, L)) < 0.02] = 15
    masks[rng.random((S, L)) < 0.002] = 5
    weights = rng.integers(1, 5, size=L)
    rates = SubstModel.discrete_gamma(0.5, K)
    pr, pf = synth.gtr_params(rng)
    sm = SubstModel.jc69() if model == "JC69" else SubstModel.gtr(pr, pf)
    ref, gref = _oracle(masks, weights, peel, blens, rates, model, True, pr, pf)
    eng = PruneEngine(masks, weights)
    for scale in (False, True):
        ll, g = eng.loglik(_dev(peel), _dev(blens), sm, rates=rates, mix=True, grad=True, scale=scale)
        assert ll.item() == pytest.approx(ref, rel=1e-9)
        np.testing.assert_allclose(g[0].cpu().numpy(), gref, rtol=1e-9, atol=1e-9 * np.abs(gref).max())
    ll0, _ = eng.loglik(_dev(peel), _dev(blens), sm, rates=rates, mix=True, grad=False)
    assert ll0.item() == pytest.approx(ref, rel=1e-9)
    # caller-supplied matrices: d lnL / d P through the same cherry paths
    from oracle import phylo

    t = torch.tensor(blens)[:, None] * torch.tensor(rates)[None, :]
    fr = torch.full([4], 0.25, dtype=torch.float64) if model == "JC69" else torch.tensor(pf)
    mats = (phylo.jc69_p_t_torch(t) if model == "JC69" else phylo.gtr_p_t_torch(t, torch.tensor(pr), torch.tensor(pf))).detach()
    mats = mats.reshape(2 * S - 2, K, 4, 4).clone().requires_grad_(True)
    llm = phylo.treelikelihood_torch([torch.tensor(phylo.mask_to_partial(m)) for m in masks], torch.tensor(weights), peel, mats, fr, gamma_mix=True)
    llm.backward()
    out = eng.loglik_mats(_dev(peel), mats.detach()[None].cuda(), fr.numpy(), grad=True, mix=True)
    assert out["lnL"].item() == pytest.approx(ref, rel=1e-9)
    gm = mats.grad.numpy()
    np.testing.assert_allclose(out["grad_mats"][0].cpu().numpy(), gm, rtol=1e-9, atol=1e-9 * np.abs(gm).max())


def test_batched_path_applies_the_soft_nj_flag_policy():
    """A tripped exactness guard (flag bit 2) or an invalid selection (bit 3) must not be consumed
    silently by the batched routes: the device step poisons those samples with NaN and counts them
    (dodo_nj_guard); ``HotPath.resolve`` / ``BaseModel.log_likelihood_batch`` redo guard-tripped samples
    with the exhaustive N^2 soft argmin -- the reference's result in every regime (soft.py:38-53 over the
    whole padded matrix, oracle.nj.nj_soft) -- and raise on an invalid selection, exactly as the one-tree
    route peeler.nj_torch does."""
    from dodonaphy_b200 import engine, synth
    from dodonaphy_b200.base_model import BaseModel
    from dodonaphy_b200.engine import SubstModel
    from dodonaphy_b200.pipeline import HotPath
    from oracle import hyp, nj
    from oracle.phylo import mask_to_partial

    S, D, B, tau = 8, 3, 6, 1e-3
    prob = synth.make_problem(S, 120, B, D=D, seed=12)
    xs = prob["xs"] * 0.1  # distances of a few 1e-2: inactive entries are no longer invisible at tau = 1e-3
    x = _dev(xs)
    pdm = engine.pdm_forward(x, -1.0, "torch")
    _, _, flags = engine.nj_soft(pdm, tau, return_flags=True)
    tripped = ((flags & 4) != 0).any(dim=1).cpu().numpy()
    assert tripped.any() and int((flags & 8).sum().item()) == 0, "the case is meant to trip the guard only"
    hp = HotPath(prob["masks"], prob["weights"], SubstModel.jc69(), tau=tau)
    out = hp.value_and_grad(x)
    n_guard, n_invalid = out["bad"].tolist()
    assert n_guard == int(tripped.sum()) and n_invalid == 0
    assert np.array_equal(torch.isnan(out["lnL"]).cpu().numpy(), tripped)
    assert np.array_equal(torch.isnan(out["grad_blens"]).all(dim=1).cpu().numpy(), tripped)
    ok = hp.resolve(x, out)
    assert torch.isfinite(ok["lnL"]).all() and torch.isfinite(ok["grad_x"]).all() and ok["bad"].tolist() == [0, 0]
    # the same through the autograd front end, against the oracle's soft NJ per sample
    tips = [torch.tensor(mask_to_partial(m)) for m in prob["masks"]]
    bm = BaseModel("vi", tips, torch.tensor(prob["weights"]), D, soft_temp=tau)
    locs = torch.tensor(xs, requires_grad=True)
    ll, peel, blens = bm.log_likelihood_batch(locs)
    ll.sum().backward()
    assert torch.isfinite(ll).all() and torch.isfinite(locs.grad).all()
    np.testing.assert_array_equal(peel.cpu().numpy(), ok["peel"].cpu().numpy())
    np.testing.assert_allclose(ll.detach().numpy(), ok["lnL"].cpu().numpy(), rtol=1e-12)
    for b in range(B):
        p_ref, b_ref = nj.nj_soft(hyp.get_pdm_torch(xs[b]), tau)
        assert np.array_equal(peel[b].cpu().numpy(), p_ref), b
        np.testing.assert_allclose(blens[b].detach().cpu().numpy(), b_ref, rtol=1e-9, atol=1e-15)
    # an invalid selection raises instead of returning the lnL of a garbage topology
    hot = HotPath(prob["masks"], prob["weights"], SubstModel.jc69(), tau=50.0)
    bad = hot.value_and_grad(x)
    if bad["bad"].tolist()[1] > 0:
        with pytest.raises(RuntimeError):
            hot.resolve(x, bad)
    bm_hot = BaseModel("vi", tips, torch.tensor(prob["weights"]), D, soft_temp=50.0)
    d = np.full((B, S, D), np.nan)
    with pytest.raises(RuntimeError):
        bm_hot.log_likelihood_batch(torch.tensor(d))


def test_graphs_own_their_workspaces():
    """A captured step keeps replaying raw workspace pointers: growing the owner's buffers afterwards (a
    bigger batch, the batched Jacobian's B*(2S-2) replays) must not free what the graph addresses."""
    from dodonaphy_b200 import synth
    from dodonaphy_b200.engine import SubstModel
    from dodonaphy_b200.pipeline import GraphedStep, HotPath

    S, D = 10, 3
    prob = synth.make_problem(S, 200, 24, D=D, seed=4)
    hp = HotPath(prob["masks"], prob["weights"], SubstModel.jc69())
    x_small, x_big = _dev(prob["xs"][:4]), _dev(prob["xs"])
    want = {k: v.clone() for k, v in hp.value_and_grad(x_small).items() if k in ("lnL", "grad_x")}
    g = GraphedStep(hp, x_small, grad=True)
    hp.value_and_grad(x_big)                 # larger pdm / nj workspaces for the owner
    hp.blens_jacobian(x_big)                 # ... and much larger ones
    torch.cuda.empty_cache()
    filler = [torch.full((1 << 20,), 7.0, device="cuda") for _ in range(8)]  # recycle whatever was freed
    out = g.run(x_small)
    torch.cuda.synchronize()
    assert torch.equal(out["lnL"], want["lnL"]) and torch.equal(out["grad_x"], want["grad_x"])
    assert all(float(f[0]) == 7.0 and float(f[-1]) == 7.0 for f in filler)
    assert g.hot.ws is not hp.ws
