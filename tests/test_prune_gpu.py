"""GPU parity: K3 (fused P(t) + pruning forward/backward) through the C ABI against
(1) fixtures produced by the real reference (tests/golden), (2) the CPU oracle on seeded inputs.
Tolerance: 1e-9 relative (BASELINE.json north_star) on lnL and on the gradient (relative to the
largest gradient component of the tree)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _engine(masks, weights):
    from dodonaphy_b200.engine import PruneEngine

    return PruneEngine(masks, weights)


def _dev(a, dtype=None):
    return torch.as_tensor(a, dtype=dtype).cuda()


def _assert_grad(ours, ref):
    scale = np.abs(ref).max()
    np.testing.assert_allclose(ours, ref, rtol=RTOL, atol=RTOL * scale)


@pytest.mark.parametrize("tree", ["mb", "nj"])
def test_ds1_known_answers(tree):
    """-6904.632 (test/test_phylo.py:239) and -7130.96813417 (test/test_bito.py:35)."""
    from dodonaphy_b200.engine import SubstModel

    g = load_golden("ds1.npz")
    eng = _engine(g["masks"], g["weights"])
    ll, grad = eng.loglik(_dev(g[f"{tree}_peel"]), _dev(g[f"{tree}_blens"]), SubstModel.jc69(), grad=True)
    expect = {"mb": -6904.632, "nj": -7130.96813417}[tree]
    assert ll.item() == pytest.approx(expect, abs=1e-3 if tree == "mb" else 1e-7)
    assert ll.item() == pytest.approx(float(g[f"{tree}_lnL_torch"]), rel=RTOL)
    _assert_grad(grad[0].cpu().numpy(), g[f"{tree}_g_blens"])
    # forward-only kernel gives the same value
    ll2, none = eng.loglik(_dev(g[f"{tree}_peel"]), _dev(g[f"{tree}_blens"]), SubstModel.jc69(), grad=False)
    assert none is None
    assert ll2.item() == pytest.approx(ll.item(), rel=1e-14)
    # GTR
    m = SubstModel.gtr(g["gtr_rates"], g["gtr_freqs"])
    ll, grad = eng.loglik(_dev(g[f"{tree}_peel"]), _dev(g[f"{tree}_blens"]), m, grad=True)
    assert ll.item() == pytest.approx(float(g[f"{tree}_gtr_lnL"]), rel=RTOL)
    _assert_grad(grad[0].cpu().numpy(), g[f"{tree}_gtr_g_blens"])


def test_reference_fixtures():
    from dodonaphy_b200.engine import SubstModel

    g = load_golden("prune.npz")
    for i in range(int(g["n"])):
        eng = _engine(g[f"{i}_masks"], g[f"{i}_weights"])
        peel, blens = _dev(g[f"{i}_peel"]), _dev(g[f"{i}_blens"])
        ll, grad = eng.loglik(peel, blens, SubstModel.jc69(), grad=True)
        assert ll.item() == pytest.approx(float(g[f"{i}_jc_lnL"]), rel=RTOL), i
        assert ll.item() == pytest.approx(float(g[f"{i}_lnL_cython"]), rel=RTOL), i
        _assert_grad(grad[0].cpu().numpy(), g[f"{i}_jc_g_blens"])
        m = SubstModel.gtr(g[f"{i}_rates"], g[f"{i}_freqs"])
        ll, grad = eng.loglik(peel, blens, m, grad=True)
        assert ll.item() == pytest.approx(float(g[f"{i}_gtr_lnL"]), rel=RTOL), i
        _assert_grad(grad[0].cpu().numpy(), g[f"{i}_gtr_g_blens"])
        # caller-supplied matrices (the "mats" argument of calculate_treelikelihood)
        mats = _dev(g[f"{i}_gtr_mats"]).reshape(1, -1, 1, 4, 4)
        ll3, _ = eng.loglik(peel, None, m, mats=mats)
        assert ll3.item() == pytest.approx(float(g[f"{i}_gtr_lnL"]), rel=RTOL), i


@pytest.mark.parametrize("S,L,B,K", [(2, 5, 1, 1), (3, 33, 2, 1), (8, 100, 5, 1), (16, 257, 4, 4), (64, 1000, 6, 4), (40, 70, 3, 2)])
def test_batched_vs_oracle(S, L, B, K):
    """Seeded synthetic batches (different topology per sample), JC69 and GTR, K categories in both
    semantics (reference broadcasting = independent categories summed; MIX = discrete gamma)."""
    from dodonaphy_b200 import synth
    from dodonaphy_b200.engine import SubstModel
    from oracle import hyp, nj, phylo

    prob = synth.make_problem(S, L, B, D=3, seed=100 + S)
    tips = [phylo.mask_to_partial(m) for m in prob["masks"]]
    peels, blens = [], []
    for b in range(B):
        p, bl = nj.nj_hard(hyp.get_pdm_np(prob["xs"][b]))
        peels.append(p)
        blens.append(bl)
    peels, blens = np.stack(peels), np.stack(blens)
    eng = _engine(prob["masks"], prob["weights"])
    rates = phylo.discrete_gamma_rates(0.5, K)
    for model_name in ("JC69", "GTR"):
        if model_name == "JC69":
            m = SubstModel.jc69()
            pfun = lambda t: phylo.jc69_p_t_torch(t)
            freqs = np.full(4, 0.25)
        else:
            m = SubstModel.gtr(prob["rates"], prob["freqs"])
            pfun = lambda t: phylo.gtr_p_t_torch(t, torch.tensor(prob["rates"]), torch.tensor(prob["freqs"]))
            freqs = prob["freqs"]
        for mix in ((False, True) if K > 1 else (False,)):
            ll, grad = eng.loglik(_dev(peels), _dev(blens), m, rates=rates, mix=mix, grad=True)
            ll_f, _ = eng.loglik(_dev(peels), _dev(blens), m, rates=rates, mix=mix, grad=False)
            for b in range(B):
                bt = torch.tensor(blens[b], requires_grad=True)
                mats = pfun(bt[:, None] * torch.tensor(rates)[None, :])
                ref = phylo.treelikelihood_torch([torch.tensor(t) for t in tips], torch.tensor(prob["weights"]),
                                                 peels[b], mats, torch.tensor(freqs), gamma_mix=mix)
                ref.backward()
                assert ll[b].item() == pytest.approx(ref.item(), rel=RTOL), (model_name, mix, b)
                assert ll_f[b].item() == pytest.approx(ref.item(), rel=RTOL)
                _assert_grad(grad[b].cpu().numpy(), bt.grad.numpy())


def test_shared_topology_and_determinism():
    from dodonaphy_b200 import synth
    from dodonaphy_b200.engine import SubstModel
    from oracle import hyp, nj

    prob = synth.make_problem(24, 300, 4, D=3, seed=7)
    peel, bl = nj.nj_hard(hyp.get_pdm_np(prob["x"]))
    eng = _engine(prob["masks"], prob["weights"])
    blens = np.stack([bl * (1 + 0.1 * b) for b in range(4)])
    m = SubstModel.gtr(prob["rates"], prob["freqs"])
    ll1, g1 = eng.loglik(_dev(peel), _dev(blens), m, grad=True)
    ll2, g2 = eng.loglik(_dev(np.stack([peel] * 4)), _dev(blens), m, grad=True)
    assert torch.equal(ll1, ll2) and torch.equal(g1, g2)
    ll3, g3 = eng.loglik(_dev(peel), _dev(blens), m, grad=True)
    assert torch.equal(ll1, ll3) and torch.equal(g1, g3)  # run-to-run bit reproducible


def test_argument_errors():
    from dodonaphy_b200.engine import PruneEngine, SubstModel

    eng = PruneEngine(np.ones((4, 10), dtype=np.uint8), np.ones(10))
    with pytest.raises(ValueError):
        eng.loglik(_dev(np.zeros((3, 3), dtype=np.int64)), _dev(np.ones(5)), SubstModel.jc69())
    with pytest.raises(ValueError):
        eng.loglik(_dev(np.zeros((2, 3), dtype=np.int64)), _dev(np.ones(6)), SubstModel.jc69())
    with pytest.raises(ValueError):
        PruneEngine(np.ones((4, 10), dtype=np.uint8), np.ones(9))


def _nj_batch(prob):
    from oracle import hyp, nj

    peels, blens = zip(*[nj.nj_hard(hyp.get_pdm_np(x)) for x in prob["xs"]])
    return np.stack(peels), np.stack(blens)


@pytest.mark.parametrize("S,L,K", [(12, 300, 1), (64, 700, 4), (130, 260, 2)])
def test_rescaling_is_bit_identical_when_nothing_underflows(S, L, K):
    """Power-of-two rescaling never touches a mantissa: with and without it the results are equal
    bit for bit (small-table and large-table kernel variants, all rate-category modes)."""
    from dodonaphy_b200 import synth
    from dodonaphy_b200.engine import SubstModel

    prob = synth.make_problem(S, L, 3, D=3, seed=300 + S)
    peels, blens = _nj_batch(prob)
    eng = _engine(prob["masks"], prob["weights"])
    m = SubstModel.gtr(prob["rates"], prob["freqs"])
    rates = SubstModel.discrete_gamma(0.5, K)
    for mix in (False, True):
        a, ga = eng.loglik(_dev(peels), _dev(blens), m, rates=rates, mix=mix, grad=True, scale=False)
        b, gb = eng.loglik(_dev(peels), _dev(blens), m, rates=rates, mix=mix, grad=True, scale=True)
        assert torch.equal(a, b) and torch.equal(ga, gb)
        c, _ = eng.loglik(_dev(peels), _dev(blens), m, rates=rates, mix=mix, grad=False, scale=True)
        assert torch.equal(a, c)


@pytest.mark.parametrize("K,mix", [(1, False), (4, True), (3, False)])
def test_rescaling_where_the_reference_underflows(K, mix):
    """Long branches on 400 taxa: the unscaled product (the reference's arithmetic) underflows to
    -inf; the rescaled kernel must agree with the log-space oracle (lnL and gradient)."""
    from dodonaphy_b200 import synth
    from dodonaphy_b200.engine import SubstModel
    from oracle import phylo

    S, L = 400, 40
    prob = synth.make_problem(S, L, 1, D=4, seed=77)
    rng = np.random.default_rng(5)
    masks = (1 << rng.integers(0, 4, size=(S, L))).astype(np.uint8)  # i.i.d. columns: tiny site likelihoods
    peels, blens = _nj_batch(prob)
    blens = blens * 40.0 + 0.5
    eng = _engine(masks, prob["weights"])
    m = SubstModel.gtr(prob["rates"], prob["freqs"])
    rates = SubstModel.discrete_gamma(0.5, K)
    raw, _ = eng.loglik(_dev(peels), _dev(blens), m, rates=rates, mix=mix, grad=False, scale=False)
    assert not torch.isfinite(raw).all(), "test input no longer underflows without rescaling"
    ll, g = eng.loglik(_dev(peels), _dev(blens), m, rates=rates, mix=mix, grad=True, scale=True)
    bt = torch.tensor(blens[0], requires_grad=True)
    mats = phylo.gtr_p_t_torch(bt[:, None] * torch.tensor(rates)[None, :], torch.tensor(prob["rates"]), torch.tensor(prob["freqs"]))
    ref = phylo.treelikelihood_rescaled_torch([torch.tensor(phylo.mask_to_partial(x)) for x in masks], torch.tensor(prob["weights"]),
                                              peels[0], mats, torch.tensor(prob["freqs"]), gamma_mix=mix)
    ref.backward()
    assert torch.isfinite(ll).all()
    assert ll.item() == pytest.approx(ref.item(), rel=RTOL)
    _assert_grad(g[0].cpu().numpy(), bt.grad.numpy())


@pytest.mark.parametrize("seed", range(6))
def test_tabulated_recomputed_and_stored_cherries_agree_bit_for_bit(seed):
    """Random tree sizes and category counts: the unscaled kernel tabulates cherries while shared
    memory lasts, the rescaled one recomputes them -- lnL and gradient must not differ in a single bit
    (same operations in the same order), and both must match the oracle to 1e-9."""
    from dodonaphy_b200 import synth
    from dodonaphy_b200.engine import SubstModel
    from oracle import phylo

    rng = np.random.default_rng(4000 + seed)
    S = int(rng.integers(4, 90))
    K = int(rng.choice([1, 2, 4, 8]))
    L = int(rng.integers(40, 400))
    prob = synth.make_problem(S, L, 2, D=3, seed=500 + seed)
    prob["masks"][rng.random(prob["masks"].shape) < 0.01] = 6  # a few two-state codes
    peels, blens = _nj_batch(prob)
    eng = _engine(prob["masks"], prob["weights"])
    model = SubstModel.jc69() if seed % 2 else SubstModel.gtr(prob["rates"], prob["freqs"])
    rates = SubstModel.discrete_gamma(0.7, K)
    a, ga = eng.loglik(_dev(peels), _dev(blens), model, rates=rates, mix=True, grad=True, scale=False)
    b, gb = eng.loglik(_dev(peels), _dev(blens), model, rates=rates, mix=True, grad=True, scale=True)
    assert torch.equal(a, b) and torch.equal(ga, gb)
    tips = [torch.tensor(phylo.mask_to_partial(m)) for m in prob["masks"]]
    for i in range(2):
        bt = torch.tensor(blens[i], requires_grad=True)
        t = bt[:, None] * torch.tensor(rates)[None, :]
        if seed % 2:
            mats, fr = phylo.jc69_p_t_torch(t), torch.full([4], 0.25, dtype=torch.float64)
        else:
            mats = phylo.gtr_p_t_torch(t, torch.tensor(prob["rates"]), torch.tensor(prob["freqs"]))
            fr = torch.tensor(prob["freqs"])
        ll = phylo.treelikelihood_torch(tips, torch.tensor(prob["weights"]), peels[i], mats, fr, gamma_mix=True)
        ll.backward()
        assert a[i].item() == pytest.approx(ll.item(), rel=1e-9)
        g = bt.grad.numpy()
        np.testing.assert_allclose(ga[i].cpu().numpy(), g, rtol=1e-9, atol=1e-9 * np.abs(g).max())


@pytest.mark.parametrize("S,L,K", [(5, 40, 1), (27, 300, 1), (64, 700, 4), (130, 260, 2), (230, 100, 3)])
def test_model_gradient_in_one_launch(S, L, K):
    """The learnable-GTR launch (dodo_prune_model: per-thread accumulation of the 4 x 4 matrix W of the
    Daleckii-Krein chain rule) against (1) the branch-length kernel (lnL bit for bit, d lnL / d blens 1e-11),
    (2) the explicit route d lnL / d P per edge -> dodo_model_chain (W, d lnL / d pi: 1e-10 of the largest
    entry), (3) torch autograd through the oracle's GTR_p_t restatement down to the rate / frequency leaves.
    Small-table, large-table and rescaled variants, both rate-category semantics."""
    from dodonaphy_b200 import synth
    from dodonaphy_b200.engine import SubstModel
    from oracle import phylo

    prob = synth.make_problem(S, L, 3, D=3, seed=500 + S)
    peels, blens = _nj_batch(prob)
    eng = _engine(prob["masks"], prob["weights"])
    m = SubstModel.gtr(prob["rates"], prob["freqs"])
    rates = SubstModel.discrete_gamma(0.5, K)
    for mix in ((False, True) if K > 1 else (False,)):
        for scale in (False, True):
            new = eng.loglik_model_grad(_dev(peels), _dev(blens), m, rates=rates, mix=mix, scale=scale)
            ll, g = eng.loglik(_dev(peels), _dev(blens), m, rates=rates, mix=mix, grad=True, scale=scale)
            assert torch.equal(new["lnL"], ll)
            _assert_grad_tol(new["grad_blens"].cpu().numpy(), g.cpu().numpy(), 1e-11)
            if not scale:
                old = eng.loglik_model_grad(_dev(peels), _dev(blens), m, rates=rates, mix=mix, via_mats=True)
                for key in ("grad_pi", "grad_blens"):
                    _assert_grad_tol(new[key].cpu().numpy(), old[key].cpu().numpy(), 1e-10)
                # W: all but the column of the eigenvalue 0 (eigh: the last), which only meets the row sums of dQ
                _assert_grad_tol(new["W"][..., :3].cpu().numpy(), old["W"][..., :3].cpu().numpy(), 1e-10)
                assert float(new["W"][..., 3].abs().max()) == 0.0
                assert abs(m.lam[3]) < 1e-12
    # an eigen-system handed over in another order (zero eigenvalue first): W comes back in that order
    perm = [3, 0, 1, 2]
    m2 = SubstModel(m.kind, m.U[:, perm], m.lam[perm], m.Uinv[perm, :], m.Q, m.pi)
    w1 = eng.loglik_model_grad(_dev(peels), _dev(blens), m, rates=rates, mix=(K > 1), scale=False)["W"].cpu().numpy()
    w2 = eng.loglik_model_grad(_dev(peels), _dev(blens), m2, rates=rates, mix=(K > 1), scale=False)["W"].cpu().numpy()
    _assert_grad_tol(w2, w1[:, perm][:, :, perm], 1e-12)
    # autograd through the reference's arithmetic (phylomodel.py:96-114 restated), one sample, down to Q
    rr = torch.tensor(prob["rates"], requires_grad=True)
    ff = torch.tensor(prob["freqs"], requires_grad=True)
    bt = torch.tensor(blens[0])
    mats = phylo.gtr_p_t_torch(bt[:, None] * torch.tensor(rates)[None, :], rr, ff)
    ref = phylo.treelikelihood_torch([torch.tensor(phylo.mask_to_partial(x)) for x in prob["masks"]], torch.tensor(prob["weights"]),
                                     peels[0], mats, ff.detach(), gamma_mix=(K > 1))
    ref.backward()
    new = eng.loglik_model_grad(_dev(peels), _dev(blens), m, rates=rates, mix=(K > 1), scale=False)
    U, Ui = torch.tensor(m.U), torch.tensor(m.Uinv)
    dQ = Ui.T @ new["W"][0].cpu() @ U.T
    rr2 = torch.tensor(prob["rates"], requires_grad=True)
    ff2 = torch.tensor(prob["freqs"], requires_grad=True)
    iu = torch.triu_indices(4, 4, 1)
    Q = torch.zeros((4, 4), dtype=torch.float64)
    Q[iu[0], iu[1]] = rr2
    Q[iu[1], iu[0]] = rr2
    Q = Q * ff2.unsqueeze(0)
    Q = Q + torch.diag(-torch.sum(Q, dim=1))
    Q = Q / (-torch.sum(torch.diagonal(Q) * ff2))  # phylomodel.py:116-128
    (Q * dQ).sum().backward()
    _assert_grad_tol(rr2.grad.numpy(), rr.grad.numpy(), 1e-8)
    _assert_grad_tol(ff2.grad.numpy(), ff.grad.numpy(), 1e-8)


def _assert_grad_tol(ours, ref, tol):
    scale = np.abs(ref).max()
    np.testing.assert_allclose(ours, ref, rtol=0, atol=tol * scale)


def test_model_gradient_where_the_reference_underflows():
    """Learnable GTR on 400 taxa with long branches: the unscaled product underflows (the reference returns -inf and
    the explicit d lnL / d P route has no rescaling); the rescaled model-gradient launch must agree with the
    log-space oracle under autograd: lnL, d lnL / d blens and the gradient to the rate / frequency parameters."""
    from dodonaphy_b200 import synth
    from dodonaphy_b200.engine import SubstModel
    from oracle import phylo

    S, L, K = 400, 40, 2
    prob = synth.make_problem(S, L, 1, D=4, seed=77)
    rng = np.random.default_rng(5)
    masks = (1 << rng.integers(0, 4, size=(S, L))).astype(np.uint8)
    peels, blens = _nj_batch(prob)
    blens = blens * 40.0 + 0.5
    eng = _engine(masks, prob["weights"])
    m = SubstModel.gtr(prob["rates"], prob["freqs"])
    rates = SubstModel.discrete_gamma(0.5, K)
    raw, _ = eng.loglik(_dev(peels), _dev(blens), m, rates=rates, mix=True, grad=False, scale=False)
    assert not torch.isfinite(raw).all(), "test input no longer underflows without rescaling"
    out = eng.loglik_model_grad(_dev(peels), _dev(blens), m, rates=rates, mix=True, scale=True)
    rr = torch.tensor(prob["rates"], requires_grad=True)
    ff = torch.tensor(prob["freqs"], requires_grad=True)
    bt = torch.tensor(blens[0], requires_grad=True)
    mats = phylo.gtr_p_t_torch(bt[:, None] * torch.tensor(rates)[None, :], rr, ff)
    ref = phylo.treelikelihood_rescaled_torch([torch.tensor(phylo.mask_to_partial(x)) for x in masks], torch.tensor(prob["weights"]),
                                              peels[0], mats, ff.detach(), gamma_mix=True)
    ref.backward()
    assert torch.isfinite(out["lnL"]).all()
    assert out["lnL"].item() == pytest.approx(ref.item(), rel=RTOL)
    _assert_grad(out["grad_blens"][0].cpu().numpy(), bt.grad.numpy())
    dQ = torch.tensor(m.Uinv).T @ out["W"][0].cpu() @ torch.tensor(m.U).T
    rr2 = torch.tensor(prob["rates"], requires_grad=True)
    ff2 = torch.tensor(prob["freqs"], requires_grad=True)
    iu = torch.triu_indices(4, 4, 1)
    Q = torch.zeros((4, 4), dtype=torch.float64)
    Q[iu[0], iu[1]] = rr2
    Q[iu[1], iu[0]] = rr2
    Q = Q * ff2.unsqueeze(0)
    Q = Q + torch.diag(-torch.sum(Q, dim=1))
    Q = Q / (-torch.sum(torch.diagonal(Q) * ff2))
    (Q * dQ).sum().backward()
    _assert_grad_tol(rr2.grad.numpy(), rr.grad.numpy(), 1e-7)
    _assert_grad_tol(ff2.grad.numpy(), ff.grad.numpy(), 1e-7)
