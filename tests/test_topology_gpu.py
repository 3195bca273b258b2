"""GPU parity: K1 (pairwise hyperbolic distance) and K2 (hard / soft neighbour joining) through
the C ABI, against fixtures produced by the real reference (tests/golden) and the CPU oracle.

Bars (BASELINE.json north_star): NJ peels bit-exact; hard-NJ branch lengths bit-exact too (same
operation order); floating point elsewhere within 1e-9 relative."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.as_tensor(a).cuda()


# ------------------------------------------------------------------ K1
def test_pdm_reference_fixtures():
    from dodonaphy_b200 import engine

    g = load_golden("pdm.npz")
    for i in range(int(g["n"])):
        x, curv, mats, C = g[f"{i}_x"], float(g[f"{i}_curv"]), bool(g[f"{i}_matsumoto"]), g[f"{i}_C"]
        for lift in ("up", "wrap"):
            ours = engine.pdm_forward(_dev(x), curv, "torch", lift, mats).cpu().numpy()
            np.testing.assert_allclose(ours, g[f"{i}_pdm_torch_{lift}"], rtol=1e-9, atol=1e-12)
            assert np.array_equal(ours, ours.T)
            gx, gc = engine.pdm_backward(_dev(x), _dev(C), curv, lift, mats)
            ref = g[f"{i}_gx_{lift}"]
            np.testing.assert_allclose(gx.cpu().numpy(), ref, rtol=1e-7, atol=1e-9 * np.abs(ref).max())
            assert gc.item() == pytest.approx(float(g[f"{i}_gc_{lift}"][0]), rel=1e-9)
        ours = engine.pdm_forward(_dev(x), curv, "numpy", "up", mats).cpu().numpy()
        ref = g[f"{i}_pdm_np"]
        off = ~np.eye(len(x), dtype=bool)
        near = ref < 1e-6  # coincident points: acosh(1 + O(eps)) is rounding noise of size sqrt(eps)
        np.testing.assert_allclose(ours[off & ~near], ref[off & ~near], rtol=1e-9, atol=1e-13)
        # the NumPy twin leaves acosh(1+eps) ~ 2e-8 rounding noise on the diagonal (never read by NJ)
        np.testing.assert_allclose(ours[~off | near], ref[~off | near], rtol=0, atol=1e-7)


def test_pdm_batched_and_euclidean():
    from dodonaphy_b200 import engine
    from oracle import hyp

    rng = np.random.default_rng(11)
    x = rng.normal(0, 0.2, (5, 33, 4))
    ours = engine.pdm_forward(_dev(x), -0.7).cpu().numpy()
    for b in range(5):
        np.testing.assert_allclose(ours[b], hyp.get_pdm_torch(x[b], -0.7), rtol=1e-10, atol=1e-13)
    eu = engine.pdm_forward(_dev(x[0]), 0.0, "numpy").cpu().numpy()
    np.testing.assert_allclose(eu, hyp.get_pdm_np(x[0], 0.0), rtol=1e-12, atol=1e-15)
    with pytest.raises(NotImplementedError):
        engine.pdm_forward(_dev(x[0]), -1.0, "numpy", "wrap")
    with pytest.raises(ValueError):
        engine.pdm_forward(_dev(np.zeros((3, 40))), -1.0)


# ------------------------------------------------------------------ K2 hard
def test_nj_hard_reference_fixtures_bit_exact():
    from dodonaphy_b200 import engine

    g = load_golden("nj_hard.npz")
    for i in range(int(g["n"])):
        peel, blens = engine.nj_hard(_dev(g[f"{i}_pdm"]))
        assert np.array_equal(peel.cpu().numpy(), g[f"{i}_peel"]), i
        assert np.array_equal(blens.cpu().numpy(), g[f"{i}_blens"]), i


@pytest.mark.parametrize("S,B", [(7, 9), (64, 16), (150, 3), (300, 2)])
def test_nj_hard_batched_vs_oracle_bit_exact(S, B):
    from dodonaphy_b200 import engine
    from oracle import hyp, nj

    rng = np.random.default_rng(S)
    xs = rng.normal(0, 0.05, (B, S, 5))
    pd = np.stack([hyp.get_pdm_np(x) for x in xs])
    peel, blens = engine.nj_hard(_dev(pd))
    for b in range(B if S <= 150 else 1):
        p_ref, b_ref = nj.nj_hard(pd[b])
        assert np.array_equal(peel[b].cpu().numpy(), p_ref)
        assert np.array_equal(blens[b].cpu().numpy(), b_ref)
    # every peel is a valid tree
    from dodonaphy_b200.engine import validate_peel

    validate_peel(peel.cpu().numpy(), S)


# ------------------------------------------------------------------ K2 soft
@pytest.mark.parametrize("exhaustive", [False, True])
def test_nj_soft_reference_fixtures(exhaustive):
    from dodonaphy_b200 import engine

    g = load_golden("nj_soft.npz")
    for i in range(int(g["n"])):
        tau = float(g[f"{i}_tau"])
        if tau < 1e-12 and not exhaustive:
            continue  # tau = 1e-20: rounding-level ties, only the exhaustive enumeration is claimed
        peel, blens, flags = engine.nj_soft(_dev(g[f"{i}_pdm"]), tau, exhaustive=exhaustive, return_flags=True)
        if tau >= 1e-12:
            assert int((flags & 12).sum().item()) == 0, (i, flags)
        assert np.array_equal(peel.cpu().numpy(), g[f"{i}_peel"]), (i, tau)
        np.testing.assert_allclose(blens.cpu().numpy(), g[f"{i}_blens"], rtol=1e-9, atol=1e-15)
        gp = engine.nj_soft_backward(peel, flags, _dev(g[f"{i}_c"])).cpu().numpy()
        ref = g[f"{i}_g_pdm"]
        np.testing.assert_allclose(2 * gp, ref + ref.T, rtol=1e-9, atol=1e-12)


def test_nj_soft_reference_fixture_at_the_headline_size():
    """The reference's own soft NJ on one 64-taxon tree of the C2 workload (tests/golden/nj_soft_c2.npz):
    peel bit-exact, branch lengths and the gradient w.r.t. the distances to 1e-9; and the whole device
    chain from the locations (distance kernel -> soft NJ) gives the same peel."""
    from dodonaphy_b200 import engine

    g = load_golden("nj_soft_c2.npz")
    tau = float(g["tau"])
    peel, blens, flags = engine.nj_soft(_dev(g["pdm"]), tau, return_flags=True)
    assert int((flags & 12).sum().item()) == 0
    assert np.array_equal(peel.cpu().numpy(), g["peel"])
    np.testing.assert_allclose(blens.cpu().numpy(), g["blens"], rtol=1e-9, atol=1e-15)
    gp = engine.nj_soft_backward(peel, flags, _dev(g["c"])).cpu().numpy()
    ref = g["g_pdm"]
    np.testing.assert_allclose(2 * gp, ref + ref.T, rtol=1e-9, atol=1e-12)
    pdm_dev = engine.pdm_forward(_dev(g["x"]), -1.0, "torch")
    np.testing.assert_allclose(pdm_dev.cpu().numpy(), g["pdm"], rtol=1e-10, atol=1e-14)
    peel2, blens2 = engine.nj_soft(pdm_dev, tau)
    assert np.array_equal(peel2.cpu().numpy().reshape(g["peel"].shape), g["peel"])
    np.testing.assert_allclose(blens2.cpu().numpy().reshape(-1), g["blens"], rtol=1e-9, atol=1e-15)


@pytest.mark.parametrize("S,B", [(27, 4), (64, 8), (130, 2)])
def test_nj_soft_batched_vs_oracle(S, B):
    from dodonaphy_b200 import engine
    from oracle import hyp, nj

    rng = np.random.default_rng(1000 + S)
    xs = rng.normal(0, 0.05, (B, S, 5))
    pd = np.stack([hyp.get_pdm_torch(x) for x in xs])
    peel, blens, flags = engine.nj_soft(_dev(pd), 1e-8, return_flags=True)
    assert int((flags & 12).sum().item()) == 0
    c = rng.normal(size=(B, 2 * S - 2))
    gp = engine.nj_soft_backward(peel, flags, _dev(c)).cpu().numpy()
    for b in range(B):
        p_ref, b_ref = nj.nj_soft(pd[b], 1e-8)
        assert np.array_equal(peel[b].cpu().numpy(), p_ref)
        np.testing.assert_allclose(blens[b].cpu().numpy(), b_ref, rtol=1e-9, atol=1e-15)
        if b == 0 and S <= 64:
            t = torch.tensor(pd[b], requires_grad=True)
            _, bt = nj.nj_soft_torch(t, 1e-8)
            (bt * torch.tensor(c[b])).sum().backward()
            ref = t.grad.numpy()
            np.testing.assert_allclose(2 * gp[b], ref + ref.T, rtol=1e-9, atol=1e-12)


def test_nj_soft_all_ties():
    """test/test_peeler.py:343-348: six equidistant taxa must still give a valid peel."""
    from dodonaphy_b200 import engine
    from dodonaphy_b200.engine import validate_peel

    pd = np.ones((6, 6)) - np.eye(6)
    peel, blens = engine.nj_soft(_dev(pd), 1e-8)
    validate_peel(peel.cpu().numpy(), 6)
    assert peel.cpu().numpy().tolist() == [[4, 5, 6], [3, 6, 7], [2, 7, 8], [1, 8, 9], [0, 9, 10]]


@pytest.mark.gpu
def test_lean_soft_nj_agrees_with_the_general_kernel(monkeypatch):
    """Trees of up to 160 taxa go through nj_soft_lean_kernel (matrix resident in shared memory, two barriers per
    join); everything it does not handle is handed to the general kernel.  Both must return the same peel
    and flags, and branch lengths equal to rounding (the lean kernel derives the new node's row sum from the old
    ones instead of summing the new row: both are the exact sum to ~1e-32 before the final rounding)."""
    import torch

    from dodonaphy_b200 import engine

    rng = np.random.default_rng(11)
    cases = []
    for S, B in ((3, 4), (4, 4), (5, 3), (17, 5), (33, 4), (64, 6), (65, 3), (130, 2), (160, 2)):
        cases.append(torch.tensor(rng.normal(0, 0.08, (B, S, 4))).cuda())
    dup = rng.normal(0, 0.08, (2, 12, 3))
    dup[:, 5] = dup[:, 2]  # two identical taxa: exact ties in Q at every join
    dup[:, 9] = dup[:, 2]
    cases.append(torch.tensor(dup).cuda())
    for x in cases:
        pdm = engine.pdm_forward(x, -1.0, "torch")
        for tau in (1e-8, 1e-4):
            monkeypatch.delenv("DODO_NJ_NO_LEAN", raising=False)
            p1, b1, f1 = engine.nj_soft(pdm, tau, return_flags=True)
            monkeypatch.setenv("DODO_NJ_NO_LEAN", "1")
            p0, b0, f0 = engine.nj_soft(pdm, tau, return_flags=True)
            assert torch.equal(p0, p1) and torch.equal(f0, f1)
            np.testing.assert_allclose(b1.cpu().numpy(), b0.cpu().numpy(), rtol=1e-12, atol=1e-15)
    # handed over as a whole: an asymmetric matrix and a temperature outside the one-candidate regime
    pdm = engine.pdm_forward(cases[3], -1.0, "torch").clone()
    pdm[:, 1, 2] += 1e-3
    for tau in (1e-8, 0.01):
        monkeypatch.delenv("DODO_NJ_NO_LEAN", raising=False)
        p1, b1, f1 = engine.nj_soft(pdm, tau, return_flags=True)
        monkeypatch.setenv("DODO_NJ_NO_LEAN", "1")
        p0, b0, f0 = engine.nj_soft(pdm, tau, return_flags=True)
        assert torch.equal(p0, p1) and torch.equal(f0, f1) and torch.equal(b0, b1)


@pytest.mark.gpu
def test_lean_hard_nj_is_bit_identical_to_the_general_kernel(monkeypatch):
    """Trees of up to 224 taxa go through nj_hard_lean_kernel (lower triangle resident in shared memory); peel AND
    branch lengths must equal the general kernel's -- and hence Cpeeler.nj_np's -- bit for bit, ties included."""
    from dodonaphy_b200 import engine, synth

    rng = np.random.default_rng(12)
    cases = [torch.tensor(rng.normal(0, 0.08, (B, S, 4))).cuda() for S, B in ((3, 4), (4, 4), (5, 3), (17, 5), (48, 3), (49, 3), (129, 2), (200, 3), (224, 2))]
    dup = rng.normal(0, 0.08, (2, 12, 3))
    dup[:, 5] = dup[:, 2]
    dup[:, 9] = dup[:, 2]
    cases.append(torch.tensor(dup).cuda())
    mats = [engine.pdm_forward(x, -1.0, "numpy") for x in cases]
    mats.append(torch.tensor(np.stack([synth.det_distances(60, 3, seed) for seed in (1, 2)])).cuda())  # bit-reproducible distances
    asym = mats[3].clone()
    asym[:, 1, 2] += 1e-3  # handed over to the general kernel as a whole
    mats.append(asym)
    for pdm in mats:
        monkeypatch.delenv("DODO_NJ_NO_LEAN", raising=False)
        p1, b1 = engine.nj_hard(pdm)
        monkeypatch.setenv("DODO_NJ_NO_LEAN", "1")
        p0, b0 = engine.nj_hard(pdm)
        assert torch.equal(p0, p1) and torch.equal(b0, b1)


@pytest.mark.gpu
@pytest.mark.parametrize("S,B", [(101, 3), (200, 8), (200, 20), (224, 40)])
def test_hard_nj_on_thread_block_clusters_is_bit_identical(monkeypatch, S, B):
    """Above 100 taxa and with few trees per GPU one tree is joined by a thread-block CLUSTER (2, 4 or 8 CTAs: the
    scan split over the cluster, results exchanged through distributed shared memory, one cluster barrier per join).
    Every cluster size -- forced, and the library's own choice -- must return the one-CTA kernel's peel and branch
    lengths bit for bit, exact ties (duplicated taxa) included."""
    from dodonaphy_b200 import engine

    x = np.random.default_rng(100 + S).normal(0, 0.06, (B, S, 3))
    x[:, 7] = x[:, 3]  # duplicated taxon: q ties broken by node id
    pdm = engine.pdm_forward(torch.tensor(x).cuda(), -1.0, "numpy")
    monkeypatch.setenv("DODO_NJ_CLUSTER", "1")
    p0, b0 = engine.nj_hard(pdm)
    for cl in ("2", "4", "8", None):
        if cl is None:
            monkeypatch.delenv("DODO_NJ_CLUSTER")
        else:
            monkeypatch.setenv("DODO_NJ_CLUSTER", cl)
        p1, b1 = engine.nj_hard(pdm)
        assert torch.equal(p0, p1) and torch.equal(b0, b1), cl
    monkeypatch.setenv("DODO_NJ_NO_LEAN", "1")
    p2, b2 = engine.nj_hard(pdm)
    assert torch.equal(p0, p2) and torch.equal(b0, b2)


@pytest.mark.gpu
def test_lean_soft_nj_backward_equals_the_replay(monkeypatch):
    """nj_soft_bwd_lean_kernel (trees up to 96 taxa: d blens / d pdm from the bilinear-form structure, one thread per
    matrix entry) against the reverse replay of the joins (nj_soft_bwd_kernel), incl. clamped branches."""
    from dodonaphy_b200 import engine

    rng = np.random.default_rng(21)
    seen_clamped = False
    for S, B in ((2, 3), (3, 4), (4, 4), (7, 5), (33, 4), (64, 6), (80, 2), (96, 2)):
        x = rng.normal(0, 0.08, (B, S, 4))
        if S >= 7:
            x[:, 3] = x[:, 1]  # coincident taxa: clamped branches (flags bit0 / bit1)
        pdm = engine.pdm_forward(torch.tensor(x).cuda(), -1.0, "torch")
        peel, blens, flags = engine.nj_soft(pdm, 1e-8, return_flags=True)
        if S >= 4:  # the clamp flags are inputs of the backward: set some by hand (a clamped branch passes no gradient)
            flags = flags.clone()
            flags[:, 1] |= 1
            flags[:, 2] |= 2
            flags[0, S - 2] |= 3
        seen_clamped = seen_clamped or bool((flags & 3).any())
        g = torch.tensor(rng.normal(size=(B, 2 * S - 2))).cuda()
        monkeypatch.delenv("DODO_NJ_NO_LEAN", raising=False)
        g1 = engine.nj_soft_backward(peel, flags, g)
        monkeypatch.setenv("DODO_NJ_NO_LEAN", "1")
        g0 = engine.nj_soft_backward(peel, flags, g)
        ref = g0.cpu().numpy()
        np.testing.assert_allclose(g1.cpu().numpy(), ref, rtol=1e-10, atol=1e-13 * max(1.0, np.abs(ref).max()))
    assert seen_clamped
