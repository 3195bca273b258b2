"""CPU tests: the oracle restatements (oracle/) against the golden vectors the reference's own
tests hold for the hot path and against fixtures produced by running the real reference
(tests/golden/make_golden.py).  No GPU, no /root/reference needed."""
import numpy as np
import pytest
import torch

from oracle import hyp, nj, phylo

from conftest import golden_problem, load_golden


# ------------------------------------------------------------------ DS1 known answers
def _ds1():
    g = load_golden("ds1.npz")
    tips = [phylo.mask_to_partial(m) for m in g["masks"]]
    return g, tips


def test_ds1_mrbayes_tree_lnl():
    """test/test_phylo.py:214-273: lnL = -6904.632 +- 1e-3 (torch and Cython-NumPy paths)."""
    g, tips = _ds1()
    ll = phylo.compute_LL_np(tips, g["weights"], g["mb_peel"], g["mb_blens"])
    assert ll == pytest.approx(-6904.632, abs=1e-3)
    assert ll == pytest.approx(float(g["mb_lnL_torch"]), rel=1e-13)
    assert ll == pytest.approx(float(g["mb_lnL_cython"]), rel=1e-13)


def test_ds1_nj_tree_lnl():
    """test/test_bito.py:35,136 and test/test_hmap.py:112: lnL = -7130.96813417."""
    g, tips = _ds1()
    ll = phylo.compute_LL_np(tips, g["weights"], g["nj_peel"], g["nj_blens"])
    assert ll == pytest.approx(-7130.96813417, abs=1e-7)
    assert ll == pytest.approx(float(g["nj_lnL_torch"]), rel=1e-13)


@pytest.mark.parametrize("tree", ["mb", "nj"])
def test_ds1_gradients(tree):
    g, tips = _ds1()
    tt = [torch.tensor(t) for t in tips]
    w = torch.tensor(g["weights"])
    b = torch.tensor(g[f"{tree}_blens"], requires_grad=True)
    ll = phylo.treelikelihood_torch(tt, w, g[f"{tree}_peel"], phylo.jc69_p_t_torch(b), torch.full([4], 0.25, dtype=torch.float64))
    ll.backward()
    np.testing.assert_allclose(b.grad.numpy(), g[f"{tree}_g_blens"], rtol=1e-10, atol=1e-9)
    # GTR
    b = torch.tensor(g[f"{tree}_blens"], requires_grad=True)
    fr = torch.tensor(g["gtr_freqs"])
    mats = phylo.gtr_p_t_torch(b, torch.tensor(g["gtr_rates"]), fr)
    ll = phylo.treelikelihood_torch(tt, w, g[f"{tree}_peel"], mats, fr)
    ll.backward()
    assert ll.item() == pytest.approx(float(g[f"{tree}_gtr_lnL"]), rel=1e-12)
    np.testing.assert_allclose(b.grad.numpy(), g[f"{tree}_gtr_g_blens"], rtol=1e-9, atol=1e-8)
    np.testing.assert_allclose(phylo.gtr_p_t(g[f"{tree}_blens"], g["gtr_rates"], g["gtr_freqs"]),
                               g[f"{tree}_gtr_mats"], rtol=1e-12, atol=1e-15)


# ------------------------------------------------------------------ P(t) properties
def test_gtr_equals_jc69():
    """test/test_phylo.py:276-283."""
    b = np.array([0.1])
    jc = phylo.jc69_p_t(b)
    gtr = phylo.gtr_p_t(b, np.full(6, 1.0 / 6.0), np.full(4, 0.25))
    np.testing.assert_allclose(jc[:, 0], gtr, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("blens_in,size", [([[1.0, 1.0]], [1, 2, 4, 4]), ([[1.0]], [1, 1, 4, 4]),
                                           ([[1.0], [1.0]], [2, 1, 4, 4]),
                                           ([[1.0, 1.0], [1.0, 1.0], [1.0, 1.0]], [3, 2, 4, 4])])
def test_gtr_mats_size(blens_in, size):
    """test/test_phylo.py:286-300."""
    rates = np.random.default_rng(0).dirichlet(np.ones(6))
    freqs = np.random.default_rng(1).dirichlet(np.ones(4))
    m = phylo.gtr_p_t(np.array(blens_in), rates, freqs)
    assert list(m.shape) == size
    np.testing.assert_allclose(m.sum(-1), 1.0, rtol=1e-12)


def test_discrete_gamma_rates():
    r4 = phylo.discrete_gamma_rates(0.5, 4)
    assert r4.mean() == pytest.approx(1.0, rel=1e-12)
    # Yang 1994 table, alpha = 0.5, K = 4
    np.testing.assert_allclose(r4, [0.03338775, 0.25191592, 0.82026848, 2.89442785], rtol=1e-6)
    big = phylo.discrete_gamma_rates(1e6, 4)
    np.testing.assert_allclose(big, 1.0, atol=2e-3)


def test_compress_alignment_iupac():
    """test/test_phylo.py:12-102: IUPAC codes map to the expected partials."""
    seqs = ["ACGTRYMWSKBDHVN?-x", "AAAAAAAAAAAAAAAAAA"]
    tips, w, masks = phylo.compress_alignment(seqs)
    col = {c: i for i, c in enumerate(sorted(set(zip(*seqs))))}
    # every column distinct -> 18 patterns, weight 1
    assert w.tolist() == [1] * 18
    expect = {"A": [1, 0, 0, 0], "C": [0, 1, 0, 0], "G": [0, 0, 1, 0], "T": [0, 0, 0, 1], "R": [1, 0, 1, 0],
              "Y": [0, 1, 0, 1], "M": [1, 1, 0, 0], "W": [1, 0, 0, 1], "S": [0, 1, 1, 0], "K": [0, 0, 1, 1],
              "B": [0, 1, 1, 1], "D": [1, 0, 1, 1], "H": [1, 1, 0, 1], "V": [1, 1, 1, 0], "N": [1, 1, 1, 1],
              "?": [1, 1, 1, 1], "-": [1, 1, 1, 1], "x": [1, 1, 1, 1]}
    for (c0, c1), i in col.items():
        assert tips[0][:, i].tolist() == expect[c0]
        assert tips[1][:, i].tolist() == expect["A"]


# ------------------------------------------------------------------ pruning fixtures
def test_prune_fixtures():
    g = load_golden("prune.npz")
    for i in range(int(g["n"])):
        tips = [phylo.mask_to_partial(m) for m in g[f"{i}_masks"]]
        w, peel, blens = g[f"{i}_weights"], g[f"{i}_peel"], g[f"{i}_blens"]
        ll = phylo.compute_LL_np(tips, w, peel, blens)
        assert ll == pytest.approx(float(g[f"{i}_jc_lnL"]), rel=1e-12)
        assert ll == pytest.approx(float(g[f"{i}_lnL_cython"]), rel=1e-12)
        mats = phylo.gtr_p_t(blens, g[f"{i}_rates"], g[f"{i}_freqs"])
        np.testing.assert_allclose(mats, g[f"{i}_gtr_mats"], rtol=1e-11, atol=1e-15)
        ll = phylo.treelikelihood(tips, w, peel, mats, g[f"{i}_freqs"])
        assert ll == pytest.approx(float(g[f"{i}_gtr_lnL"]), rel=1e-12)
        # gradients via the torch twin
        tt = [torch.tensor(t) for t in tips]
        b = torch.tensor(blens, requires_grad=True)
        fr = torch.tensor(g[f"{i}_freqs"])
        m = phylo.gtr_p_t_torch(b, torch.tensor(g[f"{i}_rates"]), fr)
        m.retain_grad()
        phylo.treelikelihood_torch(tt, torch.tensor(w), peel, m, fr).backward()
        np.testing.assert_allclose(b.grad.numpy(), g[f"{i}_gtr_g_blens"], rtol=1e-8, atol=1e-7)
        np.testing.assert_allclose(m.grad.numpy(), g[f"{i}_gtr_g_mats"], rtol=1e-9, atol=1e-7)


def test_gamma_mixture_limits():
    """K>1 is new functionality (parity unpinned): K=1 == reference semantics, alpha->inf == K=1."""
    g = load_golden("prune.npz")
    i = 1
    tips = [phylo.mask_to_partial(m) for m in g[f"{i}_masks"]]
    w, peel, blens = g[f"{i}_weights"], g[f"{i}_peel"], g[f"{i}_blens"]
    one = phylo.treelikelihood_gamma(tips, w, peel, phylo.jc69_p_t(blens), np.full(4, 0.25))
    assert one == pytest.approx(float(g[f"{i}_jc_lnL"]), rel=1e-13)
    rates = phylo.discrete_gamma_rates(1e7, 4)
    many = phylo.treelikelihood_gamma(tips, w, peel, phylo.jc69_p_t(blens[:, None] * rates[None, :]), np.full(4, 0.25))
    assert many == pytest.approx(one, rel=1e-6)


# ------------------------------------------------------------------ distances
def test_pdm_fixtures():
    g = load_golden("pdm.npz")
    for i in range(int(g["n"])):
        x, curv, mats, C = g[f"{i}_x"], float(g[f"{i}_curv"]), bool(g[f"{i}_matsumoto"]), g[f"{i}_C"]
        np.testing.assert_allclose(hyp.get_pdm_torch(x, curv, mats, "up"), g[f"{i}_pdm_torch_up"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(hyp.get_pdm_torch(x, curv, mats, "wrap"), g[f"{i}_pdm_torch_wrap"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(hyp.get_pdm_np(x, curv, mats), g[f"{i}_pdm_np"], rtol=1e-7, atol=1e-10)
        gx, gc = hyp.get_pdm_torch_vjp(x, C, curv, mats)
        np.testing.assert_allclose(gx, g[f"{i}_gx_up"], rtol=1e-7, atol=1e-8)
        np.testing.assert_allclose(gc, g[f"{i}_gc_up"][0], rtol=1e-9)


def test_lorentz_norm_after_lift():
    """test/test_hyperboloid.py:115-148 property: <s,s>_L = -1 after either lift."""
    x = np.random.default_rng(3).normal(0, 0.4, (9, 4))
    for lift in (hyp.lift_up, hyp.lift_wrap):
        s = lift(x)
        np.testing.assert_allclose(-s[:, 0] ** 2 + np.sum(s[:, 1:] ** 2, axis=1), -1.0, rtol=1e-12)


# ------------------------------------------------------------------ NJ
def test_nj_hard_fixtures_bit_exact():
    g = load_golden("nj_hard.npz")
    for i in range(int(g["n"])):
        peel, blens = nj.nj_hard(g[f"{i}_pdm"])
        assert np.array_equal(peel, g[f"{i}_peel"]), i
        assert np.array_equal(blens, g[f"{i}_blens"]), i


def test_nj_soft_fixtures():
    g = load_golden("nj_soft.npz")
    for i in range(int(g["n"])):
        tau = float(g[f"{i}_tau"])
        peel, blens = nj.nj_soft(g[f"{i}_pdm"], tau)
        assert np.array_equal(peel, g[f"{i}_peel"]), (i, tau)
        np.testing.assert_allclose(blens, g[f"{i}_blens"], rtol=1e-12, atol=1e-15)
        pdm = torch.tensor(g[f"{i}_pdm"], requires_grad=True)
        p2, b2 = nj.nj_soft_torch(pdm, tau)
        assert np.array_equal(p2, g[f"{i}_peel"])
        (b2 * torch.tensor(g[f"{i}_c"])).sum().backward()
        ours = pdm.grad.numpy()
        ref = g[f"{i}_g_pdm"]
        np.testing.assert_allclose(ours + ours.T, ref + ref.T, rtol=1e-9, atol=1e-12)


def test_nj_soft_fixture_at_the_headline_size():
    """One 64-taxon tree (the C2 shape) through the reference's own soft NJ: the restatement must give
    the same peel, branch lengths and gradient."""
    g = load_golden("nj_soft_c2.npz")
    tau = float(g["tau"])
    peel, blens = nj.nj_soft(g["pdm"], tau)
    assert np.array_equal(peel, g["peel"])
    np.testing.assert_allclose(blens, g["blens"], rtol=1e-12, atol=1e-15)
    pdm = torch.tensor(g["pdm"], requires_grad=True)
    p2, b2 = nj.nj_soft_torch(pdm, tau)
    assert np.array_equal(p2, g["peel"])
    (b2 * torch.tensor(g["c"])).sum().backward()
    ours, ref = pdm.grad.numpy(), g["g_pdm"]
    np.testing.assert_allclose(ours + ours.T, ref + ref.T, rtol=1e-9, atol=1e-12)


def test_prune_fixture_at_the_headline_size():
    """The restatement of the pruning arithmetic against the reference's answers at 64 taxa x 10 000
    patterns (JC69; GTR; GTR with the category axis summed as the reference's broadcasting does)."""
    from dodonaphy_b200 import synth

    g = load_golden("prune_c2.npz")
    prob = synth.make_problem(64, 10000, 128, D=5, seed=1002)
    assert int(prob["masks"].astype(np.int64).sum()) == int(g["mask_sum"])
    tips = [torch.tensor(phylo.mask_to_partial(m)) for m in prob["masks"]]
    w = torch.tensor(prob["weights"])
    for key, K in (("jc", 1), ("gtr", 1), ("gtrk4", 4)):
        bt = torch.tensor(g["blens"], requires_grad=True)
        t = bt[:, None] * (torch.tensor(g["cat_rates"]) if K == 4 else torch.ones(1, dtype=torch.float64))[None, :]
        if key == "jc":
            mats, fr = phylo.jc69_p_t_torch(t), torch.full([4], 0.25, dtype=torch.float64)
        else:
            mats, fr = phylo.gtr_p_t_torch(t, torch.tensor(g["rates"]), torch.tensor(g["freqs"])), torch.tensor(g["freqs"])
        ll = phylo.treelikelihood_torch(tips, w, g["peel"], mats, fr, gamma_mix=False)
        ll.backward()
        assert ll.item() == pytest.approx(float(g[f"{key}_lnL"]), rel=1e-11)
        ref = g[f"{key}_g_blens"]
        np.testing.assert_allclose(bt.grad.numpy(), ref, rtol=1e-9, atol=1e-10 * np.abs(ref).max())


def test_hard_nj_fixtures_at_the_config_sizes():
    """tests/golden/configs.npz: the reference's Cpeeler.nj_np on bit-reproducible distances at 200, 500
    and 2000 taxa -- the C restatement returns the same peel and branch lengths bit for bit."""
    from dodonaphy_b200 import synth
    from oracle import cbuild

    g = load_golden("configs.npz")
    cases = [(synth.det_distances(200, 5, 3000) * 0.05, "c3_0"), (synth.det_distances(200, 5, 3001) * 0.05, "c3_1"),
             (synth.det_distances(500, 5, 5000) * 0.02, "c5"), (synth.det_distances(2000, 5, 4000), "c4")]
    for pdm, key in cases:
        peel, blens = cbuild.nj_hard_c(pdm)
        assert np.array_equal(peel, g[f"{key}_peel"]), key
        assert np.array_equal(blens, g[f"{key}_blens"]), key


def test_whole_chain_fixture_at_the_headline_size():
    """tests/golden/chain_c2.npz: one VI sample of the headline workload end to end in the reference.  The
    restatements chained the same way (distance -> soft NJ -> GTR P(t) -> pruning -> autograd) give the
    same topology, lnL and gradient w.r.t. the 64 x 5 locations."""
    from dodonaphy_b200 import synth

    g = load_golden("chain_c2.npz")
    prob = golden_problem("c2")
    masks = prob["masks"]
    assert int((masks.astype(np.int64) * (np.arange(masks.shape[1]) % 97 + 1)).sum()) == int(g["mask_dot"])
    x = torch.tensor(g["x"], requires_grad=True)
    z = torch.sqrt((x * x).sum(1) + 1)
    H = torch.minimum(x @ x.T - torch.outer(z, z), torch.tensor(-1.0000001, dtype=torch.float64))
    pdm = torch.acosh(-H) * (1 - torch.eye(64, dtype=torch.float64))
    peel, blens = nj.nj_soft_torch(pdm, 1e-8)
    assert np.array_equal(peel, g["peel"])
    tips = [torch.tensor(phylo.mask_to_partial(m)) for m in masks]
    mats = phylo.gtr_p_t_torch(blens[:, None], torch.tensor(g["rates"]), torch.tensor(g["freqs"]))
    ll = phylo.treelikelihood_torch(tips, torch.tensor(prob["weights"]), peel, mats, torch.tensor(g["freqs"]), gamma_mix=False)
    ll.backward()
    assert ll.item() == pytest.approx(float(g["lnL"]), rel=1e-11)
    ref = g["g_x"]
    np.testing.assert_allclose(x.grad.numpy(), ref, rtol=1e-8, atol=1e-9 * np.abs(ref).max())


def test_q_matrix_wikipedia():
    """test/test_peeler.py:228-303: Q of the 5-taxon textbook example."""
    D = np.array([[0, 5, 9, 9, 8], [5, 0, 10, 10, 9], [9, 10, 0, 8, 7], [9, 10, 8, 0, 3], [8, 9, 7, 3, 0]], dtype=np.float64)
    Q = nj.compute_Q(D, np.zeros(5, dtype=bool))
    expect = np.array([[0, -50, -38, -34, -34], [-50, 0, -38, -34, -34], [-38, -38, 0, -40, -40],
                       [-34, -34, -40, 0, -48], [-34, -34, -40, -48, 0]], dtype=np.float64)
    np.fill_diagonal(Q, 0.0)
    np.testing.assert_array_equal(Q, expect)


def _dogbone_pdm(matsumoto=False):
    """Inputs of test/test_peeler.py:87-95: 4 points at Poincare radius .5, lifted to the
    hyperboloid and (a quirk of that test) handed to Chyp_np.get_pdm as if they were R^3 points."""
    theta = np.array([np.pi / 10, -np.pi / 10, np.pi * 6 / 8, -np.pi * 6 / 8])
    poin = 0.5 * np.stack([np.cos(theta), np.sin(theta)], axis=1)
    return hyp.get_pdm_np(hyp.poincare_to_hyper(poin), matsumoto=matsumoto)


DOGBONE_TOPOLOGIES = [[[1, 0, 4], [3, 2, 5], [5, 4, 6]], [[0, 1, 4], [2, 3, 5], [4, 5, 6]], [[2, 3, 4], [1, 4, 5], [0, 5, 6]],
                      [[0, 1, 4], [4, 2, 5], [5, 3, 6]], [[2, 3, 4], [0, 4, 5], [1, 5, 6]], [[2, 3, 4], [0, 1, 5], [5, 4, 6]],
                      [[0, 1, 4], [4, 3, 5], [2, 5, 6]]]


def test_nj_dogbone():
    """test/test_peeler.py:87-132,306-340: 4-taxon example; sum of blens 3.29274740 (hard, soft),
    2.0318 with the Matsumoto adjustment; topology in the reference's list of acceptable peels."""
    pd = _dogbone_pdm()
    peel, blens = nj.nj_hard(pd)
    assert blens.sum() == pytest.approx(3.29274740, abs=1e-3)
    assert peel.tolist() in DOGBONE_TOPOLOGIES
    peel_s, blens_s = nj.nj_soft(pd, 1e-4)
    assert blens_s.sum() == pytest.approx(3.29274740, abs=0.05)
    assert blens_s.sum() == pytest.approx(blens.sum(), abs=1e-9)
    peel_m, blens_m = nj.nj_hard(_dogbone_pdm(matsumoto=True))
    assert blens_m.sum() == pytest.approx(2.0318, abs=1e-3)
    assert peel_m.tolist() in DOGBONE_TOPOLOGIES


def test_softsort_values():
    """test/test_soft.py:21-32: soft sort of [2,1,3]... pinned values at tau=.5."""
    s = np.array([3.0, 2.0, 1.0])
    P = nj.soft_sort_dense(s, 0.5)
    np.testing.assert_allclose(P @ s, [2.8509, 2.0, 1.1491], atol=1e-4)


def test_soft_argmin_closed_form_matches_dense():
    rng = np.random.default_rng(5)
    for n in (3, 6, 11):
        for tau in (1e-8, 1e-4, 0.3):
            Q = np.tril(rng.normal(size=(n, n)))
            Q[rng.integers(1, n), 0] = Q.min()  # inject an exact tie
            assert nj.soft_argmin(Q, tau) == pytest.approx(nj.soft_argmin_dense(Q, tau), rel=1e-12, abs=1e-12)


def test_geodesic_oracle_against_reference_outputs():
    """oracle/geodesic.py vs peeler.make_soft_peel_tips / make_hard_peel_geodesic run on the reference."""
    import torch

    from oracle import geodesic as G

    g = load_golden("geodesic.npz")
    for i in range(int(g["n_soft"])):
        x = g[f"s{i}_x"]
        peel, il, bl = G.soft_peel_tips(x)
        np.testing.assert_array_equal(peel, g[f"s{i}_peel"])
        np.testing.assert_allclose(il, g[f"s{i}_int"], rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(bl, g[f"s{i}_blens"], rtol=1e-11, atol=1e-13)
        if x.shape[0] <= 12:
            xt = torch.tensor(x, requires_grad=True)
            _, il2, bl2 = G.soft_peel_tips_torch(xt)
            (gr,) = torch.autograd.grad((bl2 * torch.tensor(g[f"s{i}_wb"])).sum() + (il2 * torch.tensor(g[f"s{i}_wi"])).sum(), xt)
            np.testing.assert_allclose(gr.numpy(), g[f"s{i}_grad"], rtol=1e-9, atol=1e-11 * np.abs(g[f"s{i}_grad"]).max())
    for i in range(int(g["n_hard"])):
        if g[f"h{i}_sheet"].shape[0] > 30:
            continue
        peel, il = G.hard_peel_geodesic(g[f"h{i}_sheet"])
        np.testing.assert_array_equal(peel, g[f"h{i}_peel"])
        np.testing.assert_allclose(il, g[f"h{i}_int"], rtol=1e-12, atol=1e-14)
