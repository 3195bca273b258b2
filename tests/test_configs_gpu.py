"""GPU parity at the shapes of BASELINE.json's other configurations (sizes the oracle finishes in
seconds) plus size-independent properties at the full shapes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.as_tensor(a).cuda()


def test_c3_shape_mcmc_chains_forward_only():
    """configs[2]: 200 taxa, JC69, independent chains, gradient-free path
    (Chyp_np.get_pdm -> Cpeeler.nj_np -> Cphylo.compute_LL_np)."""
    from dodonaphy_b200 import engine, synth
    from dodonaphy_b200.engine import SubstModel
    from dodonaphy_b200.pipeline import HotPath
    from oracle import cbuild, phylo

    S, L, B = 200, 3000, 4
    prob = synth.make_problem(S, L, B, D=5, seed=1003)
    hp = HotPath(prob["masks"], prob["weights"], SubstModel.jc69())
    out = hp.value(_dev(prob["xs"]))
    tips = [phylo.mask_to_partial(m) for m in prob["masks"]]
    pdm = engine.pdm_forward(_dev(prob["xs"]), -1.0, "numpy").cpu().numpy()
    for b in range(B):
        peel, blens = cbuild.nj_hard_c(pdm[b])
        assert np.array_equal(out["peel"][b].cpu().numpy(), peel)
        assert np.array_equal(out["blens"][b].cpu().numpy(), blens)
        assert out["lnL"][b].item() == pytest.approx(phylo.compute_LL_np(tips, prob["weights"], peel, blens), rel=1e-9)


@pytest.mark.parametrize("S,B", [(700, 2), (2000, 2)])
def test_c4_shape_topology_stress_hard_nj_bit_exact(S, B):
    """configs[3]: pairwise distance + batched NJ at up to 2000 taxa; hard NJ bit-exact against the
    C restatement of Cpeeler.nj_np (the Cython original needs ~4 min per 2000-taxon tree)."""
    from dodonaphy_b200 import engine
    from dodonaphy_b200.engine import validate_peel
    from oracle import cbuild, hyp

    rng = np.random.default_rng(1004)
    xs = rng.normal(0, 0.05, (B, S, 5))
    pdm = engine.pdm_forward(_dev(xs), -1.0, "numpy")
    off = ~np.eye(S, dtype=bool)
    np.testing.assert_allclose(pdm[0].cpu().numpy()[off], hyp.get_pdm_np(xs[0])[off], rtol=1e-10, atol=1e-14)
    peel, blens = engine.nj_hard(pdm)
    p_ref, b_ref = cbuild.nj_hard_c(pdm[0].cpu().numpy())
    assert np.array_equal(peel[0].cpu().numpy(), p_ref)
    assert np.array_equal(blens[0].cpu().numpy(), b_ref)
    validate_peel(peel.cpu().numpy(), S)


def test_c4_shape_soft_nj_medium_vs_oracle_and_large_properties():
    from dodonaphy_b200 import engine
    from dodonaphy_b200.engine import validate_peel
    from oracle import hyp, nj

    rng = np.random.default_rng(1005)
    S = 300
    x = rng.normal(0, 0.05, (S, 5))
    pdm = hyp.get_pdm_torch(x)
    peel, blens, flags = engine.nj_soft(_dev(pdm), 1e-8, return_flags=True)
    assert int((flags & 12).sum().item()) == 0
    p_ref, b_ref = nj.nj_soft(pdm, 1e-8)
    assert np.array_equal(peel.cpu().numpy(), p_ref)
    np.testing.assert_allclose(blens.cpu().numpy(), b_ref, rtol=1e-9, atol=1e-15)
    # 2000 taxa: the reference cannot run at all (its SoftSort matrix would hold (3999^2)^2 doubles);
    # check what does not depend on size: a valid tree, exactness guard silent, and the NJ identity
    # that the two branches of every join add up to the distance between the joined nodes.
    S = 2000
    xs = rng.normal(0, 0.05, (1, S, 5))
    pdm = engine.pdm_forward(_dev(xs), -1.0, "torch")
    peel, blens, flags = engine.nj_soft(pdm, 1e-8, return_flags=True)
    assert int((flags & 12).sum().item()) == 0
    validate_peel(peel.cpu().numpy(), S)
    pl, bl, d = peel[0].cpu().numpy(), blens[0].cpu().numpy(), pdm[0].cpu().numpy()
    first = pl[0]
    assert bl[first[0]] + bl[first[1]] == pytest.approx(d[first[0], first[1]], rel=1e-12)
    assert (bl > 0).all()


def test_c5_shape_long_alignment_site_blocks():
    """configs[4]: 500 taxa, one tree, long alignment split into site blocks: per-block lnL and
    gradients add up to the whole (what the NCCL all_reduce does across GPUs), rescaling on, and the
    whole agrees with the log-space oracle."""
    from dodonaphy_b200 import synth
    from dodonaphy_b200.dist import shard_bounds
    from dodonaphy_b200.engine import PruneEngine, SubstModel
    from oracle import cbuild, hyp, phylo

    S, L = 500, 6000
    prob = synth.make_problem(S, L, 1, D=5, seed=1006)
    peel, blens = cbuild.nj_hard_c(hyp.get_pdm_np(prob["xs"][0]))
    m = SubstModel.jc69()
    whole = PruneEngine(prob["masks"], prob["weights"])
    ll, g = whole.loglik(_dev(peel), _dev(blens), m, grad=True)
    parts_ll, parts_g = 0.0, 0.0
    for r in range(4):
        lo, hi = shard_bounds(L, r, 4)
        eng = PruneEngine(np.ascontiguousarray(prob["masks"][:, lo:hi]), prob["weights"][lo:hi])
        l, gg = eng.loglik(_dev(peel), _dev(blens), m, grad=True)
        parts_ll = parts_ll + l
        parts_g = parts_g + gg
    assert parts_ll.item() == pytest.approx(ll.item(), rel=1e-12)
    np.testing.assert_allclose(parts_g.cpu().numpy(), g.cpu().numpy(), rtol=1e-10, atol=1e-10 * g.abs().max().item())
    bt = torch.tensor(blens, requires_grad=True)
    ref = phylo.treelikelihood_rescaled_torch([torch.tensor(phylo.mask_to_partial(x)) for x in prob["masks"]],
                                              torch.tensor(prob["weights"]), peel, phylo.jc69_p_t_torch(bt), torch.full([4], 0.25, dtype=torch.float64))
    ref.backward()
    assert ll.item() == pytest.approx(ref.item(), rel=1e-9)
    np.testing.assert_allclose(g[0].cpu().numpy(), bt.grad.numpy(), rtol=1e-9, atol=1e-9 * np.abs(bt.grad.numpy()).max())


def test_c2_full_size_headline_workload():
    """configs[1] at full size (64 taxa x 10 000 patterns x 128 samples, GTR+G4, soft NJ): two samples
    against the oracle end to end, plus size-independent properties for the whole batch: lnL additive
    over site blocks, forward-only == forward of the gradient kernel, directional finite difference
    of lnL along the gradient w.r.t. the locations."""
    from dodonaphy_b200 import synth
    from dodonaphy_b200.engine import PruneEngine, SubstModel
    from dodonaphy_b200.pipeline import HotPath
    from oracle import hyp, nj, phylo

    S, L, B, D, K = 64, 10000, 128, 5, 4
    prob = synth.make_problem(S, L, B, D=D, seed=1002)
    model = SubstModel.gtr(prob["rates"], prob["freqs"])
    rates = SubstModel.discrete_gamma(0.5, K)
    hp = HotPath(prob["masks"], prob["weights"], model, rates=rates, mix=True, tau=1e-8)
    x = _dev(prob["xs"])
    out = hp.value_and_grad(x)
    assert torch.isfinite(out["lnL"]).all() and int((out["flags"] & 12).sum().item()) == 0
    tips = [torch.tensor(phylo.mask_to_partial(m)) for m in prob["masks"]]
    for b in (0, 77):
        peel, blens = nj.nj_soft(hyp.get_pdm_torch(prob["xs"][b]), 1e-8)
        assert np.array_equal(out["peel"][b].cpu().numpy(), peel)
        bt = torch.tensor(blens, requires_grad=True)
        mats = phylo.gtr_p_t_torch(bt[:, None] * torch.tensor(rates)[None, :], torch.tensor(prob["rates"]), torch.tensor(prob["freqs"]))
        ref = phylo.treelikelihood_torch(tips, torch.tensor(prob["weights"]), peel, mats, torch.tensor(prob["freqs"]), gamma_mix=True)
        ref.backward()
        assert out["lnL"][b].item() == pytest.approx(ref.item(), rel=1e-9)
        g = out["grad_blens"][b].cpu().numpy()
        np.testing.assert_allclose(g, bt.grad.numpy(), rtol=1e-9, atol=1e-9 * np.abs(g).max())
    # additivity over site blocks (what site sharding relies on)
    halves = 0.0
    for lo, hi in ((0, 4321), (4321, L)):
        eng = PruneEngine(np.ascontiguousarray(prob["masks"][:, lo:hi]), prob["weights"][lo:hi])
        halves = halves + eng.loglik(out["peel"], out["blens"], model, rates=rates, mix=True)[0]
    np.testing.assert_allclose(halves.cpu().numpy(), out["lnL"].cpu().numpy(), rtol=1e-12)
    # forward-only kernel agrees with the gradient kernel's value bit for bit
    fwd, _ = hp.prune.loglik(out["peel"], out["blens"], model, rates=rates, mix=True, grad=False)
    assert torch.equal(fwd, out["lnL"])
    # directional derivative w.r.t. the locations (through soft NJ and the distance kernel)
    d = out["grad_x"] / out["grad_x"].abs().amax(dim=(1, 2), keepdim=True)
    h = 1e-7
    plus = hp.value_and_grad(x + h * d)
    minus = hp.value_and_grad(x - h * d)
    same = (torch.equal(plus["peel"], out["peel"]) and torch.equal(minus["peel"], out["peel"]))
    fd = (plus["lnL"] - minus["lnL"]) / (2 * h)
    an = (out["grad_x"] * d).sum(dim=(1, 2))
    ok = torch.isclose(fd, an, rtol=2e-4, atol=1e-3)
    assert same or ok.float().mean().item() > 0.9
    assert ok.float().mean().item() > 0.9, (fd[:4], an[:4])


def test_c2_full_size_against_the_reference_itself():
    """64 taxa x 10 000 patterns (sample 0 of the bench's problem): the REFERENCE's lnL and autograd
    gradients (tests/golden/prune_c2.npz, generated by importing the reference) for JC69, GTR and GTR
    with the four gamma rates as the reference's category axis (per-category lnL summed) -- through
    the engine (device-formed P) and through the reference-shaped API (caller's mats, torch autograd
    down to the unconstrained GTR leaves).  1e-9 relative."""
    from conftest import golden_problem, load_golden
    from dodonaphy_b200 import phylo, synth
    from dodonaphy_b200.engine import PruneEngine, SubstModel
    from dodonaphy_b200.phylomodel import PhyloModel
    from oracle import phylo as ophylo

    g = load_golden("prune_c2.npz")
    prob = golden_problem("c2")
    masks, weights = prob["masks"], prob["weights"]
    assert int((masks.astype(np.int64) * (np.arange(masks.shape[1]) % 97 + 1)).sum()) == int(g["mask_dot"])
    assert int(weights.sum()) == int(g["weight_sum"])
    np.testing.assert_array_equal(prob["rates"], g["rates"])
    peel, blens = g["peel"], g["blens"]
    eng = PruneEngine(masks, weights)

    def check(ll, grad, key):
        assert ll.item() == pytest.approx(float(g[f"{key}_lnL"]), rel=1e-9)
        ref = g[f"{key}_g_blens"]
        np.testing.assert_allclose(grad.cpu().numpy().reshape(-1), ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())

    gtr = SubstModel.gtr(prob["rates"], prob["freqs"])
    for scale in (False, True):
        check(*eng.loglik(_dev(peel), _dev(blens), SubstModel.jc69(), grad=True, scale=scale), "jc")
        check(*eng.loglik(_dev(peel), _dev(blens), gtr, grad=True, scale=scale), "gtr")
        check(*eng.loglik(_dev(peel), _dev(blens), gtr, rates=g["cat_rates"], mix=False, grad=True, scale=scale), "gtrk4")
    # the reference-shaped call: calculate_treelikelihood(partials, weights, peel, mats, freqs) + autograd
    partials = [torch.tensor(ophylo.mask_to_partial(m)) for m in masks]
    pm = PhyloModel("GTR", torch.tensor(prob["rates"]), torch.tensor(prob["freqs"]))
    bt = torch.tensor(blens, requires_grad=True)
    ll = phylo.calculate_treelikelihood(partials, torch.tensor(weights), peel, pm.get_transition_mats(bt), pm.freqs)
    ll.backward()
    check(ll, bt.grad, "gtr")
    for ours, key in ((pm._sub_rates.grad, "gtr_g_raw_rates"), (pm._freqs.grad, "gtr_g_raw_freqs")):
        np.testing.assert_allclose(ours.numpy(), g[key], rtol=1e-8, atol=1e-9 * np.abs(g[key]).max())


def _same_alignment(masks, checksum):
    """The stored answers belong to one particular alignment: regenerated from its seed on the committed
    guide tree (conftest.golden_problem), it is the same on every host."""
    assert int(masks.astype(np.int64).sum()) == int(checksum)


def test_configs_2_to_4_against_the_reference_itself():
    """tests/golden/configs.npz: the reference's own Cpeeler.nj_np / Cphylo.compute_LL_np / torch path at
    the shapes of configs[2..4].  Hard NJ on bit-reproducible distances: peel AND branch lengths equal
    bit for bit (200, 500 and 2000 taxa); lnL (and the gradient at 500 taxa) to 1e-9."""
    from conftest import golden_problem, load_golden
    from dodonaphy_b200 import engine, synth
    from dodonaphy_b200.engine import PruneEngine, SubstModel

    g = load_golden("configs.npz")
    jc = SubstModel.jc69()
    # configs[2]: 200 taxa, chains
    prob = synth.make_problem(200, 20000, 64, D=5, seed=1003)
    eng = None
    for i in range(2):
        pdm = synth.det_distances(200, 5, 3000 + i) * 0.05
        peel, blens = engine.nj_hard(_dev(pdm))
        assert np.array_equal(peel.cpu().numpy().reshape(-1, 3), g[f"c3_{i}_peel"])
        assert np.array_equal(blens.cpu().numpy().reshape(-1), g[f"c3_{i}_blens"])
        if int(prob["masks"].astype(np.int64).sum()) == int(g["c3_mask_sum"]):
            eng = eng or PruneEngine(prob["masks"], prob["weights"])
            for scale in (False, True):
                ll, _ = eng.loglik(peel, blens, jc, grad=False, scale=scale)
                assert ll.item() == pytest.approx(float(g[f"c3_{i}_lnL"]), rel=1e-9)
    # configs[3]: one 2000-taxon tree
    peel, blens = engine.nj_hard(_dev(synth.det_distances(2000, 5, 4000)))
    assert np.array_equal(peel.cpu().numpy().reshape(-1, 3), g["c4_peel"])
    assert np.array_equal(blens.cpu().numpy().reshape(-1), g["c4_blens"])
    # configs[4]: 500 taxa
    pdm = synth.det_distances(500, 5, 5000) * 0.02
    peel, blens = engine.nj_hard(_dev(pdm))
    assert np.array_equal(peel.cpu().numpy().reshape(-1, 3), g["c5_peel"])
    assert np.array_equal(blens.cpu().numpy().reshape(-1), g["c5_blens"])
    # likelihood answers: on the reference's NJ tree of the sample's own locations (on an unrelated tree
    # the reference's unscaled product is -inf at 500 taxa)
    prob = golden_problem("c5")
    _same_alignment(prob["masks"], g["c5_mask_sum"])
    ll, gr = PruneEngine(prob["masks"], prob["weights"]).loglik(_dev(g["c5_ll_peel"]), _dev(g["c5_ll_blens"]), jc, grad=True)
    assert ll.item() == pytest.approx(float(g["c5_lnL"]), rel=1e-9)
    assert ll.item() == pytest.approx(float(g["c5_lnL_cython"]), rel=1e-9)
    ref = g["c5_g_blens"]
    np.testing.assert_allclose(gr.cpu().numpy().reshape(-1), ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())


def test_c2_one_sample_end_to_end_against_the_reference_itself():
    """tests/golden/chain_c2.npz: ONE VI sample of the headline workload through the reference's own
    chain -- locations (64 x 5) -> Chyp_torch.get_pdm -> peeler.nj_torch(tau=1e-8) -> GTR P(t) ->
    calculate_treelikelihood over 10 000 patterns -> torch autograd back to the locations.  The fused
    device path must reproduce the topology bit for bit and lnL, branch lengths and d lnL / d locations
    to 1e-9 (K = 1 and the four gamma rates as the reference's summed category axis)."""
    from conftest import golden_problem, load_golden
    from dodonaphy_b200 import synth
    from dodonaphy_b200.engine import SubstModel
    from dodonaphy_b200.pipeline import HotPath

    g = load_golden("chain_c2.npz")
    prob = golden_problem("c2")
    masks, weights = prob["masks"], prob["weights"]
    assert int((masks.astype(np.int64) * (np.arange(masks.shape[1]) % 97 + 1)).sum()) == int(g["mask_dot"])
    assert int(weights.sum()) == int(g["weight_sum"])
    gtr = SubstModel.gtr(g["rates"], g["freqs"])
    x = _dev(g["x"])[None]
    for rates, key_ll, key_g in ((None, "lnL", "g_x"), (g["cat_rates"], "k4_lnL", "k4_g_x")):
        hp = HotPath(masks, weights, gtr, rates=rates, mix=False, tau=1e-8)
        out = hp.value_and_grad(x)
        assert np.array_equal(out["peel"][0].cpu().numpy(), g["peel"])
        np.testing.assert_allclose(out["blens"][0].cpu().numpy(), g["blens"], rtol=1e-9, atol=1e-15)
        assert out["lnL"][0].item() == pytest.approx(float(g[key_ll]), rel=1e-9)
        ref = g[key_g]
        np.testing.assert_allclose(out["grad_x"][0].cpu().numpy(), ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())


def test_mcmc_chain_end_to_end_against_the_reference_itself():
    """tests/golden/chains.npz: the reference's gradient-free chain (Chyp_np.get_pdm -> Cpeeler.nj_np ->
    Cphylo.compute_LL_np) from the locations to lnL at the configs[2] shape (200 taxa x 20 000
    patterns) and the configs[0] shape (DS1: 27 taxa x 934 patterns).  lnL and tree length do not
    depend on how the exactly-tied last joins are ordered: 1e-9 through the whole device chain."""
    from conftest import golden_problem, load_golden
    from dodonaphy_b200 import synth
    from dodonaphy_b200.engine import SubstModel
    from dodonaphy_b200.pipeline import HotPath

    g = load_golden("chains.npz")
    d = load_golden("ds1.npz")
    hp = HotPath(d["masks"], d["weights"], SubstModel.jc69())
    out = hp.value(_dev(np.stack([g[f"c1_{i}_x"] for i in range(4)])))
    for i in range(4):
        assert out["lnL"][i].item() == pytest.approx(float(g[f"c1_{i}_lnL"]), rel=1e-9)
        assert out["blens"][i].sum().item() == pytest.approx(float(g[f"c1_{i}_len"]), rel=1e-9)
    prob = golden_problem("c3")
    _same_alignment(prob["masks"], g["c3_mask_sum"])
    hp = HotPath(prob["masks"], prob["weights"], SubstModel.jc69())
    out = hp.value(_dev(np.stack([g[f"c3_{i}_x"] for i in range(4)])))
    for i in range(4):
        assert out["lnL"][i].item() == pytest.approx(float(g[f"c3_{i}_lnL"]), rel=1e-9)
        assert out["blens"][i].sum().item() == pytest.approx(float(g[f"c3_{i}_len"]), rel=1e-9)


def test_c4_shape_distances_against_the_reference_itself():
    """tests/golden/pdm_c4.npz: 4000 sampled entries (plus row sums and the grand total) of the reference's
    own Chyp_np.get_pdm / Chyp_torch.get_pdm matrices for 2000 seeded taxa (configs[3])."""
    from conftest import golden_problem, load_golden
    from dodonaphy_b200 import engine

    g = load_golden("pdm_c4.npz")
    x = np.random.default_rng(int(g["seed"])).normal(0, 0.05, (2000, 5))
    ii, jj = g["ii"], g["jj"]
    dn = engine.pdm_forward(_dev(x), -1.0, "numpy").cpu().numpy().reshape(2000, 2000)
    dt = engine.pdm_forward(_dev(x), -1.0, "torch").cpu().numpy().reshape(2000, 2000)
    off = ii != jj  # the NumPy variant's diagonal is sqrt(2 eps)-sized rounding noise in the reference
    np.testing.assert_allclose(dn[ii, jj][off], g["np_vals"][off], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(dt[ii, jj], g["torch_vals"], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose((dn.sum(1) - np.diag(dn))[:16], g["np_row_sum"] - 0.0, rtol=1e-9, atol=1e-6)
    assert float(dt.sum()) == pytest.approx(float(g["torch_total"]), rel=1e-10)


def test_c5_shape_alignment_ingestion_properties():
    """configs[4], the step before the path: 500 taxa x 1M sites through the radix-sort kernels (csrc/align.cu;
    phylo.compress_alignment, phylo.py:16-58).  Too large for the host oracle, so size-independent properties: the
    weights add up to the number of sites, the patterns come out strictly increasing in the reference's order
    (lexicographic, taxon 0 most significant), they are exactly the distinct columns the input was tiled from, and
    ingesting the patterns themselves returns them unchanged with weight 1 (idempotence)."""
    from dodonaphy_b200 import _lib

    S, n, n_base = 500, 1_000_000, 20_000
    g = torch.Generator(device="cuda").manual_seed(11)
    letters = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device="cuda")  # A C G T
    base = letters[torch.randint(0, 4, (S, n_base), generator=g, device="cuda")]
    raw = base.repeat(1, n // n_base).contiguous()
    lut = torch.full((256,), 15, dtype=torch.uint8)
    lut[65], lut[67], lut[71], lut[84] = 1, 2, 4, 8
    lut = lut.cuda()
    lib = _lib.load()

    def ingest(cols):
        m = int(cols.shape[1])
        ws = torch.empty(int(lib.dodo_alignment_workspace_bytes(S, m)), dtype=torch.uint8, device="cuda")
        nu = torch.zeros(1, dtype=torch.int32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.dodo_alignment_sort(cols.data_ptr(), S, m, ws.data_ptr(), nu.data_ptr(), st))
        U = int(nu.item())
        masks = torch.empty((S, U), dtype=torch.uint8, device="cuda")
        w = torch.empty(U, dtype=torch.int64, device="cuda")
        _lib.check(lib.dodo_alignment_gather(cols.data_ptr(), S, m, ws.data_ptr(), U, lut.data_ptr(), masks.data_ptr(), w.data_ptr(), st))
        return masks, w

    masks, w = ingest(raw)
    U = int(masks.shape[1])
    assert int(w.sum().item()) == n
    # strictly increasing: at the first taxon where neighbouring patterns differ the left one is smaller
    # (the mask codes 1, 2, 4, 8 keep the order of the bytes A < C < G < T)
    differ = masks[:, 1:] != masks[:, :-1]
    assert bool(differ.any(dim=0).all())
    first = differ.to(torch.uint8).argmax(dim=0)
    left = masks[:, :-1].gather(0, first[None])[0]
    right = masks[:, 1:].gather(0, first[None])[0]
    assert bool((left < right).all())
    # exactly the distinct columns of the tile, each with the right multiplicity
    ref_cols, ref_counts = torch.unique(lut[base.long()].T.contiguous(), dim=0, return_counts=True)
    assert U == int(ref_cols.shape[0])
    assert torch.equal(masks, ref_cols.T.contiguous())
    assert torch.equal(w, ref_counts * (n // n_base))
    # idempotence
    inv = torch.zeros(16, dtype=torch.uint8, device="cuda")
    inv[1], inv[2], inv[4], inv[8] = 65, 67, 71, 84
    again, w1 = ingest(inv[masks.long()].contiguous())
    assert torch.equal(again, masks) and bool((w1 == 1).all())
