"""GPU parity of the geodesic tree connector (peeler.py:14-152) through the reference-shaped API and
the C ABI: against outputs of the reference itself (tests/golden/geodesic.npz), against the oracle on
fresh seeded inputs, and the expectations of the reference's own tests (test/test_peeler.py:22-85,401-470)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_soft_geodesic_matches_reference_outputs_and_gradient():
    from dodonaphy_b200 import peeler

    g = load_golden("geodesic.npz")
    for i in range(int(g["n_soft"])):
        x = torch.tensor(g[f"s{i}_x"], requires_grad=True)
        peel, il, bl = peeler.make_soft_peel_tips(x, connector="geodesics", curvature=-torch.ones(1))
        np.testing.assert_array_equal(peel, g[f"s{i}_peel"])
        np.testing.assert_allclose(il.detach().numpy(), g[f"s{i}_int"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(bl.detach().numpy(), g[f"s{i}_blens"], rtol=1e-9, atol=1e-12)
        assert il.requires_grad and bl.requires_grad  # test/test_peeler.py:466-467
        (gx,) = torch.autograd.grad((bl * torch.tensor(g[f"s{i}_wb"])).sum() + (il * torch.tensor(g[f"s{i}_wi"])).sum(), x)
        scale = np.abs(g[f"s{i}_grad"]).max()
        np.testing.assert_allclose(gx.numpy(), g[f"s{i}_grad"], rtol=1e-8, atol=1e-9 * scale)


def test_hard_geodesic_matches_reference_outputs():
    from dodonaphy_b200 import peeler

    g = load_golden("geodesic.npz")
    for i in range(int(g["n_hard"])):
        peel, il = peeler.make_hard_peel_geodesic(g[f"h{i}_sheet"])
        np.testing.assert_array_equal(peel, g[f"h{i}_peel"])
        np.testing.assert_allclose(il, g[f"h{i}_int"], rtol=1e-9, atol=1e-12)
    # test/test_peeler.py:22-31 (dogbone)
    np.testing.assert_array_equal(peeler.make_hard_peel_geodesic(g["h0_sheet"])[0], [[1, 0, 4], [3, 2, 5], [4, 5, 6]])


def test_batched_geodesic_against_oracle():
    """A batch of seeded embeddings through the C ABI in one launch vs the oracle tree by tree."""
    from dodonaphy_b200 import engine
    from oracle import geodesic as G

    rng = np.random.default_rng(91)
    B, S, D = 5, 17, 4
    x = rng.normal(0, 0.1, (B, S, D))
    out = engine.geo_soft(torch.tensor(x))
    wb = rng.normal(size=(B, 2 * S - 2))
    gx = engine.geo_soft_backward(out["peel"], out["slots"], out["tape"], torch.tensor(wb).cuda())
    sheet = np.concatenate((np.sqrt((x * x).sum(-1, keepdims=True) + 1), x), -1)
    hp, hil = engine.geo_hard(torch.tensor(sheet))
    assert int(out["flags"].max()) == 0
    for b in range(B):
        xt = torch.tensor(x[b], requires_grad=True)
        peel, il, bl = G.soft_peel_tips_torch(xt)
        np.testing.assert_array_equal(out["peel"][b].cpu().numpy(), peel)
        np.testing.assert_allclose(out["int_locs"][b].cpu().numpy(), il.detach().numpy(), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(out["blens"][b].cpu().numpy(), bl.detach().numpy(), rtol=1e-9, atol=1e-12)
        (gref,) = torch.autograd.grad((bl * torch.tensor(wb[b])).sum(), xt)
        np.testing.assert_allclose(gx[b].cpu().numpy(), gref.numpy(), rtol=1e-8, atol=1e-9 * float(gref.abs().max()))
        p2, il2 = G.hard_peel_geodesic(sheet[b])
        np.testing.assert_array_equal(hp[b].cpu().numpy(), p2)
        np.testing.assert_allclose(hil[b].cpu().numpy(), il2, rtol=1e-9, atol=1e-12)


def test_geodesic_argument_errors():
    from dodonaphy_b200 import peeler

    x = torch.zeros((4, 2), dtype=torch.float64)
    with pytest.raises(NotImplementedError):
        peeler.make_soft_peel_tips(x, connector="nj")
    with pytest.raises(NotImplementedError):
        peeler.make_soft_peel_tips(x, curvature=-2 * torch.ones(1))
    with pytest.raises(ValueError):
        peeler.make_soft_peel_tips(torch.zeros((4, 16), dtype=torch.float64))  # D > 15


def _tips(masks):
    from oracle.phylo import mask_to_partial

    return [torch.tensor(mask_to_partial(m)) for m in masks]


def test_pipeline_with_geodesic_connector_against_oracle():
    """locations -> geodesic connector -> lnL, and d lnL / d locations, for a batch (vi.py:479-482 +
    base_model.py:219-247 + backward) vs the oracle composition under torch autograd."""
    from dodonaphy_b200 import synth
    from dodonaphy_b200.engine import SubstModel
    from dodonaphy_b200.pipeline import HotPath
    from oracle import geodesic as G
    from oracle.phylo import jc69_p_t_torch, treelikelihood_torch

    S, D, L, B = 9, 3, 80, 4
    prob = synth.make_problem(S, L, B, D=D, seed=92)
    x = np.random.default_rng(93).normal(0, 0.2, (B, S, D))
    hp = HotPath(prob["masks"], prob["weights"], SubstModel.jc69(), connector="geodesics")
    out = hp.value_and_grad(torch.tensor(x).cuda())
    tips, w = _tips(prob["masks"]), torch.tensor(prob["weights"], dtype=torch.float64)
    for b in range(B):
        xt = torch.tensor(x[b], requires_grad=True)
        peel, _, bl = G.soft_peel_tips_torch(xt)
        ll = treelikelihood_torch(tips, w, peel, jc69_p_t_torch(bl), torch.full((4,), 0.25, dtype=torch.float64))
        (gref,) = torch.autograd.grad(ll, xt)
        np.testing.assert_array_equal(out["peel"][b].cpu().numpy(), peel)
        np.testing.assert_allclose(float(out["lnL"][b]), float(ll.detach()), rtol=1e-10)
        np.testing.assert_allclose(out["grad_x"][b].cpu().numpy(), gref.numpy(), rtol=1e-8, atol=1e-9 * float(gref.abs().max()))


def test_callers_with_geodesic_connector():
    """HMAP (fused pass == call-by-call route), VI (one batched ELBO step) and the MCMC-side
    ``connect`` (chain.py:176-185) with connector='geodesics'."""
    from dodonaphy_b200 import synth
    from dodonaphy_b200.base_model import BaseModel
    from dodonaphy_b200.hmap import HMAP
    from dodonaphy_b200.vi import DodonaphyVI

    S, D, L = 8, 3, 60
    prob = synth.make_problem(S, L, 1, D=D, seed=94)
    tips, w = _tips(prob["masks"]), torch.tensor(prob["weights"])
    x0 = np.random.default_rng(95).normal(0, 0.2, (S, D))
    m = HMAP(tips, w, dim=D, soft_temp=1e-8, prior="gammadir", connector="geodesics")
    m.init_embedding_params(x0)
    res = {}
    for fused in (True, False):
        m.params_optim["leaf_loc"].grad = None
        loss = m.compute_loss(fused=fused)
        loss.sum().backward()
        res[fused] = (float(loss.detach()), m.params_optim["leaf_loc"].grad.clone(), np.array(m.peel))
    np.testing.assert_allclose(res[True][0], res[False][0], rtol=1e-11)
    np.testing.assert_array_equal(res[True][2], res[False][2])
    np.testing.assert_allclose(res[True][1].numpy(), res[False][1].numpy(), rtol=1e-8, atol=1e-9)
    hist = m.learn(epochs=5, learn_rate=0.01)
    assert len(hist) == 6 and np.isfinite(hist).all()

    torch.manual_seed(4)
    vi = DodonaphyVI(tips, w, dim=D, connector="geodesics")
    vi.set_params_optim(x0, torch.full((S, D), -8.0, dtype=torch.float64))
    elbo = vi.elbo_siwae(importance=3)
    (-elbo).backward()
    assert torch.isfinite(elbo) and torch.isfinite(vi.params_optim["leaf_mu"].grad).all()
    assert vi.params_optim["leaf_mu"].grad.abs().sum() > 0

    g = load_golden("geodesic.npz")
    S2 = g["chain_x"].shape[0]
    prob2 = synth.make_problem(S2, 30, 1, D=3, seed=96)
    chain = BaseModel("mcmc", _tips(prob2["masks"]), torch.tensor(prob2["weights"]), dim=3, soft_temp=None,
                      connector="geodesics", require_grad=False)
    peel, blens, int_x = chain.connect(g["chain_x"])
    np.testing.assert_array_equal(peel, g["chain_peel"])
    np.testing.assert_allclose(int_x, g["chain_int"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(blens, g["chain_blens"], rtol=1e-9, atol=1e-12)
    assert np.isfinite(chain.compute_LL(peel, blens))
