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
