"""Generate the golden fixtures in this directory by RUNNING THE REAL REFERENCE.

Run in the build container only (needs /root/reference and oracle/_ref, see oracle/build_ref.py):

    python tests/golden/make_golden.py

Outputs (committed; small):
  ds1.npz        DS1 alignment (27 taxa, 934 patterns) as tip masks + weights, the MrBayes tree and
                 the NJ tree of the reference's own tests as (peel, blens), and the reference's
                 lnL / gradients on them (pins -6904.632, test/test_phylo.py:214-273, and
                 -7130.96813417, test/test_bito.py:35).
  pdm.npz        Chyp_torch.get_pdm / Chyp_np.get_pdm outputs + torch-autograd gradients.
  nj_hard.npz    Cpeeler.nj_np outputs.
  nj_soft.npz    peeler.nj_torch outputs + autograd gradient of a random linear functional of blens.
  nj_soft_c2.npz the same for ONE 64-taxon tree (the headline shape; ~19 GB and minutes in the reference).
  prune_c2.npz   lnL + gradients of the reference at the full headline size (64 taxa x 10 000 patterns).
  chain_c2.npz   one headline-size VI sample end to end in the reference: locations -> distances -> soft NJ
                 -> GTR P(t) -> lnL, and autograd back to the locations.
  blens_jac.npz  the reference's autograd Jacobian of the decoded branch lengths w.r.t. the locations (vi.py:376-387).
  vi.npz         the reference's own DodonaphyVI object: elbo_siwae, its per-sample terms and gradients (vi.py:315-341).
  laplace.npz    the reference's own Laplace object: log posterior and its autograd Hessian (laplace.py:9-25,40).
  hmap.npz       the reference's own HMAP object: MAP loss, its parts and autograd gradients (hmap.py:185-192).
  chains.npz     the reference's gradient-free MCMC chain end to end (locations -> lnL) at the configs[0] and [2] shapes.
  pdm_c4.npz     sampled entries of the reference's 2000-taxon distance matrices (NumPy and torch variants).
  configs.npz    the reference at the shapes of configs[2..4]: 200- and 500-taxon trees + lnL, one 2000-taxon NJ tree.
  prune.npz      phylo.calculate_treelikelihood lnL + autograd gradients (JC69 and GTR).
  pdm_jac.npz    Chyp_np.get_pdm(get_jacobian=True) pair Jacobians.
  charts.npz     tangent-space / sheet / ball chart changes of Chyp_torch (t02p, p2t0, ...).
  geodesic.npz   peeler.make_soft_peel_tips (+ autograd gradient) and peeler.make_hard_peel_geodesic.

The tests never read /root/reference; they read these files.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refload  # noqa: E402
from dodonaphy_b200 import newick, synth  # noqa: E402

torch.set_default_dtype(torch.float64)
REF = "/root/reference"


class _TN:
    def __init__(self, labels):
        self._l = labels

    def labels(self):
        return self._l


class _Aln:
    def __init__(self, labels, seqs):
        self.taxon_namespace = _TN(labels)
        self._s = seqs

    def sequences(self):
        return self._s


def masks_from_partials(parts):
    out = []
    for p in parts:
        a = p.numpy()
        out.append((a[0] + 2 * a[1] + 4 * a[2] + 8 * a[3]).astype(np.uint8))
    return np.stack(out)


def ref_partials(r, masks):
    from oracle.phylo import mask_to_partial

    S, L = masks.shape
    parts = [torch.tensor(mask_to_partial(m)) for m in masks]
    parts += [torch.zeros((1, 4, L)) for _ in range(S - 1)]
    return parts


def ref_ll_and_grads(r, masks, weights, peel, blens, model, rates=None, freqs=None):
    """lnL + autograd gradients through the reference's own functions."""
    b = torch.tensor(blens, requires_grad=True)
    pm = r.phylomodel.PhyloModel(model)
    if model == "GTR":
        # set the constrained parameters directly on the transforms' inverse (phylomodel.py:23-45)
        pm._freqs = pm.freqs_transform.inv(torch.tensor(freqs)).clone().detach().requires_grad_(True)
        pm._sub_rates = pm.sub_rates_transform.inv(torch.tensor(rates)).clone().detach().requires_grad_(True)
    mats = pm.get_transition_mats(b)
    mats.retain_grad()
    fr = pm.freqs
    if model == "GTR":
        fr.retain_grad()
    ll = r.phylo.calculate_treelikelihood(ref_partials(r, masks), torch.tensor(weights), peel, mats, fr)
    ll.backward()
    out = dict(lnL=ll.item(), g_blens=b.grad.numpy().copy(), g_mats=mats.grad.numpy().copy(),
               mats=mats.detach().numpy().copy())
    if model == "GTR":
        out["g_raw_freqs"] = pm._freqs.grad.numpy().copy()
        out["g_raw_rates"] = pm._sub_rates.grad.numpy().copy()
        out["raw_freqs"] = pm._freqs.detach().numpy().copy()
        out["raw_rates"] = pm._sub_rates.detach().numpy().copy()
    return out


def make_ds1(r):
    labels, seqs = newick.read_nexus_alignment(f"{REF}/test/data/ds1/dna.nex")
    parts, w = r.phylo.compress_alignment(_Aln(labels, seqs))
    masks = masks_from_partials(parts)
    weights = w.numpy().astype(np.int64)
    out = dict(masks=masks, weights=weights)
    nwk, tr = newick.read_nexus_tree(f"{REF}/test/data/ds1/mb_tree300000.nex")
    trees = {"mb": newick.newick_to_peel(nwk, labels, tr),
             "nj": newick.newick_to_peel(open(f"{REF}/test/data/ds1/dna.nj.newick").read(), labels)}
    rng = np.random.default_rng(20)
    rates, freqs = synth.gtr_params(rng)
    out["gtr_rates"], out["gtr_freqs"] = rates, freqs
    for name, (peel, blens) in trees.items():
        out[f"{name}_peel"], out[f"{name}_blens"] = peel, blens
        pn = [p.numpy() for p in parts] + [np.zeros((1, 4, masks.shape[1])) for _ in range(26)]
        out[f"{name}_lnL_cython"] = float(r.Cphylo.compute_LL_np(pn, weights, peel, blens))
        jc = ref_ll_and_grads(r, masks, weights, peel, blens, "JC69")
        out[f"{name}_lnL_torch"] = jc["lnL"]
        out[f"{name}_g_blens"] = jc["g_blens"]
        g = ref_ll_and_grads(r, masks, weights, peel, blens, "GTR", rates, freqs)
        for k, v in g.items():
            out[f"{name}_gtr_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "ds1.npz"), **out)
    print("ds1:", out["mb_lnL_torch"], out["nj_lnL_torch"], out["mb_gtr_lnL"])


def make_pdm(r):
    rng = np.random.default_rng(31)
    out = {}
    cases = [(5, 2, -1.0, False), (12, 3, -1.0, False), (27, 3, -0.5, False), (64, 5, -1.0, False),
             (17, 4, -2.3, True), (64, 5, -1.0, True), (130, 5, -1.0, False)]
    out["n"] = len(cases)
    for i, (S, D, curv, mats) in enumerate(cases):
        x = rng.normal(0, 0.05 if i % 2 == 0 else 0.3, (S, D))
        if i == 1:
            x[3] = x[2]  # coincident points: clamp + zero-gradient path
        C = rng.normal(size=(S, S))
        out[f"{i}_x"], out[f"{i}_curv"], out[f"{i}_matsumoto"], out[f"{i}_C"] = x, curv, mats, C
        for proj in ("up", "wrap"):
            xt = torch.tensor(x, requires_grad=True)
            ct = torch.tensor([curv], requires_grad=True)
            Dm = r.Chyp_torch.get_pdm(xt, curvature=ct, matsumoto=mats, projection=proj)
            (Dm * torch.tensor(C)).sum().backward()
            out[f"{i}_pdm_torch_{proj}"] = Dm.detach().numpy().copy()
            out[f"{i}_gx_{proj}"] = xt.grad.numpy().copy()
            out[f"{i}_gc_{proj}"] = ct.grad.numpy().copy()
        out[f"{i}_pdm_np"] = r.Chyp_np.get_pdm(x, curvature=curv, matsumoto=mats)
    np.savez_compressed(os.path.join(HERE, "pdm.npz"), **out)
    print("pdm cases:", len(cases))


def make_nj_hard(r):
    rng = np.random.default_rng(32)
    out = {}
    sizes = [2, 3, 4, 5, 8, 13, 27, 40, 64, 100]
    out["n"] = len(sizes) + 2
    for i, S in enumerate(sizes):
        x = rng.normal(0, 0.05, (S, 3 if S < 30 else 5))
        pdm = np.ascontiguousarray(r.Chyp_np.get_pdm(x))
        peel, blens = r.Cpeeler.nj_np(pdm)
        out[f"{i}_pdm"], out[f"{i}_peel"], out[f"{i}_blens"] = pdm, np.asarray(peel, dtype=np.int64), blens
    # exact-tie inputs: equidistant taxa, and an integer matrix (test/test_peeler.py:228-303)
    i = len(sizes)
    pdm = np.ones((6, 6)) - np.eye(6)
    peel, blens = r.Cpeeler.nj_np(np.ascontiguousarray(pdm))
    out[f"{i}_pdm"], out[f"{i}_peel"], out[f"{i}_blens"] = pdm, np.asarray(peel, dtype=np.int64), blens
    i += 1
    pdm = np.array([[0, 5, 9, 9, 8], [5, 0, 10, 10, 9], [9, 10, 0, 8, 7], [9, 10, 8, 0, 3], [8, 9, 7, 3, 0]], dtype=np.float64)
    peel, blens = r.Cpeeler.nj_np(np.ascontiguousarray(pdm))
    out[f"{i}_pdm"], out[f"{i}_peel"], out[f"{i}_blens"] = pdm, np.asarray(peel, dtype=np.int64), blens
    np.savez_compressed(os.path.join(HERE, "nj_hard.npz"), **out)
    print("nj_hard cases:", out["n"])


def make_nj_soft(r):
    rng = np.random.default_rng(33)
    out = {}
    cases = [(3, 1e-8), (4, 1e-8), (5, 1e-8), (8, 1e-8), (12, 1e-8), (12, 1e-4), (17, 1e-8), (21, 1e-8), (27, 1e-8)]
    n = 0
    for S, tau in cases:
        x = rng.normal(0, 0.05, (S, 3))
        pdm = r.Chyp_torch.get_pdm(torch.tensor(x)).detach().clone().requires_grad_(True)
        peel, blens = r.peeler.nj_torch(pdm, tau=tau)
        c = rng.normal(size=2 * S - 2)
        (blens * torch.tensor(c)).sum().backward()
        out[f"{n}_pdm"], out[f"{n}_tau"], out[f"{n}_peel"] = pdm.detach().numpy().copy(), tau, np.asarray(peel, dtype=np.int64)
        out[f"{n}_blens"], out[f"{n}_c"], out[f"{n}_g_pdm"] = blens.detach().numpy().copy(), c, pdm.grad.numpy().copy()
        n += 1
    # all-ties case of test/test_peeler.py:343-348 and the 17-taxon matrix of test/test_peel_data.txt
    pdm = torch.tensor(np.ones((6, 6)) - np.eye(6), requires_grad=True)
    peel, blens = r.peeler.nj_torch(pdm, tau=1e-8)
    c = rng.normal(size=10)
    (blens * torch.tensor(c)).sum().backward()
    out[f"{n}_pdm"], out[f"{n}_tau"], out[f"{n}_peel"] = pdm.detach().numpy().copy(), 1e-8, np.asarray(peel, dtype=np.int64)
    out[f"{n}_blens"], out[f"{n}_c"], out[f"{n}_g_pdm"] = blens.detach().numpy().copy(), c, pdm.grad.numpy().copy()
    n += 1
    d1 = np.genfromtxt(f"{REF}/test/test_peel_data.txt", dtype=np.double, delimiter=", ")
    til = np.tril_indices(17, -1)
    d2 = np.zeros((17, 17))
    d2[til[0], til[1]] = d1
    d2[til[1], til[0]] = d1
    for tau in (1e-8, 1e-20):
        pdm = torch.tensor(d2, requires_grad=True)
        peel, blens = r.peeler.nj_torch(pdm, tau=tau)
        c = rng.normal(size=2 * 17 - 2)
        (blens * torch.tensor(c)).sum().backward()
        out[f"{n}_pdm"], out[f"{n}_tau"], out[f"{n}_peel"] = pdm.detach().numpy().copy(), tau, np.asarray(peel, dtype=np.int64)
        out[f"{n}_blens"], out[f"{n}_c"], out[f"{n}_g_pdm"] = blens.detach().numpy().copy(), c, pdm.grad.numpy().copy()
        n += 1
    out["n"] = n
    np.savez_compressed(os.path.join(HERE, "nj_soft.npz"), **out)
    print("nj_soft cases:", n)


def make_nj_soft_c2(r):
    """ONE 64-taxon tree through the reference's own soft NJ (peeler.nj_torch, tau = 1e-8) at the
    shape the headline metric is quoted on (BASELINE.json configs[1]: 64 taxa, d5; sample 0 of the
    bench's synthetic problem).  The reference needs ~19 GB and minutes for this single tree
    (SURVEY.md Appendix B), so it is not part of the default run: python make_golden.py nj_soft_c2"""
    prob = synth.make_problem(64, 8, 2, D=5, seed=1002)
    rng = np.random.default_rng(64)
    x = prob["xs"][0]
    pdm = r.Chyp_torch.get_pdm(torch.tensor(x)).detach().clone().requires_grad_(True)
    peel, blens = r.peeler.nj_torch(pdm, tau=1e-8)
    c = rng.normal(size=2 * 64 - 2)
    (blens * torch.tensor(c)).sum().backward()
    np.savez_compressed(os.path.join(HERE, "nj_soft_c2.npz"), x=x, pdm=pdm.detach().numpy().copy(), tau=1e-8,
                        peel=np.asarray(peel, dtype=np.int64), blens=blens.detach().numpy().copy(), c=c,
                        g_pdm=pdm.grad.numpy().copy())
    print("nj_soft_c2: peel tail", np.asarray(peel)[-3:].tolist())


def make_prune(r):
    out = {}
    cases = [(4, 40, 41), (8, 120, 42), (27, 300, 43), (64, 500, 44), (130, 200, 45)]
    n = 0
    for S, L, seed in cases:
        prob = synth.make_problem(S, L, B=1, D=5, seed=seed)
        pdm = np.ascontiguousarray(r.Chyp_np.get_pdm(prob["xs"][0]))
        peel, blens = r.Cpeeler.nj_np(pdm)
        peel = np.asarray(peel, dtype=np.int64)
        out[f"{n}_masks"], out[f"{n}_weights"], out[f"{n}_peel"], out[f"{n}_blens"] = prob["masks"], prob["weights"], peel, blens
        out[f"{n}_rates"], out[f"{n}_freqs"] = prob["rates"], prob["freqs"]
        jc = ref_ll_and_grads(r, prob["masks"], prob["weights"], peel, blens, "JC69")
        for k, v in jc.items():
            out[f"{n}_jc_{k}"] = v
        g = ref_ll_and_grads(r, prob["masks"], prob["weights"], peel, blens, "GTR", prob["rates"], prob["freqs"])
        for k, v in g.items():
            out[f"{n}_gtr_{k}"] = v
        # MCMC twin
        pn = [p.numpy() for p in ref_partials(r, prob["masks"])]
        out[f"{n}_lnL_cython"] = float(r.Cphylo.compute_LL_np(pn, prob["weights"], peel, blens))
        n += 1
    out["n"] = n
    np.savez_compressed(os.path.join(HERE, "prune.npz"), **out)
    print("prune cases:", n)


def make_chain_c2(r):
    """The whole differentiable chain of ONE headline-size VI sample through the reference:
    locations (64 x 5) -> Chyp_torch.get_pdm -> peeler.nj_torch(tau = 1e-8) -> GTR P(t) ->
    phylo.calculate_treelikelihood over 10 000 patterns -> torch autograd back to the locations
    (vi.py:467-491 / base_model.py:219-247).  K = 1 and the four gamma rates as the reference's summed
    category axis.  ~19 GB and minutes (the soft NJ); python make_golden.py chain_c2"""
    from dodonaphy_b200.engine import SubstModel

    prob = synth.make_problem(64, 10000, 128, D=5, seed=1002)
    masks, weights = prob["masks"], prob["weights"]
    x = torch.tensor(prob["xs"][0], requires_grad=True)
    pdm = r.Chyp_torch.get_pdm(x)
    peel, blens = r.peeler.nj_torch(pdm, tau=1e-8)
    pm = r.phylomodel.PhyloModel("GTR")
    pm._freqs = pm.freqs_transform.inv(torch.tensor(prob["freqs"])).clone().detach().requires_grad_(True)
    pm._sub_rates = pm.sub_rates_transform.inv(torch.tensor(prob["rates"])).clone().detach().requires_grad_(True)
    parts, w = ref_partials(r, masks), torch.tensor(weights)
    ll = r.phylo.calculate_treelikelihood(parts, w, peel, pm.get_transition_mats(blens), pm.freqs)
    ll.backward(retain_graph=True)
    out = dict(x=prob["xs"][0], peel=np.asarray(peel, dtype=np.int64), blens=blens.detach().numpy().copy(), lnL=ll.item(),
               g_x=x.grad.numpy().copy(), g_raw_rates=pm._sub_rates.grad.numpy().copy(), g_raw_freqs=pm._freqs.grad.numpy().copy(),
               rates=prob["rates"], freqs=prob["freqs"], weight_sum=int(weights.sum()),
               mask_dot=int((masks.astype(np.int64) * (np.arange(masks.shape[1]) % 97 + 1)).sum()))
    x.grad = None
    cat = SubstModel.discrete_gamma(0.5, 4)
    ll4 = r.phylo.calculate_treelikelihood(ref_partials(r, masks), w, peel,
                                           pm.get_transition_mats(blens[:, None] * torch.tensor(cat)[None, :]), pm.freqs)
    ll4.backward()
    out["cat_rates"], out["k4_lnL"], out["k4_g_x"] = np.asarray(cat), ll4.item(), x.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "chain_c2.npz"), **out)
    print("chain_c2: lnL", out["lnL"], out["k4_lnL"], "|g_x|max", np.abs(out["g_x"]).max())


def make_chains(r):
    """The gradient-free MCMC chain of the reference end to end (chain.py:145-169: Chyp_np.get_pdm ->
    Cpeeler.nj_np -> Cphylo.compute_LL_np) for four chains at the configs[2] shape (200 taxa x 20 000
    patterns) and four at the configs[0] shape (the DS1 alignment, 27 taxa, d3).  Stored: locations,
    lnL and tree length -- both are invariant to how the last exactly-tied joins are ordered, which
    depends on the last bit of the distances.  python make_golden.py chains"""
    out = {}
    prob = synth.make_problem(200, 20000, 64, D=5, seed=1003)
    out["c3_mask_sum"] = int(prob["masks"].astype(np.int64).sum())
    pn = [p.numpy() for p in ref_partials(r, prob["masks"])]
    for i in range(4):
        x = prob["xs"][i]
        peel, blens = r.Cpeeler.nj_np(np.ascontiguousarray(r.Chyp_np.get_pdm(x)))
        out[f"c3_{i}_x"], out[f"c3_{i}_lnL"] = x, float(r.Cphylo.compute_LL_np(pn, prob["weights"], np.asarray(peel), blens))
        out[f"c3_{i}_len"] = float(np.sum(blens))
    d = np.load(os.path.join(HERE, "ds1.npz"))
    pn = [p.numpy() for p in ref_partials(r, d["masks"])]
    rng = np.random.default_rng(27)
    for i in range(4):
        x = rng.normal(0, 0.05, (27, 3))
        peel, blens = r.Cpeeler.nj_np(np.ascontiguousarray(r.Chyp_np.get_pdm(x)))
        out[f"c1_{i}_x"], out[f"c1_{i}_lnL"] = x, float(r.Cphylo.compute_LL_np(pn, d["weights"], np.asarray(peel), blens))
        out[f"c1_{i}_len"] = float(np.sum(blens))
    np.savez_compressed(os.path.join(HERE, "chains.npz"), **out)
    print("chains: c3 lnL", [out[f"c3_{i}_lnL"] for i in range(4)], "c1 lnL", [out[f"c1_{i}_lnL"] for i in range(4)])


def make_blens_jac(r):
    """vi.py:376-387 with the reference's own functions: torch.autograd.functional.jacobian of the decoded
    branch lengths w.r.t. the flattened locations (Chyp_torch.get_pdm -> peeler.nj_torch(tau = 1e-8)) and
    log sqrt|det(J J^T)|, for 7, 12 and 17 taxa.  python make_golden.py blens_jac"""
    rng = np.random.default_rng(77)
    out = {"n": 3}
    for i, (S, D) in enumerate([(7, 3), (12, 2), (17, 5)]):
        x = rng.normal(0, 0.08, (S, D))

        def get_blens(flat):
            return r.peeler.nj_torch(r.Chyp_torch.get_pdm(flat.reshape(S, D)), tau=1e-8)[1]

        J = torch.autograd.functional.jacobian(get_blens, torch.tensor(x).flatten())
        out[f"{i}_x"], out[f"{i}_J"] = x, J.numpy().copy()
        out[f"{i}_logdet"] = float(torch.log(torch.sqrt(torch.abs(torch.linalg.det(J @ J.T)))))
    np.savez_compressed(os.path.join(HERE, "blens_jac.npz"), **out)
    print("blens_jac: log|det| terms", [out[f"{i}_logdet"] for i in range(3)])


def make_hmap(r):
    """The reference's own HMAP object (hmap.py:20-64,185-192,310-334): MAP loss = -(lnL + gamma-Dirichlet
    prior) at given tip locations, with torch autograd to the locations and the curvature leaf.
    Cases: 8 taxa x 60 patterns and the DS1 alignment (27 taxa), "up" and "wrap" embedders.
    python make_golden.py hmap"""
    import importlib

    hm = importlib.import_module("dodonaphy.hmap")
    d = np.load(os.path.join(HERE, "ds1.npz"))
    prob = synth.make_problem(8, 60, 1, D=3, seed=77)
    rng = np.random.default_rng(99)
    cases = [(prob["masks"], prob["weights"], prob["xs"][0], "up", "nj"), (prob["masks"], prob["weights"], prob["xs"][0], "wrap", "nj"),
             (d["masks"], d["weights"], rng.normal(0, 0.05, (27, 3)), "up", "nj"),
             (prob["masks"], prob["weights"], prob["xs"][0] * 4.0, "up", "geodesics")]  # ball points of norm ~0.3
    out = {"n": len(cases)}
    for i, (masks, weights, x, emb, conn) in enumerate(cases):
        parts = ref_partials(r, masks)[:masks.shape[0]]  # tips only: BaseModel allocates the internal nodes itself
        m = hm.HMAP(parts, torch.tensor(weights), dim=x.shape[1], soft_temp=1e-8, loss_fn="likelihood", path_write=None,
                    prior="gammadir", embedder=emb, connector=conn)
        m.params_optim["leaf_loc"] = torch.tensor(x, requires_grad=True)
        out[f"{i}_connector"] = conn
        loss = m.compute_loss()
        loss.backward()
        out[f"{i}_masks"], out[f"{i}_weights"], out[f"{i}_x"], out[f"{i}_embedder"] = masks, weights, x, emb
        out[f"{i}_loss"], out[f"{i}_ln_p"], out[f"{i}_ln_prior"] = loss.item(), m.ln_p.item(), m.ln_prior.item()
        out[f"{i}_g_x"] = m.params_optim["leaf_loc"].grad.numpy().copy()
        gc = m.params_optim["curvature"].grad  # None where the connector does not use the curvature leaf
        out[f"{i}_g_curv"] = float(gc) if gc is not None else float("nan")
        out[f"{i}_peel"] = np.asarray(m.peel, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "hmap.npz"), **out)
    print("hmap: losses", [out[f"{i}_loss"] for i in range(len(cases))], "g_curv", [out[f"{i}_g_curv"] for i in range(len(cases))])


def make_vi(r):
    """The reference's own DodonaphyVI object (vi.py:161-202,315-341): elbo_siwae over two importance
    samples, its per-sample terms (ln p, ln prior, ln q; the log-Jacobian is -inf -> 0 for both, vi.py:199-201)
    and autograd gradients w.r.t. leaf_mu / leaf_sigma.  The sampled locations are stored, so the test
    feeds the same draws instead of relying on the generator's stream.  python make_golden.py vi"""
    import importlib
    import warnings

    vim = importlib.import_module("dodonaphy.vi")
    S, D, L, imp = 8, 3, 60, 2
    prob = synth.make_problem(S, L, 1, D=D, seed=55)
    vi = vim.DodonaphyVI(ref_partials(r, prob["masks"])[:S], torch.tensor(prob["weights"]), D, embedder="up", connector="nj",
                         soft_temp=1e-8, path_write=None)
    mu, sg = torch.tensor(prob["x"]), torch.full((S, D), -9.0)
    vi.set_params_optim({"leaf_mu": mu.clone(), "leaf_sigma": sg.clone()})
    rec, orig = [], vi.rsample_tree

    def wrapped(*a, **k):
        rec.append(orig(*a, **k))
        return rec[-1]

    vi.rsample_tree = wrapped
    drawn, orig_draw = [], vi.rsample_Euclid  # the sample dict holds the MEAN under "leaf_locs": record the draws

    def draw(*a, **k):
        loc, log_q = orig_draw(*a, **k)
        drawn.append(loc.detach().numpy().copy())
        return loc, log_q

    vi.rsample_Euclid = draw
    torch.manual_seed(3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        elbo = vi.elbo_siwae(importance=imp)
    elbo.backward()
    assert all(float(t["jacobian"]) == 0.0 for t in rec)
    out = dict(masks=prob["masks"], weights=prob["weights"], mu=mu.numpy(), sigma=sg.numpy(), elbo=elbo.item(),
               x=np.stack(drawn),
               ln_p=np.array([float(t["ln_p"]) for t in rec]), ln_prior=np.array([float(t["ln_prior"]) for t in rec]),
               log_q=np.array([float(t["logQ"]) for t in rec]),
               g_mu=vi.params_optim["leaf_mu"].grad.numpy().copy(), g_sigma=vi.params_optim["leaf_sigma"].grad.numpy().copy())
    np.savez_compressed(os.path.join(HERE, "vi.npz"), **out)
    print("vi: elbo", out["elbo"], "terms", out["ln_p"], out["ln_prior"], out["log_q"])


def make_laplace(r):
    """The reference's own Laplace object (laplace.py:9-25,40): log posterior at the flattened tip
    locations and its torch.autograd.functional.hessian, 6 taxa x 50 patterns, d2.
    python make_golden.py laplace"""
    import importlib

    lp = importlib.import_module("dodonaphy.laplace")
    S, D, L = 6, 2, 50
    prob = synth.make_problem(S, L, 1, D=D, seed=79)
    m = lp.Laplace(ref_partials(r, prob["masks"])[:S], torch.tensor(prob["weights"]), dim=D, soft_temp=1e-8, loss_fn="likelihood",
                   path_write=None, prior="gammadir", embedder="up", connector="nj")
    x0 = torch.tensor(prob["xs"][0]).flatten()
    val = m.get_ln_posterior(x0)
    H = torch.autograd.functional.hessian(m.get_ln_posterior, x0, vectorize=True)
    np.savez_compressed(os.path.join(HERE, "laplace.npz"), masks=prob["masks"], weights=prob["weights"], x=prob["xs"][0],
                        ln_posterior=float(val), hessian=H.numpy().copy())
    print("laplace: ln posterior", float(val), "|H|max", float(H.abs().max()))


def make_pdm_c4(r):
    """Chyp_np.get_pdm and Chyp_torch.get_pdm of the reference at the configs[3] size (2000 taxa, d5, seeded
    locations): 4000 sampled entries of each 2000 x 2000 matrix (the full matrices are 32 MB).
    python make_golden.py pdm_c4"""
    rng = np.random.default_rng(2000)
    x = rng.normal(0, 0.05, (2000, 5))
    ii, jj = rng.integers(0, 2000, 4000), rng.integers(0, 2000, 4000)
    dn = r.Chyp_np.get_pdm(x)
    dt = r.Chyp_torch.get_pdm(torch.tensor(x)).numpy()
    np.savez_compressed(os.path.join(HERE, "pdm_c4.npz"), seed=2000, ii=ii, jj=jj, np_vals=dn[ii, jj], torch_vals=dt[ii, jj],
                        np_row_sum=dn.sum(1)[:16], torch_total=float(dt.sum()))
    print("pdm_c4: mean distance", float(dt.mean()))


def make_prune_c2(r):
    """The reference's lnL and autograd gradients at the FULL size of the headline workload (64 taxa x
    10 000 patterns, sample 0 of bench.py's synthetic problem, seed 1002): JC69, GTR, and GTR with the
    four discrete-gamma rates as a category axis (the reference's broadcasting then sums the
    per-category lnL, phylo.py:137-139).  Only the answers are stored: the alignment is regenerated
    from the seed by the test (its checksum is stored too).  python make_golden.py prune_c2"""
    from dodonaphy_b200.engine import SubstModel

    prob = synth.make_problem(64, 10000, 128, D=5, seed=1002)
    masks, weights = prob["masks"], prob["weights"]
    pdm = np.ascontiguousarray(r.Chyp_np.get_pdm(prob["xs"][0]))
    peel, blens = r.Cpeeler.nj_np(pdm)
    peel = np.asarray(peel, dtype=np.int64)
    out = dict(peel=peel, blens=np.asarray(blens), rates=prob["rates"], freqs=prob["freqs"],
               mask_sum=int(masks.astype(np.int64).sum()), mask_dot=int((masks.astype(np.int64) * (np.arange(masks.shape[1]) % 97 + 1)).sum()),
               weight_sum=int(weights.sum()))
    for name, model in (("jc", "JC69"), ("gtr", "GTR")):
        res = ref_ll_and_grads(r, masks, weights, peel, blens, model, prob["rates"], prob["freqs"])
        out[f"{name}_lnL"], out[f"{name}_g_blens"] = res["lnL"], res["g_blens"]
        if model == "GTR":
            out["gtr_g_raw_rates"], out["gtr_g_raw_freqs"] = res["g_raw_rates"], res["g_raw_freqs"]
            out["gtr_raw_rates"], out["gtr_raw_freqs"] = res["raw_rates"], res["raw_freqs"]
    cat = SubstModel.discrete_gamma(0.5, 4)
    b = torch.tensor(blens, requires_grad=True)
    pm = r.phylomodel.PhyloModel("GTR")
    pm._freqs = pm.freqs_transform.inv(torch.tensor(prob["freqs"])).clone().detach()
    pm._sub_rates = pm.sub_rates_transform.inv(torch.tensor(prob["rates"])).clone().detach()
    mats = pm.get_transition_mats(b[:, None] * torch.tensor(cat)[None, :])  # (E, K, 4, 4)
    ll = r.phylo.calculate_treelikelihood(ref_partials(r, masks), torch.tensor(weights), peel, mats, pm.freqs)
    ll.backward()
    out["cat_rates"], out["gtrk4_lnL"], out["gtrk4_g_blens"] = np.asarray(cat), ll.item(), b.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "prune_c2.npz"), **out)
    print("prune_c2: lnL", out["jc_lnL"], out["gtr_lnL"], out["gtrk4_lnL"])


det_distances = synth.det_distances  # bit-reproducible input distances (elementwise IEEE operations only)


def configs_c5_likelihood(r, prob):
    """500 taxa: on a tree unrelated to the alignment the reference's unscaled product underflows to
    -inf, so the likelihood answers use the reference's own NJ tree of the sample's locations (stored;
    the bit-reproducible 500-taxon NJ case above is separate)."""
    pdm = np.ascontiguousarray(r.Chyp_np.get_pdm(prob["xs"][0]))
    peel, blens = r.Cpeeler.nj_np(pdm)
    peel = np.asarray(peel, dtype=np.int64)
    pn = [p.numpy() for p in ref_partials(r, prob["masks"])]
    res = ref_ll_and_grads(r, prob["masks"], prob["weights"], peel, blens, "JC69")
    return dict(c5_ll_peel=peel, c5_ll_blens=np.asarray(blens), c5_lnL=res["lnL"], c5_g_blens=res["g_blens"],
                c5_lnL_cython=float(r.Cphylo.compute_LL_np(pn, prob["weights"], peel, blens)))


def make_configs_c5(r):
    """Refresh only the 500-taxon likelihood answers inside configs.npz (seconds, not minutes)."""
    old = dict(np.load(os.path.join(HERE, "configs.npz")))
    prob = synth.make_problem(500, 20000, 1, D=5, seed=1005)
    old.update(configs_c5_likelihood(r, prob))
    print("c5 lnL", old["c5_lnL"], old["c5_lnL_cython"])
    np.savez_compressed(os.path.join(HERE, "configs.npz"), **old)


def make_configs(r):
    """The reference itself at the shapes of BASELINE.json configs[2..4] (python make_golden.py configs):
    c3  200 taxa x 20 000 patterns, JC69, two chains: Cpeeler.nj_np tree + Cphylo.compute_LL_np value;
    c4  ONE 2000-taxon tree through Cpeeler.nj_np (minutes of Python-object arithmetic);
    c5  500 taxa x 20 000 patterns, JC69: compute_LL_np value and the torch path's lnL + autograd gradient.
    Distances come from det_distances (bit-reproducible input); alignments from the seeded generator."""
    out = {}
    prob = synth.make_problem(200, 20000, 64, D=5, seed=1003)
    out["c3_mask_sum"] = int(prob["masks"].astype(np.int64).sum())
    pn = [p.numpy() for p in ref_partials(r, prob["masks"])]
    for i in range(2):
        pdm = det_distances(200, 5, 3000 + i) * 0.05
        peel, blens = r.Cpeeler.nj_np(np.ascontiguousarray(pdm))
        peel = np.asarray(peel, dtype=np.int64)
        out[f"c3_{i}_peel"], out[f"c3_{i}_blens"] = peel, np.asarray(blens)
        out[f"c3_{i}_lnL"] = float(r.Cphylo.compute_LL_np(pn, prob["weights"], peel, blens))
    prob = synth.make_problem(500, 20000, 1, D=5, seed=1005)
    out["c5_mask_sum"] = int(prob["masks"].astype(np.int64).sum())
    pdm = det_distances(500, 5, 5000) * 0.02
    peel, blens = r.Cpeeler.nj_np(np.ascontiguousarray(pdm))
    peel = np.asarray(peel, dtype=np.int64)
    out["c5_peel"], out["c5_blens"] = peel, np.asarray(blens)
    out.update(configs_c5_likelihood(r, prob))
    print("c3 lnL", out["c3_0_lnL"], out["c3_1_lnL"], "c5 lnL", out["c5_lnL"], out["c5_lnL_cython"])
    import time
    t0 = time.time()
    pdm = det_distances(2000, 5, 4000)
    peel, blens = r.Cpeeler.nj_np(np.ascontiguousarray(pdm))
    out["c4_peel"], out["c4_blens"] = np.asarray(peel, dtype=np.int64), np.asarray(blens)
    print("c4: reference nj_np at 2000 taxa took %.0f s" % (time.time() - t0))
    np.savez_compressed(os.path.join(HERE, "configs.npz"), **out)


def make_pdm_jac(r):
    """Chyp_np.get_pdm(..., get_jacobian=True) (Chyp_np.pyx:369-410): analytic Jacobian of the pair
    distances w.r.t. the sheet coordinates."""
    rng = np.random.default_rng(34)
    out = {}
    cases = [(4, 2, -1.0), (9, 3, -0.7), (20, 5, -1.0)]
    out["n"] = len(cases)
    for i, (S, D, curv) in enumerate(cases):
        x = rng.normal(0, 0.1, (S, D))
        Dm, J = r.Chyp_np.get_pdm(x, curvature=curv, get_jacobian=True)
        out[f"{i}_x"], out[f"{i}_curv"], out[f"{i}_pdm"], out[f"{i}_jac"] = x, curv, Dm, J
    np.savez_compressed(os.path.join(HERE, "pdm_jac.npz"), **out)
    print("pdm_jac cases:", len(cases))


def make_charts(r):
    """Chart changes around the wrap embedding (Chyp_torch.pyx:39-117,212-362,425-457) on seeded rows."""
    C = r.Chyp_torch
    g = torch.Generator().manual_seed(5)
    x = 0.3 * torch.randn(6, 3, dtype=torch.float64, generator=g)
    mu = 0.2 * torch.randn(6, 3, dtype=torch.float64, generator=g)
    out = {"x": x.numpy(), "mu": mu.numpy()}
    out["t02p"] = C.t02p(x).numpy()
    p, jac = C.t02p(x, mu, get_jacobian=True)
    out["t02p_mu"], out["t02p_mu_jac"] = p.numpy(), jac.numpy()
    q, qjac = C.p2t0(p, mu, get_jacobian=True)
    out["p2t0_mu"], out["p2t0_mu_jac"] = q.numpy(), qjac.numpy()
    out["p2t0"] = C.p2t0(p).numpy()
    out["real2ball"], out["ball2real"] = C.real2ball(x, 2.0).numpy(), C.ball2real(p).numpy()
    out["ladj"] = C.real2ball_LADJ(x, 2.0).numpy()
    muh, xh = C.project_up_2d(mu), C.project_up_2d(x)
    out["tangent_to_hyper"] = np.stack([C.tangent_to_hyper(muh[i], x[i], 3).numpy() for i in range(6)])
    out["transport"] = np.stack([C.parallel_transport(xh[i] - muh[i], muh[i], xh[(i + 1) % 6]).numpy() for i in range(6)])
    out["exp_inverse"] = np.stack([C.exp_map_inverse(xh[i], muh[i]).numpy() for i in range(6)])
    out["exp_map"] = np.stack([C.exponential_map(muh[i], C.exp_map_inverse(xh[i], muh[i])).numpy() for i in range(6)])
    out["h2p_jac"] = np.array([float(C.hyper_to_poincare_jacobian(xh[i])) for i in range(6)])
    # NumPy twins (Chyp_np.pyx:41-229)
    N = r.Chyp_np
    xn, mun = x.numpy(), mu.numpy()
    xhn, muhn = np.stack([N.project_up(v) for v in xn]), np.stack([N.project_up(v) for v in mun])
    out["np_tangent_to_hyper"] = np.stack([N.tangent_to_hyper(muhn[i], xn[i], 3) for i in range(6)])
    out["np_unwrap"] = N.unwrap_2d(xhn)
    out["np_exp_inverse"] = np.stack([N.exp_map_inverse(xhn[i], muhn[i]) for i in range(6)])
    out["np_transport"] = np.stack([N.parallel_transport(xhn[i] - muhn[i], muhn[i], xhn[(i + 1) % 6]) for i in range(6)])
    out["np_h2p"] = np.stack([N.hyper_to_poincare(xhn[i]) for i in range(6)])
    out["np_h2p_jac"] = np.array([N.hyper_to_poincare_jacobian(xhn[i]) for i in range(6)])
    out["np_project_up_jacobian"] = N.project_up_jacobian(xn)
    np.savez_compressed(os.path.join(HERE, "charts.npz"), **out)
    print("charts: ok")


def make_geodesic(r):
    """peeler.make_soft_peel_tips (+ autograd of a random linear functional of blens and int_locs) and
    peeler.make_hard_peel_geodesic (peeler.py:14-152), incl. the inputs of test/test_peeler.py:22-85,401-470."""
    rng = np.random.default_rng(35)
    soft = [rng.normal(0, 0.2, (6, 3)), rng.normal(0, 0.1, (12, 2)), rng.normal(0, 0.15, (27, 3)), rng.normal(0, 0.05, (40, 5)),
            0.9 * np.array([[np.cos(np.pi), np.sin(np.pi)], [np.cos(0.9 * np.pi), np.sin(0.9 * np.pi)],
                            [np.cos(-0.5 * np.pi), np.sin(-0.5 * np.pi)]]),
            np.stack([rr * np.array([np.cos(t), np.sin(t)]) for rr, t in
                      zip([0.6, 0.6, 0.5, 0.5], [np.pi * 0.2, 0, np.pi, -np.pi * 0.9])])]
    out = {"n_soft": len(soft)}
    for i, x in enumerate(soft):
        xt = torch.tensor(x, requires_grad=True)
        peel, il, bl = r.peeler.make_soft_peel_tips(xt, connector="geodesics", curvature=-torch.ones(1))
        wb, wi = rng.normal(size=bl.shape), rng.normal(size=il.shape)
        (g,) = torch.autograd.grad((bl * torch.tensor(wb)).sum() + (il * torch.tensor(wi)).sum(), xt)
        out[f"s{i}_x"], out[f"s{i}_peel"], out[f"s{i}_int"], out[f"s{i}_blens"] = x, np.asarray(peel, dtype=np.int64), il.detach().numpy(), bl.detach().numpy()
        out[f"s{i}_wb"], out[f"s{i}_wi"], out[f"s{i}_grad"] = wb, wi, g.numpy()
    dog = np.stack([0.5 * np.array([np.cos(t), np.sin(t)]) for t in (np.pi / 10, -np.pi / 10, np.pi * 6 / 8, -np.pi * 6 / 8)])
    hard = [dog, np.array([(0.05909264, 0.16842421, -0.03628194), (0.08532969, -0.07187002, 0.17884444),
                           (-0.11422830, 0.01955054, 0.14127290), (-0.06550432, 0.07029946, -0.14566249),
                           (-0.07060744, -0.12278600, -0.17569585), (0.11386343, -0.03121063, -0.18112418)]),
            rng.normal(0, 0.2, (6, 3)), rng.normal(0, 0.1, (30, 2)), rng.normal(0, 0.05, (64, 5))]
    out["n_hard"] = len(hard)
    for i, x in enumerate(hard):
        sheet = r.Chyp_np.poincare_to_hyper_2d(x)
        peel, il = r.peeler.make_hard_peel_geodesic(sheet)
        out[f"h{i}_sheet"], out[f"h{i}_peel"], out[f"h{i}_int"] = sheet, np.asarray(peel, dtype=np.int64), il
    # the MCMC route (chain.py:176-185): raw (S, D) locations go in as they are
    leaf_x = rng.normal(0, 0.1, (9, 3))
    peel, int_x = r.peeler.make_hard_peel_geodesic(leaf_x)
    out["chain_x"], out["chain_peel"], out["chain_int"] = leaf_x, np.asarray(peel, dtype=np.int64), int_x
    out["chain_blens"] = r.Cphylo.compute_branch_lengths_np(9, np.asarray(peel, dtype=np.int64), leaf_x, int_x)
    np.savez_compressed(os.path.join(HERE, "geodesic.npz"), **out)
    print("geodesic: soft", len(soft), "hard", len(hard), "dogbone peel", out["h0_peel"].tolist())


def make_hypdist(r):
    """The reference's pointwise distances (Chyp_torch.pyx:119-154, Chyp_np.pyx:279-296) on random pairs of
    ball points -- generic, near the boundary, coincident (clamped) -- with torch autograd gradients to both
    points and the curvature for the two torch forms.  python make_golden.py hypdist"""
    rng = np.random.default_rng(41)
    out = {}
    k = 0
    for D in (2, 3, 5):
        for radius in (0.3, 0.9, 0.999):
            for kappa in (-1.0, -0.6):
                n = 6
                dirs = rng.normal(size=(2, n, D))
                dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
                x1 = dirs[0] * rng.uniform(0, radius, (n, 1))
                x2 = dirs[1] * rng.uniform(0, radius, (n, 1))
                x2[0] = x1[0]  # a coincident pair: the Lorentz inner product is clamped
                res = {"x1": x1, "x2": x2, "kappa": np.array(kappa)}
                for form, fn in (("poincare", r.Chyp_torch.hyperbolic_distance), ("lorentz", r.Chyp_torch.hyperbolic_distance_lorentz)):
                    d, g1, g2, gk = [], [], [], []
                    for i in range(n):
                        a = torch.tensor(x1[i], requires_grad=True)
                        b = torch.tensor(x2[i], requires_grad=True)
                        c = torch.tensor([kappa], dtype=torch.float64, requires_grad=True)
                        v = fn(a, b, c).reshape(())
                        d.append(v.item())
                        if v.requires_grad and torch.isfinite(v):
                            ga, gb, gc = torch.autograd.grad(v, (a, b, c), allow_unused=True)
                        else:
                            ga = gb = gc = None
                        g1.append(np.zeros(D) if ga is None else ga.numpy())
                        g2.append(np.zeros(D) if gb is None else gb.numpy())
                        gk.append(0.0 if gc is None else float(gc))
                    res[f"{form}_d"], res[f"{form}_g1"], res[f"{form}_g2"], res[f"{form}_gk"] = np.array(d), np.stack(g1), np.stack(g2), np.array(gk)
                res["np_d"] = np.array([r.Chyp_np.hyperbolic_distance(x1[i], x2[i], kappa) for i in range(n)])
                for key, v in res.items():
                    out[f"{k}_{key}"] = v
                k += 1
    out["n"] = k
    np.savez_compressed(os.path.join(HERE, "hypdist.npz"), **out)
    print("hypdist:", k, "cases;", out["0_poincare_d"], out["0_lorentz_d"], out["0_np_d"])


if __name__ == "__main__":
    r = refload.load()
    assert r.compiled and r.python, "needs /root/reference and oracle/_ref (python oracle/build_ref.py)"
    makers = dict(ds1=make_ds1, pdm=make_pdm, nj_hard=make_nj_hard, nj_soft=make_nj_soft, prune=make_prune,
                  pdm_jac=make_pdm_jac, charts=make_charts, geodesic=make_geodesic, hypdist=make_hypdist)
    slow = dict(nj_soft_c2=make_nj_soft_c2, chain_c2=make_chain_c2, chains=make_chains, blens_jac=make_blens_jac, hmap=make_hmap, vi=make_vi, laplace=make_laplace, pdm_c4=make_pdm_c4, prune_c2=make_prune_c2, configs=make_configs, configs_c5=make_configs_c5)  # only when named explicitly
    for name in (sys.argv[1:] or list(makers)):  # optional: names of the fixtures to regenerate
        {**makers, **slow}[name](r)
