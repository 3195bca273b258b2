"""Run the reference's own drivers on this library, unchanged.

The reference (mattapow/dodonaphy) reaches its hot path through plain module-level imports --
``from dodonaphy import Chyp_torch, peeler, ...`` (vi.py:11, chain.py:9, hmap.py:8, base_model.py:18-22)
-- there is no plugin registry.  ``install()`` therefore binds this package's modules under the
reference's module names *before* the reference's callers are imported:

    import dodonaphy_b200.dropin as dropin
    dropin.install()                       # dodonaphy.phylo, .Cphylo, .Chyp_torch, .Chyp_np, .peeler,
    from dodonaphy.vi import DodonaphyVI   # .Cpeeler, .soft now resolve to the CUDA implementations
    from dodonaphy.chain import Chain

After that ``DodonaphyVI`` / ``Chain`` / ``DodonaphyMCMC`` / ``HMAP`` / ``Laplace`` and ``run.py`` run as they
are: every distance matrix, NJ tree and likelihood they ask for is computed by the kernels behind
include/dodonaphy_b200.h.  ``tests/test_dropin_gpu.py`` runs the reference's own classes both ways and
compares them.  Modules of the reference that are not on the path (``tree``, ``utils``, ``poincare``,
``Cutils``, ``Ctransforms``, ``phylomodel`` -- pure host code) stay the reference's own; pass
``names=ALIASES + ("phylomodel",)`` to take this package's ``PhyloModel`` as well.
"""
import importlib
import sys

ALIASES = ("phylo", "Cphylo", "Chyp_torch", "Chyp_np", "peeler", "Cpeeler", "soft")
# reference modules that bind the aliased names at import time (``from dodonaphy import ...``): if they
# were imported before install() they are dropped from sys.modules so that the next import re-binds
_CALLERS = ("base_model", "vi", "chain", "mcmc", "hmap", "laplace", "brute", "run", "utils", "peeler", "tree")


def install(package="dodonaphy", names=ALIASES):
    """Bind ``dodonaphy_b200.<name>`` as ``<package>.<name>`` for every hot-path module.  The
    reference package itself must be importable (it provides the callers)."""
    pkg = importlib.import_module(package)
    saved = getattr(pkg, "_b200_saved", None)
    if saved is None:
        saved = pkg._b200_saved = {}
    for name in names:
        mod = importlib.import_module(f"{__package__}.{name}")
        full = f"{package}.{name}"
        if name not in saved:
            saved[name] = sys.modules.get(full)
        sys.modules[full] = mod
        setattr(pkg, name, mod)
    for name in _CALLERS:
        if name not in names:
            sys.modules.pop(f"{package}.{name}", None)
            if hasattr(pkg, name):
                delattr(pkg, name)
    return pkg


def uninstall(package="dodonaphy"):
    """Undo ``install``: the reference's own modules are bound again on the next import."""
    pkg = sys.modules.get(package)
    saved = getattr(pkg, "_b200_saved", None) if pkg is not None else None
    if not saved:
        return
    for name, mod in saved.items():
        full = f"{package}.{name}"
        if mod is None:
            sys.modules.pop(full, None)
            if hasattr(pkg, name):
                delattr(pkg, name)
        else:
            sys.modules[full] = mod
            setattr(pkg, name, mod)
    saved.clear()
    for name in _CALLERS:
        sys.modules.pop(f"{package}.{name}", None)
        if hasattr(pkg, name):
            delattr(pkg, name)
