"""dodonaphy_b200 -- B200-native (sm_100a) implementation of Dodonaphy's hot path:
pairwise hyperbolic distance -> (soft|hard) neighbour joining -> Felsenstein pruning lnL + gradient.

Submodules mirror the reference's module names (``phylo``, ``Cphylo``, ``phylomodel``, ``Chyp_torch``,
``Chyp_np``, ``peeler``, ``Cpeeler``, ``base_model``); ``engine`` / ``pipeline`` are the batched device
API underneath; ``dist`` shards it over GPUs.  Importing the package does not need a GPU; calling
into it does (there is no CPU fallback).
"""
import importlib

__version__ = "0.1.0"
_SUBMODULES = ("phylo", "Cphylo", "phylomodel", "Chyp_torch", "Chyp_np", "peeler", "Cpeeler", "soft", "base_model", "vi", "mcmc", "hmap", "laplace",
               "engine", "pipeline", "dist", "priors", "synth", "newick", "tree", "build")


def __getattr__(name):
    if name in _SUBMODULES:
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
