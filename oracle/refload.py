"""Loader for the *real* reference (mattapow/dodonaphy) -- TEST INFRASTRUCTURE ONLY.

``load()`` assembles an importable ``dodonaphy`` package out of

* ``oracle/_ref/dodonaphy/*.so``  -- the reference's Cython modules compiled by
  ``oracle/build_ref.py`` (binaries; travel to the GPU box with the snapshot), and
* the reference's pure-Python modules: the unmodified files ``oracle/build_ref.py`` placed next to the
  compiled ones in the git-ignored ``oracle/_ref/dodonaphy/`` (they travel to the GPU box with the
  snapshot), else ``/root/reference/dodonaphy/*.py`` when that directory exists (this container).
  With neither, only the compiled modules (``Cphylo``, ``Cpeeler``, ``Chyp_np``, ``Chyp_torch``) load.

It returns a small namespace describing what could be loaded.  Stubs installed:
``dendropy*``, ``matplotlib*``, ``hydraPlus`` as MagicMock modules, ``np.NaN``/``np.infty`` aliases.
When the reference's ``node.py``/``edge.py`` helper classes are not available the loader supplies
stand-ins with the same attributes (``Cpeeler.pyx:7-8`` imports them at module load).
"""
import importlib
import os
import sys
import types
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO_DIR = os.path.join(HERE, "_ref", "dodonaphy")
REF_PY_DIR = REF_SO_DIR if os.path.isfile(os.path.join(REF_SO_DIR, "phylo.py")) else "/root/reference/dodonaphy"
REF_DATA_DIR = os.path.join(HERE, "_ref", "data")

_state = {}


class _Node:  # stand-in for dodonaphy/node.py (3 attributes read by Cpeeler.nj_np)
    def __init__(self, taxon=None):
        self.taxon = taxon
        self._nj_distances = {}
        self._nj_xsub = 0.0


class _Edge:  # stand-in for dodonaphy/edge.py
    def __init__(self, distance, node_1, node_2):
        self.distance = distance
        self.from_ = node_1
        self.to_ = node_2

    def __lt__(self, other):
        return self.distance < other.distance


class _StubFinder:
    """Last-resort importer: any ``dendropy*`` / ``matplotlib*`` / ``hydraPlus*`` module that is
    not installed resolves to a MagicMock package (the hot path never calls into them)."""

    ROOTS = ("dendropy", "matplotlib", "hydraPlus", "hydra")

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] not in self.ROOTS:
            return None
        import importlib.machinery

        return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__path__ = []
        m.__spec__ = spec
        return m

    def exec_module(self, module):
        return None


def load():
    """Import the reference.  Returns a namespace with attributes
    ``compiled`` (bool: Cython hot path available), ``python`` (bool: pure-Python torch path
    available) and the modules that could be imported (``Cphylo``, ``Cpeeler``, ``Chyp_np``,
    ``Chyp_torch``, and, when ``python``: ``phylo``, ``phylomodel``, ``soft``, ``peeler``)."""
    if _state:
        return types.SimpleNamespace(**_state)
    import numpy as np

    if not hasattr(np, "NaN"):
        np.NaN = np.nan
    if not hasattr(np, "infty"):
        np.infty = np.inf

    have_so = os.path.isdir(REF_SO_DIR) and any(f.endswith(".so") for f in os.listdir(REF_SO_DIR))
    have_py = os.path.isfile(os.path.join(REF_PY_DIR, "phylo.py"))
    _state.update(compiled=False, python=False)
    if not have_so:
        return types.SimpleNamespace(**_state)

    sys.meta_path.append(_StubFinder())

    pkg = types.ModuleType("dodonaphy")
    pkg.__path__ = [REF_SO_DIR] + ([REF_PY_DIR] if have_py and REF_PY_DIR != REF_SO_DIR else [])
    pkg.__package__ = "dodonaphy"
    sys.modules["dodonaphy"] = pkg
    if not have_py:
        node = types.ModuleType("dodonaphy.node")
        node.Node = _Node
        edge = types.ModuleType("dodonaphy.edge")
        edge.Edge = _Edge
        sys.modules["dodonaphy.node"] = node
        sys.modules["dodonaphy.edge"] = edge

    for m in ["Ctransforms", "Chyp_np", "Chyp_torch", "Cphylo", "Cpeeler"]:
        _state[m] = importlib.import_module("dodonaphy." + m)
    _state["compiled"] = True
    if have_py:
        for m in ["phylo", "phylomodel", "soft", "peeler"]:
            _state[m] = importlib.import_module("dodonaphy." + m)
        _state["python"] = True
    return types.SimpleNamespace(**_state)
