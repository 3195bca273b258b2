import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def golden_problem(name):
    """The seeded synthetic workload ``name`` ('c2' | 'c3' | 'c5') whose reference answers are stored
    under tests/golden/: regenerated from its seed ON the committed guide tree (guides.npz), so the
    alignment is the same on every host; callers assert its checksum instead of skipping."""
    from dodonaphy_b200 import synth

    S, L, B, seed = {"c2": (64, 10000, 128, 1002), "c3": (200, 20000, 64, 1003), "c5": (500, 20000, 1, 1005)}[name]
    g = load_golden("guides.npz")
    return synth.make_problem(S, L, B, D=5, seed=seed, guide=(g[f"{name}_peel"], g[f"{name}_blens"]))
