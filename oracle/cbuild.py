"""Build and load oracle/c/*.c (TEST INFRASTRUCTURE): gcc -O2 -ffp-contract=off -> oracle/_build/liboracle.so."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "nj_hard.c")
OUT = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    if os.path.exists(OUT) and not force and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", SRC, "-o", OUT], check=True)
    return OUT


def nj_hard_c(pdm):
    """C twin of oracle.nj.nj_hard (bit-identical to it and to Cpeeler.nj_np)."""
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_nj_hard.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    pdm = np.ascontiguousarray(pdm, dtype=np.float64)
    S = pdm.shape[0]
    peel = np.zeros((S - 1, 3), dtype=np.int64)
    blens = np.zeros(2 * S - 2, dtype=np.float64)
    rc = _lib.oracle_nj_hard(pdm.ctypes.data, S, peel.ctypes.data, blens.ctypes.data)
    if rc != 0:
        raise MemoryError("oracle_nj_hard")
    return peel, blens
