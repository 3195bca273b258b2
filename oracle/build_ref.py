"""Build recipe for ``oracle/_ref`` -- the reference's own compiled hot-path modules.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Nothing under ``dodonaphy_b200/``
imports this.

The reference (mattapow/dodonaphy) ships its MCMC hot path as Cython extension modules
(``setup.py:20-54``): ``Cphylo`` (pruning, JC69 P(t)), ``Cpeeler`` (hard NJ), ``Chyp_np`` and
``Chyp_torch`` (pairwise hyperbolic distance) and ``Ctransforms`` (imported by the two Chyp
modules).  This script compiles them *from the sources where they lie under /root/reference*
into ``oracle/_ref/dodonaphy/*.so`` (git-ignored binaries; no reference source is copied into the
repository).  We do not run the reference's own ``setup.py``.

Two mechanical shims are needed for today's toolchain (probed):

* ``np.int_t`` no longer exists in NumPy 2's pxd  ->  compiled from a patched temporary copy in
  which it is spelt ``np.npy_long`` (same C type the reference was written against);
* ``dtype=long`` at ``Cpeeler.pyx:28`` is Python-2 syntax  ->  ``language_level=2`` (the default of
  the Cython 0.29 named in the reference's requirements.txt).

The extension modules do ``from dodonaphy.node import Node`` / ``from dodonaphy.edge import Edge``
at import time.  When /root/reference is not on the box (the GPU box) ``oracle/refload.py`` supplies
tiny stand-ins for those two helper classes; in this container the real ones are imported from
/root/reference.

Usage:  python oracle/build_ref.py [--reference /root/reference] [--force]
"""
import argparse
import os
import re
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "dodonaphy")
MODULES = ["Ctransforms", "Chyp_np", "Chyp_torch", "Cphylo", "Cpeeler", "Cutils"]
# the reference's pure-Python modules (its callers: vi.py, chain.py, mcmc.py, base_model.py ... and
# its torch path phylo.py / phylomodel.py / peeler.py / soft.py).  They are placed, unmodified, next
# to the compiled modules in the git-ignored oracle/_ref/ so that they travel to the GPU box with the
# snapshot: tests/test_dropin_gpu.py runs the reference's OWN DodonaphyVI / Chain over this library's
# modules there, and bench.py --impl reference times the reference's own calculate_treelikelihood.
# Nothing is committed (oracle/_ref/ is in .gitignore).
DATA_FILES = ["test/data/ds1/dna.nex", "test/data/ds1/dna.nj.newick", "test/test_peel_data.txt"]


def ext_suffix():
    return sysconfig.get_config_var("EXT_SUFFIX")


def is_built():
    return all(os.path.exists(os.path.join(OUT, m + ext_suffix())) for m in MODULES)


def python_present():
    return os.path.exists(os.path.join(OUT, "vi.py")) and os.path.exists(os.path.join(HERE, "_ref", "data", "dna.nex"))


def copy_python(reference="/root/reference"):
    """Place the reference's pure-Python modules and the DS1 inputs in oracle/_ref (git-ignored)."""
    src = os.path.join(reference, "dodonaphy")
    if not os.path.isdir(src):
        return python_present()
    os.makedirs(OUT, exist_ok=True)
    for f in sorted(os.listdir(src)):
        if f.endswith(".py"):
            shutil.copyfile(os.path.join(src, f), os.path.join(OUT, f))
    data = os.path.join(HERE, "_ref", "data")
    os.makedirs(data, exist_ok=True)
    for rel in DATA_FILES:
        if os.path.exists(os.path.join(reference, rel)):
            shutil.copyfile(os.path.join(reference, rel), os.path.join(data, os.path.basename(rel)))
    return python_present()


def build(reference="/root/reference", force=False, verbose=False):
    """Compile the reference's Cython modules into oracle/_ref/dodonaphy/.  Returns True when
    oracle/_ref is usable afterwards, False when the reference sources are absent."""
    copy_python(reference)
    if is_built() and not force:
        return True
    src_dir = os.path.join(reference, "dodonaphy", "cython")
    if not os.path.isdir(src_dir):
        return is_built()
    import numpy as np
    from Cython.Build import cythonize  # noqa: F401  (fail early if Cython is missing)

    os.makedirs(OUT, exist_ok=True)
    py_inc = sysconfig.get_paths()["include"]
    np_inc = np.get_include()
    with tempfile.TemporaryDirectory(prefix="dodoref_") as tmp:
        pkg = os.path.join(tmp, "dodonaphy")
        os.makedirs(pkg)
        for m in MODULES:
            with open(os.path.join(src_dir, m + ".pyx"), "r", encoding="utf-8") as fh:
                text = fh.read()
            text = re.sub(r"np\.int_t", "np.npy_long", text)
            with open(os.path.join(pkg, m + ".pyx"), "w", encoding="utf-8") as fh:
                fh.write(text)
        for m in MODULES:
            pyx = os.path.join(pkg, m + ".pyx")
            c_file = os.path.join(pkg, m + ".c")
            cmd = [sys.executable, "-m", "cython", "-2", "--module-name", "dodonaphy." + m,
                   "-o", c_file, pyx]
            subprocess.run(cmd, check=True, cwd=tmp,
                           stdout=None if verbose else subprocess.DEVNULL,
                           stderr=None if verbose else subprocess.DEVNULL)
            so = os.path.join(OUT, m + ext_suffix())
            cc = ["gcc", "-O2", "-fPIC", "-shared", "-fwrapv", "-w",
                  "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
                  "-I", py_inc, "-I", np_inc, c_file, "-o", so]
            subprocess.run(cc, check=True, cwd=tmp)
    return is_built()


def clean():
    shutil.rmtree(os.path.join(HERE, "_ref"), ignore_errors=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    ok = build(args.reference, args.force, args.verbose)
    print("oracle/_ref", "ready" if ok else "NOT built (reference sources absent)")
    sys.exit(0 if ok else 1)
