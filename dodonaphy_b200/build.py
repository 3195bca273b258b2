"""Build the CUDA library in-tree:  python -m dodonaphy_b200.build

nvcc -gencode arch=compute_100a,code=sm_100a -> dodonaphy_b200/_C/libdodonaphy_b200.so
(plain C ABI, include/dodonaphy_b200.h; no torch headers, no pybind).  The built .so is git-ignored
but travels to the GPU box with the snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT_DIR = os.path.join(PKG, "_C")
LIB = os.path.join(OUT_DIR, "libdodonaphy_b200.so")
STAMP = os.path.join(OUT_DIR, "build.stamp")
SOURCES = ["common.cu", "prune.cu", "pdm.cu", "hypdist.cu", "nj.cu", "align.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--fmad=true", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--expt-relaxed-constexpr"]
# tuning knobs (compile-time), e.g. DODO_DEFINES="-DDODO_PRUNE_MINBLOCKS=3"
NVCC_FLAGS += os.environ.get("DODO_DEFINES", "").split()


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the dodonaphy_b200 CUDA library cannot be built")


def _digest():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files.append(os.path.join(os.path.dirname(PKG), "include", "dodonaphy_b200.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh():
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == _digest()


def build(force=False, verbose=False):
    if is_fresh() and not force:
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = nvcc_path()
    objs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(STAMP, "w") as fh:
        fh.write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
