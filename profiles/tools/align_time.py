"""Alignment ingestion at the long-alignment shape (500 taxa x 1M sites, configs[4]): the library's radix-sort
kernels (csrc/align.cu) next to torch.unique(dim=0) on the same device.   python profiles/tools/align_time.py [S] [n]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from dodonaphy_b200 import _lib
S = int(sys.argv[1]) if len(sys.argv) > 1 else 500
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
rng = np.random.default_rng(3)
base = rng.integers(0, 4, size=(S, 20000), dtype=np.uint8)      # 20k distinct-ish columns, repeated (as the benches do)
raw_np = np.frombuffer(b"ACGT", dtype=np.uint8)[np.tile(base, (1, -(-n // 20000)))[:, :n]]
raw = torch.as_tensor(np.ascontiguousarray(raw_np)).cuda()
lib = _lib.load()
lut = torch.full((256,), 15, dtype=torch.uint8); lut[65], lut[67], lut[71], lut[84] = 1, 2, 4, 8
lut = lut.cuda()
def ours():
    ws = torch.empty(int(lib.dodo_alignment_workspace_bytes(S, n)), dtype=torch.uint8, device="cuda")
    nu = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.dodo_alignment_sort(raw.data_ptr(), S, n, ws.data_ptr(), nu.data_ptr(), st))
    U = int(nu.item())
    masks = torch.empty((S, U), dtype=torch.uint8, device="cuda"); w = torch.empty(U, dtype=torch.int64, device="cuda")
    _lib.check(lib.dodo_alignment_gather(raw.data_ptr(), S, n, ws.data_ptr(), U, lut.data_ptr(), masks.data_ptr(), w.data_ptr(), st))
    torch.cuda.synchronize()
    return masks, w
def lib_unique():
    cols, counts = torch.unique(raw.T.contiguous(), dim=0, return_counts=True)
    m = lut[cols.T.contiguous().long()]
    torch.cuda.synchronize()
    return m, counts
for name, fn in (("radix-sort kernels", ours), ("torch.unique", lib_unique)):
    fn()
    t0 = time.perf_counter(); out = fn(); dt = time.perf_counter() - t0
    print(f"{name}: {dt * 1e3:.1f} ms, {out[0].shape[1]} patterns", flush=True)
a, b = ours(), lib_unique()
print("identical:", bool(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])))
