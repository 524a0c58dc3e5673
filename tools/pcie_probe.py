"""PCIe probe for the host-pointer pipeline (p3gpu_coset_lde_batch): rate of cudaMemcpy2DAsync between a pinned host matrix of
pitch 400 bytes (w = 100) and a compact device buffer, as a function of the chunk width, against the contiguous copy.
Usage (GPU box): python tools/pcie_probe.py"""
import time

import torch
from cuda import cudart

H, W = 1 << 20, 100


def chk(r):
    if isinstance(r, tuple):
        assert int(r[0]) == 0, r
        return r[1:] if len(r) > 2 else (r[1] if len(r) == 2 else None)
    assert int(r) == 0, r


host = torch.empty((2 * H, W), dtype=torch.int32).pin_memory()
dev = torch.empty((2 * H, W), dtype=torch.int32, device="cuda")
s = torch.cuda.Stream()
K = cudart.cudaMemcpyKind


def run(label, fn, nbytes, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{label:48s} {dt * 1e3:8.2f} ms  {nbytes / dt / 1e9:7.1f} GB/s")


for rows, name in ((H, "H2D"), (2 * H, "D2H")):
    kind = K.cudaMemcpyHostToDevice if name == "H2D" else K.cudaMemcpyDeviceToHost
    nb = rows * W * 4
    if name == "H2D":
        run(f"{name} contiguous {nb >> 20} MB", lambda: chk(cudart.cudaMemcpyAsync(dev.data_ptr(), host.data_ptr(), nb, kind, s.cuda_stream)), nb)
    else:
        run(f"{name} contiguous {nb >> 20} MB", lambda: chk(cudart.cudaMemcpyAsync(host.data_ptr(), dev.data_ptr(), nb, kind, s.cuda_stream)), nb)
    for wc in (8, 24, 48, 100):
        if name == "H2D":
            f = lambda: chk(cudart.cudaMemcpy2DAsync(dev.data_ptr(), wc * 4, host.data_ptr(), W * 4, wc * 4, rows, kind, s.cuda_stream))
        else:
            f = lambda: chk(cudart.cudaMemcpy2DAsync(host.data_ptr(), W * 4, dev.data_ptr(), wc * 4, wc * 4, rows, kind, s.cuda_stream))
        run(f"{name} 2-D {wc} cols ({wc * 4} B rows at pitch 400) x {rows} rows", f, rows * wc * 4)
