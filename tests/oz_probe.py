"""Development probe (GPU): float64 products through the tcgen05 int8 digit-plane kernel against torch float64.
    python tests/oz_probe.py            -> accuracy and speed table on stdout"""
import ctypes
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gapro_b200 import _lib  # noqa: E402


def oz_gemm(lib, A, transA, B, transB, S, ks=None, reps=1):
    dev = A.device
    M = A.shape[1] if transA else A.shape[0]
    K = A.shape[0] if transA else A.shape[1]
    N = B.shape[1] if transB else B.shape[0]
    C = torch.full((M, N), float("nan"), dtype=torch.float64, device=dev)
    ws = torch.empty(lib.gapro_ozaki_workspace_bytes(M, N, K, S) + 1024, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 1024
    ms_s, ms_g = ctypes.c_float(), ctypes.c_float()
    _lib.check(lib.gapro_ozaki_gemm(A.data_ptr(), A.stride(0), int(transA), 0 if ks is None else ks.data_ptr(),
                                    B.data_ptr(), B.stride(0), int(transB), M, N, K, S, C.data_ptr(), C.stride(0),
                                    ws.data_ptr() + off, ws.numel() - off, reps, ctypes.byref(ms_s), ctypes.byref(ms_g),
                                    torch.cuda.current_stream(dev).cuda_stream), "gapro_ozaki_gemm")
    torch.cuda.synchronize(dev)
    return C, ms_s.value, ms_g.value


def main():
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    rows = []
    for (M, N, K) in [(128, 64, 64), (64, 64, 32), (200, 130, 100), (520, 1000, 333), (1344, 1344, 1344), (4096, 4096, 4096)]:
        for tA, tB in [(0, 0), (1, 0), (0, 1)]:
            if (M > 600) and (tA or tB) and M != 1344:
                continue
            A = torch.randn((K, M) if tA else (M, K), generator=g, dtype=torch.float64, device=dev)
            B = torch.randn((K, N) if tB else (N, K), generator=g, dtype=torch.float64, device=dev)
            # wide dynamic range inside rows, like L^-1: scale columns by 10^U(-3, 2)
            A = A * torch.pow(10.0, torch.rand(A.shape[1], generator=g, dtype=torch.float64, device=dev) * 5 - 3)
            opA = A.t() if tA else A
            opB = B.t() if tB else B
            ref = opA @ opB.t()
            bound = opA.abs().max(1)[0][:, None] * opB.abs().max(1)[0][None, :] * K
            for S in (4, 5, 6, 7, 8):
                C, ms_s, ms_g = oz_gemm(lib, A, tA, B, tB, S, reps=3 if M >= 1000 else 1)
                err_b = float(((C - ref).abs() / bound).max())
                err_n = float((C - ref).abs().max() / ref.abs().max())
                tf = 2.0 * M * N * K / (ms_g * 1e-3) / 1e12
                rows.append((M, N, K, tA, tB, S, err_b, err_n, ms_s, ms_g, tf))
                print(f"M={M:5d} N={N:5d} K={K:5d} tA={tA} tB={tB} S={S}  err/bound={err_b:.2e}  err/max|C|={err_n:.2e}  "
                      f"slice {ms_s:.3f} ms  gemm {ms_g:.3f} ms  = {tf:.1f} f64-equivalent TFLOP/s", flush=True)
    # kscale path
    M = N = K = 300
    A = torch.randn(M, K, generator=g, dtype=torch.float64, device=dev)
    B = torch.randn(N, K, generator=g, dtype=torch.float64, device=dev)
    ks = torch.randn(K, generator=g, dtype=torch.float64, device=dev)
    C, _, _ = oz_gemm(lib, A, 0, B, 0, 6, ks=ks)
    ref = (A * ks) @ B.t()
    print("kscale err/max|C| = %.2e" % float((C - ref).abs().max() / ref.abs().max()))
    # the DMMA reference rate on the same product for comparison: torch f64 matmul (cuBLAS DGEMM)
    for n in (1344, 4096):
        A = torch.randn(n, n, dtype=torch.float64, device=dev)
        B = torch.randn(n, n, dtype=torch.float64, device=dev)
        for _ in range(2):
            A @ B
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            A @ B
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"cuBLAS DGEMM n={n}: {ms:.3f} ms = {2.0 * n ** 3 / (ms * 1e-3) / 1e12:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
