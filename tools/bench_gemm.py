"""tcgen05 contraction micro-benchmark: TFLOP/s per (shape, block_n, kernel variant).  Measurement tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdb200 import _lib, ops
lib = _lib.load()
dev = "cuda"
torch.manual_seed(0)


def bench(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


shapes = [  # (M, N, K, residual, out_bf16)
    (32768, 1280, 5760, False, False),
    (32768, 320, 2880, True, False),
    (32768, 320, 320, True, False),
    (32768, 960, 320, False, True),
    (8192, 640, 5760, True, False),
    (2048, 1280, 11520, True, False),
    (512, 1280, 11520, True, False),
]
for M, N, K, res, obf in shapes:
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    R = torch.randn(M, N, device=dev) if res else None
    ref = None
    for variant in (0, 1):
        lib.sdb_tc_set_pair_kernel(variant)
        for bn in (128, 160, 256):
            if N % bn and not (N < bn):
                pass
            for sk in ((1, 0) if M <= 2048 else (1,)):
                try:
                    f = lambda: ops.gemm_tc(A, W, bias, residual=R, out_dtype=torch.bfloat16 if obf else torch.float32, block_n=bn, split_k=sk)
                    out = f().float()
                    if ref is None:
                        ref = out
                    err = float((out - ref).norm() / ref.norm())
                    ms = bench(f)
                    print("M=%6d N=%5d K=%6d res=%d bf16out=%d | pair=%d bn=%3d split=%s : %8.3f us %8.1f TFLOP/s  rel-diff %.1e" % (
                        M, N, K, res, obf, variant, bn, "auto" if sk == 0 else "1", ms * 1e3, 2.0 * M * N * K / ms / 1e9, err), flush=True)
                except Exception as e:
                    print("M=%d N=%d K=%d pair=%d bn=%d FAILED: %s" % (M, N, K, variant, bn, e), flush=True)
