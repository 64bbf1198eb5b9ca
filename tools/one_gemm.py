"""Run one tcgen05 GEMM shape a few times (for ncu).  usage: one_gemm.py M N K res(0/1) bf16out(0/1) pair(0/1) [bn]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdb200 import _lib, ops
M, N, K, res, obf, pair = [int(v) for v in sys.argv[1:7]]
bn = int(sys.argv[7]) if len(sys.argv) > 7 else 0
geglu = int(sys.argv[8]) if len(sys.argv) > 8 else 0
lib = _lib.load()
lib.sdb_tc_set_pair_kernel(pair)
torch.manual_seed(0)
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
W = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
R = torch.randn(M, N, device="cuda") if res else None
if geglu:
    from sdb200.engine import PackedLinear
    pl = PackedLinear(W.float(), bias, "bf16", geglu=True)
for _ in range(4):
    if geglu:
        out = ops.gemm_tc(A, pl.w, pl.bias, out_dtype=torch.bfloat16, geglu=True, block_n=pl.block_n)
    else:
        out = ops.gemm_tc(A, W, bias, residual=R, out_dtype=torch.bfloat16 if obf else torch.float32, block_n=bn)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
