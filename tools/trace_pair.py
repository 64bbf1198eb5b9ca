"""Timeline of the persistent CTA-pair contraction kernel: builds a -DSDB_TC_TRACE copy of the library under gpurun_out/
(never shipped or loaded by the product), runs one GEMM and prints per-unit SM-clock stamps of the leader CTA of a few pairs:
  P0/P1 producer first/last TMA issue, M0w MMA warp reached the unit, M0 accumulator free, M1 first operands landed,
  M2 last MMA committed, E0 epilogue (warp 4) reached the unit, E1 epilogue done.
usage: trace_pair.py M N K res obf bn [geglu]"""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pk = os.path.join(ROOT, "stable-diffusion-from-scratch_b200")
so = os.path.join(ROOT, "variants", "libsdb200_trace.so")      # prebuilt on the dev box when present (variants/ travels)
if not os.path.exists(so):
    so = os.path.join(ROOT, "gpurun_out", "libsdb200_trace.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
if not os.path.exists(so):
    srcs = [os.path.join(pk, "csrc", f) for f in sorted(os.listdir(os.path.join(pk, "csrc"))) if f.endswith(".cu")]
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-DSDB_TC_TRACE",
                           "-shared", "-o", so] + srcs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"])
from sdb200 import _lib
lib = _lib.load(so)
_lib._lib = lib
from sdb200 import ops
from sdb200.engine import PackedLinear
M, N, K, res, obf, bn = [int(v) for v in sys.argv[1:7]]
geglu = int(sys.argv[7]) if len(sys.argv) > 7 else 0
torch.manual_seed(0)
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
W = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
R = torch.randn(M, N, device="cuda") if res else None
if geglu:
    pl = PackedLinear(W.float(), bias, "bf16", geglu=True)
    fn = lambda: ops.gemm_tc(A, pl.w, pl.bias, out_dtype=torch.bfloat16, geglu=True, block_n=pl.block_n, variant=2)
else:
    fn = lambda: ops.gemm_tc(A, W, bias, residual=R, out_dtype=torch.bfloat16 if obf else torch.float32, block_n=bn, variant=2)
for _ in range(3):
    fn()
trace = torch.zeros(74 * 64 * 8, dtype=torch.int64, device="cuda")
lib.sdb_tc_set_trace.argtypes = [C.c_void_p]
lib.sdb_tc_set_trace(trace.data_ptr())
fn()
torch.cuda.synchronize()
lib.sdb_tc_set_trace(None)
t = trace.cpu().reshape(74, 64, 8)
print("args", sys.argv[1:])
for pair in (0, 36, 73):
    nz = t[pair][t[pair] > 0]
    if nz.numel() == 0:
        continue
    base = int(nz.min())
    print("pair %d (cycles since its first stamp): unit | P0 P1 | M0w M0 M1 M2 | E0 E1" % pair)
    for u in range(64):
        r = t[pair, u]
        if int(r.max()) == 0:
            break
        f = lambda k: ("%7d" % (int(r[k]) - base)) if int(r[k]) else "      -"
        print("  %2d | %s %s | %s %s %s %s | %s %s" % (u, f(6), f(7), f(5), f(0), f(1), f(2), f(3), f(4)))
