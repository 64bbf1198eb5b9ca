"""Timeline of the one-CTA contraction kernel (measurement build -DSDB_TC_TRACE): per CTA, nanoseconds since the first CTA
started: start | setup done | dependency wait done | last MMA committed | epilogue warp reached tile | epilogue done, and the SM.
usage: trace_single.py M N K res obf bn"""
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
M, N, K, res, obf, bn = [int(v) for v in sys.argv[1:7]]
torch.manual_seed(0)
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
W = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
R = torch.randn(M, N, device="cuda") if res else None
fn = lambda: ops.gemm_tc(A, W, bias, residual=R, out_dtype=torch.bfloat16 if obf else torch.float32, block_n=bn, variant=1)
for _ in range(3):
    fn()
trace = torch.zeros(2048 * 8, dtype=torch.int64, device="cuda")
lib.sdb_tc_set_trace.argtypes = [C.c_void_p]
lib.sdb_tc_set_trace(trace.data_ptr())
fn()
torch.cuda.synchronize()
lib.sdb_tc_set_trace(None)
t = trace.cpu().reshape(2048, 8)
live = t[:, 0] > 0
n = int(live.sum())
base = int(t[live, 0].min())
print("args", sys.argv[1:], "CTAs", n, "span %.1f us" % ((int(t[live, 5].max()) - base) / 1e3))
d = lambda a, b: (t[live, a] - t[live, b]).double()
print("mean ns: setup %.0f | dep wait %.0f | mainloop (to last commit) %.0f | epilogue %.0f | CTA lifetime %.0f" % (
    d(1, 0).mean(), d(2, 1).mean(), d(3, 2).mean(), d(5, 4).mean(), d(5, 0).mean()))
print("CTA start times (us) percentiles:", [round((float(torch.quantile((t[live, 0] - base).double(), q))) / 1e3, 1) for q in (0, .25, .5, .75, 1)])
sm0 = int(t[0, 7])
print("CTAs on SM %d: (start, setup, dep, mma_done, epi_start, epi_end) us" % sm0)
for i in range(2048):
    if live[i] and int(t[i, 7]) == sm0:
        print("   cta %4d: %s" % (i, " ".join("%7.2f" % ((int(t[i, k]) - base) / 1e3) for k in range(6))))
