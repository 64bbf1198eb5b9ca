"""Per-item timeline of CTA (0,0,0) of tc_attention_kv1_kernel (measurement build with -DSDB_XATTN_TRACE under gpurun_out/, never
loaded by the product).  Softmax warp stamps (SM clock, relative): 0 reached item, 1 S ready, 2 S in registers, 3 row max,
4 P written, 5 reached output, 6 P V retired, 7 output stored.  Issuer stamps: 0 reached QK, 1 Q landed, 2 S free, 3 QK issued,
4 reached PV, 5 P published, 6 PV issued.    usage: trace_xattn.py B H Sq Sk d"""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pk = os.path.join(ROOT, "stable-diffusion-from-scratch_b200")
so = os.path.join(ROOT, "variants", "libsdb200_xtrace.so")      # prebuilt on the dev box (variants/ travels, is not tracked)
if not os.path.exists(so):
    so = os.path.join(ROOT, "gpurun_out", "libsdb200_xtrace.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
if not os.path.exists(so):
    srcs = [os.path.join(pk, "csrc", f) for f in sorted(os.listdir(os.path.join(pk, "csrc"))) if f.endswith(".cu")]
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-DSDB_XATTN_TRACE",
                           "-shared", "-o", so] + srcs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"])
from sdb200 import _lib
lib = _lib.load(so)
_lib._lib = lib
from sdb200 import ops
from sdb200.engine import head_pad
B, H, Sq, Sk, d = [int(v) for v in sys.argv[1:6]]
dp = head_pad(d)
dev = "cuda"
torch.manual_seed(0)
q = torch.zeros(B, Sq, H, dp, device=dev, dtype=torch.bfloat16); q[..., :d] = torch.randn(B, Sq, H, d, device=dev)
k = torch.zeros(B, Sk, H, dp, device=dev, dtype=torch.bfloat16); k[..., :d] = torch.randn(B, Sk, H, d, device=dev)
v = torch.zeros(B, Sk, H, dp, device=dev, dtype=torch.bfloat16); v[..., :d] = torch.randn(B, Sk, H, d, device=dev)
fn = lambda: ops.attention_tc(q, k, v, B, H, Sq, Sk, d, dp, d ** -0.5, (Sq * H * dp, H * dp, dp), (Sk * H * dp, H * dp, dp), (Sk * H * dp, H * dp, dp))
for _ in range(3):
    fn()
trace = torch.zeros(10 * 32 * 8, dtype=torch.int64, device=dev)
lib.sdb_xattn_set_trace.argtypes = [C.c_void_p]
lib.sdb_xattn_set_trace(trace.data_ptr())
fn()
torch.cuda.synchronize()
lib.sdb_xattn_set_trace(None)
t = trace.cpu().reshape(10, 32, 8)
base = int(t[t > 0].min())
print("args", sys.argv[1:])
for g in (0, 1):
    print("MMA issuer of query tile %d: item | reachQK Qfull Sfree QKissued | reachPV Pfull PVissued" % g)
    for u in range(32):
        r = t[8 + g, u]
        if int(r.max()) == 0:
            break
        print("  %2d | %s" % (u, " ".join(("%7d" % (int(x) - base)) if int(x) else "      -" for x in r[:7])))
for w in ([int(x) for x in os.environ.get('XA_WARPS', '0,4').split(',')]):
    print("softmax warp %d (query tile %d, lane quarter %d): item | reach Sready Sregs max Pwritten | reachOut PVdone stored | step" % (w + 4, w // 4, w % 4))
    prev = None
    for u in range(32):
        r = t[w, u]
        if int(r.max()) == 0:
            break
        vals = [int(x) - base for x in r]
        print("  %2d | %s | %s" % (u, " ".join("%7d" % x for x in vals), "" if prev is None else str(vals[7] - prev)))
        prev = vals[7]
