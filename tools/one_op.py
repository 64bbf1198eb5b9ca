"""Run ONE kernel configuration a few times (for ncu captures).  Measurement tool.
  one_op.py gemm M N K res(0/1) bf16out(0/1) variant(0/1/2) bn geglu(0/1)
  one_op.py conv NB H W Cin Cout ksize stride res(0/1) variant bn
  one_op.py attn B H Sq Sk d
  one_op.py gn N HW C act
  one_op.py attnw B Sq Sk d          (wide single head, VAE AttnBlock)
  one_op.py ddim n                   (fused DDIM update over n fp32 elements)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdb200 import _lib, ops
from sdb200.engine import PackedLinear, head_pad
lib = _lib.load(os.environ.get("SDB200_LIB") or None)   # measurement builds (tools/gpu_attn_variants.sh) pass their own library
torch.manual_seed(0)
kind = sys.argv[1]
v = [int(x) for x in sys.argv[2:]]
dev = "cuda"
if kind == "gemm":
    M, N, K, res, obf, variant, bn, geglu = (v + [0] * 8)[:8]
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    R = torch.randn(M, N, device=dev) if res else None
    if geglu:
        pl = PackedLinear(W.float(), bias, "bf16", geglu=True)
        fn = lambda: ops.gemm_tc(A, pl.w, pl.bias, out_dtype=torch.bfloat16, geglu=True, block_n=pl.block_n, variant=variant)
    else:
        fn = lambda: ops.gemm_tc(A, W, bias, residual=R, out_dtype=torch.bfloat16 if obf else torch.float32, block_n=bn, variant=variant)
elif kind == "conv":
    NB, H, Wd, Cin, Cout, ks, stride, res, variant, bn = (v + [0] * 10)[:10]
    x = torch.randn(NB, H, Wd, Cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(ks * ks, Cout, Cin, device=dev) / (ks * ks * Cin) ** 0.5).to(torch.bfloat16)
    bias = torch.randn(Cout, device=dev)
    OH = (H + 2 * (ks // 2) - ks) // stride + 1
    R = torch.randn(NB, OH, OH, Cout, device=dev) if res else None
    fn = lambda: ops.conv_tc(x, w, bias, ks, ks, stride=stride, pad=ks // 2, residual=R, block_n=bn, variant=variant)
elif kind == "attn":
    B, H, Sq, Sk, d = v[:5]
    dp = head_pad(d)
    q = torch.zeros(B, Sq, H, dp, device=dev, dtype=torch.bfloat16); q[..., :d] = torch.randn(B, Sq, H, d, device=dev)
    k = torch.zeros(B, Sk, H, dp, device=dev, dtype=torch.bfloat16); k[..., :d] = torch.randn(B, Sk, H, d, device=dev)
    vv = torch.zeros(B, Sk, H, dp, device=dev, dtype=torch.bfloat16); vv[..., :d] = torch.randn(B, Sk, H, d, device=dev)
    fn = lambda: ops.attention_tc(q, k, vv, B, H, Sq, Sk, d, dp, d ** -0.5, (Sq * H * dp, H * dp, dp), (Sk * H * dp, H * dp, dp), (Sk * H * dp, H * dp, dp))
elif kind == "attnw":
    B, Sq, Sk, d = v[:4]
    q = torch.randn(B * Sq, d, device=dev).to(torch.bfloat16)
    k = torch.randn(B * Sk, d, device=dev).to(torch.bfloat16)
    vv = torch.randn(B * Sk, d, device=dev).to(torch.bfloat16)
    fn = lambda: ops.attention_wide(q, k, vv, B, Sq, Sk, d, d ** -0.5, (Sq * d, d), (Sk * d, d), (Sk * d, d))
elif kind == "ddim":
    n = v[0]
    x, e = torch.randn(n, device=dev), torch.randn(n, device=dev)
    fn = lambda: ops.ddim_step(x, e, 0.9, 0.95, 0.3, 0.0, 0.4)[0]
elif kind == "gn":
    N, HW, Cc, act = v[:4]
    x = torch.randn(N, HW, 1, Cc, device=dev)
    g, b = torch.randn(Cc, device=dev), torch.randn(Cc, device=dev)
    fn = lambda: ops.groupnorm(x, g, b, 1e-5, act=act, out_dtype=torch.bfloat16)
elif kind == "ln":
    rows, Cc = v[:2]
    x = torch.randn(rows, Cc, device=dev)
    g, b = torch.randn(Cc, device=dev), torch.randn(Cc, device=dev)
    fn = lambda: ops.layernorm(x, g, b, 1e-5, out_dtype=torch.bfloat16)
else:
    raise SystemExit("unknown kind")
for _ in range(4):
    out = fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = fn()
e1.record()
torch.cuda.synchronize()
if os.environ.get("ONE_OP_GRAPH", "0") == "1":
    # 20 calls captured in one CUDA graph: launches short enough to be host-bound above (< ~20 us: three cuTensorMapEncode calls
    # and a ctypes round trip per launch) show their device time
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            for _ in range(20):
                out = fn()
    torch.cuda.synchronize()
    gr.replay()
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(5):
        gr.replay()
    g1.record()
    torch.cuda.synchronize()
    print("graph replay: %.2f us/call" % (g0.elapsed_time(g1) / 100 * 1e3))
if kind == "attn":      # accuracy of (batch 0, head 0) against fp64 softmax(q k^T) v on the same bf16 inputs
    qd, kd, vd = q[0, :, 0, :d].double(), k[0, :, 0, :d].double(), vv[0, :, 0, :d].double()
    ref = torch.softmax(qd @ kd.T * d ** -0.5, -1) @ vd
    got = out.view(B, Sq, H, d)[0, :, 0].double()
    print("rel-L2 vs fp64: %.3e" % float((got - ref).norm() / ref.norm()))
print("ok %s %s: %.2f us/call, mean|out| %.4f" % (kind, v, e0.elapsed_time(e1) / 20 * 1e3, float(out.float().abs().mean())))
