"""Per-layer timing of the SD-1.x UNet (or the VAE decoder) at batch B: every C-ABI contraction / attention /
norm launch is bracketed by CUDA events inside ONE eager call (the device is parked behind a spin so host
launch gaps do not count), aggregated per shape class.  Runs once per tcgen05 kernel variant (one-CTA vs
CTA-pair) and reports the eps difference between the two.  Measurement tool, not product code."""
import argparse, collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--latent", type=int, default=64)
ap.add_argument("--what", default="unet", choices=["unet", "vae"])
ap.add_argument("--variants", default="0,1")
ap.add_argument("--json", default=None)
a = ap.parse_args()

from sdb200 import _lib
from sdb200.pipeline import SD_UNET_CONFIG, SD_VAE_DDCONFIG
from sdb200.openai_model import UNetModel
from sdb200.autoencoder import AutoencoderKL

torch.manual_seed(0)
dev = torch.device("cuda:0")
lib = _lib.load()
if a.what == "unet":
    net = UNetModel(**SD_UNET_CONFIG, compute_mode="bf16")
    for m in net.modules():
        if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear)) and float(m.weight.detach().abs().max()) == 0.0:
            m.reset_parameters()
    net = net.to(dev)
    x = torch.randn(a.batch, 4, a.latent, a.latent, device=dev)
    t = torch.full((a.batch,), 500, device=dev)
    c = torch.randn(a.batch, 77, 768, device=dev)
    run = lambda: net(x, t, c)
else:
    net = AutoencoderKL(ddconfig=SD_VAE_DDCONFIG, embed_dim=4, compute_mode="bf16").to(dev)
    z = torch.randn(a.batch, 4, a.latent, a.latent, device=dev)
    run = lambda: net.decode(z)

NAMES = ["sdb_tc_contract", "sdb_attention_fwd", "sdb_attention_wide_fwd", "sdb_groupnorm_nhwc", "sdb_groupnorm_from_colstats", "sdb_layernorm", "sdb_cast_concat",
         "sdb_simt_contract", "sdb_skinny_linear"]


def describe(name, args):
    if name == "sdb_tc_contract":
        o = args[0]._obj
        kind = "conv%dx%d s%d" % (o.taps // max(o.kw, 1), o.kw, o.stride) if o.taps else ("geglu" if o.geglu else "gemm")
        return "%s M=%d N=%d K=%d%s%s" % (kind, o.M, o.N, o.K, " +res" if o.residual else "", " bf16out" if o.out_dtype == 1 else ""), 2.0 * o.M * o.N * o.K
    if name == "sdb_attention_fwd":
        o = args[0]._obj
        return "attn B=%d H=%d Sq=%d Sk=%d d=%d" % (o.B, o.H, o.Sq, o.Sk, o.d), 4.0 * o.B * o.H * o.Sq * o.Sk * o.d
    if name == "sdb_attention_wide_fwd":
        B, Sq, Sk, d = args[12], args[13], args[14], args[15]
        return "attn_wide B=%d Sq=%d Sk=%d d=%d" % (B, Sq, Sk, d), 4.0 * B * Sq * Sk * d
    if name == "sdb_groupnorm_nhwc":
        return "groupnorm N=%d HW=%d C=%d" % (args[4], args[5], args[1] + args[3]), 0.0
    if name == "sdb_groupnorm_from_colstats":
        return "groupnorm(colstats) N=%d HW=%d C=%d" % (args[8], args[9], args[1] + args[5]), 0.0
    if name == "sdb_layernorm":
        return "layernorm rows=%d C=%d" % (args[1], args[2]), 0.0
    if name == "sdb_cast_concat":
        return "cast_concat N=%d HW=%d C=%d up=%d" % (args[4], args[5] * args[6], args[1] + args[3], args[7]), 0.0
    if name == "sdb_simt_contract":
        o = args[0]._obj
        return "simt M=%d N=%d K=%d" % (o.M, o.N, o.K), 2.0 * o.M * o.N * o.K
    return name, 0.0


def timed_call(variant):
    lib.sdb_tc_set_pair_kernel(variant)
    for _ in range(2):
        out = run()
    torch.cuda.synchronize()
    recs = []
    orig = {}
    for n in NAMES:
        fn = getattr(lib, n)
        orig[n] = fn

        def make(n, fn):
            def w(*args):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = fn(*args)
                e1.record()
                d, fl = describe(n, args)
                recs.append((n, d, fl, e0, e1))
                return rc
            return w
        setattr(lib, n, make(n, fn))
    try:
        torch.cuda._sleep(int(4e8))
        E0, E1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        E0.record()
        out = run()
        E1.record()
        torch.cuda.synchronize()
    finally:
        for n in NAMES:
            setattr(lib, n, orig[n])
    agg = collections.OrderedDict()
    for n, d, fl, e0, e1 in recs:
        k = (n, d)
        v = agg.setdefault(k, [0, 0.0, 0.0])
        v[0] += 1
        v[1] += e0.elapsed_time(e1)
        v[2] += fl
    return out, agg, E0.elapsed_time(E1)


results = {}
outs = {}
for v in [int(s) for s in a.variants.split(",")]:
    out, agg, total = timed_call(v)
    outs[v] = out.float()
    print("==== variant pair_kernel=%d : one call %.3f ms (events incl. per-launch event overhead)" % (v, total))
    by_fn = collections.defaultdict(float)
    rows = []
    for (n, d), (cnt, ms, fl) in agg.items():
        by_fn[n] += ms
        rows.append((ms, cnt, d, fl))
    for ms, cnt, d, fl in sorted(rows, reverse=True):
        tf = " %7.1f TFLOP/s" % (fl / ms / 1e9) if fl else ""
        print("%9.3f ms x%-3d %s%s" % (ms, cnt, d, tf))
    print("by entry point:", {k: round(v_, 3) for k, v_ in by_fn.items()}, "sum %.3f" % sum(by_fn.values()))
    results[v] = {"total_ms": total, "by_fn": dict(by_fn), "rows": [(d, cnt, ms, fl) for ms, cnt, d, fl in rows]}
if a.what == "unet":
    net.compute_mode = "fp32"
    ref = run().float()
    for v in sorted(outs):
        print("variant %d vs fp32 SIMT mode: rel-L2 %.3e" % (v, float((outs[v] - ref).norm() / ref.norm())))
vs = sorted(outs)
if len(vs) == 2:
    d = (outs[vs[0]] - outs[vs[1]]).norm() / outs[vs[0]].norm()
    print("rel-L2 between variants: %.3e" % float(d))
if a.json:
    json.dump(results, open(a.json, "w"))
