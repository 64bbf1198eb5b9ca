"""Launch-plan sweep for every distinct tcgen05 contraction of one SD-1.x UNet call (or VAE decode): for each
shape, time every (kernel variant, block_n, split_k) the library accepts, on synthetic operands rotated through
several copies so weights come from DRAM as they do in the real step.  Writes JSON lines.  Measurement tool."""
import argparse, ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--latent", type=int, default=64)
ap.add_argument("--what", default="unet", choices=["unet", "vae"])
ap.add_argument("--out", default="gpurun_out/tune_tc.jsonl")
ap.add_argument("--iters", type=int, default=12)
a = ap.parse_args()

from sdb200 import _lib, ops
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

FIELDS = ["M", "N", "K", "out_dtype", "geglu", "col_group", "col_group_stride", "taps", "kw", "stride", "pad_h", "pad_w",
          "NB", "IH", "IW", "Cin", "OH", "OW", "cout_pad", "rows_per_item", "block_n", "out_sh", "out_sw", "out_oh", "out_ow", "OHF", "OWF"]
shapes = {}
orig = lib.sdb_tc_contract


def hook(args, stream):
    o = args._obj
    key = tuple(getattr(o, f) for f in FIELDS) + (bool(o.residual), bool(o.rowvec), bool(o.bias))
    shapes[key] = shapes.get(key, 0) + 1
    return orig(args, stream)


lib.sdb_tc_contract = hook
run()
torch.cuda.synchronize()
lib.sdb_tc_contract = orig
del net
torch.cuda.empty_cache()
print("distinct contractions:", len(shapes), flush=True)


def timed(fn, nrot, iters, warm=None):
    """Device time per launch of `iters` BACK-TO-BACK launches (operands rotated through nrot copies), best of 3 rounds.  The
    device is parked behind a ~1 ms spin while the host enqueues the round, so the events bracket kernel execution only —
    launch latency (~20 us for an isolated launch, more than most of these kernels run) and host gaps do not count, and
    consecutive launches overlap through programmatic dependent launch as they do inside the captured UNet graph."""
    for i in range(3):
        fn(i % nrot)
    torch.cuda.synchronize()
    best = None
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(int(3e6))
        e0.record()
        for i in range(iters):
            fn(i % nrot)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        best = us if best is None else min(best, us)
    return best


out_f = open(a.out, "w")
for key, count in sorted(shapes.items(), key=lambda kv: -kv[0][0] * kv[0][1] * kv[0][2] * kv[1]):
    d = dict(zip(FIELDS, key[:len(FIELDS)]))
    has_res, has_rowvec, has_bias = key[len(FIELDS):]
    M, N, K = d["M"], d["N"], d["K"]
    conv = d["taps"] > 0
    obf = d["out_dtype"] == 1
    n_out = N // 2 if d["geglu"] else N
    width = (N // d["col_group"]) * d["col_group_stride"] if d["col_group"] else n_out
    bytes_ = M * (K if not conv else d["Cin"]) * 2 + N * K * 2 + M * width * (2 if obf else 4) * (2 if has_res else 1)
    nrot = max(1, min(6, int(300e6 // bytes_)))
    if conv:
        A = [torch.randn(d["NB"], d["IH"], d["IW"], d["Cin"], device=dev).to(torch.bfloat16) for _ in range(nrot)]
        W = [(torch.randn(d["taps"], N, d["Cin"], device=dev) / K ** 0.5).to(torch.bfloat16) for _ in range(nrot)]
        phase = (d["out_sh"], d["out_sw"], d["out_oh"], d["out_ow"], d["OHF"], d["OWF"], d["pad_h"], d["pad_w"]) if d["out_sh"] > 1 else None
        oshape = (d["NB"], d["OHF"], d["OWF"], N) if phase else (d["NB"], d["OH"], d["OW"], N)
    else:
        A = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(nrot)]
        W = [(torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16) for _ in range(nrot)]
        oshape = (M, width)
    Asrc = A[0].clone()
    warm = (lambda i: A[i].copy_(Asrc)) if A[0].numel() * 2 <= 48e6 else None
    stats = conv and not obf
    bias = torch.randn(N, device=dev) if has_bias else None
    R = [torch.randn(oshape, device=dev) for _ in range(nrot)] if has_res else None
    rowvec = torch.randn(d["NB"], N, device=dev) if has_rowvec else None
    O = [torch.zeros(oshape, device=dev, dtype=torch.bfloat16 if obf else torch.float32) for _ in range(nrot)]
    res = []
    bns = [d["block_n"]] if d["geglu"] else [64, 128, 160, 256]
    for variant in (1, 2):
        for bn in bns:
            if variant == 2 and bn < 128:
                continue
            if bn > 64 and bn >= 2 * N and bn != 128:
                continue
            for sk in ((1,) if (d["geglu"] or d["col_group"] or M > 8192) else (1, 2, 3, 4, 6, 8, 12, 16)):
                kblocks = max(d["taps"], 1) * ((K // max(d["taps"], 1) + 63) // 64)
                if sk > 1 and kblocks // sk < 4:
                    continue
                def fn(i):
                    if conv:
                        ops.conv_tc(A[i], W[i], bias, d["taps"] // d["kw"], d["kw"], stride=d["stride"], pad=d["pad_h"], rowvec=rowvec,
                                    residual=R[i] if R else None, out=O[i], split_k=sk, block_n=bn, variant=variant, want_stats=stats and sk == 1 and not phase, phase=phase)
                    else:
                        ops.gemm_tc(A[i], W[i], bias, residual=R[i] if R else None, geglu=bool(d["geglu"]), col_group=d["col_group"],
                                    col_group_stride=d["col_group_stride"], split_k=sk, block_n=bn, out=O[i],
                                    rows_per_item=d["rows_per_item"], variant=variant)
                try:
                    us = timed(fn, nrot, a.iters, warm)
                    res.append((variant, bn, sk, round(us, 2)))
                except Exception as e:
                    res.append((variant, bn, sk, None))
    # the library's own choice
    def fn0(i):
        if conv:
            ops.conv_tc(A[i], W[i], bias, d["taps"] // d["kw"], d["kw"], stride=d["stride"], pad=d["pad_h"], rowvec=rowvec,
                        residual=R[i] if R else None, out=O[i], block_n=d["block_n"], want_stats=stats and not phase, phase=phase)
        else:
            ops.gemm_tc(A[i], W[i], bias, residual=R[i] if R else None, geglu=bool(d["geglu"]), col_group=d["col_group"],
                        col_group_stride=d["col_group_stride"], block_n=d["block_n"], out=O[i], rows_per_item=d["rows_per_item"])
    auto_us = timed(fn0, nrot, a.iters, warm)
    ok = [r for r in res if r[3] is not None]
    best = min(ok, key=lambda r: r[3])
    rec = {"shape": d, "res": has_res, "rowvec": has_rowvec, "count": count, "auto_us": round(auto_us, 2), "best": best, "all": res,
           "tflops_best": round(2.0 * M * N * K / best[3] / 1e6, 1), "nrot": nrot}
    out_f.write(json.dumps(rec) + "\n")
    out_f.flush()
    print("%s M=%d N=%d K=%d res=%d obf=%d x%d: auto %.1f us | best v=%d bn=%d sk=%d %.1f us (%.0f TF)" % (
        "conv" if conv else ("geglu" if d["geglu"] else "gemm"), M, N, K, has_res, obf, count, auto_us, best[0], best[1], best[2],
        best[3], rec["tflops_best"]), flush=True)
    del A, W, O, R
