#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE (not product, not collected by pytest) — the "existing Blackwell library kernels" bar.

SURVEY.md K1/K5 and BASELINE.md §4 ask how the hand-written sm_100a kernels compare with what the reference's own GPU
path would have used on the same box: cuDNN convolutions, cuBLASLt GEMMs, F.scaled_dot_product_attention and
flash_attn_func (FA2 compiled for sm_100), ATen GroupNorm / LayerNorm.  This script times

  1. per-class microbenchmarks at the SD-1.x shapes (conv3x3, short-K GEMM, attention, norms), library op vs sdb200 op
     on identical bf16 inputs;
  2. one whole batch-8 UNet call and one batch-8 VAE decode of the oracle restatement executed ON THE GPU in bf16
     channels-last (eager and as a CUDA graph), against the sdb200 modules.

    python tests/library_bar.py [--out profiles/r02_library_bar.txt] [--skip-whole]

Every number is a CUDA-event time over `iters` back-to-back launches after warm-up, inputs rotated through a ring of
buffers larger than L2 for the bandwidth-class ops.
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1000.0      # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-whole", action="store_true")
    a = ap.parse_args()
    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    import __graft_entry__ as ge
    ge.build()
    from sdb200 import ops
    lines = []

    def emit(s):
        print(s, flush=True)
        lines.append(s)

    emit("# library bar on %s, torch %s, cudnn %s" % (torch.cuda.get_device_name(0), torch.__version__, torch.backends.cudnn.version()))
    try:
        from flash_attn import flash_attn_func
        import flash_attn
        emit("# flash_attn %s loaded" % flash_attn.__version__)
    except Exception as e:                                           # noqa: BLE001
        flash_attn_func = None
        emit("# flash_attn unavailable: %r" % (e,))
    torch.backends.cudnn.benchmark = True
    g = torch.Generator(device=dev).manual_seed(0)
    bf = torch.bfloat16

    def rnd(*shape, dtype=bf):
        return (torch.randn(*shape, device=dev, generator=g, dtype=torch.float32) * 0.5).to(dtype)

    # ---------------- conv3x3 / conv1x1 (cuDNN, channels-last bf16) vs sdb conv_tc ----------------
    emit("\n## conv (B=8): cuDNN bf16 channels-last F.conv2d vs sdb200 conv_tc (fp32 NHWC out + bias)")
    emit("%-44s %10s %10s %8s %8s" % ("shape", "cudnn us", "sdb us", "cudnn TF", "sdb TF"))
    for (Cin, Cout, HW, k) in [(320, 320, 64, 3), (640, 640, 32, 3), (1280, 1280, 16, 3), (1280, 1280, 8, 3), (640, 320, 64, 3),
                               (2560, 1280, 16, 3), (128, 128, 512, 3), (256, 256, 256, 3), (512, 512, 128, 3), (512, 512, 64, 3),
                               (320, 320, 64, 1), (640, 640, 32, 1), (1280, 1280, 16, 1)]:
        B = 8
        x = rnd(B, Cin, HW, HW).contiguous(memory_format=torch.channels_last)
        w = (rnd(Cout, Cin, k, k) / (Cin * k * k) ** 0.5 * 2).contiguous(memory_format=torch.channels_last)
        b = rnd(Cout)
        t_lib = timeit(lambda: F.conv2d(x, w, b, padding=k // 2))
        xn = x.permute(0, 2, 3, 1).contiguous()
        wp = ops.pack_conv_weight(w, bf)
        bfp = b.float()
        t_sdb = timeit(lambda: ops.conv_tc(xn, wp, bfp, k, k, pad=k // 2))
        fl = 2.0 * B * HW * HW * Cout * Cin * k * k
        emit("%-44s %10.1f %10.1f %8.0f %8.0f" % ("conv%dx%d %d->%d @%d^2" % (k, k, Cin, Cout, HW), t_lib, t_sdb, fl / t_lib / 1e6, fl / t_sdb / 1e6))
        del x, w, xn, wp

    # ---------------- GEMM (cuBLASLt via F.linear) vs sdb gemm_tc ----------------
    emit("\n## GEMM: cuBLASLt bf16 F.linear (+bias, bf16 out) vs sdb200 gemm_tc (bf16 out / fp32 out + fp32 residual)")
    emit("%-44s %10s %10s %10s %8s %8s" % ("M x N x K", "cublas us", "sdb bf16", "sdb f32+r", "lib TF", "sdb TF"))
    for (M, N, K) in [(32768, 320, 320), (32768, 960, 320), (32768, 2560, 320), (32768, 320, 1280), (8192, 640, 640),
                      (8192, 1920, 640), (8192, 5120, 640), (8192, 640, 2560), (2048, 1280, 1280), (2048, 3840, 1280),
                      (2048, 10240, 1280), (2048, 1280, 5120)]:
        A = rnd(M, K)
        Wt = rnd(N, K) / K ** 0.5
        b = rnd(N)
        res = rnd(M, N, dtype=torch.float32)
        t_lib = timeit(lambda: F.linear(A, Wt, b))
        bfp = b.float()
        t_sdb = timeit(lambda: ops.gemm_tc(A, Wt, bfp, out_dtype=bf, rows_per_item=M // 8))
        t_sdb_r = timeit(lambda: ops.gemm_tc(A, Wt, bfp, residual=res, rows_per_item=M // 8))
        fl = 2.0 * M * N * K
        emit("%-44s %10.1f %10.1f %10.1f %8.0f %8.0f" % ("%d x %d x %d" % (M, N, K), t_lib, t_sdb, t_sdb_r, fl / t_lib / 1e6, fl / t_sdb / 1e6))

    # ---------------- attention ----------------
    emit("\n## attention (B=8, H=8): SDPA / flash_attn_func 2.x vs sdb200 attention_tc")
    emit("%-30s %10s %10s %10s %8s %8s" % ("Sq x Sk x d", "sdpa us", "fa2 us", "sdb us", "best lib TF", "sdb TF"))
    for (Sq, Sk, d) in [(4096, 4096, 40), (1024, 1024, 80), (256, 256, 160), (4096, 77, 40), (1024, 77, 80), (256, 77, 160),
                        (9216, 9216, 40)]:
        B, H = 8, 8
        q, k, v = rnd(B, Sq, H, d), rnd(B, Sk, H, d), rnd(B, Sk, H, d)
        scale = d ** -0.5
        qt, kt, vt = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
        t_sdpa = timeit(lambda: F.scaled_dot_product_attention(qt, kt, vt, scale=scale))
        t_fa = float("nan")
        if flash_attn_func is not None:
            try:
                t_fa = timeit(lambda: flash_attn_func(q, k, v, softmax_scale=scale, causal=False))
            except Exception as e:                                   # noqa: BLE001
                emit("# flash_attn_func failed at d=%d: %r" % (d, str(e)[:100]))
        dp = (d + 63) // 64 * 64
        C = H * d
        q2, k2, v2 = q.reshape(B * Sq, C), k.reshape(B * Sk, C), v.reshape(B * Sk, C)
        t_sdb = timeit(lambda: ops.attention_tc(q2, k2, v2, B, H, Sq, Sk, d, dp, scale, (Sq * C, C, d), (Sk * C, C, d), (Sk * C, C, d), dense=True))
        fl = 4.0 * B * H * Sq * Sk * d
        best = min(x for x in (t_sdpa, t_fa) if x == x)
        emit("%-30s %10.1f %10.1f %10.1f %8.0f %8.0f" % ("%d x %d x %d" % (Sq, Sk, d), t_sdpa, t_fa, t_sdb, fl / best / 1e6, fl / t_sdb / 1e6))

    # ---------------- norms ----------------
    emit("\n## norms (B=8): ATen vs sdb200 (fp32 NHWC in, bf16 out); GB/s = (4 B in + 2 B out) per element")
    emit("%-40s %10s %10s %10s" % ("shape", "aten us", "sdb us", "sdb GB/s"))
    for (Cc, HW) in [(320, 64), (640, 32), (1280, 16), (128, 512), (256, 256)]:
        B = 8
        ring = [rnd(B, HW, HW, Cc, dtype=torch.float32) for _ in range(max(2, int(300e6 // (B * HW * HW * Cc * 4)) + 1))]
        gam, bet = rnd(Cc, dtype=torch.float32), rnd(Cc, dtype=torch.float32)
        xb = ring[0].permute(0, 3, 1, 2).to(bf).contiguous(memory_format=torch.channels_last)
        gb, bb = gam.to(bf), bet.to(bf)
        t_lib = timeit(lambda: F.silu(F.group_norm(xb, 32, gb, bb, 1e-5)))
        i = [0]

        def sdb_gn():
            i[0] = (i[0] + 1) % len(ring)
            return ops.groupnorm(ring[i[0]], gam, bet, 1e-5, act=1, out_dtype=bf)
        t_sdb = timeit(sdb_gn)
        emit("%-40s %10.1f %10.1f %10.0f" % ("GroupNorm32+SiLU C=%d @%d^2" % (Cc, HW), t_lib, t_sdb, B * HW * HW * Cc * 6 / t_sdb / 1e3))
        del ring
    for (rows, Cc) in [(32768, 320), (8192, 640), (2048, 1280)]:
        ring = [rnd(rows, Cc, dtype=torch.float32) for _ in range(max(2, int(300e6 // (rows * Cc * 4)) + 1))]
        gam, bet = rnd(Cc, dtype=torch.float32), rnd(Cc, dtype=torch.float32)
        xb = ring[0].to(bf)
        gb, bb = gam.to(bf), bet.to(bf)
        t_lib = timeit(lambda: F.layer_norm(xb, (Cc,), gb, bb, 1e-5))
        i = [0]

        def sdb_ln():
            i[0] = (i[0] + 1) % len(ring)
            return ops.layernorm(ring[i[0]], gam, bet, 1e-5, out_dtype=bf)
        t_sdb = timeit(sdb_ln)
        emit("%-40s %10.1f %10.1f %10.0f" % ("LayerNorm rows=%d C=%d" % (rows, Cc), t_lib, t_sdb, rows * Cc * 6 / t_sdb / 1e3))
        del ring

    if not a.skip_whole:
        whole(emit, dev, flash_attn_func)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            f.write("\n".join(lines) + "\n")


def whole(emit, dev, flash_attn_func):
    """The oracle restatement (the reference's arithmetic, functional PyTorch) on the GPU in bf16 channels-last: every conv is
    cuDNN, every Linear cuBLASLt, attention SDPA or FA2, norms ATen — the library-kernel version of the hot path."""
    from oracle import restate as R
    from oracle import weights as W
    from oracle.golden import load_golden
    from sdb200.autoencoder import AutoencoderKL
    from sdb200.openai_model import UNetModel
    bf = torch.bfloat16
    B = 8
    gu, gv = load_golden("unet_sd.pt"), load_golden("vae_sd_z16.pt")
    sdu32 = W.make_state_dict(gu["key_shapes"], 31)
    sdv32 = W.make_state_dict(gv["key_shapes"], 51)

    def to_lib(sd):
        out = {}
        for k, v in sd.items():
            v = v.to(dev, bf)
            if v.dim() == 4:
                v = v.contiguous(memory_format=torch.channels_last)
            out[k] = v
        return out
    sdu, sdv = to_lib(sdu32), to_lib(sdv32)
    x = W.seeded_randn((B, 4, 64, 64), 3).to(dev)
    ctx = W.seeded_randn((B, 77, 768), 4).to(dev)
    t = torch.full((B,), 500, device=dev, dtype=torch.long)
    xb = x.to(bf).contiguous(memory_format=torch.channels_last)
    cb = ctx.to(bf)

    def attn_lib(kind):
        def cross_attention(sd, p, xx, context, heads):
            context = xx if context is None else context
            q, k, v = R._lin(sd, p + ".to_q", xx), R._lin(sd, p + ".to_k", context), R._lin(sd, p + ".to_v", context)
            b, n, inner = q.shape
            d = inner // heads
            if kind == "fa2":
                o = flash_attn_func(q.view(b, n, heads, d), k.view(b, -1, heads, d), v.view(b, -1, heads, d), softmax_scale=d ** -0.5)
                o = o.reshape(b, n, inner)
            else:
                o = F.scaled_dot_product_attention(q.view(b, n, heads, d).transpose(1, 2), k.view(b, -1, heads, d).transpose(1, 2),
                                                   v.view(b, -1, heads, d).transpose(1, 2), scale=d ** -0.5).transpose(1, 2).reshape(b, n, inner)
            return R._lin(sd, p + ".to_out.0", o)
        return cross_attention

    def vae_attn_sdpa(sd, p, xx):
        h = R._gn(sd, p + ".norm", xx, 1e-6)
        q, k, v = R._conv(sd, p + ".q", h), R._conv(sd, p + ".k", h), R._conv(sd, p + ".v", h)
        b, c, hh, ww = q.shape
        o = F.scaled_dot_product_attention(q.reshape(b, 1, c, hh * ww).transpose(2, 3), k.reshape(b, 1, c, hh * ww).transpose(2, 3),
                                           v.reshape(b, 1, c, hh * ww).transpose(2, 3), scale=c ** -0.5)
        o = o.transpose(2, 3).reshape(b, c, hh, ww)
        return xx + R._conv(sd, p + ".proj_out", o)

    emit("\n## whole modules, batch 8, bf16: oracle restatement on GPU library kernels vs sdb200")
    orig_ca, orig_va = R.cross_attention, R.vae_attn_block
    try:
        kinds = ["sdpa"] + (["fa2"] if flash_attn_func is not None else [])
        for kind in kinds:
            R.cross_attention = attn_lib(kind)
            with torch.no_grad():
                def unet_lib():
                    return R.unet_forward(sdu, R.SD_UNET_CFG, xb, t, cb)
                try:
                    t_eager = timeit(unet_lib, iters=5, warm=3)
                except Exception as e:                               # noqa: BLE001
                    emit("# library UNet (%s) failed: %r" % (kind, str(e)[:200]))
                    continue
                t_graph = float("nan")
                try:
                    gph = torch.cuda.CUDAGraph()
                    s = torch.cuda.Stream()
                    s.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(s):
                        unet_lib()
                    torch.cuda.current_stream().wait_stream(s)
                    # (the restatement builds the timestep frequencies with torch.arange on the host: inside a capture the
                    # factory functions must default to the GPU, or the host-to-device copy of that table aborts the capture)
                    with torch.cuda.graph(gph), torch.device(dev):
                        unet_lib()
                    t_graph = timeit(gph.replay, iters=10, warm=3)
                except Exception as e:                               # noqa: BLE001
                    emit("# library UNet graph capture (%s) failed: %r" % (kind, str(e)[:200]))
                emit("UNet step B=8 64x64, library kernels (%s attention): eager %.2f ms, CUDA graph %.2f ms" % (kind, t_eager / 1e3, t_graph / 1e3))
        R.vae_attn_block = vae_attn_sdpa
        z = W.seeded_randn((B, 4, 64, 64), 5).to(dev)
        zb = z.to(bf).contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            def dec_lib():
                return R.autoencoder_decode(sdv, R.SD_VAE_DDCONFIG, zb)
            try:
                t_dec = timeit(dec_lib, iters=3, warm=2)
                emit("VAE decode B=8 64x64->512x512, library kernels: eager %.2f ms" % (t_dec / 1e3))
            except Exception as e:                                   # noqa: BLE001
                emit("# library VAE decode failed: %r" % (str(e)[:200],))
    finally:
        R.cross_attention, R.vae_attn_block = orig_ca, orig_va

    net = UNetModel(**gu["cfg"])
    net.load_state_dict(sdu32)
    net = net.to(dev)
    net.compute_mode = "bf16"
    t_e = timeit(lambda: net(x, t, ctx), iters=5, warm=3)
    net.use_cuda_graph = True
    t_g = timeit(lambda: net(x, t, ctx), iters=10, warm=3)
    emit("UNet step B=8 64x64, sdb200: eager %.2f ms, CUDA graph %.2f ms" % (t_e / 1e3, t_g / 1e3))
    vae = AutoencoderKL(ddconfig=gv["ddconfig"], embed_dim=4)
    vae.load_state_dict(sdv32, strict=False)
    vae = vae.to(dev)
    vae.compute_mode = "bf16"
    z = W.seeded_randn((B, 4, 64, 64), 5).to(dev)
    t_v = timeit(lambda: vae.decode(z), iters=3, warm=2)
    emit("VAE decode B=8 64x64->512x512, sdb200: %.2f ms" % (t_v / 1e3))


if __name__ == "__main__":
    main()
