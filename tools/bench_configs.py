"""BASELINE.json configs[2] (VAE decode, batch 16) and configs[4] (96x96 latent UNet step, batch sweep 1-32) as JSON lines:
device-timed (CUDA events, 3 warm-up + 5 timed calls, CUDA-graph replay for the UNet), model TFLOP/s from SURVEY.md §8d's
algorithmic work per sample and the fraction of the measured sustained bf16 peak.  Measurement tool, not product code."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdb200.pipeline import SD_UNET_CONFIG, SD_VAE_DDCONFIG
from sdb200.openai_model import UNetModel
from sdb200.autoencoder import AutoencoderKL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", 1341.2)
except Exception:
    PEAK = 1341.2
GFLOP = {64: 803.27, 96: 2148.12}          # UNet step per sample
VAE_GFLOP = 2514.52                         # decode per image
dev = torch.device("cuda:0")
torch.manual_seed(0)


def timed(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


net = UNetModel(**SD_UNET_CONFIG, compute_mode="bf16")
for m in net.modules():
    if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear)) and float(m.weight.detach().abs().max()) == 0.0:
        m.reset_parameters()
net = net.to(dev)
net.use_cuda_graph = True
for latent, batches in ((96, (1, 2, 4, 8, 16, 32)), (64, (1, 8, 16))):
    for B in batches:
        x = torch.randn(B, 4, latent, latent, device=dev)
        t = torch.full((B,), 500, device=dev)
        c = torch.randn(B, 77, 768, device=dev)
        ms = timed(lambda: net(x, t, c))
        tf = B * GFLOP[latent] / ms
        print(json.dumps({"config": "UNet step, latent %dx%dx4 (%dpx), ctx 77x768, bf16, CUDA graph" % (latent, latent, latent * 8), "batch": B,
                          "ms_per_step": round(ms, 3), "model_tflops": round(tf, 1), "frac_of_tensor_peak": round(tf / PEAK, 4),
                          "peak_tflops": PEAK}), flush=True)
        net._graphs.clear()
        torch.cuda.empty_cache()
del net
torch.cuda.empty_cache()
vae = AutoencoderKL(ddconfig=SD_VAE_DDCONFIG, embed_dim=4, compute_mode="bf16").to(dev)
for B in (8, 16):
    z = torch.randn(B, 4, 64, 64, device=dev)
    ms = timed(lambda: vae.decode(z), warm=2, reps=3)
    tf = B * VAE_GFLOP / ms
    print(json.dumps({"config": "VAE decode 64x64x4 -> 512x512x3, bf16 (micro-batches of %d)" % vae.micro_batch, "batch": B, "ms": round(ms, 2),
                      "images_per_s": round(B / ms * 1e3, 1), "model_tflops": round(tf, 1), "frac_of_tensor_peak": round(tf / PEAK, 4)}), flush=True)

# classifier-free guidance 7.5 (two UNet calls per DDIM step; SURVEY.md §8d asks for it to be reported separately)
if "--cfg" in sys.argv:
    from sdb200.pipeline import LatentDiffusion
    del vae
    torch.cuda.empty_cache()
    ld = LatentDiffusion(compute_mode="bf16")
    for m in ld.modules():
        if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear)) and float(m.weight.detach().abs().max()) == 0.0:
            m.reset_parameters()
    ld = ld.to(dev)
    ld.model.diffusion_model.use_cuda_graph = True
    B = 8
    c = torch.randn(B, 77, 768, device=dev)
    uc = torch.randn(B, 77, 768, device=dev)
    x_T = torch.randn(B, 4, 64, 64, device=dev)
    for scale in (1.0, 7.5):
        fn = lambda: ld.txt2img(c, B, ddim_steps=50, shape=(4, 64, 64), x_T=x_T, unconditional_guidance_scale=scale,
                                unconditional_conditioning=uc if scale != 1.0 else None)
        ms = timed(fn, warm=1, reps=1)
        print(json.dumps({"config": "txt2img DDIM-50 + decode through LatentDiffusion.txt2img, batch 8, guidance scale %.1f" % scale,
                          "batch": B, "ms": round(ms, 1), "images_per_s": round(B / ms * 1e3, 2),
                          "unet_calls_per_step": 1 if scale == 1.0 else 2}), flush=True)
