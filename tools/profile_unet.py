"""One eager (no CUDA graph) SD-1.x UNet call at batch B for ncu: warm-up calls, then one call inside
cudaProfilerStart/Stop (use ncu --profile-from-start off)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--latent", type=int, default=64)
ap.add_argument("--mode", default="bf16")
ap.add_argument("--what", default="unet", choices=["unet", "vae"])
a = ap.parse_args()
from sdb200 import _lib
from sdb200.pipeline import SD_UNET_CONFIG, SD_VAE_DDCONFIG
from sdb200.openai_model import UNetModel
from sdb200.autoencoder import AutoencoderKL
torch.manual_seed(0)
dev = torch.device("cuda:0")
if a.what == "unet":
    net = UNetModel(**SD_UNET_CONFIG, compute_mode=a.mode)
    for m in net.modules():
        if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear)) and float(m.weight.abs().max()) == 0.0:
            m.reset_parameters()
    net = net.to(dev)
    x = torch.randn(a.batch, 4, a.latent, a.latent, device=dev)
    t = torch.full((a.batch,), 500, device=dev)
    c = torch.randn(a.batch, 77, 768, device=dev)
    run = lambda: net(x, t, c)
else:
    net = AutoencoderKL(ddconfig=SD_VAE_DDCONFIG, embed_dim=4, compute_mode=a.mode).to(dev)
    z = torch.randn(a.batch, 4, a.latent, a.latent, device=dev)
    run = lambda: net.decode(z)
for _ in range(2):
    run()
torch.cuda.synchronize()
lib = _lib.load()
l0 = lib.sdb_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
run()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches", lib.sdb_launch_count() - l0, "ms", e0.elapsed_time(e1))
