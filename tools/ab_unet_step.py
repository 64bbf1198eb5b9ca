"""Same-box A/B of the SD-1.x UNet step (batch 8, bf16 mode, CUDA-graph replay): run once per environment setting inside ONE
gpurun call — clocks differ by 2-3 % between boxes (power capping), so numbers from different calls do not compare.
usage: [ENV=...] python tools/ab_unet_step.py [label]"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdb200.pipeline import SD_UNET_CONFIG
from sdb200.openai_model import UNetModel

label = sys.argv[1] if len(sys.argv) > 1 else ""
torch.manual_seed(0)
dev = torch.device("cuda:0")
net = UNetModel(**SD_UNET_CONFIG, compute_mode="bf16")
for m in net.modules():
    if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear)) and float(m.weight.detach().abs().max()) == 0.0:
        m.reset_parameters()
net = net.to(dev)
net.use_cuda_graph = True
B = int(os.environ.get("AB_BATCH", "8"))
x = torch.randn(B, 4, 64, 64, device=dev)
t = torch.full((B,), 500, device=dev)
c = torch.randn(B, 77, 768, device=dev)
for _ in range(5):
    net(x, t, c)
torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        net(x, t, c)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
print("%-28s UNet step %.3f ms (best of 5 x 20 replays), sm clock after %s MHz" % (label, best, clk))
