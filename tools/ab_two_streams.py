"""Experiment: one batch-8 UNet graph vs two batch-4 UNet graphs replayed concurrently on two streams (same weights object is not
shared: two model copies, so weight traffic doubles — a pessimistic bound for an in-model batch split)."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdb200.pipeline import SD_UNET_CONFIG
from sdb200.openai_model import UNetModel
torch.manual_seed(0)
dev = torch.device("cuda:0")
net = UNetModel(**SD_UNET_CONFIG, compute_mode="bf16")
for m in net.modules():
    if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear)) and float(m.weight.detach().abs().max()) == 0.0:
        m.reset_parameters()
net = net.to(dev)
net.use_cuda_graph = True
net2 = copy.deepcopy(net)
net2.use_cuda_graph = True
def inputs(B):
    return torch.randn(B, 4, 64, 64, device=dev), torch.full((B,), 500, device=dev), torch.randn(B, 77, 768, device=dev)
x8, t8, c8 = inputs(8)
xa, ta, ca = inputs(4)
xb, tb, cb = inputs(4)
for _ in range(4):
    net(x8, t8, c8)
torch.cuda.synchronize()
def timeit(fn, n=20):
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best
print("one graph, batch 8: %.3f ms" % timeit(lambda: net(x8, t8, c8)))
for _ in range(4):
    net(xa, ta, ca); net2(xb, tb, cb)
torch.cuda.synchronize()
print("batch 4 alone: %.3f ms" % timeit(lambda: net(xa, ta, ca)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def two():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        net(xa, ta, ca)
    with torch.cuda.stream(s2):
        net2(xb, tb, cb)
    cur.wait_stream(s1); cur.wait_stream(s2)
for _ in range(3):
    two()
print("two concurrent batch-4 graphs: %.3f ms per pair" % timeit(two))
