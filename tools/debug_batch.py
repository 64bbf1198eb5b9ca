"""Which of (sample alone, sample inside a batch) deviates at a ragged latent: compare both bf16 results with the fp32-mode run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import weights as W
from oracle.golden import load_golden
from sdb200.openai_model import UNetModel
latent = int(sys.argv[1]) if len(sys.argv) > 1 else 80
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g = load_golden("unet_sd.pt")
net = UNetModel(**g["cfg"], compute_mode="bf16")
net.load_state_dict(W.make_state_dict(g["key_shapes"], g["seed"]))
net = net.cuda()
x = W.seeded_randn((B, 4, latent, latent), 7).cuda()
ctx = W.seeded_randn((B, 77, 768), 8).cuda()
t = torch.tensor([981, 500, 21][:B], device="cuda")
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
full = net(x, t, ctx)
full2 = net(x, t, ctx)
print("deterministic (same batch twice):", bool(torch.equal(full, full2)))
if os.environ.get("QUICK"):
    a1 = net(x[1:2], t[1:2], ctx[1:2])
    print("QUICK in-batch[1] vs alone %.3e" % rel(full[1], a1[0]))
    sys.exit(0)
alone = [net(x[i:i + 1], t[i:i + 1], ctx[i:i + 1]) for i in range(B)]
net.compute_mode = "fp32"
ref = net(x, t, ctx)
ref1 = [net(x[i:i + 1], t[i:i + 1], ctx[i:i + 1]) for i in range(B)]
for i in range(B):
    print("sample %d: bf16 in-batch vs fp32 %.3e | bf16 alone vs fp32 %.3e | bf16 in-batch vs alone %.3e | fp32 in-batch vs alone %.3e"
          % (i, rel(full[i], ref[i]), rel(alone[i][0], ref[i]), rel(full[i], alone[i][0]), rel(ref[i], ref1[i][0])))
# where does it start: hook the per-block outputs
net.compute_mode = "bf16"
taps = {}
orig = net._run_block
def rb(seq, P, mode, h, x1, emb_all, context):
    out = orig(seq, P, mode, h, x1, emb_all, context)
    taps.setdefault("cur", []).append(out)
    return out
net._run_block = rb
taps["cur"] = []; net(x, t, ctx); a = taps["cur"]
taps["cur"] = []; net(x[1:2], t[1:2], ctx[1:2]); b = taps["cur"]
for k, (u, v) in enumerate(zip(a, b)):
    print("block %2d shape %s: in-batch[1] vs alone %.3e" % (k, tuple(u.shape), rel(u[1].float(), v[0].float())))
