"""GPU: the parity claims of BASELINE.json at the FULL Stable-Diffusion-1.x shapes (configs[1] and configs[2]).

The fp32 (SIMT) mode is pinned to the reference at these shapes by test_gpu_models.py (unet_sd / vae_sd_z16 fixtures,
<= 1e-5).  Here the bf16 tensor-core mode runs the whole free-running pipeline — DDIM-50 on a 64x64x4 latent with a
77x768 context, then the 512x512x3 decode — next to the fp32 mode on the same weights, noise and conditioning:
per-step eps (teacher-forced on the fp32 trajectory) <= 1e-2, decoded image PSNR >= 40 dB.
"""
import pytest
import torch

from gpu_util import rel
from oracle import restate as R
from oracle import weights as W
from oracle.golden import load_golden

pytestmark = pytest.mark.gpu


def test_sd_ddim50_decode_bf16_vs_fp32(cuda):
    from sdb200.pipeline import LatentDiffusion
    gu, gv = load_golden("unet_sd.pt"), load_golden("vae_sd_z16.pt")
    ld = LatentDiffusion(unet_config=gu["cfg"], first_stage_config=gv["ddconfig"], compute_mode="fp32")
    ld.model.diffusion_model.load_state_dict(W.make_state_dict(gu["key_shapes"], gu["seed"]))
    r = ld.first_stage_model.load_state_dict(W.make_state_dict(gv["key_shapes"], gv["seed"]), strict=False)
    assert not r.unexpected_keys
    ld = ld.cuda()
    unet, vae = ld.model.diffusion_model, ld.first_stage_model
    B, S = 1, 50
    x_T = W.seeded_randn((B, 4, 64, 64), 2).cuda()
    ctx = W.seeded_randn((B, 77, 768), 3).cuda()

    # fp32 mode: free-running trajectory, recording (x_t, t, eps) of every step
    rec = []
    apply = ld.apply_model

    def recording(x, t, c):
        e = apply(x, t, c)
        rec.append((x.clone(), t.clone(), e.clone()))
        return e
    ld.apply_model = recording
    z32, img32 = ld.txt2img(ctx, B, ddim_steps=S, shape=(4, 64, 64), x_T=x_T)
    ld.apply_model = apply
    assert len(rec) == S and torch.isfinite(img32).all() and img32.shape == (B, 3, 512, 512)

    unet.compute_mode = "bf16"
    vae.compute_mode = "bf16"
    unet.use_cuda_graph = True
    worst = max(rel(ld.apply_model(x, t, ctx), e) for x, t, e in rec[::7] + [rec[-1]])      # teacher-forced per-step eps
    z16, img16 = ld.txt2img(ctx, B, ddim_steps=S, shape=(4, 64, 64), x_T=x_T)
    psnr = R.psnr_255(img16.cpu(), img32.cpu())
    print("SD-1.x DDIM-50 + decode, bf16 vs fp32 mode: worst teacher-forced eps rel-L2 %.3e, final latent rel-L2 %.3e, "
          "image PSNR %.1f dB" % (worst, rel(z16, z32), psnr))
    assert worst <= 1e-2
    assert psnr >= 40.0


def test_sd_vae_decode_full_size(cuda):
    """configs[2] shape: 64x64x4 -> 512x512x3, bf16 against the fp32 mode, and micro-batch invariance at full size."""
    from sdb200.autoencoder import AutoencoderKL
    g = load_golden("vae_sd_z16.pt")
    vae = AutoencoderKL(ddconfig=g["ddconfig"], embed_dim=4, compute_mode="fp32")
    vae.load_state_dict(W.make_state_dict(g["key_shapes"], g["seed"]), strict=False)
    vae = vae.cuda()
    z = W.seeded_randn((3, 4, 64, 64), 4).cuda()
    ref = vae.decode(z[:1])
    vae.compute_mode = "bf16"
    img = vae.decode(z)
    assert img.shape == (3, 3, 512, 512) and torch.isfinite(img).all()
    psnr = R.psnr_255(img[:1].cpu(), ref.cpu())
    print("SD VAE decode 512x512: bf16 vs fp32 mode rel-L2 %.3e, PSNR %.1f dB" % (rel(img[:1], ref), psnr))
    assert psnr >= 40.0
    vae.micro_batch = 2
    assert torch.equal(vae.decode(z), img)
