"""GPU: bandwidth / elementwise kernels of libsdb200.so against plain torch math and the reference's
DDIM step fixtures (bit-exact)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import restate as R
from oracle.golden import load_golden
from gpu_util import bf16_round, nchw, nhwc, randn, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,H,W,C0,C1,act,eps", [
    (2, 16, 16, 320, 0, 1, 1e-5), (1, 8, 8, 640, 1280, 1, 1e-5), (3, 5, 7, 128, 0, 0, 1e-6),
    (2, 64, 64, 320, 640, 1, 1e-5), (1, 1, 1, 512, 512, 1, 1e-5), (2, 32, 32, 2560, 0, 0, 1e-6), (1, 128, 128, 128, 0, 1, 1e-6)])
def test_groupnorm(cuda, N, H, W, C0, C1, act, eps):
    from sdb200 import ops
    x0 = randn(N, H, W, C0, seed=1) * 2 + 0.5
    x1 = randn(N, H, W, C1, seed=2) if C1 else None
    C = C0 + C1
    g, b = randn(C, seed=3) * 0.1 + 1, randn(C, seed=4) * 0.1
    xc = torch.cat([x0, x1], -1) if C1 else x0
    ref = F.group_norm(nchw(xc).double(), 32, g.double(), b.double(), eps)
    if act:
        ref = F.silu(ref)
    ref = nhwc(ref)
    out32 = ops.groupnorm(x0, g, b, eps, act=act, out_dtype=torch.float32, x1=x1, exact=True)
    assert rel(out32, ref) < 2e-6
    out16 = ops.groupnorm(x0, g, b, eps, act=act, out_dtype=torch.bfloat16, x1=x1)
    assert rel(out16, ref) < 4e-3


@pytest.mark.parametrize("rows,C", [(64, 320), (1000, 640), (77, 1280), (5, 64)])
def test_layernorm(cuda, rows, C):
    from sdb200 import ops
    x = randn(rows, C, seed=5) * 3 + 1
    g, b = randn(C, seed=6) * 0.1 + 1, randn(C, seed=7) * 0.1
    ref = F.layer_norm(x.double(), (C,), g.double(), b.double(), 1e-5)
    assert rel(ops.layernorm(x, g, b, 1e-5), ref) < 2e-6
    assert rel(ops.layernorm(x, g, b, 1e-5, out_dtype=torch.bfloat16), ref) < 4e-3


def test_cast_concat_and_upsample(cuda):
    from sdb200 import ops
    x0, x1 = randn(2, 6, 5, 64, seed=8), randn(2, 6, 5, 128, seed=9)
    out = ops.cast_concat(x0, x1, up=1, out_dtype=torch.float32)
    assert torch.equal(out, torch.cat([x0, x1], -1))
    up = ops.cast_concat(x0, None, up=2, out_dtype=torch.bfloat16)
    ref = nhwc(F.interpolate(nchw(x0), scale_factor=2, mode="nearest")).to(torch.bfloat16)
    assert torch.equal(up, ref)
    bl = ops.upsample_bilinear2x(x0)
    refb = nhwc(F.interpolate(nchw(x0), scale_factor=2.0, mode="bilinear", align_corners=True))
    assert rel(bl, refb) < 1e-6
    one = randn(2, 1, 1, 64, seed=10)
    assert rel(ops.upsample_bilinear2x(one), nhwc(F.interpolate(nchw(one), scale_factor=2.0, mode="bilinear", align_corners=True))) < 1e-6


def test_small_elementwise(cuda):
    from sdb200 import ops
    x = randn(7, 3, 5, 32, seed=11)
    rv = randn(7, 100, seed=12)
    assert torch.equal(ops.add_rowvec(x, rv[:, 10:42]), x + rv[:, None, None, 10:42])
    assert torch.equal(ops.add(x, x * 2), x + x * 2)
    assert rel(ops.activation(x, 1), F.silu(x)) < 1e-6
    assert rel(ops.activation(x, 2), F.gelu(x)) < 1e-6
    h = randn(9, 2 * 48, seed=13)
    a, gate = h.chunk(2, -1)
    assert rel(ops.geglu(h), a * F.gelu(gate)) < 1e-6
    s = randn(33, 77, seed=14) * 4
    assert rel(ops.softmax_rows(s, 0.37), torch.softmax(s.double() * 0.37, -1)) < 1e-6
    big = randn(3, 4096, seed=15) * 4
    assert rel(ops.softmax_rows(big, 0.1, out_dtype=torch.bfloat16), torch.softmax(big.double() * 0.1, -1)) < 4e-3
    xi = randn(3, 5, 6, 7, seed=16)
    assert torch.equal(ops.nchw_to_nhwc(xi), nhwc(xi))
    assert torch.equal(ops.nhwc_to_nchw(nhwc(xi)), xi)
    assert torch.equal(ops.nchw_to_nhwc(xi, out_dtype=torch.bfloat16), nhwc(xi).to(torch.bfloat16))
    padded = ops.nchw_to_nhwc(xi, out_dtype=torch.bfloat16, pad_to=32)          # latent -> 32-channel tensor-core operand
    assert padded.shape == (3, 6, 7, 32) and torch.equal(padded[..., :5], nhwc(xi).to(torch.bfloat16))
    assert float(padded[..., 5:].abs().max()) == 0.0
    # split operand: bf16(x) | bf16(x - bf16(x)) | zeros — the two halves together carry x to ~2^-17
    sp = ops.nchw_to_nhwc(xi, out_dtype=torch.bfloat16, pad_to=32, split=True)
    hi = nhwc(xi).to(torch.bfloat16)
    assert sp.shape == (3, 6, 7, 32) and torch.equal(sp[..., :5], hi)
    assert torch.equal(sp[..., 5:10], (nhwc(xi) - hi.float()).to(torch.bfloat16))
    assert float(sp[..., 10:].abs().max()) == 0.0
    assert rel(sp[..., :5].float() + sp[..., 5:10].float(), nhwc(xi)) < 2e-5


def test_timestep_embedding_and_skinny(cuda):
    from sdb200 import ops
    t = torch.tensor([981, 1, 500, 0], dtype=torch.long)
    ref = R.timestep_embedding(t, 320)
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half).cuda()
    got = ops.timestep_embedding(t.float().cuda(), freqs, round_fp16=False)
    assert rel(got, ref) < 1e-6
    got16 = ops.timestep_embedding(t.float().cuda(), freqs, round_fp16=True)
    assert float((got16.cpu() - ref.half().float()).abs().max()) <= 2e-3    # at most one fp16 ulp from the reference's rounding
    tbl = randn(1000, 128, seed=17)
    assert torch.equal(ops.gather_rows(tbl, t.cuda()), tbl[t.cuda()])
    for M in (1, 4, 8, 13, 40):
        x, Wt, b = randn(M, 1280, seed=18), randn(640, 1280, seed=19) * 0.03, randn(640, seed=20)
        ref = F.linear(F.silu(x.double()), Wt.double(), b.double())
        assert rel(ops.skinny_linear(x, Wt, b, act_in=1), ref) < 2e-6
        assert rel(ops.skinny_linear(x, Wt, b, act_out=1), F.silu(F.linear(x.double(), Wt.double(), b.double()))) < 2e-6
        assert rel(ops.skinny_linear(x, Wt, None, act_out=2), F.gelu(F.linear(x.double(), Wt.double()))) < 2e-6
        # bf16 weight storage (the bf16 mode's emb_layers matrix): exact against the same bf16-rounded weights
        W16 = Wt.to(torch.bfloat16)
        assert rel(ops.skinny_linear(x, W16, b, act_in=1), F.linear(F.silu(x.double()), W16.double(), b.double())) < 2e-6
    x, W16 = randn(8, 1288, seed=21), (randn(77, 1288, seed=22) * 0.03).to(torch.bfloat16)       # K not a multiple of 256, ragged N
    assert rel(ops.skinny_linear(x, W16), F.linear(x.double(), W16.double())) < 2e-6
    # many columns (the 20160 x 1280 emb_layers matrix): four columns per warp, ragged last group
    for M in (3, 8):
        x, Wb, bb = randn(M, 1280, seed=23), randn(9603, 1280, seed=24) * 0.03, randn(9603, seed=25)
        assert rel(ops.skinny_linear(x, Wb, bb, act_in=1), F.linear(F.silu(x.double()), Wb.double(), bb.double())) < 2e-6
        W16 = Wb.to(torch.bfloat16)
        assert rel(ops.skinny_linear(x, W16, bb, act_in=1), F.linear(F.silu(x.double()), W16.double(), bb.double())) < 2e-6


def test_ddim_step_bit_exact_vs_reference(cuda):
    """Every single-step fixture (eta = 0 / > 0, CFG on/off) produced by the unmodified reference."""
    from sdb200.ddim import DDIMSampler
    g = load_golden("ddim.pt")
    n = 0
    for key, st in g.items():
        if not key.startswith("step."):
            continue
        S = int(key.split(".S")[1].split(".")[0])
        eta = float(key.split(".eta")[1].split(".cfg")[0])
        cfg = float(key.split(".cfg")[1].split(".i")[0])
        from oracle.make_golden import toy_model_fn
        calls = []

        class Shim(R.ModelShim):
            def apply_model(self, x, t, c):
                calls.append(1)
                return toy_model_fn(x.cpu(), t.cpu(), c.cpu()).cuda()     # eps computed like the fixture did (CPU)

        shim = Shim(None, R.sd_alphas_cumprod(), device="cuda")
        shim.betas = shim.betas.cuda()
        s = DDIMSampler(shim)
        s.make_schedule(S, ddim_eta=eta, verbose=False)
        torch.manual_seed(0)
        import sdb200.ddim as dd
        orig = torch.randn
        try:
            torch.randn = lambda *a, **k: st["noise"].cuda()               # the reference's draw for this step
            xp, p0 = s.p_sample_ddim(st["x"].cuda(), st["c"].cuda(), torch.full((2,), st["step"], device="cuda"),
                                     index=st["index"], unconditional_guidance_scale=cfg,
                                     unconditional_conditioning=None if st["uc"] is None else st["uc"].cuda())
        finally:
            torch.randn = orig
        assert torch.equal(p0.cpu(), st["pred_x0"]), key
        assert torch.equal(xp.cpu(), st["x_prev"]), key
        assert len(calls) == 1        # classifier-free guidance is ONE call on the concatenated 2B batch, like ddim.py:176-179
        n += 1
    assert n == 12


def test_ddim_trajectory_vs_reference(cuda):
    from oracle.make_golden import toy_model_fn
    from sdb200.ddim import DDIMSampler
    g = load_golden("ddim.pt")

    class Shim(R.ModelShim):
        def apply_model(self, x, t, c):
            return toy_model_fn(x.cpu(), t.cpu(), c.cpu()).cuda()

    for S, cfg in ((10, 1.0), (50, 1.0), (10, 5.0)):
        t = g["traj.S%d.cfg%g" % (S, cfg)]
        shim = Shim(None, R.sd_alphas_cumprod(), device="cuda")
        shim.betas = shim.betas.cuda()
        z, inter = DDIMSampler(shim).sample(S, 3, (4, 8, 8), conditioning=t["c"].cuda(), verbose=False, x_T=t["x_T"].cuda(), eta=0.,
                                            unconditional_guidance_scale=cfg,
                                            unconditional_conditioning=None if t["uc"] is None else t["uc"].cuda())
        # The toy eps-model runs on THIS host's CPU (libm/SLEEF sin/cos can differ in the last ulp between
        # hosts), so bit-exactness is checked against the oracle sampler run here with the same eps-model ...
        zo, _ = R.DDIMOracle(R.ModelShim(toy_model_fn, R.sd_alphas_cumprod())).sample(
            S, 3, (4, 8, 8), conditioning=t["c"], eta=0., x_T=t["x_T"], unconditional_guidance_scale=cfg,
            unconditional_conditioning=t["uc"])
        assert torch.equal(z.cpu(), zo)                                       # whole trajectory, bit for bit
        assert len(inter["x_inter"]) == t["n_inter"]
        # ... and against the reference's committed trajectory up to that host-libm difference
        assert rel(z, t["z"]) < 1e-6 and rel(inter["pred_x0"][-1], t["last_pred_x0"]) < 1e-6


def test_ddim_inpainting_branch_vs_reference(cuda):
    """mask / x0 branch of ddim_sampling (ldm/diffusion/ddim.py:144-149) against trajectories of the UNMODIFIED reference sampler
    (tests/golden/ddim_mask.pt): q_sample + blend run as the fused sdb_inpaint_blend kernel, whole trajectory bit for bit
    against the oracle sampler run here, and at the host-libm level against the committed reference trajectory."""
    from oracle.make_golden import toy_model_fn
    from sdb200.ddim import DDIMSampler
    g = load_golden("ddim_mask.pt")

    class Shim(R.ModelShim):
        def apply_model(self, x, t, c):
            return toy_model_fn(x.cpu(), t.cpu(), c.cpu()).cuda()

    for key, t in g.items():
        S = int(key.split(".S")[1].split(".")[0])
        fixed = t["q_noise"]
        shim = Shim(None, R.sd_alphas_cumprod(), device="cuda")
        shim.betas = shim.betas.cuda()
        base = shim.q_sample
        shim.q_sample = lambda xs, ts, noise=None: base(xs.cpu(), ts.cpu(), noise=fixed).cuda()      # honoured as given: only the blend is fused
        z, _ = DDIMSampler(shim).sample(S, 2, (4, 8, 8), conditioning=t["c"].cuda(), verbose=False, x_T=t["x_T"].cuda(), eta=0.,
                                        mask=t["mask"].cuda(), x0=t["x0"].cuda())
        orc_shim = R.ModelShim(toy_model_fn, R.sd_alphas_cumprod())
        ob = orc_shim.q_sample
        orc_shim.q_sample = lambda xs, ts, noise=None: ob(xs, ts, noise=fixed)
        zo, _ = R.DDIMOracle(orc_shim).sample(S, 2, (4, 8, 8), conditioning=t["c"], eta=0., x_T=t["x_T"], mask=t["mask"], x0=t["x0"])
        assert torch.equal(z.cpu(), zo), key
        assert rel(z, t["z"]) < 1e-6, key


def test_inpaint_blend_and_q_sample_fused(cuda):
    """sdb_inpaint_blend with per-sample q_sample coefficients (the LatentDiffusion fast path) against the eager arithmetic of
    ddim.py:146-149 + ddpm.py:411-412 on the same draws: bit-exact, for a [B,1,H,W] and a [B,C,H,W] mask; and
    LatentDiffusion.q_sample (uniform default draw as the reference writes it) against the oracle's restatement."""
    from sdb200 import ops
    from sdb200.pipeline import LatentDiffusion
    ld = LatentDiffusion(first_stage_config=False, unet=torch.nn.Linear(1, 1)).cuda()
    B = 3
    x0, img, noise = randn(B, 4, 16, 16, seed=1), randn(B, 4, 16, 16, seed=2), torch.rand(B, 4, 16, 16, generator=torch.Generator().manual_seed(3)).cuda()
    ts = torch.tensor([981, 500, 1], device="cuda")
    a, c = ld.q_sample_coefficients(ts)
    shim = R.ModelShim(None, R.sd_alphas_cumprod())
    want_q = shim.q_sample(x0.cpu(), ts.cpu(), noise=noise.cpu())
    assert torch.equal(ld.q_sample(x0, ts, noise=noise).cpu(), want_q)
    for cm in (1, 4):
        mask = (randn(B, cm, 16, 16, seed=4) > 0).float() * 0.75
        want = want_q * mask.cpu() + (1. - mask.cpu()) * img.cpu()
        got = ops.inpaint_blend(x0, noise, a, c, mask.contiguous(), img)
        assert torch.equal(got.cpu(), want), cm
    torch.manual_seed(5)
    u = torch.rand_like(x0)
    torch.manual_seed(5)
    assert torch.equal(ld.q_sample(x0, ts), ld.q_sample(x0, ts, noise=u))          # default draw = torch.rand_like (ddpm.py:409)


def test_ddim_noise_dropout_quantize_and_rng(cuda):
    """The optional branches of p_sample_ddim (ldm/diffusion/ddim.py:198-204) and the sampler's RNG consumption:
    noise_dropout, quantize_denoised (duck-typed first_stage_model.quantize) against the eager arithmetic on the same
    draws; after sample() at eta = 0 the CUDA generator has advanced exactly as if randn(shape) had been drawn every
    step, which is what the reference does (noise_like, util.py:264-267)."""
    from sdb200.ddim import DDIMSampler
    fn = lambda x, t, c: 0.3 * x + 0.05
    shim = R.ModelShim(fn, R.sd_alphas_cumprod(), device="cuda")
    shim.betas = shim.betas.cuda()

    class FS:
        @staticmethod
        def quantize(p):
            return torch.round(p * 4) / 4, None, None
    shim.first_stage_model = FS()
    s = DDIMSampler(shim)
    s.make_schedule(20, ddim_eta=0.7, verbose=False)
    x = randn(2, 4, 8, 8, seed=7)
    index = 11
    ts = torch.full((2,), int(s.ddim_timesteps[index]), device="cuda")
    a_t = torch.full((2, 1, 1, 1), s.ddim_alphas[index], device="cuda")
    a_prev = torch.full((2, 1, 1, 1), s.ddim_alphas_prev[index], device="cuda")
    sigma_t = torch.full((2, 1, 1, 1), s.ddim_sigmas[index], device="cuda")
    s1m = torch.full((2, 1, 1, 1), s.ddim_sqrt_one_minus_alphas[index], device="cuda")
    for kw in (dict(noise_dropout=0.25), dict(quantize_denoised=True), dict(quantize_denoised=True, noise_dropout=0.5, temperature=0.8)):
        torch.manual_seed(11)
        xp, p0 = s.p_sample_ddim(x, None, ts, index=index, **kw)
        torch.manual_seed(11)
        e_t = fn(x, ts, None)
        pred_x0 = (x - s1m * e_t) / a_t.sqrt()
        if kw.get("quantize_denoised"):
            pred_x0 = FS.quantize(pred_x0)[0]
        dir_xt = (1. - a_prev - sigma_t ** 2).sqrt() * e_t
        noise = sigma_t * torch.randn(x.shape, device="cuda") * kw.get("temperature", 1.)
        if kw.get("noise_dropout", 0.) > 0.:
            noise = torch.nn.functional.dropout(noise, p=kw["noise_dropout"])
        want = a_prev.sqrt() * pred_x0 + dir_xt + noise
        assert torch.equal(p0, pred_x0) and torch.equal(xp, want), kw
    # RNG consumption at eta = 0
    smp = DDIMSampler(shim)
    assert smp.consume_rng_like_reference
    torch.manual_seed(3)
    smp.sample(10, 2, (4, 8, 8), conditioning=None, verbose=False, x_T=x, eta=0.)
    after = torch.randn(4, device="cuda")
    torch.manual_seed(3)
    for _ in range(10):
        torch.randn((2, 4, 8, 8), device="cuda")
    assert torch.equal(after, torch.randn(4, device="cuda"))


def test_ddim_step_rejects_wrong_dtype(cuda):
    """Every operand is read as n contiguous fp32 values: an fp16 e_uncond (what a UNet returns for fp16 latents) must be
    rejected by the op and normalised by the sampler, never silently mis-read."""
    from sdb200 import ops
    from sdb200._lib import SdbError
    x = randn(2, 4, 8, 8, seed=1)
    with pytest.raises(SdbError):
        ops.ddim_step(x, x, 1.0, 1.0, 0.5, 0.0, 0.5, e_uncond=x.half())
    with pytest.raises(SdbError):
        ops.ddim_step(x, x, 1.0, 1.0, 0.5, 0.3, 0.5, noise=x[:1])
