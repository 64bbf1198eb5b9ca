"""GPU: the drop-in modules against the golden fixtures produced by the unmodified reference.
Tolerances are BASELINE.json's: per-step eps rel-L2 <= 1e-5 (fp32 mode) / <= 1e-2 (bf16 mode);
decoded images PSNR >= 40 dB vs the fp32 reference."""
import pytest
import torch

from gpu_util import rel
from oracle import restate as R
from oracle import weights as W
from oracle.golden import load_golden

pytestmark = pytest.mark.gpu

EPS_TOL = {"fp32": 1e-5, "bf16": 1e-2}


def _unet(g, mode):
    from sdb200.openai_model import UNetModel
    net = UNetModel(**g["cfg"], compute_mode=mode)
    net.load_state_dict(W.make_state_dict(g["key_shapes"], g["seed"]))
    return net.cuda()


@pytest.mark.parametrize("name", ["unet_tiny", "unet_sd"])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_unet_eps_parity(cuda, name, mode):
    g = load_golden(name + ".pt")
    net = _unet(g, mode)
    x = W.seeded_randn(g["x_shape"], g["seed"] + 1).cuda()
    ctx = W.seeded_randn(g["ctx_shape"], g["seed"] + 2).cuda()
    eps = net(x, g["t"].cuda(), ctx)
    assert eps.shape == tuple(g["x_shape"]) and eps.dtype == torch.float32
    e_ref, e_64 = rel(eps, g["eps_ref"]), rel(eps, g["eps_f64"])
    print("%s %s: rel-L2 vs reference fp32 %.3e, vs float64 %.3e" % (name, mode, e_ref, e_64))
    assert e_ref <= EPS_TOL[mode] and e_64 <= EPS_TOL[mode]


def test_unet_context_cache_survives_freed_and_reallocated_conditioning(cuda):
    """prompt 1's conditioning is freed, prompt 2's same-shaped conditioning is handed the same address by
    the caching allocator (and has _version 0 again).  The eager path's K/V cache must project prompt 2 afresh: the output
    has to equal a fresh model's, for fp32 contexts and for contexts that need a dtype conversion."""
    g = load_golden("unet_tiny.pt")
    x = W.seeded_randn(g["x_shape"], g["seed"] + 1).cuda()
    t = g["t"].cuda()
    for dt in (torch.float32, torch.float16):
        net = _unet(g, "bf16")
        ctx_a = W.seeded_randn(g["ctx_shape"], 1001).cuda().to(dt)
        ptr_a = ctx_a.data_ptr()
        out_a = net(x, t, ctx_a)
        del ctx_a
        ctx_b = (W.seeded_randn(g["ctx_shape"], 1002).cuda() * 2).to(dt)
        recycled = ctx_b.data_ptr() == ptr_a
        out_b = net(x, t, ctx_b)
        fresh = _unet(g, "bf16")(x, t, ctx_b)
        print("context cache: dtype %s, address recycled by the allocator: %s" % (dt, recycled))
        assert torch.equal(out_b, fresh)
        assert not torch.equal(out_b, out_a)
        assert torch.equal(net(x, t, ctx_b), out_b)           # and the second call hits the cache with the same result


@pytest.mark.parametrize("name", ["unet_var_legacy", "unet_var_neworder"])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_unet_variant_eps_parity(cuda, name, mode):
    """'next' row f4: the non-transformer UNetModel variants (AttentionBlock in both attention orders, use_scale_shift_norm,
    resblock_updown, num_classes) against the reference's eps on the same weights and inputs."""
    from sdb200.openai_model import UNetModel
    g = load_golden(name + ".pt")
    net = UNetModel(**g["cfg"], compute_mode=mode)
    net.load_state_dict(W.make_state_dict(g["key_shapes"], g["seed"]))
    net = net.cuda()
    x = W.seeded_randn(g["x_shape"], g["seed"] + 1).cuda()
    y = g["y"].cuda() if g["y"] is not None else None
    eps = net(x, g["t"].cuda(), None, y)
    e = rel(eps, g["eps_f64"])
    print("%s %s: eps rel-L2 %.3e" % (name, mode, e))
    assert eps.shape == g["eps_ref"].shape
    assert e <= EPS_TOL[mode]
    if y is not None:
        with pytest.raises(AssertionError):
            net(x, g["t"].cuda(), None)            # class-conditional model needs y (model.py:561-563)


@pytest.mark.parametrize("latent,B", [(80, 2), (96, 2), (40, 3)])
def test_unet_sd_ragged_latents_graph(cuda, latent, B):
    """BASELINE.json configs[4] shapes (96x96 latent) and other latents that are not multiples of 64: feature maps of
    12x12 / 10x10 / 5x5 pixels give partial tiles and odd tile counts (the CTA-pair kernel then runs a phantom m-tile).
    Inside a CUDA graph's private pool an out-of-bounds read faults instead of passing silently; graph replay must equal
    eager execution bit for bit, and sample i of the batch must equal the same sample run alone bit for bit (every kernel /
    tile / statistics-path choice is made from the per-sample geometry, never from the batch size: in bf16 mode two paths
    that differ by 1e-7 decorrelate to the bf16 noise level within a few layers)."""
    g = load_golden("unet_sd.pt")
    net = _unet(g, "bf16")
    x = W.seeded_randn((B, 4, latent, latent), 7).cuda()
    ctx = W.seeded_randn((B, 77, 768), 8).cuda()
    t = torch.tensor([981, 500, 21][:B], device="cuda")
    eager = net(x, t, ctx)
    torch.cuda.synchronize()
    assert torch.isfinite(eager).all() and float(eager.abs().max()) > 0.1
    net.use_cuda_graph = True
    for _ in range(2):
        out = net(x, t, ctx)
        torch.cuda.synchronize()
        assert torch.equal(out, eager)
    net.use_cuda_graph = False
    assert torch.equal(net(x[1:2], t[1:2], ctx[1:2]), eager[1:2])


def test_unet_batch_invariance_and_graph(cuda):
    """Size-independent properties at the benchmark shape: sample i of a batch equals the same sample run
    alone (nothing reduces over the batch), and CUDA-graph replay equals eager execution bit for bit."""
    g = load_golden("unet_sd.pt")
    net = _unet(g, "bf16")
    B = 4
    x = W.seeded_randn((B, 4, 64, 64), 5).cuda()
    ctx = W.seeded_randn((B, 77, 768), 6).cuda()
    t = torch.tensor([981, 500, 21, 1], device="cuda")
    full = net(x, t, ctx)
    assert torch.isfinite(full).all()
    for i in (0, 3):
        one = net(x[i:i + 1].contiguous(), t[i:i + 1].contiguous(), ctx[i:i + 1].contiguous())
        assert torch.equal(one[0], full[i])
    net.use_cuda_graph = True
    g1 = net(x, t, ctx)
    g2 = net(x, t, ctx)
    assert torch.equal(g1, full) and torch.equal(g2, full)
    # the context's K / V projections live in their own graph, replayed only for new conditioning: a different context,
    # an in-place update of the same tensor, and a return to the first one must all give the eager result
    ctx2 = W.seeded_randn((B, 77, 768), 7).cuda()
    net.use_cuda_graph = False
    full2 = net(x, t, ctx2)
    net.use_cuda_graph = True
    assert torch.equal(net(x, t, ctx2), full2)
    assert torch.equal(net(x, t, ctx), full)
    ctx.copy_(ctx2)
    assert torch.equal(net(x, t, ctx), full2)


@pytest.mark.parametrize("name", ["vae_tiny", "vae_sd_z16"])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_vae_decode_parity(cuda, name, mode):
    from sdb200.autoencoder import AutoencoderKL
    g = load_golden(name + ".pt")
    vae = AutoencoderKL(ddconfig=g["ddconfig"], embed_dim=4, compute_mode=mode)
    missing = vae.load_state_dict(W.make_state_dict(g["key_shapes"], g["seed"]), strict=False)
    assert not missing.unexpected_keys and all(k.startswith(("encoder.", "quant_conv.")) for k in missing.missing_keys)
    vae = vae.cuda()
    z = W.seeded_randn(g["z_shape"], g["seed"] + 1).cuda()
    img = vae.decode(z)
    e = rel(img, g["img_ref"])
    psnr = R.psnr_255(img.cpu(), g["img_ref"])
    print("%s %s: rel-L2 %.3e, PSNR %.1f dB" % (name, mode, e, psnr))
    assert img.shape == g["img_ref"].shape
    assert psnr >= 40.0
    assert e <= (1e-5 if mode == "fp32" else 2e-2)
    if name == "vae_tiny":    # micro-batching must not change anything
        vae.micro_batch = 1
        assert torch.equal(vae.decode(z), img)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_vae_encode_parity(cuda, mode):
    """'next' row f3: AutoencoderKL.encode (Encoder with the asymmetric-pad stride-2 convs, quant_conv, posterior)
    and the img2img entry DDIMSampler.stochastic_encode / decode against the reference's outputs."""
    from sdb200.autoencoder import AutoencoderKL
    from sdb200.ddim import DDIMSampler
    from oracle.make_golden import toy_model_fn
    g = load_golden("vae_enc_tiny.pt")
    vae = AutoencoderKL(ddconfig=g["ddconfig"], embed_dim=4, compute_mode=mode)
    r = vae.load_state_dict(W.make_state_dict(g["key_shapes"], g["seed"]), strict=False)
    assert not r.unexpected_keys and all(k.startswith(("decoder.", "post_quant_conv.")) for k in r.missing_keys)
    vae = vae.cuda()
    x = W.seeded_randn(g["x_shape"], g["seed"] + 1).cuda()
    post = vae.encode(x)
    tol = 1e-5 if mode == "fp32" else 2e-2
    em, el, es = rel(post.mean, g["mean_ref"]), rel(post.logvar, g["logvar_ref"]), rel(post.std, g["std_ref"])
    print("vae_enc_tiny %s: mean %.3e logvar %.3e std %.3e" % (mode, em, el, es))
    assert post.mean.shape == g["mean_ref"].shape
    assert em <= tol and el <= tol and es <= tol
    assert rel(post.var, g["std_ref"] ** 2) <= 2 * tol
    assert rel(post.sample(g["noise"].cuda()), g["z_ref"]) <= tol
    assert torch.equal(post.mode(), post.mean)
    assert rel(post.kl(), g["kl_ref"]) <= 10 * tol
    vae.micro_batch = 1
    assert torch.equal(vae.encode(x).mean, post.mean)
    if mode == "fp32":
        # img2img entry on the reference's latent: stochastic_encode is bit-exact, decode follows the reference trajectory
        shim = R.ModelShim(toy_model_fn, R.sd_alphas_cumprod(), device="cuda")
        shim.betas = shim.betas.cuda()
        smp = DDIMSampler(shim)
        smp.make_schedule(20, ddim_eta=0.0, verbose=False)
        ts = torch.full((x.shape[0],), g["t_enc"], device="cuda", dtype=torch.long)
        zt = smp.stochastic_encode(g["z_ref"].cuda(), ts, noise=g["enc_noise"].cuda())
        assert torch.equal(zt.cpu(), g["zt_ref"])
        zdec = smp.decode(zt, None, g["t_enc"])
        assert rel(zdec, g["zdec_ref"]) <= 1e-5


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_ddpm_unet_parity(cuda, mode):
    from sdb200.ddpm_unet import UNet
    g = load_golden("ddpm_unet.pt")
    net = UNet(image_size=32, input_channels=3, compute_mode=mode)
    net.load_state_dict(W.make_state_dict(g["key_shapes"], g["seed"]))
    net = net.cuda()
    y = net(W.seeded_randn(g["x_shape"], 12).cuda(), g["t"].cuda())
    e = rel(y, g["y_ref"])
    print("ddpm_unet %s: rel-L2 %.3e" % (mode, e))
    assert e <= (1e-5 if mode == "fp32" else 2e-2)
    for st in g["c1_steps"]:   # teacher-forced per-step eps along the reference's own C1 trajectory
        e_t = net(st["x_t"].cuda(), torch.full((4,), st["t"], device="cuda"))
        assert rel(e_t, st["e_t"]) <= (1e-5 if mode == "fp32" else 2e-2)


def test_c1_ddim50_end_to_end(cuda):
    """BASELINE.json configs[0]: DDPM UNet 32x32x3, DDIM-50, batch 4 — sampler + UNet on the GPU vs the
    reference's final samples."""
    from sdb200.ddim import DDIMSampler
    from sdb200.ddpm_unet import UNet
    g = load_golden("ddpm_unet.pt")
    net = UNet(image_size=32, input_channels=3, compute_mode="fp32")
    net.load_state_dict(W.make_state_dict(g["key_shapes"], g["seed"]))
    net = net.cuda()
    shim = R.ModelShim(None, R.ddpm_alphas_cumprod(), device="cuda")
    shim.betas = shim.betas.cuda()
    shim.apply_model = lambda x, t, c: net(x, t)
    x_T = W.seeded_randn((4, 3, 32, 32), g["c1_x_T_seed"]).cuda()
    z, _ = DDIMSampler(shim).sample(S=50, batch_size=4, shape=(3, 32, 32), conditioning=None, verbose=False, x_T=x_T, eta=0.)
    e = rel(z, g["c1_z_ref"])
    print("C1 DDIM-50 B=4 fp32: final-sample rel-L2 vs reference %.3e" % e)
    assert e <= 1e-4


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_tiny_pipeline_psnr(cuda, mode):
    """UNet DDIM-10 + VAE decode (tiny nets) vs the CPU oracle run on the same inputs: free-running image
    PSNR >= 40 dB and teacher-forced per-step eps within tolerance."""
    from sdb200.pipeline import LatentDiffusion
    gu, gv, ge = load_golden("unet_tiny.pt"), load_golden("vae_tiny.pt"), load_golden("vae_enc_tiny.pt")
    sdu = W.make_state_dict(gu["key_shapes"], gu["seed"])
    sdv = W.make_state_dict(gv["key_shapes"], gv["seed"])
    sdv.update(W.make_state_dict(ge["key_shapes"], ge["seed"]))     # encoder + quant_conv: the whole reference AutoencoderKL key set
    ld = LatentDiffusion(unet_config=gu["cfg"], first_stage_config=gv["ddconfig"], compute_mode=mode)
    ld.model.diffusion_model.load_state_dict(sdu)
    ld.first_stage_model.load_state_dict(sdv, strict=True)
    ld = ld.cuda()
    B, S = 2, 10
    x_T = W.seeded_randn((B, 4, 8, 8), 71)
    ctx = W.seeded_randn((B, 7, 64), 72)
    rec = []
    orc = R.DDIMOracle(R.ModelShim(lambda x, t, c: R.unet_forward(sdu, gu["cfg"], x, t, c), R.sd_alphas_cumprod()))
    with torch.no_grad():
        z_ref, _ = orc.sample(S, B, (4, 8, 8), conditioning=ctx, x_T=x_T, record=rec)
        img_ref = R.autoencoder_decode(sdv, gv["ddconfig"], z_ref / 0.18215)
    z, img = ld.txt2img(ctx.cuda(), B, ddim_steps=S, shape=(4, 8, 8), x_T=x_T.cuda())
    psnr = R.psnr_255(img.cpu(), img_ref)
    worst = max(rel(ld.apply_model(x_t.cuda(), torch.full((B,), t, device="cuda"), ctx.cuda()), e_t) for x_t, t, e_t in rec)
    print("tiny pipeline %s: latent rel-L2 %.3e, PSNR %.1f dB, worst teacher-forced eps rel-L2 %.3e" % (mode, rel(z, z_ref), psnr, worst))
    assert psnr >= 40.0
    assert worst <= EPS_TOL[mode]
    # img2img ('next' row f3): encode the oracle's image, noise it to step t_enc, denoise, decode — vs the oracle doing the same
    n_post, n_enc = W.seeded_randn((B, 4, 8, 8), 73), W.seeded_randn((B, 4, 8, 8), 74)
    t_enc = 6
    with torch.no_grad():
        mean, _, std = R.autoencoder_encode(sdv, gv["ddconfig"], img_ref)
        z0 = 0.18215 * (mean + std * n_post)
        orc.make_schedule(S, ddim_eta=0.0)
        zt = orc.stochastic_encode(z0, torch.full((B,), t_enc - 1, dtype=torch.long), n_enc)
        z2_ref = orc.decode(zt, ctx, t_enc)
        img2_ref = R.autoencoder_decode(sdv, gv["ddconfig"], z2_ref / 0.18215)
    post = ld.encode_first_stage(img_ref.cuda())
    z0_gpu = ld.get_first_stage_encoding(post, noise=n_post.cuda())
    assert rel(z0_gpu, z0) <= (1e-5 if mode == "fp32" else 2e-2)
    z2, img2 = ld.img2img(img_ref.cuda(), ctx.cuda(), strength=t_enc / S, ddim_steps=S, noise=n_enc.cuda(),
                          posterior_noise=n_post.cuda())
    psnr2 = R.psnr_255(img2.cpu(), img2_ref)
    print("tiny img2img %s: latent rel-L2 %.3e, PSNR %.1f dB" % (mode, rel(z2, z2_ref), psnr2))
    assert psnr2 >= 40.0
