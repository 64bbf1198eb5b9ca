"""TEST INFRASTRUCTURE — CPU restatement (the parity oracle) of the reference's sampling hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (sdb200) never does.  Every function restates, in plain functional
PyTorch (fp32 or fp64, any device but meant for CPU), the arithmetic of the reference file:line it
cites.  It takes the reference's own state-dict keys, so the same weights drive the reference, this
oracle and the CUDA path.  It is pinned against the unmodified reference (imported from
/root/reference under the harness shims of oracle/ref_harness.py) by oracle/make_golden.py; the
resulting fixtures live in tests/golden/ and are re-checked by tests/test_oracle_golden.py.
The reference itself has no golden vectors or assertions for this path.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------------
# shared pieces
# ----------------------------------------------------------------------------------------------------


def timestep_embedding(timesteps, dim, max_period=10000):
    """openai_model/utils.py:225-245 — [cos | sin], freqs built in fp32 on the host."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
    freqs = freqs.to(device=timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(sd, prefix, x, eps, groups=32):
    return F.group_norm(x, groups, sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def _conv(sd, prefix, x, stride=1, padding=0):
    return F.conv2d(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"), stride=stride, padding=padding)


def _lin(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


# ----------------------------------------------------------------------------------------------------
# openai_model UNetModel
# ----------------------------------------------------------------------------------------------------

SD_UNET_CFG = dict(  # Diffusion/config.yaml:31-44
    image_size=32, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
    num_res_blocks=2, channel_mult=(1, 2, 4, 4), num_heads=8, use_spatial_transformer=True,
    transformer_depth=1, context_dim=768, use_checkpoint=False, legacy=False)


def unet_layout(cfg):
    """Block structure produced by UNetModel.__init__ (openai_model/model.py:362-532).

    Returns (input_blocks, middle_block, output_blocks): lists of layer lists; each layer is a tuple
    ("conv", cin, cout) | ("res", cin, cout, updown) | ("st", ch, heads, dim_head) | ("attn", ch, heads, new_order)
    | ("down", ch) | ("up", ch); updown in (None, "up", "down") (resblock_updown, model.py:421-436,505-520).
    """
    mc = cfg["model_channels"]
    mult = tuple(cfg.get("channel_mult", (1, 2, 4, 8)))
    nrb = cfg["num_res_blocks"]
    attn_res = list(cfg["attention_resolutions"])
    num_heads = cfg.get("num_heads", -1)
    num_head_channels = cfg.get("num_head_channels", -1)
    num_heads_upsample = cfg.get("num_heads_upsample", -1)
    if num_heads_upsample == -1:
        num_heads_upsample = num_heads
    legacy = cfg.get("legacy", True)
    use_st = cfg.get("use_spatial_transformer", False)
    new_order = cfg.get("use_new_attention_order", False)
    updown = cfg.get("resblock_updown", False)

    def attn_layer(ch, heads_arg):
        # model.py:388-408 — `num_heads` itself is reassigned when num_head_channels is given
        nonlocal num_heads
        if num_head_channels == -1:
            dh = ch // num_heads
        else:
            num_heads = ch // num_head_channels
            dh = num_head_channels
        if legacy:
            dh = ch // num_heads if use_st else num_head_channels
        if use_st:
            return ("st", ch, num_heads, dh)
        # AttentionBlock.__init__, attention.py:559-574: num_head_channels == -1 -> the given head count
        h_arg = num_heads if heads_arg is None else heads_arg
        nh = h_arg if dh == -1 else ch // dh
        return ("attn", ch, nh, new_order)

    inputs = [[("conv", cfg["in_channels"], mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            layers = [("res", ch, m * mc, None)]
            ch = m * mc
            if ds in attn_res:
                layers.append(attn_layer(ch, None))
            inputs.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            inputs.append([("res", ch, ch, "down")] if updown else [("down", ch)])
            chans.append(ch)
            ds *= 2
    middle = [("res", ch, ch, None), attn_layer(ch, None), ("res", ch, ch, None)]
    outputs = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, mc * m, None)]
            ch = mc * m
            if ds in attn_res:
                layers.append(attn_layer(ch, num_heads_upsample))
            if level and i == nrb:
                layers.append(("res", ch, ch, "up") if updown else ("up", ch))
                ds //= 2
            outputs.append(layers)
    return inputs, middle, outputs


def resblock(sd, p, x, emb, scale_shift=False, updown=None):
    """ResBlock._forward, openai_model/model.py:232-252, including the up/down (h_upd / x_upd = nearest x2 or
    avg_pool2d(2), :184-191) and use_scale_shift_norm (:244-248) branches."""
    h = F.silu(_gn(sd, p + ".in_layers.0", x, 1e-5))
    if updown == "up":
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    elif updown == "down":
        h = F.avg_pool2d(h, kernel_size=2, stride=2)
        x = F.avg_pool2d(x, kernel_size=2, stride=2)
    h = _conv(sd, p + ".in_layers.2", h, padding=1)
    emb_out = _lin(sd, p + ".emb_layers.1", F.silu(emb)).type(h.dtype)[..., None, None]
    if scale_shift:
        scale, shift = torch.chunk(emb_out, 2, dim=1)
        h = _gn(sd, p + ".out_layers.0", h, 1e-5) * (1 + scale) + shift
        h = F.silu(h)
    else:
        h = h + emb_out
        h = F.silu(_gn(sd, p + ".out_layers.0", h, 1e-5))
    h = _conv(sd, p + ".out_layers.3", h, padding=1)
    if (p + ".skip_connection.weight") in sd:
        x = _conv(sd, p + ".skip_connection", x)
    return x + h


def attention_block(sd, p, x, heads, new_order):
    """AttentionBlock._forward (openai_model/attention.py:588-599) as the reference executes it:
    both attention classes view the qkv conv output as [N, T, 3, H, ch] (channel = s*H*ch + h*ch + c);
    QKVAttentionLegacy (:497-523) passes softmax_scale = ch**-0.25 to flash_attn_qkvpacked_func,
    FlashAttention (:372-399) passes ch**-0.5 and then permutes the [N,T,H,ch] result to [N,H,T,ch]
    BEFORE reshaping it to [N,T,H*ch] (so tokens and heads are interleaved) — restated as executed."""
    b, c = x.shape[:2]
    xf = x.reshape(b, c, -1)
    T = xf.shape[-1]
    hn = F.group_norm(xf, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5)
    qkv = F.conv1d(hn, sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    ch = c // heads
    q, k, v = qkv.permute(0, 2, 1).reshape(b, T, 3, heads, ch).unbind(2)            # each [N,T,H,ch]
    scale = ch ** -0.5 if new_order else 1 / math.sqrt(math.sqrt(ch))
    w = torch.softmax(torch.einsum("bthc,bshc->bhts", q, k) * scale, dim=-1)
    out = torch.einsum("bhts,bshc->bthc", w, v)                                      # [N,T,H,ch]
    if new_order:
        out = out.permute(0, 2, 1, 3).reshape(b, T, heads * ch).permute(0, 2, 1)
    else:
        out = out.reshape(b, T, -1).permute(0, 2, 1)
    h = F.conv1d(out, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (xf + h).reshape(x.shape)


def cross_attention(sd, p, x, context, heads):
    """CrossAttention.forward, openai_model/attention.py:63-117; flash_attn_func(q,k,v,
    softmax_scale=dim_head**-0.5, causal=False) == softmax(q k^T * scale) v per head."""
    context = x if context is None else context
    q = _lin(sd, p + ".to_q", x)
    k = _lin(sd, p + ".to_k", context)
    v = _lin(sd, p + ".to_v", context)
    b, n, inner = q.shape
    d = inner // heads
    scale = d ** -0.5
    q = q.view(b, n, heads, d).transpose(1, 2)
    k = k.view(b, -1, heads, d).transpose(1, 2)
    v = v.view(b, -1, heads, d).transpose(1, 2)
    attn = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * scale, dim=-1)
    out = torch.matmul(attn, v).transpose(1, 2).reshape(b, n, inner)
    return _lin(sd, p + ".to_out.0", out)


def basic_transformer_block(sd, p, x, context, heads):
    """BasicTransformerBlock._forward, openai_model/attention.py:233-257; GEGLU :129-141, FF :146-172."""
    C = x.shape[-1]

    def ln(name, t):
        return F.layer_norm(t, (C,), sd[p + "." + name + ".weight"], sd[p + "." + name + ".bias"], 1e-5)

    x = cross_attention(sd, p + ".attn1", ln("norm1", x), None, heads) + x
    x = cross_attention(sd, p + ".attn2", ln("norm2", x), context, heads) + x
    h = _lin(sd, p + ".ff.net.0.proj", ln("norm3", x))
    a, gate = h.chunk(2, dim=-1)
    h = a * F.gelu(gate)
    x = _lin(sd, p + ".ff.net.2", h) + x
    return x


def spatial_transformer(sd, p, x, context, heads, depth=1):
    """SpatialTransformer.forward, openai_model/attention.py:336-363 (Normalize eps=1e-6, :10-11)."""
    b, c, hh, ww = x.shape
    x_in = x
    x = _gn(sd, p + ".norm", x, 1e-6)
    x = _conv(sd, p + ".proj_in", x)
    x = x.permute(0, 2, 3, 1).reshape(b, hh * ww, -1)
    for i in range(depth):
        x = basic_transformer_block(sd, "%s.transformer_blocks.%d" % (p, i), x, context, heads)
    x = x.reshape(b, hh, ww, -1).permute(0, 3, 1, 2)
    x = _conv(sd, p + ".proj_out", x)
    return x + x_in


def _run_layers(sd, prefix, layers, h, emb, context, depth, scale_shift=False):
    for j, layer in enumerate(layers):
        p = "%s.%d" % (prefix, j)
        kind = layer[0]
        if kind == "conv":
            h = _conv(sd, p, h, padding=1)
        elif kind == "res":
            h = resblock(sd, p, h, emb, scale_shift=scale_shift, updown=layer[3])
        elif kind == "st":
            h = spatial_transformer(sd, p, h, context, layer[2], depth)
        elif kind == "attn":
            h = attention_block(sd, p, h, layer[2], layer[3])
        elif kind == "down":   # Downsample.forward, model.py:95-97: conv3x3 stride 2 pad 1
            h = _conv(sd, p + ".op", h, stride=2, padding=1)
        elif kind == "up":     # Upsample.forward, model.py:119-131: nearest x2 then conv3x3
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = _conv(sd, p + ".conv", h, padding=1)
    return h


def unet_forward(sd, cfg, x, timesteps, context, taps=None, y=None):
    """UNetModel.forward, openai_model/model.py:550-595.  `emb = self.time_embed(t_emb.half())`
    (:566) rounds the sinusoidal embedding through fp16 before the MLP; harness shim S-3 only casts it
    back to the weights' dtype, so that rounding IS part of the reference's result and is restated
    here.  `taps`, if a dict, receives intermediate activations keyed by block name."""
    inputs, middle, outputs = unet_layout(cfg)
    depth = cfg.get("transformer_depth", 1)
    dt = sd["time_embed.0.weight"].dtype
    t_emb = timestep_embedding(timesteps, cfg["model_channels"]).half().to(dt)
    emb = _lin(sd, "time_embed.2", F.silu(_lin(sd, "time_embed.0", t_emb)))
    if cfg.get("num_classes") is not None:      # model.py:567-569
        emb = emb + sd["label_emb.weight"][y]
    ss = cfg.get("use_scale_shift_norm", False)
    h = x.to(dt)
    context = context.to(dt) if context is not None else None
    hs = []
    for i, layers in enumerate(inputs):
        h = _run_layers(sd, "input_blocks.%d" % i, layers, h, emb, context, depth, ss)
        hs.append(h)
        if taps is not None:
            taps["input_blocks.%d" % i] = h
    h = _run_layers(sd, "middle_block", middle, h, emb, context, depth, ss)
    if taps is not None:
        taps["middle_block"] = h
    for i, layers in enumerate(outputs):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_layers(sd, "output_blocks.%d" % i, layers, h, emb, context, depth, ss)
        if taps is not None:
            taps["output_blocks.%d" % i] = h
    h = F.silu(_gn(sd, "out.0", h, 1e-5))
    return _conv(sd, "out.2", h, padding=1)


# ----------------------------------------------------------------------------------------------------
# ldm VAE Decoder + AutoencoderKL.decode
# ----------------------------------------------------------------------------------------------------

SD_VAE_DDCONFIG = dict(  # Diffusion/config.yaml:51-64
    double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=(1, 2, 4, 4),
    num_res_blocks=2, attn_resolutions=[], dropout=0.0)


def _swish(x):   # nonlinearity, ldm/modules/diffusionmodules/model.py:35-37
    return x * torch.sigmoid(x)


def vae_resnet_block(sd, p, x):
    """ResnetBlock.forward with temb=None, ldm/modules/diffusionmodules/model.py:123-143."""
    h = _conv(sd, p + ".conv1", _swish(_gn(sd, p + ".norm1", x, 1e-6)), padding=1)
    h = _conv(sd, p + ".conv2", _swish(_gn(sd, p + ".norm2", h, 1e-6)), padding=1)
    if (p + ".nin_shortcut.weight") in sd:
        x = _conv(sd, p + ".nin_shortcut", x)
    elif (p + ".conv_shortcut.weight") in sd:
        x = _conv(sd, p + ".conv_shortcut", x, padding=1)
    return x + h


def vae_attn_block(sd, p, x):
    """AttnBlock.forward, ldm/modules/diffusionmodules/model.py:180-204 (1 head, d = C)."""
    h_ = _gn(sd, p + ".norm", x, 1e-6)
    q, k, v = _conv(sd, p + ".q", h_), _conv(sd, p + ".k", h_), _conv(sd, p + ".v", h_)
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, hh * ww)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(sd, p + ".proj_out", h_)


def decoder_forward(sd, ddconfig, z, prefix=""):
    """Decoder.forward, ldm/modules/diffusionmodules/model.py:541-574."""
    ch_mult = tuple(ddconfig["ch_mult"])
    nres = len(ch_mult)
    nrb = ddconfig["num_res_blocks"]
    attn_resolutions = list(ddconfig.get("attn_resolutions", []))
    curr_res = ddconfig["resolution"] // 2 ** (nres - 1)
    p = prefix
    h = _conv(sd, p + "conv_in", z, padding=1)
    h = vae_resnet_block(sd, p + "mid.block_1", h)
    h = vae_attn_block(sd, p + "mid.attn_1", h)
    h = vae_resnet_block(sd, p + "mid.block_2", h)
    for i_level in reversed(range(nres)):
        for i_block in range(nrb + 1):
            h = vae_resnet_block(sd, "%sup.%d.block.%d" % (p, i_level, i_block), h)
            if curr_res in attn_resolutions:
                h = vae_attn_block(sd, "%sup.%d.attn.%d" % (p, i_level, i_block), h)
        if i_level != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")   # Upsample.forward, model.py:55-59
            h = _conv(sd, "%sup.%d.upsample.conv" % (p, i_level), h, padding=1)
            curr_res *= 2
    h = _swish(_gn(sd, p + "norm_out", h, 1e-6))
    h = _conv(sd, p + "conv_out", h, padding=1)
    if ddconfig.get("tanh_out", False):
        h = torch.tanh(h)
    return h


def autoencoder_decode(sd, ddconfig, z):
    """AutoencoderKL.decode, ldm/models/autoencoder.py:337-340: post_quant_conv then Decoder."""
    return decoder_forward(sd, ddconfig, _conv(sd, "post_quant_conv", z), prefix="decoder.")


def encoder_forward(sd, ddconfig, x, prefix=""):
    """Encoder.forward, ldm/modules/diffusionmodules/model.py:434-465 (Downsample.forward :74-81: zero pad (0,1,0,1),
    then a stride-2 3x3 conv without padding)."""
    ch_mult = tuple(ddconfig["ch_mult"])
    nres = len(ch_mult)
    nrb = ddconfig["num_res_blocks"]
    attn_resolutions = list(ddconfig.get("attn_resolutions", []))
    curr_res = ddconfig["resolution"]
    p = prefix
    h = _conv(sd, p + "conv_in", x, padding=1)
    for i_level in range(nres):
        for i_block in range(nrb):
            h = vae_resnet_block(sd, "%sdown.%d.block.%d" % (p, i_level, i_block), h)
            if curr_res in attn_resolutions:
                h = vae_attn_block(sd, "%sdown.%d.attn.%d" % (p, i_level, i_block), h)
        if i_level != nres - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)
            h = _conv(sd, "%sdown.%d.downsample.conv" % (p, i_level), h, stride=2, padding=0)
            curr_res //= 2
    h = vae_resnet_block(sd, p + "mid.block_1", h)
    h = vae_attn_block(sd, p + "mid.attn_1", h)
    h = vae_resnet_block(sd, p + "mid.block_2", h)
    h = _swish(_gn(sd, p + "norm_out", h, 1e-6))
    return _conv(sd, p + "conv_out", h, padding=1)


def autoencoder_encode(sd, ddconfig, x):
    """AutoencoderKL.encode, ldm/models/autoencoder.py:331-335: Encoder, quant_conv (1x1), then the moments of a
    DiagonalGaussianDistribution (ldm/modules/distributions/distributions.py:24-35): returns (mean, logvar clamped
    to [-30, 20], std)."""
    moments = _conv(sd, "quant_conv", encoder_forward(sd, ddconfig, x, prefix="encoder."))
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    return mean, logvar, torch.exp(0.5 * logvar)


def q_sample_ddim(x0, noise, ddim_alphas, ddim_sqrt_one_minus_alphas, t):
    """DDIMSampler.stochastic_encode, ldm/diffusion/ddim.py:208-222 (use_original_steps=False): extract_into_tensor of
    sqrt(ddim_alphas) and ddim_sqrt_one_minus_alphas at index t, then sa * x0 + sb * noise."""
    sa = torch.sqrt(torch.as_tensor(ddim_alphas)).to(torch.float32 if x0.dtype != torch.float64 else torch.float64)
    sb = torch.as_tensor(ddim_sqrt_one_minus_alphas).to(sa.dtype)
    shape = (-1,) + (1,) * (x0.dim() - 1)
    return sa.gather(-1, t).reshape(shape) * x0 + sb.gather(-1, t).reshape(shape) * noise


# ----------------------------------------------------------------------------------------------------
# DDIM sampler
# ----------------------------------------------------------------------------------------------------


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2):
    """DDIM/diffusion_modules.py:21-44 ('linear' = linspace in sqrt space, float64)."""
    if schedule == "linear":
        betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    elif schedule == "sqrt_linear":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64)
    elif schedule == "sqrt":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64) ** 0.5
    else:
        raise ValueError("schedule '%s' unknown." % schedule)
    return betas.numpy()


def make_ddim_timesteps(method, num_ddim, num_ddpm):
    """DDIM/diffusion_modules.py:46-60."""
    if method == "uniform":
        c = num_ddpm // num_ddim
        ts = np.asarray(list(range(0, num_ddpm, c)))
    elif method == "quad":
        ts = ((np.linspace(0, np.sqrt(num_ddpm * .8), num_ddim)) ** 2).astype(int)
    else:
        raise NotImplementedError('There is no ddim discretization method called "%s"' % method)
    return ts + 1


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta):
    """DDIM/diffusion_modules.py:63-74 — alphacums is a CPU fp32 torch tensor, as in ddim.py:44."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


class ModelShim:
    """Duck-typed `model` the sampler needs (ldm/diffusion/ddim.py:16,28-34,120,174)."""

    def __init__(self, fn, alphas_cumprod, device="cpu"):
        ac = np.asarray(alphas_cumprod, dtype=np.float64)
        self.num_timesteps = ac.shape[0]
        self.device = torch.device(device)
        self.alphas_cumprod = torch.tensor(ac, dtype=torch.float32)
        self.alphas_cumprod_prev = torch.tensor(np.append(1., ac[:-1]), dtype=torch.float32)
        self.betas = torch.tensor(1. - ac / np.append(1., ac[:-1]), dtype=torch.float32)
        self._ac64 = ac
        self.fn = fn

    def apply_model(self, x, t, c):
        return self.fn(x, t, c)

    def q_sample(self, x_start, t, noise=None):
        """LatentDiffusion.q_sample, ldm/diffusion/ddpm.py:407-412, as written: the default draw is torch.rand_like
        (uniform); sqrt_alphas_cumprod = to_torch(np.sqrt(alphas_cumprod)) (:212-213), gathered per sample (util.py:96-99)."""
        if noise is None:
            noise = torch.rand_like(x_start)
        ac = self.alphas_cumprod.double().numpy() if not hasattr(self, "_ac64") else self._ac64
        sa = torch.tensor(np.sqrt(ac), dtype=torch.float32).gather(-1, t).reshape(-1, *((1,) * (x_start.dim() - 1)))
        sc = torch.tensor(np.sqrt(1. - ac), dtype=torch.float32).gather(-1, t).reshape(-1, *((1,) * (x_start.dim() - 1)))
        return sa * x_start + sc * noise


def sd_alphas_cumprod():
    """Diffusion/config.yaml:5-9 linear(sqrt-space) schedule, T=1000."""
    betas = make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.0120)
    return np.cumprod(1. - betas, axis=0)


def ddpm_alphas_cumprod():
    """DDPM/ddpm.py:18-27 with DDPM/train.py:64-66 values: linspace(1e-4, 1e-2, 1000) in fp32."""
    betas = torch.linspace(1e-4, 1e-2, 1000)
    return torch.cumprod(1 - betas, dim=0).double().numpy()


class DDIMOracle:
    """Restatement of DDIMSampler (DDIM/ddim.py:12-206 == ldm/diffusion/ddim.py), CPU tensors."""

    def __init__(self, model):
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0.):
        self.ddim_timesteps = make_ddim_timesteps(ddim_discretize, ddim_num_steps, self.ddpm_num_timesteps)
        ac = self.model.alphas_cumprod
        assert ac.shape[0] == self.ddpm_num_timesteps
        sig, al, alp = make_ddim_sampling_parameters(ac.cpu(), self.ddim_timesteps, ddim_eta)
        self.ddim_sigmas, self.ddim_alphas, self.ddim_alphas_prev = sig, al, alp
        self.ddim_sqrt_one_minus_alphas = np.sqrt(1. - al)

    def coefficients(self, index, b=1):
        """The four [b,1,1,1] tensors of ddim.py:191-194 (dtype as torch.full infers them)."""
        a_t = torch.full((b, 1, 1, 1), self.ddim_alphas[index])
        a_prev = torch.full((b, 1, 1, 1), self.ddim_alphas_prev[index])
        sigma_t = torch.full((b, 1, 1, 1), self.ddim_sigmas[index])
        s1m = torch.full((b, 1, 1, 1), self.ddim_sqrt_one_minus_alphas[index])
        return a_t, a_prev, sigma_t, s1m

    def p_sample_ddim(self, x, c, t, index, temperature=1., unconditional_guidance_scale=1.,
                      unconditional_conditioning=None, noise=None):
        """ddim.py:168-206. `noise`: pre-drawn N(0,1) tensor (the reference draws torch.randn here)."""
        b = x.shape[0]
        if unconditional_conditioning is None or unconditional_guidance_scale == 1.:
            e_t = self.model.apply_model(x, t, c)
        else:
            x_in, t_in = torch.cat([x] * 2), torch.cat([t] * 2)
            c_in = torch.cat([unconditional_conditioning, c])
            e_t_uncond, e_t = self.model.apply_model(x_in, t_in, c_in).chunk(2)
            e_t = e_t_uncond + unconditional_guidance_scale * (e_t - e_t_uncond)
        a_t, a_prev, sigma_t, s1m = self.coefficients(index, b)
        pred_x0 = (x - s1m * e_t) / a_t.sqrt()
        dir_xt = (1. - a_prev - sigma_t ** 2).sqrt() * e_t
        if noise is None:
            noise = torch.randn(x.shape)
        nz = sigma_t * noise * temperature
        x_prev = a_prev.sqrt() * pred_x0 + dir_xt + nz
        return x_prev, pred_x0, e_t

    def sample(self, S, batch_size, shape, conditioning=None, eta=0., x_T=None, temperature=1.,
               unconditional_guidance_scale=1., unconditional_conditioning=None, log_every_t=100,
               record=None, mask=None, x0=None):
        """sample + ddim_sampling, ddim.py:56-165. `record`, if a list, receives (x_t, t, e_t) per step."""
        self.make_schedule(S, ddim_eta=eta)
        C, H, W = shape
        img = torch.randn((batch_size, C, H, W)) if x_T is None else x_T
        timesteps = self.ddim_timesteps
        inter = {"x_inter": [img], "pred_x0": [img]}
        time_range = np.flip(timesteps)
        total = timesteps.shape[0]
        for i, step in enumerate(time_range):
            index = total - i - 1
            ts = torch.full((batch_size,), int(step), dtype=torch.long)
            if mask is not None:          # inpainting branch, ddim.py:144-149
                img_orig = self.model.q_sample(x0, ts)
                img = img_orig * mask + (1. - mask) * img
            x_in = img
            img, pred_x0, e_t = self.p_sample_ddim(
                img, conditioning, ts, index, temperature=temperature,
                unconditional_guidance_scale=unconditional_guidance_scale,
                unconditional_conditioning=unconditional_conditioning)
            if record is not None:
                record.append((x_in, int(step), e_t))
            if index % log_every_t == 0 or index == total - 1:
                inter["x_inter"].append(img)
                inter["pred_x0"].append(pred_x0)
        return img, inter


    def stochastic_encode(self, x0, t, noise):
        """ddim.py:208-222 with the current (make_schedule) tables."""
        return q_sample_ddim(x0, noise, self.ddim_alphas, self.ddim_sqrt_one_minus_alphas, t)

    def decode(self, x_latent, cond, t_start, unconditional_guidance_scale=1., unconditional_conditioning=None):
        """ddim.py:224-243: DDIM steps t_start-1 ... 0 starting from x_latent."""
        timesteps = self.ddim_timesteps[:t_start]
        total = timesteps.shape[0]
        x_dec = x_latent
        for i, step in enumerate(np.flip(timesteps)):
            index = total - i - 1
            ts = torch.full((x_latent.shape[0],), int(step), dtype=torch.long)
            x_dec, _, _ = self.p_sample_ddim(x_dec, cond, ts, index, unconditional_guidance_scale=unconditional_guidance_scale,
                                             unconditional_conditioning=unconditional_conditioning)
        return x_dec


# ----------------------------------------------------------------------------------------------------
# DDPM small UNet (parity config C1)
# ----------------------------------------------------------------------------------------------------


def ddpm_pe_matrix(dimension=128, max_timesteps=1000):
    """TransformerPositionalEmbedding.__init__, DDPM/models/layers.py:10-25 (interleaved sin/cos table)."""
    pe = torch.zeros(max_timesteps, dimension)
    even = torch.arange(0, dimension, 2)
    log_term = torch.log(torch.tensor(10000.0)) / dimension
    div = torch.exp(even * -log_term)
    ts = torch.arange(max_timesteps).unsqueeze(1)
    pe[:, 0::2] = torch.sin(ts * div)
    pe[:, 1::2] = torch.cos(ts * div)
    return pe


def _ddpm_convblock(sd, p, x, groups=32):   # ConvBlock, layers.py:37-48: conv -> GN -> SiLU
    x = _conv(sd, p + ".conv", x, padding=1)
    return F.silu(F.group_norm(x, groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5))


def _ddpm_resnet(sd, p, x, temb):           # ResNetBlock.forward, layers.py:323-338
    h = _ddpm_convblock(sd, p + ".block1", x)
    te = _lin(sd, p + ".time_embedding_projectile.1", F.silu(temb))[:, :, None, None]
    h = _ddpm_convblock(sd, p + ".block2", te + h)
    res = _conv(sd, p + ".residual_conv", x) if (p + ".residual_conv.weight") in sd else x
    return h + res


def _ddpm_attn(sd, p, x, heads=4, groups=32):   # SelfAttentionBlock.forward, layers.py:160-200
    b, c, hh, ww = x.shape
    t = x.view(b, c, hh * ww).transpose(1, 2)
    q, k, v = _lin(sd, p + ".query_projection", t), _lin(sd, p + ".key_projection", t), _lin(sd, p + ".value_projection", t)
    d = c // heads
    q = q.view(b, -1, heads, d).transpose(1, 2)
    k = k.view(b, -1, heads, d).transpose(1, 2)
    v = v.view(b, -1, heads, d).transpose(1, 2)
    a = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5), dim=-1)
    o = torch.matmul(a, v).permute(0, 2, 1, 3).contiguous().view(b, hh * ww, c)
    o = _lin(sd, p + ".final_projection", o).transpose(-1, -2).reshape(b, c, hh, ww)
    return F.group_norm(o + x, groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5)


# (kind, attention?, has down/upsample) for DDPM/models/unet.py:33-49
_DDPM_DOWN = [(False, True), (False, True), (False, True), (True, True), (False, True)]
_DDPM_UP = [False, True, False, False, False]


def ddpm_unet_forward(sd, x, time, pe=None):
    """UNet.forward, DDPM/models/unet.py:57-81."""
    pe = ddpm_pe_matrix() if pe is None else pe
    dt = sd["initial_conv.weight"].dtype
    temb = pe.to(dt)[time]
    temb = _lin(sd, "positional_encoding.3", F.gelu(_lin(sd, "positional_encoding.1", temb)))
    x0 = _conv(sd, "initial_conv", x.to(dt), padding=1)
    skips = [x0]
    h = x0
    for i, (attn, down) in enumerate(_DDPM_DOWN):
        p = "downsample_blocks.%d" % i
        for j in range(2):
            h = _ddpm_resnet(sd, "%s.resnet_blocks.%d" % (p, j), h, temb)
            if attn:
                h = _ddpm_attn(sd, "%s.attention_blocks.%d" % (p, j), h)
        if down:
            h = _conv(sd, p + ".downsample.conv", h, stride=2, padding=1)
        skips.append(h)
    skips = list(reversed(skips))
    for j in range(2):
        h = _ddpm_resnet(sd, "bottleneck.resnet_blocks.%d" % j, h, temb)
        h = _ddpm_attn(sd, "bottleneck.attention_blocks.%d" % j, h)
    for i, attn in enumerate(_DDPM_UP):
        p = "upsample_blocks.%d" % i
        h = torch.cat([h, skips[i]], dim=1)
        for j in range(2):
            h = _ddpm_resnet(sd, "%s.resnet_blocks.%d" % (p, j), h, temb)
            if attn:
                h = _ddpm_attn(sd, "%s.attention_blocks.%d" % (p, j), h)
        h = F.interpolate(h, scale_factor=2.0, mode="bilinear", align_corners=True)   # UpsampleBlock, layers.py:68-72
        h = _conv(sd, p + ".upsample.conv", h, padding=1)
    h = torch.cat([h, skips[-1]], dim=1)
    h = F.silu(F.group_norm(h, 32, sd["output_conv.0.weight"], sd["output_conv.0.bias"], 1e-5))
    return _conv(sd, "output_conv.2", h, padding=1)


# ----------------------------------------------------------------------------------------------------
# CLIP text tower behind FrozenCLIPEmbedder (SURVEY.md §8 'next' row f2)
# ----------------------------------------------------------------------------------------------------


def clip_text_forward(sd, cfg, input_ids):
    """`CLIPTextModel(input_ids).last_hidden_state` as FrozenCLIPEmbedder.forward uses it (clip_encoder/modules.py:246-252).
    The model itself lives in a third-party dependency (HuggingFace transformers, `transformers==4.49.0` in the reference's
    req.txt; the image has 5.5): its published algorithm is restated here — token + learned position embedding, pre-LayerNorm
    transformer layers with causal multi-head self-attention (q scaled by d^-1/2) and a quick-GELU MLP, final LayerNorm —
    and pinned against the installed library by oracle/make_golden.py (tests/golden/clip_text_*.pt)."""
    D, H, eps = cfg["hidden_size"], cfg["num_attention_heads"], cfg["layer_norm_eps"]
    d = D // H
    B, S = input_ids.shape
    p = "text_model."
    x = sd[p + "embeddings.token_embedding.weight"][input_ids] + sd[p + "embeddings.position_embedding.weight"][:S][None]
    mask = torch.full((S, S), float("-inf"), dtype=x.dtype).triu(1)
    for i in range(cfg["num_hidden_layers"]):
        q_ = "%sencoder.layers.%d." % (p, i)
        h = F.layer_norm(x, (D,), sd[q_ + "layer_norm1.weight"], sd[q_ + "layer_norm1.bias"], eps)
        q = _lin(sd, q_ + "self_attn.q_proj", h).view(B, S, H, d).transpose(1, 2) * d ** -0.5
        k = _lin(sd, q_ + "self_attn.k_proj", h).view(B, S, H, d).transpose(1, 2)
        v = _lin(sd, q_ + "self_attn.v_proj", h).view(B, S, H, d).transpose(1, 2)
        a = torch.softmax(q @ k.transpose(-1, -2) + mask, dim=-1) @ v
        x = x + _lin(sd, q_ + "self_attn.out_proj", a.transpose(1, 2).reshape(B, S, D))
        h = F.layer_norm(x, (D,), sd[q_ + "layer_norm2.weight"], sd[q_ + "layer_norm2.bias"], eps)
        h = _lin(sd, q_ + "mlp.fc1", h)
        h = h * torch.sigmoid(1.702 * h) if cfg["hidden_act"] == "quick_gelu" else F.gelu(h)
        x = x + _lin(sd, q_ + "mlp.fc2", h)
    return F.layer_norm(x, (D,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], eps)


# ----------------------------------------------------------------------------------------------------
# metrics
# ----------------------------------------------------------------------------------------------------


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def psnr_255(img, ref):
    """image_degradation/utils_image.py:621-635 on the clamp((x+1)/2)*255 post-processing
    (DDPM/utils.py:12-16): 20*log10(255/sqrt(MSE))."""
    a = ((img.double().clamp(-1, 1) + 1) / 2 * 255)
    b = ((ref.double().clamp(-1, 1) + 1) / 2 * 255)
    mse = float(((a - b) ** 2).mean())
    if mse == 0:
        return float("inf")
    return 20 * math.log10(255.0 / math.sqrt(mse))
