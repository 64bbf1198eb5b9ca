"""TEST INFRASTRUCTURE — runs the UNMODIFIED reference (from /root/reference) as the ground truth.

Usable only where /root/reference exists (the build container, never the GPU box).  Nothing on disk
under the reference is touched; the reference's hard CUDA/fp16 assumptions are neutralised by
harness-side shims:
  S-1  stub `omegaconf.listconfig.ListConfig` (lazy import at openai_model/model.py:322)
  S-2  `openai_model.attention.flash_attn_func` -> exact-math SDPA adapter (module global, :106)
  S-3  `model.float()` + a pre-hook on `time_embed` undoing `t_emb.half()` (model.py:566)
  S-4  stdout redirected (forward() prints whole tensors)
  S-5  the `ldm` tree's mixed absolute/relative imports resolved by aliasing `ldm.X` as `X`
  S-6  `DDIMSampler.register_buffer` override (as written it forces `.to("cuda")`, ddim.py:19-23)
"""
import contextlib
import importlib
import io
import os
import sys
import types

import torch
import torch.nn.functional as F

REF = os.environ.get("SDB_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "openai_model"))


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def _stub_omegaconf():
    if "omegaconf" in sys.modules:
        return
    oc, lc = types.ModuleType("omegaconf"), types.ModuleType("omegaconf.listconfig")
    lc.ListConfig = type("ListConfig", (list,), {})
    oc.listconfig = lc
    sys.modules.update({"omegaconf": oc, "omegaconf.listconfig": lc})


def _stub_flash_attn():
    try:
        import flash_attn  # noqa: F401
    except Exception:
        fa = types.ModuleType("flash_attn")
        fa.flash_attn_func = None
        fa.flash_attn_qkvpacked_func = None
        sys.modules["flash_attn"] = fa


def _sdpa_adapter(q, k, v, dropout_p=0.0, softmax_scale=None, causal=False):
    return F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                          scale=softmax_scale).transpose(1, 2)


def _sdpa_qkvpacked_adapter(qkv, dropout_p=0.0, softmax_scale=None, causal=False):
    q, k, v = qkv.unbind(2)                                     # [B,S,3,H,D] -> three [B,S,H,D]
    return _sdpa_adapter(q, k, v, softmax_scale=softmax_scale)


def build_unet(cfg, state_dict=None, dtype=torch.float32):
    """Construct the reference UNetModel under shims S-1..S-4; optionally load `state_dict`."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    _stub_omegaconf()
    _stub_flash_attn()
    import openai_model.attention as oa
    import openai_model.model as om
    oa.flash_attn_func = _sdpa_adapter
    oa.flash_attn_qkvpacked_func = _sdpa_qkvpacked_adapter      # AttentionBlock variants (attention.py:387,514)
    with quiet():
        net = om.UNetModel(**cfg)
    net = net.to(dtype).eval()
    net.dtype = dtype
    net.time_embed.register_forward_pre_hook(lambda m, i: (i[0].to(m[0].weight.dtype),))
    if state_dict is not None:
        net.load_state_dict({k: v.to(dtype) for k, v in state_dict.items()}, strict=True)
    return net


def run_unet(net, x, t, ctx, y=None):
    with torch.no_grad(), quiet():
        return net(x, t, ctx, y) if y is not None else net(x, t, ctx)


def build_decoder(ddconfig, state_dict=None, dtype=torch.float32):
    """Construct ldm.modules.diffusionmodules.model.Decoder under shim S-5."""
    build_decoder_module()
    with quiet():
        dec = sys.modules["modules.diffusionmodules.model"].Decoder(**ddconfig)
    dec = dec.to(dtype).eval()
    if state_dict is not None:
        dec.load_state_dict({k: v.to(dtype) for k, v in state_dict.items()}, strict=True)
    return dec


def build_decoder_module():
    """Shim S-5: import the `ldm` tree under both of the package spellings its files use."""
    for pth in (REF, os.path.join(REF, "ldm")):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    for name in ["modules", "modules.diffusionmodules", "modules.diffusionmodules.util", "modules.attention",
                 "modules.distributions", "modules.distributions.distributions", "modules.diffusionmodules.model"]:
        if name not in sys.modules:
            sys.modules[name] = importlib.import_module("ldm." + name)


def build_encoder(ddconfig, state_dict=None, dtype=torch.float32):
    """Construct ldm.modules.diffusionmodules.model.Encoder (same shim as build_decoder)."""
    build_decoder_module()
    with quiet():
        enc = sys.modules["modules.diffusionmodules.model"].Encoder(**ddconfig)
    enc = enc.to(dtype).eval()
    if state_dict is not None:
        enc.load_state_dict({k: v.to(dtype) for k, v in state_dict.items()}, strict=True)
    return enc


def gaussian_distribution_class():
    build_decoder_module()
    return sys.modules["modules.distributions.distributions"].DiagonalGaussianDistribution


def ddim_module():
    pth = os.path.join(REF, "DDIM")
    if pth not in sys.path:
        sys.path.insert(0, pth)
    import ddim
    return ddim


def make_cpu_sampler(model):
    ddim = ddim_module()

    class CPUSampler(ddim.DDIMSampler):
        def register_buffer(self, name, attr):   # S-6
            setattr(self, name, attr)

    return CPUSampler(model)


def build_ddpm_unet(state_dict=None):
    """DDPM/models/unet.py UNet. Must run in a process that has not imported the ldm `modules` alias."""
    pth = os.path.join(REF, "DDPM")
    if pth not in sys.path:
        sys.path.insert(0, pth)
    from models.unet import UNet
    net = UNet(image_size=32, input_channels=3).eval()
    if state_dict is not None:
        net.load_state_dict(state_dict, strict=True)
    return net
