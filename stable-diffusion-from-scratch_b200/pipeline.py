"""Minimal, Lightning-free sampling glue around the hot path.

`LatentDiffusion` exposes exactly what `DDIMSampler` and the reference's `sample_log` /
`decode_first_stage` call sites need (ldm/diffusion/ddpm.py:1326-1448, 1083-1156, 1814-1826):
schedule buffers, `apply_model(x, t, cond)`, `decode_first_stage(z)`.  It owns a `UNetModel`
(as `model.diffusion_model`, like DiffusionWrapper, ddpm.py:1992-2034) and an `AutoencoderKL`.
"""
import numpy as np
import torch
from torch import nn

from .autoencoder import AutoencoderKL
from .ddim import DDIMSampler
from .openai_model import UNetModel

SD_UNET_CONFIG = dict(  # Diffusion/config.yaml:31-44
    image_size=32, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
    num_res_blocks=2, channel_mult=(1, 2, 4, 4), num_heads=8, use_spatial_transformer=True,
    transformer_depth=1, context_dim=768, use_checkpoint=False, legacy=False)
SD_VAE_DDCONFIG = dict(  # Diffusion/config.yaml:51-64
    double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=(1, 2, 4, 4),
    num_res_blocks=2, attn_resolutions=[], dropout=0.0)


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2):
    """ldm/modules/diffusionmodules/util.py:21-44."""
    if schedule == "linear":
        betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    elif schedule == "sqrt_linear":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64)
    elif schedule == "sqrt":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64) ** 0.5
    else:
        raise ValueError(f"schedule '{schedule}' unknown.")
    return betas.numpy()


class DiffusionWrapper(nn.Module):
    """ldm/diffusion/ddpm.py:1992-2034, crossattn conditioning only."""

    def __init__(self, diffusion_model, conditioning_key="crossattn"):
        super().__init__()
        self.diffusion_model = diffusion_model
        self.conditioning_key = conditioning_key
        assert conditioning_key in (None, "crossattn")

    def forward(self, x, t, c_concat=None, c_crossattn=None):
        if self.conditioning_key is None:
            return self.diffusion_model(x, t)
        cc = c_crossattn[0] if len(c_crossattn) == 1 else torch.cat(c_crossattn, 1)
        return self.diffusion_model(x, t, context=cc)


class LatentDiffusion(nn.Module):
    def __init__(self, unet_config=None, first_stage_config=None, timesteps=1000, linear_start=0.00085, linear_end=0.0120,
                 beta_schedule="linear", scale_factor=0.18215, conditioning_key="crossattn", compute_mode=None,
                 unet=None, first_stage_model=None, cond_stage_model=None):
        super().__init__()
        # optional text conditioner (Diffusion/config.yaml:69-70: FrozenCLIPEmbedder); None = the caller supplies `context` tensors
        self.cond_stage_model = cond_stage_model
        unet = unet if unet is not None else UNetModel(**(unet_config or SD_UNET_CONFIG), compute_mode=compute_mode)
        self.model = DiffusionWrapper(unet, conditioning_key)
        if first_stage_model is None and first_stage_config is not False:
            first_stage_model = AutoencoderKL(ddconfig=(first_stage_config or SD_VAE_DDCONFIG), embed_dim=4, compute_mode=compute_mode)
        self.first_stage_model = first_stage_model
        self.scale_factor = scale_factor
        self.parameterization = "eps"
        self.num_timesteps = int(timesteps)
        betas = make_beta_schedule(beta_schedule, timesteps, linear_start=linear_start, linear_end=linear_end)
        alphas_cumprod = np.cumprod(1. - betas, axis=0)
        alphas_cumprod_prev = np.append(1., alphas_cumprod[:-1])
        to_t = lambda a: torch.tensor(a, dtype=torch.float32)
        self.register_buffer("betas", to_t(betas))
        self.register_buffer("alphas_cumprod", to_t(alphas_cumprod))
        self.register_buffer("alphas_cumprod_prev", to_t(alphas_cumprod_prev))
        self.register_buffer("sqrt_alphas_cumprod", to_t(np.sqrt(alphas_cumprod)))           # ldm/diffusion/ddpm.py:212-213
        self.register_buffer("sqrt_one_minus_alphas_cumprod", to_t(np.sqrt(1. - alphas_cumprod)))

    @property
    def device(self):
        return self.betas.device

    def apply_model(self, x_noisy, t, cond, return_ids=False):
        """ldm/diffusion/ddpm.py:1326-1448 without the fold/unfold patch branch."""
        if isinstance(cond, dict):
            pass
        else:
            if not isinstance(cond, list):
                cond = [cond]
            cond = {"c_crossattn": cond}
        return self.model(x_noisy, t, **cond)

    @torch.no_grad()
    def get_learned_conditioning(self, c):
        """ldm/diffusion/ddpm.py `get_learned_conditioning`: cond_stage_model.encode(c) (text or token ids -> [B, 77, 768])."""
        assert self.cond_stage_model is not None, "no cond_stage_model was given"
        if hasattr(self.cond_stage_model, "encode") and callable(self.cond_stage_model.encode):
            return self.cond_stage_model.encode(c)
        return self.cond_stage_model(c)

    def q_sample_coefficients(self, t):
        """extract_into_tensor(sqrt_alphas_cumprod, t, .) and (sqrt_one_minus_alphas_cumprod, t, .) as two [B] fp32 vectors
        (ldm/diffusion/ddpm.py:411-412; ldm/modules/diffusionmodules/util.py:96-99)."""
        t = t.to(self.sqrt_alphas_cumprod.device)
        return self.sqrt_alphas_cumprod.gather(-1, t).contiguous(), self.sqrt_one_minus_alphas_cumprod.gather(-1, t).contiguous()

    @torch.no_grad()
    def q_sample(self, x_start, t, noise=None):
        """ldm/diffusion/ddpm.py:407-412 as written — the default draw is torch.rand_like (uniform), not randn —
        on the fused q_sample kernel (sdb_q_sample): sqrt(a_t) * x_start + sqrt(1 - a_t) * noise per sample."""
        from . import ops
        if noise is None:
            noise = torch.rand_like(x_start)
        a, c = self.q_sample_coefficients(t)
        out = ops.q_sample(x_start.float().contiguous(), noise.float().contiguous(), a, c)
        return out if x_start.dtype == torch.float32 else out.to(x_start.dtype)

    @torch.no_grad()
    def decode_first_stage(self, z):
        """ldm/diffusion/ddpm.py:1083-1156: z / scale_factor then first_stage_model.decode."""
        z = 1. / self.scale_factor * z
        return self.first_stage_model.decode(z)

    @torch.no_grad()
    def encode_first_stage(self, x):
        """ldm/diffusion/ddpm.py `encode_first_stage`: first_stage_model.encode(x) -> posterior."""
        return self.first_stage_model.encode(x)

    @torch.no_grad()
    def get_first_stage_encoding(self, encoder_posterior, noise=None):
        """ldm/diffusion/ddpm.py `get_first_stage_encoding`: scale_factor * posterior.sample()."""
        if hasattr(encoder_posterior, "sample"):
            z = encoder_posterior.sample() if noise is None else encoder_posterior.sample(noise)
        elif isinstance(encoder_posterior, torch.Tensor):
            z = encoder_posterior
        else:
            raise NotImplementedError(f"encoder_posterior of type '{type(encoder_posterior)}' not yet implemented")
        return self.scale_factor * z

    @torch.no_grad()
    def img2img(self, image, cond, strength=0.75, ddim_steps=50, eta=0.0, unconditional_guidance_scale=1.0,
                unconditional_conditioning=None, noise=None, posterior_noise=None):
        """Encode -> DDIMSampler.stochastic_encode to step t_enc = strength * ddim_steps -> DDIMSampler.decode -> VAE
        decode (the mask-free img2img use of ldm/diffusion/ddim.py:208-243): latents and [-1,1] images.
        `posterior_noise` / `noise`: optional pre-drawn N(0,1) tensors for the posterior sample and the forward noising."""
        sampler = DDIMSampler(self)
        sampler.make_schedule(ddim_num_steps=ddim_steps, ddim_eta=eta, verbose=False)
        t_enc = max(1, min(ddim_steps, int(strength * ddim_steps)))
        z0 = self.get_first_stage_encoding(self.encode_first_stage(image), noise=posterior_noise)
        ts = torch.full((z0.shape[0],), t_enc - 1, device=z0.device, dtype=torch.long)
        zt = sampler.stochastic_encode(z0, ts, noise=noise)
        z = sampler.decode(zt, cond, t_enc, unconditional_guidance_scale=unconditional_guidance_scale,
                           unconditional_conditioning=unconditional_conditioning)
        return z, self.decode_first_stage(z)

    @torch.no_grad()
    def sample_log(self, cond, batch_size, ddim=True, ddim_steps=50, shape=(4, 64, 64), **kwargs):
        """ldm/diffusion/ddpm.py:1814-1826 (DDIM branch)."""
        assert ddim, "only the DDIM branch is on the hot path"
        sampler = DDIMSampler(self)
        return sampler.sample(ddim_steps, batch_size, shape, cond, verbose=False, **kwargs)

    @torch.no_grad()
    def txt2img(self, cond, batch_size, ddim_steps=50, shape=(4, 64, 64), x_T=None, eta=0.0,
                unconditional_guidance_scale=1.0, unconditional_conditioning=None):
        """DDIM-`ddim_steps` + VAE decode: latents and [-1,1] images."""
        z, _ = self.sample_log(cond, batch_size, ddim=True, ddim_steps=ddim_steps, shape=shape, x_T=x_T, eta=eta,
                               unconditional_guidance_scale=unconditional_guidance_scale,
                               unconditional_conditioning=unconditional_conditioning)
        return z, self.decode_first_stage(z)
