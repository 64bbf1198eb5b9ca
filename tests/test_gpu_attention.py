"""GPU: fused tcgen05 attention vs fp64 softmax(q k^T) v on the bf16-rounded inputs."""
import pytest
import torch

from gpu_util import randn, rel

pytestmark = pytest.mark.gpu


def _ref(q, k, v, scale):
    qh, kh, vh = (t.double().transpose(1, 2) for t in (q, k, v))       # [B,H,S,d]
    o = torch.softmax(qh @ kh.transpose(-1, -2) * scale, -1) @ vh
    return o.transpose(1, 2).reshape(q.shape[0], q.shape[1], -1)


@pytest.mark.parametrize("B,H,Sq,Sk,d", [
    (1, 1, 128, 128, 64), (2, 8, 256, 256, 40), (1, 8, 1024, 1024, 80), (2, 8, 256, 77, 160), (1, 8, 64, 64, 160),
    (1, 2, 4096, 4096, 40), (2, 8, 1024, 77, 80), (1, 4, 16, 16, 64), (2, 3, 200, 300, 40), (1, 8, 4096, 77, 40)])
def test_attention_tc(cuda, B, H, Sq, Sk, d):
    from sdb200 import ops
    from sdb200.engine import head_pad
    dp = head_pad(d)
    q = randn(B, Sq, H, d, seed=1).to(torch.bfloat16)
    k = randn(B, Sk, H, d, seed=2).to(torch.bfloat16)
    v = randn(B, Sk, H, d, seed=3).to(torch.bfloat16)
    scale = d ** -0.5
    ref = _ref(q, k, v, scale)

    def padded(t):
        p = torch.zeros(t.shape[0], t.shape[1], H, dp, dtype=torch.bfloat16, device=t.device)
        p[..., :d] = t
        return p

    qp, kp, vp = padded(q), padded(k), padded(v)
    out = ops.attention_tc(qp, kp, vp, B, H, Sq, Sk, d, dp, scale,
                           (Sq * H * dp, H * dp, dp), (Sk * H * dp, H * dp, dp), (Sk * H * dp, H * dp, dp))
    assert out.shape == (B, Sq, H * d)
    assert rel(out, ref) < 1e-2, (B, H, Sq, Sk, d)
    # dense heads (d channels in memory, q | k | v side by side in one [rows, 3*H*d] projection output): the pad channels of
    # the tiles come from TMA's out-of-bounds zero fill — same bits as the padded layout
    if Sq == Sk:
        qkv = torch.cat([q.reshape(B * Sq, H * d), k.reshape(B * Sk, H * d), v.reshape(B * Sk, H * d)], 1).contiguous()
        W3 = 3 * H * d
        st = (Sq * W3, W3, d)
        out_d = ops.attention_tc(qkv, qkv[:, H * d:], qkv[:, 2 * H * d:], B, H, Sq, Sk, d, dp, scale, st, st, st, dense=True)
    else:
        out_d = ops.attention_tc(q.contiguous(), k.contiguous(), v.contiguous(), B, H, Sq, Sk, d, dp, scale,
                                 (Sq * H * d, H * d, d), (Sk * H * d, H * d, d), (Sk * H * d, H * d, d), dense=True)
    assert torch.equal(out_d, out)


def test_attention_tc_large_logits(cuda):
    """Rows whose running max keeps growing exercise the lazy-rescale path."""
    from sdb200 import ops
    B, H, S, d, dp = 1, 2, 512, 64, 64
    q = (randn(B, S, H, d, seed=4) * 3).to(torch.bfloat16)
    k = (randn(B, S, H, d, seed=5) * 3).to(torch.bfloat16)
    k = k * torch.linspace(0.2, 3.0, S, device=k.device).view(1, S, 1, 1).to(torch.bfloat16)   # later keys -> larger logits
    v = randn(B, S, H, d, seed=6).to(torch.bfloat16)
    out = ops.attention_tc(q, k, v, B, H, S, S, d, dp, 0.5, (S * H * d, H * d, d), (S * H * d, H * d, d), (S * H * d, H * d, d))
    assert rel(out, _ref(q, k, v, 0.5)) < 1e-2


def test_attention_tc_9216_tokens(cuda):
    """BASELINE configs[4] (96x96 latent): S = 9216 self-attention at d = 40, 72 key tiles per query tile, against fp64 math
    (computed per head to bound the checker's memory: one 9216 x 9216 fp64 score matrix is 680 MB)."""
    from sdb200 import ops
    B, H, S, d, dp = 1, 8, 9216, 40, 64
    q = randn(B, S, H, d, seed=11).to(torch.bfloat16)
    k = randn(B, S, H, d, seed=12).to(torch.bfloat16)
    v = randn(B, S, H, d, seed=13).to(torch.bfloat16)
    scale = d ** -0.5
    C = H * d
    out = ops.attention_tc(q, k, v, B, H, S, S, d, dp, scale, (S * C, C, d), (S * C, C, d), (S * C, C, d), dense=True)
    worst = 0.0
    for h in range(H):
        ref = _ref(q[:, :, h:h + 1], k[:, :, h:h + 1], v[:, :, h:h + 1], scale)
        worst = max(worst, rel(out[:, :, h * d:(h + 1) * d], ref))
        del ref
    print("attention S=9216 d=40: worst per-head rel-L2 %.3e" % worst)
    assert worst < 1e-2


@pytest.mark.parametrize("B,Sq,Sk,d", [(1, 128, 64, 256), (2, 256, 256, 512), (1, 4096, 4096, 512), (2, 200, 300, 512), (1, 1024, 1024, 256), (3, 77, 130, 512)])
def test_attention_wide(cuda, B, Sq, Sk, d):
    """tc_attention_wide_kernel (the VAE AttnBlock's single head of d = C channels, ldm/modules/diffusionmodules/model.py:180-204)
    against fp64 softmax(q k^T d^-1/2) v on the same bf16 inputs, with q | k | v as column blocks of one projection output
    (the layout the decoder uses) and ragged sequence lengths."""
    from sdb200 import ops
    qkv = randn(B * max(Sq, Sk), 3 * d, seed=1).to(torch.bfloat16)
    q = qkv[:B * Sq, :d]
    k = qkv[:B * Sk, d:2 * d]
    v = qkv[:B * Sk, 2 * d:]
    scale = d ** -0.5
    W3 = 3 * d
    out = ops.attention_wide(q, k, v, B, Sq, Sk, d, scale, (Sq * W3, W3), (Sk * W3, W3), (Sk * W3, W3))
    qd = q.reshape(B, Sq, d).double()
    kd = k.reshape(B, Sk, d).double()
    vd = v.reshape(B, Sk, d).double()
    ref = torch.softmax(qd @ kd.transpose(1, 2) * scale, -1) @ vd
    err = rel(out, ref)
    print("attention_wide B=%d Sq=%d Sk=%d d=%d: rel-L2 %.3e" % (B, Sq, Sk, d, err))
    assert out.shape == (B, Sq, d) and err < 1e-2


def test_attention_wide_large_logits(cuda):
    """growing logits exercise the lazy rescale of the 256-column O accumulator"""
    from sdb200 import ops
    B, S, d = 1, 512, 512
    q = (randn(B * S, d, seed=4) * 1.5).to(torch.bfloat16)
    k = (randn(B * S, d, seed=5) * 1.5).to(torch.bfloat16) * torch.linspace(0.2, 3.0, S, device="cuda").view(S, 1).to(torch.bfloat16)
    v = randn(B * S, d, seed=6).to(torch.bfloat16)
    out = ops.attention_wide(q, k, v, B, S, S, d, 0.2, (S * d, d), (S * d, d), (S * d, d))
    ref = torch.softmax(q.double().view(B, S, d) @ k.double().view(B, S, d).transpose(1, 2) * 0.2, -1) @ v.double().view(B, S, d)
    assert rel(out, ref) < 1e-2


@pytest.mark.parametrize("B,H,Sq,Sk,d", [
    (8, 8, 4096, 77, 40), (8, 8, 1024, 77, 80), (16, 8, 4096, 77, 40), (2, 8, 300, 77, 40), (1, 2, 700, 128, 80), (3, 5, 513, 50, 64),
    (1, 8, 100, 8, 40), (2, 4, 256, 96, 80), (1, 1, 2304, 77, 40), (1, 8, 9216, 77, 40)])
def test_attention_short_keys(cuda, B, H, Sq, Sk, d):
    """tc_attention_kv1_kernel (one key tile, CTAs walk query items; the cross-attention over the text tokens,
    openai_model/attention.py:99-112) against fp64 math AND against the key-tile-walking kernel on the same inputs, with
    ragged query counts, every key-count class (<= 64, 65..80, 81..128) and both head layouts."""
    from sdb200 import _lib, ops
    from sdb200.engine import head_pad
    lib = _lib.load()
    dp = head_pad(d)
    q = randn(B, Sq, H, d, seed=21).to(torch.bfloat16)
    k = randn(B, Sk, H, d, seed=22).to(torch.bfloat16)
    v = randn(B, Sk, H, d, seed=23).to(torch.bfloat16)
    scale = d ** -0.5
    sq, sk = (Sq * H * d, H * d, d), (Sk * H * d, H * d, d)

    def run():
        return ops.attention_tc(q, k, v, B, H, Sq, Sk, d, dp, scale, sq, sk, sk, dense=True)
    prev = lib.sdb_attention_set_short_key_kernel(1)
    try:
        out = run()
        lib.sdb_attention_set_short_key_kernel(0)
        old = run()
    finally:
        lib.sdb_attention_set_short_key_kernel(prev)
    worst = 0.0
    for b in range(B):
        worst = max(worst, rel(out[b:b + 1], _ref(q[b:b + 1], k[b:b + 1], v[b:b + 1], scale)))
    assert worst < 1e-2, worst
    assert rel(out, old.float()) < 5e-3
    # zero-padded heads in memory give the same bits as the dense layout
    if d != dp:
        def padded(t):
            p = torch.zeros(t.shape[0], t.shape[1], H, dp, dtype=torch.bfloat16, device=t.device)
            p[..., :d] = t
            return p
        out_p = ops.attention_tc(padded(q), padded(k), padded(v), B, H, Sq, Sk, d, dp, scale,
                                 (Sq * H * dp, H * dp, dp), (Sk * H * dp, H * dp, dp), (Sk * H * dp, H * dp, dp))
        assert torch.equal(out_p, out)


def test_attention_short_keys_peaked_rows(cuda):
    """One dominant key per row (logit gap ~40): the exact row max keeps the other 76 probabilities at their true tiny values."""
    from sdb200 import ops
    B, H, Sq, Sk, d, dp = 2, 8, 512, 77, 40, 64
    q = (randn(B, Sq, H, d, seed=31) * 4).to(torch.bfloat16)
    k = (randn(B, Sk, H, d, seed=32) * 4).to(torch.bfloat16)
    v = randn(B, Sk, H, d, seed=33).to(torch.bfloat16)
    st, sk = (Sq * H * d, H * d, d), (Sk * H * d, H * d, d)
    out = ops.attention_tc(q, k, v, B, H, Sq, Sk, d, dp, 0.5, st, sk, sk, dense=True)
    assert rel(out, _ref(q, k, v, 0.5)) < 1e-2
