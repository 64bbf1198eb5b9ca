"""GPU: tcgen05 implicit-GEMM convolution (4-D TMA boxes, zero-fill padding, traversal-stride 2)."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import nchw, nhwc, randn, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,H,W,Cin,Cout,k,stride", [
    (2, 64, 64, 64, 128, 3, 1), (1, 32, 32, 320, 320, 3, 1), (2, 16, 16, 128, 64, 3, 1), (3, 8, 8, 64, 160, 3, 1),
    (1, 8, 8, 128, 128, 3, 1), (1, 128, 128, 64, 32, 3, 1), (2, 24, 24, 64, 64, 3, 1), (1, 12, 20, 64, 64, 3, 1),
    (2, 16, 16, 192, 96, 1, 1), (1, 4, 4, 64, 64, 3, 1), (5, 1, 1, 128, 128, 3, 1), (2, 32, 32, 64, 4, 3, 1),
    (2, 32, 32, 64, 64, 3, 2), (1, 64, 64, 128, 128, 3, 2), (3, 8, 8, 64, 64, 3, 2), (1, 16, 16, 2560, 1280, 3, 1)])
def test_conv_tc(cuda, N, H, W, Cin, Cout, k, stride):
    from sdb200 import ops
    x = randn(N, H, W, Cin, seed=1).to(torch.bfloat16)
    w = (randn(Cout, Cin, k, k, seed=2) * (Cin * k * k) ** -0.5).to(torch.bfloat16)
    b = randn(Cout, seed=3)
    pad = k // 2
    ref = nhwc(F.conv2d(nchw(x.double()), w.double(), b.double(), stride=stride, padding=pad))
    wp = ops.pack_conv_weight(w, torch.bfloat16)
    out = ops.conv_tc(x, wp, b, k, k, stride=stride, pad=pad)
    assert out.shape == ref.shape
    # fp32 accumulation over K = k*k*Cin products (up to 23 040): the bound grows ~ sqrt(K) * 2^-24
    tol = 1e-5 if k * k * Cin <= 8192 else 3e-5
    assert rel(out, ref) < tol, (N, H, W, Cin, Cout, k, stride)
    rv = randn(N, Cout + 16, seed=4)
    res = randn(*ref.shape, seed=5)
    out2 = ops.conv_tc(x, wp, b, k, k, stride=stride, pad=pad, rowvec=rv[:, 8:8 + Cout], residual=res)
    assert rel(out2, ref + rv[:, None, None, 8:8 + Cout].double() + res.double()) < tol
    if Cout % 8 == 0:
        out3 = ops.conv_tc(x, wp, b, k, k, stride=stride, pad=pad, out_dtype=torch.bfloat16)
        assert rel(out3, ref) < 4e-3


def test_conv_tc_split_k(cuda):
    from sdb200 import ops
    x = randn(2, 8, 8, 1280, seed=1).to(torch.bfloat16)
    w = (randn(1280, 1280, 3, 3, seed=2) * (1280 * 9) ** -0.5).to(torch.bfloat16)
    ref = nhwc(F.conv2d(nchw(x.double()), w.double(), None, padding=1))
    out = ops.conv_tc(x, ops.pack_conv_weight(w, torch.bfloat16), None, 3, 3, pad=1, split_k=4)
    assert rel(out, ref) < 1e-5


@pytest.mark.parametrize("N,H,W,Cin,Cout,k,variant,bn", [
    (2, 32, 32, 64, 320, 3, 1, 160), (2, 32, 32, 64, 320, 3, 2, 160), (3, 16, 16, 128, 640, 1, 2, 256),
    (1, 64, 64, 64, 96, 3, 1, 0), (3, 24, 20, 64, 128, 3, 2, 128), (2, 16, 16, 64, 1280, 3, 0, 0)])
def test_conv_tc_column_statistics_feed_groupnorm(cuda, N, H, W, Cin, Cout, k, variant, bn):
    """The conv epilogue's per-slot column sums equal the sums of what it stored, and GroupNorm computed from them
    (one pass over the tensor) matches GroupNorm computed from the tensor — alone and as one half of a channel concat."""
    from sdb200 import ops
    x = randn(N, H, W, Cin, seed=1).to(torch.bfloat16)
    w = (randn(Cout, Cin, k, k, seed=2) * (Cin * k * k) ** -0.5).to(torch.bfloat16)
    b = randn(Cout, seed=3)
    res = randn(N, H, W, Cout, seed=5)
    wp = ops.pack_conv_weight(w, torch.bfloat16)
    out = ops.conv_tc(x, wp, b, k, k, pad=k // 2, residual=res, want_stats=True, variant=variant, block_n=bn)
    plain = ops.conv_tc(x, wp, b, k, k, pad=k // 2, residual=res, variant=variant, block_n=bn)
    assert torch.equal(out, plain)
    # 24x20 images: the batch-independent tile choice (nominal batch of 8) packs several images into one tile there, the
    # epilogue emits no statistics and ops.groupnorm takes its two-pass path — the GroupNorm results below must hold either way
    if (H, W) != (24, 20):
        cs, slots, spi = out._sdb_cs[:3]
        assert cs.shape == (2, slots, Cout) and slots >= N * spi
        per_sample = cs[:, :N * spi].double().reshape(2, N, spi, Cout).sum(2)
        o = out.double().reshape(N, H * W, Cout)
        assert rel(per_sample[0], o.sum(1)) < 1e-5
        assert rel(per_sample[1], (o * o).sum(1)) < 1e-5
    g, be = randn(Cout, seed=6) * 0.1 + 1, randn(Cout, seed=7) * 0.1
    ref = nhwc(F.silu(F.group_norm(nchw(out).double(), 32, g.double(), be.double(), 1e-5)))
    got = ops.groupnorm(out, g, be, 1e-5, act=1, out_dtype=torch.float32, exact=True)
    assert rel(got, ref) < 2e-6
    got16, raw = ops.groupnorm(out, g, be, 1e-5, act=1, out_dtype=torch.bfloat16, want_raw=True)
    assert rel(got16, ref) < 4e-3 and torch.equal(raw, out.to(torch.bfloat16))
    # concat with a second producer's output (different channel count)
    w2 = (randn(64, Cin, k, k, seed=8) * (Cin * k * k) ** -0.5).to(torch.bfloat16)
    out2 = ops.conv_tc(x, ops.pack_conv_weight(w2, torch.bfloat16), None, k, k, pad=k // 2, want_stats=True)
    C = Cout + 64
    g2, be2 = randn(C, seed=9) * 0.1 + 1, randn(C, seed=10) * 0.1
    ref2 = nhwc(F.group_norm(nchw(torch.cat([out, out2], -1)).double(), 32, g2.double(), be2.double(), 1e-6))
    assert rel(ops.groupnorm(out, g2, be2, 1e-6, out_dtype=torch.float32, x1=out2, exact=True), ref2) < 2e-6


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 16, 16, 64, 128), (1, 32, 32, 128, 64), (3, 8, 8, 64, 96), (2, 24, 20, 64, 320)])
def test_conv_up2_subpixel(cuda, N, H, W, Cin, Cout):
    """nearest-2x upsample + conv3x3 as four folded 2x2 phase convs == the direct form; its per-phase column statistics
    feed GroupNorm (alone and concatenated with an ordinary conv's output)."""
    from sdb200 import ops
    x = randn(N, H, W, Cin, seed=1).to(torch.bfloat16)
    w = randn(Cout, Cin, 3, 3, seed=2) * (Cin * 9) ** -0.5
    b = randn(Cout, seed=3)
    up = F.interpolate(nchw(x.double()), scale_factor=2, mode="nearest")
    ref = nhwc(F.conv2d(up, w.double(), b.double(), padding=1))
    out = ops.conv_up2_tc(x, ops.fold_upsample_weights(w, torch.bfloat16), b, want_stats=True)
    assert out.shape == ref.shape
    # folded taps are sums of up to four fp32 weights rounded once to bf16 (the direct form rounds each tap): ~2^-9 relative
    assert rel(out, ref) < 4e-3
    # exactness of the decomposition itself: same folded bf16 weights applied by the reference
    wf = ops.fold_upsample_weights(w, torch.bfloat16)
    chk = torch.zeros_like(ref)
    xp = F.pad(nchw(x.double()), (1, 1, 1, 1))
    for p_, (py, px) in enumerate(ops.UP2_PHASES):
        k = wf[p_].double().reshape(2, 2, Cout, Cin).permute(2, 3, 0, 1)
        y = F.conv2d(xp[:, :, py:py + H + 1, px:px + W + 1], k)
        chk[:, py::2, px::2, :] = nhwc(y) + b.double()
    assert rel(out, chk) < 1e-5
    if getattr(out, "_sdb_cs", None) is not None:
        g, be = randn(Cout, seed=6) * 0.1 + 1, randn(Cout, seed=7) * 0.1
        refn = nhwc(F.silu(F.group_norm(nchw(out).double(), 32, g.double(), be.double(), 1e-5)))
        assert rel(ops.groupnorm(out, g, be, 1e-5, act=1, out_dtype=torch.float32, exact=True), refn) < 2e-6
        x2 = randn(N, 2 * H, 2 * W, 64, seed=8).to(torch.bfloat16)
        w2 = (randn(64, 64, 3, 3, seed=9) * (64 * 9) ** -0.5).to(torch.bfloat16)
        o2 = ops.conv_tc(x2, ops.pack_conv_weight(w2, torch.bfloat16), None, 3, 3, pad=1, want_stats=True)
        C = Cout + 64
        g2, be2 = randn(C, seed=10) * 0.1 + 1, randn(C, seed=11) * 0.1
        ref2 = nhwc(F.group_norm(nchw(torch.cat([o2, out], -1)).double(), 32, g2.double(), be2.double(), 1e-6))
        assert rel(ops.groupnorm(o2, g2, be2, 1e-6, out_dtype=torch.float32, x1=out, exact=True), ref2) < 2e-6


@pytest.mark.parametrize("N,H,W,Cin,Cout,sk", [(8, 8, 8, 1280, 1280, 3), (2, 16, 16, 512, 512, 2), (1, 8, 8, 640, 320, 5)])
def test_split_k_phase_conv_touches_only_its_own_pixels(cuda, N, H, W, Cin, Cout, sk):
    """One sub-pixel phase of an upsampling conv run with split-K: the fixed-order reduction must visit this phase's pixels only.
    The workspace is poisoned with NaN (a same-sized block is filled and handed back to the allocator just before the call) and
    the other three phases' pixels of `out` hold a sentinel that has to survive."""
    from sdb200 import _lib, ops
    x = randn(N, H, W, Cin, seed=1).to(torch.bfloat16)
    w = (randn(4, Cout, Cin, seed=2) * (Cin * 4) ** -0.5).to(torch.bfloat16)
    b = randn(Cout, seed=3)
    for p_, (py, px) in enumerate(ops.UP2_PHASES):
        ph = (2, 2, py, px, 2 * H, 2 * W, 1 - py, 1 - px)
        out = torch.full((N, 2 * H, 2 * W, Cout), 7.0, device="cuda")
        probe = _lib.TcArgs()
        ops._fill_conv_args(probe, x, w, b, out, 2, 2, 1, ph)
        probe.split_k = sk
        need = _lib.load().sdb_tc_workspace_bytes(probe)
        assert need > 0
        poison = torch.full((need // 4,), float("nan"), device="cuda")
        del poison                                     # the next same-sized torch.empty (the workspace) gets this block back
        ops.conv_tc(x, w, b, 2, 2, stride=1, pad=0, out=out, phase=ph, split_k=sk)
        ref = ops.conv_tc(x, w, b, 2, 2, stride=1, pad=0, out=torch.full_like(out, 7.0), phase=ph, split_k=1)
        torch.cuda.synchronize()
        mine = out[:, py::2, px::2]
        assert torch.isfinite(out).all()
        assert rel(mine, ref[:, py::2, px::2]) < 1e-5
        other = out.clone()
        other[:, py::2, px::2] = 7.0
        assert float((other - 7.0).abs().max()) == 0.0, "the reduction wrote pixels of another phase"
