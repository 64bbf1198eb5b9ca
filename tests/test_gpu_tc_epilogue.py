"""GPU: the CTA-pair kernel's TMA epilogue (swizzled smem staging + cp.async.bulk.tensor stores, residual tiles by TMA loads)
against (a) fp32 torch math on the same bf16-rounded operands and (b) the register-store epilogue it replaces — the two
epilogues add bias / residual / time-embedding row in the same order, so their outputs must be BIT-IDENTICAL; the GroupNorm
column statistics are summed in a different (fixed) order and are compared to tolerance."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import nchw, nhwc, randn, rel

pytestmark = pytest.mark.gpu


def _both(fn):
    """run fn() with the TMA epilogue and with the register-store epilogue"""
    from sdb200 import _lib
    lib = _lib.load()
    prev = lib.sdb_tc_set_tma_epilogue(1)
    try:
        a = fn()
        lib.sdb_tc_set_tma_epilogue(0)
        b = fn()
    finally:
        lib.sdb_tc_set_tma_epilogue(prev)
    return a, b


# ragged M (rows past the last 128 / 256-row tile), N not a multiple of the tile width, tiny M, many waves
GEMM_SHAPES = [(32768, 320, 320), (8192, 640, 640), (2048, 1280, 1280), (1000, 320, 320), (77, 640, 768), (616, 960, 320),
               (4096, 1920, 640), (300, 200, 64), (8192, 136, 320)]


@pytest.mark.parametrize("variant", [2, 1], ids=["pair", "single"])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_epilogues_agree(cuda, M, N, K, variant):
    from sdb200 import ops
    A = randn(M, K, seed=1).to(torch.bfloat16)
    W = (randn(N, K, seed=2) * K ** -0.5).to(torch.bfloat16)
    bias, res = randn(N, seed=3), randn(M, N, seed=4)
    ref = A.float() @ W.float().T + bias
    for kw, tol in ((dict(residual=res), 2e-5), (dict(), 2e-5), (dict(out_dtype=torch.bfloat16), 4e-3)):
        a, b = _both(lambda: ops.gemm_tc(A, W, bias, variant=variant, **kw))
        want = ref + res if "residual" in kw else ref
        assert rel(a, want) < tol, (kw.keys(), rel(a, want))
        assert torch.equal(a, b), kw.keys()
    # no bias, output written into a wider buffer (row pitch > N): the q | k | v layout
    wide = torch.zeros((M, N + 64), dtype=torch.bfloat16, device="cuda")
    a, b = _both(lambda: ops.gemm_tc(A, W, None, out=wide[:, 32:32 + N], out_dtype=torch.bfloat16, ldc=N + 64, variant=variant).clone())
    assert torch.equal(a, b) and rel(a, A.float() @ W.float().T) < 4e-3
    assert float(wide[:, :32].abs().max()) == 0.0 and float(wide[:, 32 + N:].abs().max()) == 0.0     # nothing written outside the columns


@pytest.mark.parametrize("M,C", [(32768, 320), (8192, 640), (2048, 1280), (520, 320)])
def test_geglu_epilogues_agree(cuda, M, C):
    from sdb200 import engine
    from sdb200.engine import PackedLinear
    A = randn(M, C, seed=1).to(torch.bfloat16)
    W = randn(8 * C, C, seed=2) * C ** -0.5
    bias = randn(8 * C, seed=3)
    pl = PackedLinear(W, bias, "bf16", geglu=True)
    h = A.float() @ W.to(torch.bfloat16).float().T + bias
    ref = h[:, :4 * C] * F.gelu(h[:, 4 * C:])
    a, b = _both(lambda: engine.linear(A, pl, out_dtype=torch.bfloat16, rows_per_item=M))
    assert rel(a, ref) < 5e-3
    assert torch.equal(a, b)
    from sdb200 import ops
    for variant in (1, 2):
        a2, b2 = _both(lambda: ops.gemm_tc(A, pl.w, pl.bias, out_dtype=torch.bfloat16, geglu=True, block_n=pl.block_n, variant=variant))
        assert torch.equal(a2, b2) and rel(a2, ref) < 5e-3


@pytest.mark.parametrize("N,H,W,Cin,Cout,k", [(8, 64, 64, 320, 320, 1), (8, 32, 32, 640, 640, 1), (2, 16, 16, 1280, 1280, 1), (3, 40, 24, 128, 256, 3),
                                              (2, 8, 8, 320, 640, 1), (1, 96, 96, 320, 320, 1), (2, 64, 64, 128, 128, 3), (5, 20, 12, 256, 320, 3)])
@pytest.mark.parametrize("variant", [2, 1], ids=["pair", "single"])
def test_conv_epilogues_agree(cuda, N, H, W, Cin, Cout, k, variant):
    from sdb200 import ops
    x = randn(N, H, W, Cin, seed=1).to(torch.bfloat16)
    w = (randn(Cout, Cin, k, k, seed=2) * (Cin * k * k) ** -0.5).to(torch.bfloat16)
    b = randn(Cout, seed=3)
    rv = randn(N, 2 * Cout, seed=4)
    res = randn(N, H, W, Cout, seed=5)
    wp = ops.pack_conv_weight(w, torch.bfloat16)
    base = nhwc(F.conv2d(nchw(x.float()), w.float(), b, padding=k // 2))
    for kw, want, tol in ((dict(rowvec=rv[:, Cout:], residual=res), base + rv[:, None, None, Cout:] + res, 3e-5),
                          (dict(residual=res), base + res, 3e-5), (dict(rowvec=rv[:, Cout:]), base + rv[:, None, None, Cout:], 3e-5),
                          (dict(), base, 3e-5), (dict(out_dtype=torch.bfloat16), base, 4e-3)):
        a, bb = _both(lambda: ops.conv_tc(x, wp, b, k, k, pad=k // 2, variant=variant, **kw))
        assert rel(a, want) < tol, (kw.keys(), rel(a, want))
        assert torch.equal(a, bb), kw.keys()


@pytest.mark.parametrize("N,H,W,C", [(2, 64, 64, 320), (3, 32, 32, 640), (2, 16, 16, 1280), (1, 40, 40, 320)])
@pytest.mark.parametrize("variant", [2, 1], ids=["pair", "single"])
def test_colstats_from_tma_epilogue(cuda, N, H, W, C, variant):
    """GroupNorm fed by the statistics the TMA epilogue emits == GroupNorm that measures the tensor itself."""
    from sdb200 import ops
    x = randn(N, H, W, C, seed=1).to(torch.bfloat16)
    w = (randn(C, C, 1, 1, seed=2) * C ** -0.5).to(torch.bfloat16)
    b = randn(C, seed=3)
    res = randn(N, H, W, C, seed=5)
    wp = ops.pack_conv_weight(w, torch.bfloat16)
    g, be = randn(C, seed=6), randn(C, seed=7)

    def run():
        out = ops.conv_tc(x, wp, b, 1, 1, pad=0, residual=res, variant=variant, want_stats=True)
        had = getattr(out, "_sdb_cs", None) is not None
        y = ops.groupnorm(out, g, be, 1e-5, act=1, out_dtype=torch.float32)
        return out, y, had
    (o1, y1, had1), (o0, y0, had0) = _both(run)
    assert had1 == had0 and (had1 or (H * W) % 128 != 0)     # statistics need tiles that do not straddle samples
    assert torch.equal(o1, o0)
    ref = F.silu(F.group_norm(nchw(o1).double(), 32, g.double(), be.double(), 1e-5))
    assert rel(nchw(y1), ref) < 2e-5 and rel(nchw(y0), ref) < 2e-5


def test_split_k_partials_through_tma(cuda):
    """split-K: the per-split partial tiles leave through the TMA epilogue into the workspace, the fixed-order reduction follows."""
    from sdb200 import ops
    M, N, K = 512, 1280, 5120
    A = randn(M, K, seed=1).to(torch.bfloat16)
    W = (randn(N, K, seed=2) * K ** -0.5).to(torch.bfloat16)
    bias, res = randn(N, seed=3), randn(M, N, seed=4)
    ref = A.float() @ W.float().T + bias + res
    for sk in (2, 5):
        for variant in (2, 1):
            a, b = _both(lambda: ops.gemm_tc(A, W, bias, residual=res, split_k=sk, variant=variant))
            assert rel(a, ref) < 2e-5
            assert torch.equal(a, b)
    x = randn(8, 8, 8, 1280, seed=5).to(torch.bfloat16)
    w = (randn(1280, 1280, 3, 3, seed=6) * (1280 * 9) ** -0.5).to(torch.bfloat16)
    wp = ops.pack_conv_weight(w, torch.bfloat16)
    a, b = _both(lambda: ops.conv_tc(x, wp, bias, 3, 3, pad=1, split_k=8, variant=2))
    assert torch.equal(a, b)
    assert rel(a, nhwc(F.conv2d(nchw(x.float()), w.float(), bias, padding=1))) < 3e-5


@pytest.mark.parametrize("N,H,W,Cin,Cout,sk", [(8, 8, 8, 1280, 1280, 8), (8, 8, 8, 1280, 1280, 1), (3, 8, 8, 640, 1280, 1), (8, 16, 16, 1280, 1280, 3),
                                               (2, 16, 16, 640, 640, 1), (5, 8, 16, 320, 640, 2)])
def test_colstats_small_maps_and_split_k(cuda, N, H, W, Cin, Cout, sk):
    """8x8 / 16x16 feature maps: a 128-row tile holds two samples (statistics per 32-row quarter still belong to one), and
    split-K launches get their statistics from the fixed-order reduction kernel.  The GroupNorm that consumes them must equal
    the GroupNorm that measures the tensor itself, and sample i of the batch must equal the same sample run alone."""
    from sdb200 import ops
    x = randn(N, H, W, Cin, seed=1).to(torch.bfloat16)
    w = (randn(Cout, Cin, 3, 3, seed=2) * (Cin * 9) ** -0.5).to(torch.bfloat16)
    b = randn(Cout, seed=3)
    res = randn(N, H, W, Cout, seed=5)
    wp = ops.pack_conv_weight(w, torch.bfloat16)
    g, be = randn(Cout, seed=6), randn(Cout, seed=7)
    for variant in (1, 2):
        out = ops.conv_tc(x, wp, b, 3, 3, pad=1, residual=res, variant=variant, split_k=sk, want_stats=True)
        assert getattr(out, "_sdb_cs", None) is not None, (variant, sk)
        ref_out = nhwc(F.conv2d(nchw(x.float()), w.float(), b, padding=1)) + res
        assert rel(out, ref_out) < 3e-5
        y = ops.groupnorm(out, g, be, 1e-5, act=1, out_dtype=torch.float32)
        ref = F.silu(F.group_norm(nchw(out).double(), 32, g.double(), be.double(), 1e-5))
        assert rel(nchw(y), ref) < 2e-5, (variant, sk)
        out1 = ops.conv_tc(x[1:2].contiguous(), wp, b, 3, 3, pad=1, residual=res[1:2].contiguous(), variant=variant, split_k=sk, want_stats=True)
        y1 = ops.groupnorm(out1, g, be, 1e-5, act=1, out_dtype=torch.float32)
        assert torch.equal(out1[0], out[1]) and torch.equal(y1[0], y[1]), (variant, sk)


@pytest.mark.parametrize("tma", [1, 0], ids=["tma_epilogue", "register_epilogue"])
def test_tail_split_is_bit_identical(cuda, tma):
    """One-CTA kernel, 160-column tiles: the tiles of the last partial wave issued as 96 | 64-column sub-tiles (sdb_tc_set_tail_split)
    give the same bits as whole tiles — GEMM with residual, ragged M and N, bf16 output, 3x3 conv with time-embedding row, residual
    and GroupNorm column statistics — through both epilogues."""
    from sdb200 import _lib, ops
    lib = _lib.load()
    prev_t, prev_e = lib.sdb_tc_set_tail_split(1), lib.sdb_tc_set_tma_epilogue(tma)

    def both(fn):
        lib.sdb_tc_set_tail_split(1)
        a = fn()
        lib.sdb_tc_set_tail_split(0)
        b = fn()
        torch.cuda.synchronize()
        return a, b
    try:
        for M, N, K in ((32768, 320, 320), (32768, 320, 1280), (25000, 300, 640), (40000, 640, 320)):
            A = randn(M, K, seed=1).to(torch.bfloat16)
            W = (randn(N, K, seed=2) * K ** -0.5).to(torch.bfloat16)
            bias, res = randn(N, seed=3), randn(M, N, seed=4)
            for kw in (dict(residual=res), dict(out_dtype=torch.bfloat16), dict()):
                a, b = both(lambda: ops.gemm_tc(A, W, bias, variant=1, block_n=160, **kw))
                assert torch.equal(a, b), (M, N, K, kw.keys())
            want = A.float() @ W.float().T + bias + res
            assert rel(ops.gemm_tc(A, W, bias, residual=res, variant=1, block_n=160), want) < 2e-5
        x = randn(8, 64, 64, 320, seed=5).to(torch.bfloat16)
        wp = ops.pack_conv_weight((randn(320, 320, 3, 3, seed=6) * (320 * 9) ** -0.5).to(torch.bfloat16), torch.bfloat16)
        b, rv, res = randn(320, seed=7), randn(8, 320, seed=8), randn(8, 64, 64, 320, seed=9)
        g, be = randn(320, seed=10) * 0.1 + 1, randn(320, seed=11) * 0.1

        def conv():
            o = ops.conv_tc(x, wp, b, 3, 3, pad=1, rowvec=rv, residual=res, variant=1, block_n=160, want_stats=True)
            return o, ops.groupnorm(o, g, be, 1e-5, act=1, out_dtype=torch.bfloat16)
        (o1, y1), (o0, y0) = both(conv)
        assert torch.equal(o1, o0) and torch.equal(y1, y0)
    finally:
        lib.sdb_tc_set_tail_split(prev_t)
        lib.sdb_tc_set_tma_epilogue(prev_e)


@pytest.mark.parametrize("variant", [2, 1], ids=["pair", "single"])
def test_b_const_is_bit_identical(cuda, variant):
    """sdb_tc_args.b_const (weight tiles requested before the programmatic-dependent-launch wait) changes when the loads are
    issued, never what is computed: GEMM and 3x3 conv, back-to-back launches so that a predecessor is actually draining."""
    from sdb200 import ops
    A = randn(8192, 640, seed=1).to(torch.bfloat16)
    W = (randn(640, 640, seed=2) * 640 ** -0.5).to(torch.bfloat16)
    bias, res = randn(640, seed=3), randn(8192, 640, seed=4)
    x = randn(4, 32, 32, 320, seed=5).to(torch.bfloat16)
    wp = ops.pack_conv_weight((randn(320, 320, 3, 3, seed=6) * (320 * 9) ** -0.5).to(torch.bfloat16), torch.bfloat16)
    cb = randn(320, seed=7)
    torch.cuda.synchronize()                       # the weights above are complete before any launch may prefetch them
    outs = {}
    for flag in (False, True, False, True):
        g = [ops.gemm_tc(A, W, bias, residual=res, variant=variant, b_const=flag) for _ in range(4)]
        c = [ops.conv_tc(x, wp, cb, 3, 3, pad=1, variant=variant, b_const=flag) for _ in range(4)]
        outs.setdefault(flag, []).append((g, c))
    torch.cuda.synchronize()
    g0, c0 = outs[False][0][0][0], outs[False][0][1][0]
    for flag in (False, True):
        for g, c in outs[flag]:
            for t in g:
                assert torch.equal(t, g0)
            for t in c:
                assert torch.equal(t, c0)
