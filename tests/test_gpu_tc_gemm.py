"""GPU: tcgen05 GEMM (TMA + UMMA + TMEM) vs fp64 math on the bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import bf16_round, randn, rel

pytestmark = pytest.mark.gpu


def _ab(M, N, K, seed=0):
    A = randn(M, K, seed=seed + 1).to(torch.bfloat16)
    W = (randn(N, K, seed=seed + 2) * K ** -0.5).to(torch.bfloat16)
    return A, W


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 128, 64, 128), (256, 320, 320, 0), (100, 640, 768, 0), (4096, 320, 1280, 160), (512, 1280, 5120, 0),
    (8, 1280, 1280, 0), (300, 4, 320, 0), (300, 64, 128, 64), (1000, 2560, 320, 256), (77, 100, 72, 0), (130, 48, 200, 0)])
def test_gemm_tc_fp32_out(cuda, M, N, K, bn):
    from sdb200 import ops
    A, W = _ab(M, N, K)
    bias, res = randn(N, seed=3), randn(M, N, seed=4)
    ref = A.double() @ W.double().T + bias.double() + res.double()
    out = ops.gemm_tc(A, W, bias, residual=res, block_n=bn)
    assert out.shape == (M, N)
    assert rel(out, ref) < 1e-5, (M, N, K, bn)


def test_gemm_tc_bf16_out_and_remap(cuda):
    from sdb200 import ops
    M, K, H, d, dp = 300, 320, 8, 40, 64
    A, W = _ab(M, 3 * H * d, K)
    ref = (A.double() @ W.double().T)
    out = ops.gemm_tc(A, W, out_dtype=torch.bfloat16)
    assert rel(out, ref) < 4e-3
    pad = ops.gemm_tc(A, W, out_dtype=torch.bfloat16, col_group=d, col_group_stride=dp)
    assert pad.shape == (M, 3 * H * dp)
    pv = pad.view(M, 3 * H, dp)
    assert torch.equal(pv[:, :, :d].reshape(M, -1), out)
    assert float(pv[:, :, d:].abs().max()) == 0.0


@pytest.mark.parametrize("M,inner,K", [(256, 1280, 320), (100, 2560, 640), (64, 5120, 1280)])
def test_gemm_tc_geglu(cuda, M, inner, K):
    from sdb200 import engine, ops
    A, W = _ab(M, 2 * inner, K)
    b = randn(2 * inner, seed=5) * 0.1
    h = A.double() @ W.double().T + b.double()
    ref = h[:, :inner] * F.gelu(h[:, inner:])
    pl = engine.PackedLinear(W.float(), b, "bf16", geglu=True)
    out = engine.linear(A, pl, out_dtype=torch.bfloat16)
    assert out.shape == (M, inner)
    assert rel(out, ref) < 5e-3


def test_gemm_tc_split_k(cuda):
    from sdb200 import ops
    A, W = _ab(64, 1280, 23040 // 4)
    bias = randn(1280, seed=3)
    ref = A.double() @ W.double().T + bias.double()
    for sk in (2, 5):
        assert rel(ops.gemm_tc(A, W, bias, split_k=sk), ref) < 1e-5
