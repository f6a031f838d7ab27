"""GPU: the GEMM kernels through the C ABI against the oracle's linear() on the same seeded operands."""
import pytest
import torch

from oracle import mumpy_oracle as orc
from tests import util

pytestmark = pytest.mark.gpu

SHAPES = [  # (M, N, K) drawn from the model: qkv / proj / fc1 / fc2 / merge / embed, with ragged M and K tails
    (49, 96, 96), (3136, 288, 96), (3136, 384, 96), (3136, 96, 384), (9408, 384, 128), (2352, 1024, 256),
    (588, 1536, 512), (588, 512, 2048), (147, 3072, 1024), (147, 768, 2560), (1000, 256, 576), (130, 64, 72),
]


def _ops():
    import mumpy_b200
    return mumpy_b200.ops


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("epi", ["plain", "bias_gelu", "bias_residual"])
def test_linear_fp32_exact(M, N, K, epi):
    ops = _ops()
    a, w = util.seeded_input((M, K), 1), util.seeded_input((N, K), 2) / K ** 0.5
    bias = util.seeded_input((N,), 3) if epi != "plain" else None
    res = util.seeded_input((M, N), 4) if epi == "bias_residual" else None
    ref = orc.linear(a, w, bias)
    if epi == "bias_gelu":
        ref = orc.gelu(ref)
    if res is not None:
        ref = ref + res
    out = ops.linear(a.cuda(), w.cuda(), None if bias is None else bias.cuda(), None if res is None else res.cuda(),
                     act=ops.ACT_GELU if epi == "bias_gelu" else ops.ACT_NONE)
    assert util.maxabs(out, ref) < 2e-5


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("epi", ["plain", "bias_gelu_bf16out", "bias_residual"])
def test_linear_tcgen05_bf16(M, N, K, epi):
    """bf16 operands, fp32 accumulation: compare with the fp32 oracle on the bf16-rounded operands.
    Tolerance: products are exact in fp32, so only summation order differs (<= ~K * 2^-24 relative) plus one bf16
    rounding of the result when the output is bf16 (2^-9 relative)."""
    ops = _ops()
    a = util.seeded_input((M, K), 1).bfloat16()
    w = (util.seeded_input((N, K), 2) / K ** 0.5).bfloat16()
    bias = util.seeded_input((N,), 3) if epi != "plain" else None
    res = util.seeded_input((M, N), 4) if epi == "bias_residual" else None
    ref = orc.linear(a.float(), w.float(), bias)
    if epi == "bias_gelu_bf16out":
        ref = orc.gelu(ref)
    if res is not None:
        ref = ref + res
    bf16_out = epi == "bias_gelu_bf16out"
    out = ops.linear(a.cuda(), w.cuda(), None if bias is None else bias.cuda(), None if res is None else res.cuda(),
                     act=ops.ACT_GELU if bf16_out else ops.ACT_NONE, out_dtype=torch.bfloat16 if bf16_out else torch.float32)
    tol = 2e-2 if bf16_out else 1e-4
    assert util.maxabs(out.float(), ref) < tol * max(1.0, float(ref.abs().max()))
