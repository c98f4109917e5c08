"""tcgen05 / TMEM / TMA GEMM (csrc/tc_gemm.cu) against a float64 matmul: every operand-major combination, ragged sizes,
split-K, both tile widths, 3xTF32 (the path's precision) and 1xTF32."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _operand(rows, K, mn_major, gen):
    """matrix [rows, K] stored K-major ([rows, ldK]) or MN-major ([K, ldR]); returns (storage, ld, logical fp64 [rows, K])"""
    pad = lambda v: (v + 3) // 4 * 4 + 4
    if mn_major:
        st = torch.randn(K, pad(rows), device="cuda", generator=gen)
        return st, st.shape[1], st[:, :rows].t().double()
    st = torch.randn(rows, pad(K), device="cuda", generator=gen)
    return st, st.shape[1], st[:, :K].double()


def _split(x):
    from cae_tools_b200.engine import ops
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    ops.tc_split(x, hi, lo)
    return hi, lo


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 1), (0, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K,splits,tile_n", [(128, 128, 32, 1, 128), (200, 136, 100, 1, 128), (384, 520, 1000, 3, 128),
                                                 (130, 300, 264, 2, 256), (64, 72, 40, 1, 128),
                                                 (256, 512, 4608, 1, 256), (256, 256, 4608, 2, 128)])
def test_tc_gemm_3xtf32(a_mn, b_mn, M, N, K, splits, tile_n):
    from cae_tools_b200.engine import ops
    gen = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + a_mn * 2 + b_mn)
    A, lda, Ad = _operand(M, K, a_mn, gen)
    B, ldb, Bd = _operand(N, K, b_mn, gen)
    ah, al = _split(A)
    bh, bl = _split(B)
    # hi = A rounded to TF32, lo = (A - hi) rounded to TF32: both exactly representable in TF32 (low 13 mantissa bits zero),
    # and what the pair drops is at most 2^-22 |A|
    for t in (ah, al):
        assert int((t.view(torch.int32) & 0x1FFF).abs().max()) == 0
    assert bool(((ah.double() + al.double() - A.double()).abs() <= 2.0 ** -22 * A.double().abs()).all())
    ldc = (N + 3) // 4 * 4
    Cb = torch.full((splits, M, ldc), float("nan"), device="cuda")
    ops.tc_gemm(M, N, K, ah, al, lda, a_mn, bh, bl, ldb, b_mn, Cb, ldc, splits=splits, split_stride=M * ldc, tile_n=tile_n)
    torch.cuda.synchronize()
    got = Cb[:, :, :N].double().sum(0)
    want = Ad @ Bd.t()
    err = float((got - want).abs().max() / want.abs().max())
    # 3xTF32 products are good to ~2^-22.  The tensor core's fp32 accumulation truncates once per MMA; the TMEM accumulator
    # only ever holds 64 K elements before it is promoted into round-to-nearest register sums, so the error no longer grows
    # with K: measured 4 - 7e-7 of the max-norm at K = 32 ... 8192 (a cuBLAS fp32 GEMM: 2e-7 ... 2.6e-6); it was 3e-6 at
    # K = 1000 and linear in K when one accumulator ran over the whole K range.
    assert err < 1.5e-6, err


def test_tc_gemm_1xtf32_is_tf32_accurate():
    from cae_tools_b200.engine import ops
    gen = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 256, 256, 512
    A, lda, Ad = _operand(M, K, 0, gen)
    B, ldb, Bd = _operand(N, K, 0, gen)
    Cb = torch.zeros(M, N, device="cuda")
    ops.tc_gemm(M, N, K, A, None, lda, 0, B, None, ldb, 0, Cb, N)
    torch.cuda.synchronize()
    want = Ad @ Bd.t()
    err = float((Cb.double() - want).abs().max() / want.abs().max())
    assert 1e-6 < err < 5e-3, err     # TF32 inputs: ~1e-3; anything much smaller would mean the fp32 path ran
