"""K4 tcgen05 GEMM (through the C ABI test hook) against the plain SIMT comparator on the same bf16 inputs (every element) and
against an fp64 host evaluation of sample rows (independent of any CUDA code)."""
import pytest

pytestmark = pytest.mark.gpu

# epilogues: 0 bias->bf16, 1 bias+gelu->bf16, 2 bias->f32, 3 bias+residual f32, 6 argmax (+ logits)
CASES = [
    # M,    N,    K,    block_n, epilogue
    (128, 256, 64, 256, 2),
    (128, 128, 128, 128, 2),
    (256, 512, 384, 256, 2),
    (300, 200, 192, 64, 2),      # ragged M and N
    (1500, 1152, 384, 256, 0),   # tiny QKV shape, partial last N tile
    (3000, 384, 1536, 128, 3),   # fc2 + residual
    (1000, 1536, 384, 256, 1),   # fc1 + GELU
    (64, 2304, 768, 64, 2),      # decode-shaped (M < 128)
    (2, 384, 384, 32, 2),
    (5, 51865, 384, 128, 6),     # logits + argmax, N not a multiple of anything
    (20000, 768, 768, 128, 3),   # many tiles per CTA (persistent loop, both accumulator stages, phase wrap)
    # block_n 512 = the CTA-pair kernel (tcgen05.mma.cta_group::2, 256 x 256 tiles)
    (256, 256, 64, 512, 2),
    (300, 200, 192, 512, 2),     # ragged M (odd number of 128-row tiles) and N
    (1500, 1152, 384, 512, 0),
    (3000, 384, 1536, 512, 3),
    (1000, 1536, 384, 512, 1),
    (48000, 768, 768, 512, 3),   # encoder-sized: many tile pairs per cluster
]


@pytest.mark.parametrize("M,N,K,block_n,epi", CASES)
def test_gemm_matches_simt(pkg, M, N, K, block_n, epi):
    diff, ref = pkg.selftest_gemm(M, N, K, block_n, epi, seed=M + N + K)
    # fp32 accumulation on identical bf16 inputs: only summation order differs (bf16 outputs add 2^-9 relative)
    tol = 2e-3 * max(ref, 1.0) if epi in (2, 3, 6) else 1.2e-2 * max(ref, 1.0)
    assert ref > 0.1, "comparator output is degenerate"
    assert diff <= tol, "tcgen05 GEMM differs from the SIMT comparator: max|diff| %g (max|ref| %g)" % (diff, ref)
