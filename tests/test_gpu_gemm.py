"""tcgen05 GEMM (mmgl_gemm_bf16) against a plain PyTorch fp32 matmul of the same bf16-rounded operands.

Floating-point kernel -> torch fp32 reference (TF32 off).  Tolerances: fp32 output 2e-5 rel-L2 (accumulation
order only); bf16 output 4e-3 rel-L2 (one bf16 rounding, 2^-9 relative, of the fp32 result).
"""
import pytest
import torch

from util import BF16, assert_close, randn

pytestmark = pytest.mark.gpu

TOL_F32 = 2e-5
TOL_BF16 = 4e-3


def _K():
    from mmgl_b200 import _capi
    return _capi


def _operands(gen, m, n, k, a_t, b_t, pad=0):
    """bf16 operands stored in the requested majorness; returns (a_store, b_store, A[M,K] fp32, B[N,K] fp32)."""
    a = randn(gen, m, k).to(BF16)
    b = randn(gen, n, k).to(BF16)
    a_store = a.t().contiguous() if a_t else a
    b_store = b.t().contiguous() if b_t else b
    return a_store, b_store, a.float(), b.float()


@pytest.mark.parametrize("a_t,b_t", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("bn", [64, 128, 192, 256])
@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (300, 200, 136), (1, 8, 8), (77, 520, 1000), (72, 64, 37)])
def test_gemm_majors_and_tails(a_t, b_t, bn, m, n, k):
    if (a_t and m % 8) or (b_t and n % 8) or (not a_t and k % 8) or (not b_t and k % 8):
        pytest.skip("the contiguous dim of each operand must be a multiple of 8 (16-byte rows)")
    torch.backends.cuda.matmul.allow_tf32 = False
    gen = torch.Generator().manual_seed(m * 7 + n * 3 + k + bn)
    a_s, b_s, a, b = _operands(gen, m, n, k, a_t, b_t)
    out = torch.full((m, n), float("nan"), dtype=torch.float32, device="cuda")
    _K().gemm(a_s, b_s, out, a_t=a_t, b_t=b_t, block_n=bn, raster=1 + (m + bn) % 2)
    assert_close("fp32 out", out, a @ b.t(), TOL_F32)
    out16 = torch.empty((m, n), dtype=BF16, device="cuda")
    _K().gemm(a_s, b_s, out16, a_t=a_t, b_t=b_t, block_n=bn)
    assert_close("bf16 out", out16, a @ b.t(), TOL_BF16)


@pytest.mark.parametrize("a_t,b_t", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("bn", [128, 256])
@pytest.mark.parametrize("m,n,k", [(256, 256, 64), (512, 384, 320), (300, 200, 136), (1, 8, 8), (1000, 520, 1000)])
def test_gemm_cta_pair_kernel(a_t, b_t, bn, m, n, k):
    """tcgen05.mma.cta_group::2: two CTAs share one 256 x BN tile (forced with pair=2), all operand majors + tails."""
    if (a_t and m % 8) or (b_t and n % 8) or (not a_t and k % 8) or (not b_t and k % 8):
        pytest.skip("the contiguous dim of each operand must be a multiple of 8 (16-byte rows)")
    gen = torch.Generator().manual_seed(m * 5 + n * 3 + k + bn)
    a_s, b_s, a, b = _operands(gen, m, n, k, a_t, b_t)
    out = torch.full((m, n), float("nan"), dtype=torch.float32, device="cuda")
    _K().gemm(a_s, b_s, out, a_t=a_t, b_t=b_t, block_n=bn, pair=2, raster=1 + (m + bn) % 2)
    assert_close("fp32 out", out, a @ b.t(), TOL_F32)


@pytest.mark.parametrize("m,n,k,a_t,b_t", [
    (5120, 2048, 2048, False, False),   # 160 pair tiles on 74 pairs: 2 full waves + 12 tail tiles cut 6 ways
    (512, 2048, 1024, False, False),    # 16 tiles, no full wave: every tile cut 4 ways
    (2048, 2048, 5120, True, True),     # wgrad shape, 64 tiles cut in 1 (no split possible) .. heuristic decides
    (1300, 1000, 520, False, True),     # ragged M / N / K tails inside split tiles
    (256, 256, 64, False, False),       # one k-block: cannot be split
])
def test_gemm_stream_k_tail(m, n, k, a_t, b_t):
    """The last partial wave of the CTA-pair kernel is cut into K-slices (fp32 partial tiles + fix-up by the owner):
    results must match the plain data-parallel schedule to fp32 accumulation-order accuracy."""
    gen = torch.Generator().manual_seed(m + n + k)
    a_s, b_s, a, b = _operands(gen, m, n, k, a_t, b_t)
    ref = a @ b.t()
    for sk in (2, 1):
        out = torch.full((m, n), float("nan"), dtype=torch.float32, device="cuda")
        _K().gemm(a_s, b_s, out, a_t=a_t, b_t=b_t, block_n=256, pair=2, stream_k=sk)
        assert_close(f"stream_k={sk}", out, ref, TOL_F32)
    # fused epilogue on the owner slice + repeated launches (flag reset)
    bias = randn(gen, n)
    res = randn(gen, m, n).to(BF16)
    out16 = torch.empty((m, n), dtype=BF16, device="cuda")
    for _ in range(3):
        _K().gemm(a_s, b_s, out16, a_t=a_t, b_t=b_t, block_n=256, pair=2, stream_k=2, bias=bias, residual=res, relu=True)
    assert_close("epilogue", out16, torch.relu(ref + bias) + res.float(), TOL_BF16)


@pytest.mark.parametrize("m,n,k,a_t,b_t", [
    (64, 768, 4608, True, True),        # LoRA dA = dt^T x: 6 tiles of 148 SMs without the split
    (768, 64, 4608, True, True),        # LoRA dB = dy^T t
    (64, 768, 4608, False, False),
    (104, 200, 2048 + 40, True, False),  # ragged everything (strides stay 16-byte multiples), K tail in the last slice
])
def test_gemm_skinny_output_is_split_along_k(m, n, k, a_t, b_t):
    """stream_k=2 with the tile heuristic: few output tiles over a long K run as K-slices (every slice parks its partial,
    a reduce kernel sums them and runs the epilogue); the result must equal the data-parallel schedule, fp32 and bf16
    outputs, with a scaled epilogue."""
    gen = torch.Generator().manual_seed(m * 3 + n + k)
    a_s, b_s, a, b = _operands(gen, m, n, k, a_t, b_t)
    ref = a @ b.t()
    for _ in range(2):                  # second launch: the slice flags were reset
        out = torch.full((m, n), float("nan"), dtype=torch.float32, device="cuda")
        _K().gemm(a_s, b_s, out, a_t=a_t, b_t=b_t, alpha=0.5, stream_k=2)
        assert_close("split", out, 0.5 * ref, TOL_F32)
    plain = torch.empty((m, n), dtype=torch.float32, device="cuda")
    _K().gemm(a_s, b_s, plain, a_t=a_t, b_t=b_t, alpha=0.5, stream_k=1)
    assert_close("split vs data-parallel", out, plain, TOL_F32)
    out16 = torch.empty((m, n), dtype=BF16, device="cuda")
    _K().gemm(a_s, b_s, out16, a_t=a_t, b_t=b_t, stream_k=2)
    assert_close("bf16", out16, ref, TOL_BF16)


def test_gemm_cta_pair_epilogue_and_second_operand():
    gen = torch.Generator().manual_seed(77)
    m, n, k0, k1 = 700, 512, 256, 64
    a0_s, b0_s, a0, b0 = _operands(gen, m, n, k0, False, False)
    a1_s, b1_s, a1, b1 = _operands(gen, m, n, k1, False, False)
    bias = randn(gen, n)
    gate = torch.tensor([0.3], device="cuda")
    res = randn(gen, m, n).to(BF16)
    out = torch.empty((m, n), dtype=BF16, device="cuda")
    aux = torch.empty((m, n), dtype=BF16, device="cuda")
    _K().gemm(a0_s, b0_s, out, a1=a1_s, b1=b1_s, bias=bias, aux=aux, gate=gate, residual=res, pair=2)
    pre = a0 @ b0.t() + a1 @ b1.t() + bias
    assert_close("aux", aux, pre, TOL_BF16)
    assert_close("out", out, res.float() + torch.tanh(gate) * pre, TOL_BF16)


@pytest.mark.parametrize("m,n,k", [(640, 2048, 2048), (2560, 8192, 2048), (2560, 2048, 8192), (64, 4096, 2048)])
def test_gemm_production_shapes_heuristic_tile(m, n, k):
    """OPT-1.3B gated cross-attention block shapes (q/out proj, fc1, fc2, K|V proj), heuristic BLOCK_N."""
    gen = torch.Generator().manual_seed(5)
    a_s, b_s, a, b = _operands(gen, m, n, k, False, False)
    out = torch.empty((m, n), dtype=BF16, device="cuda")
    _K().gemm(a_s, b_s, out)
    assert_close("bf16 out", out, a @ b.t(), TOL_BF16)


@pytest.mark.parametrize("m,n,k,a_t,b_t,fp32", [
    (10240, 2048, 2048, False, False, False),   # q / out projection at per-GPU batch 16 (M = 16 x 640): pair-kernel tiles
    (10240, 8192, 2048, False, False, False),   # fc1
    (10240, 2048, 8192, False, False, False),   # fc2
    (10240, 2048, 8192, False, True, False),    # dgrad of fc1
    (8192, 2048, 10240, True, True, True),      # wgrad of fc1: K = tokens = 10240, fp32 master-gradient output
    (2048, 8192, 10240, True, True, True),      # wgrad of fc2
    (2048, 2048, 10240, True, True, True),      # wgrad of q / out projection
    (1024, 2048, 2048, False, False, False),    # K / V projection of the bank (16 x 64 rows)
])
def test_gemm_at_benchmarked_shapes(m, n, k, a_t, b_t, fp32):
    """the GEMM shapes of the timed cfg2 step at batch 16 (the earlier production-shape test stopped
    at M = 2560), forward, dgrad and wgrad layouts, against an fp32 torch matmul on the same bf16 operands (on the GPU:
    cuBLAS fp32 without TF32 is the reference arithmetic here)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    gen = torch.Generator().manual_seed(m + n + k)
    a_s, b_s, a, b = _operands(gen, m, n, k, a_t, b_t)
    out = torch.empty((m, n), dtype=torch.float32 if fp32 else BF16, device="cuda")
    _K().gemm(a_s, b_s, out, a_t=a_t, b_t=b_t)
    assert_close("out", out, a @ b.t(), TOL_F32 if fp32 else TOL_BF16)


def test_gemm_wgrad_shape_fp32():
    """dW[N_out,K_in] = dy^T x: both operands MN-major, long K (= tokens), fp32 master-weight output."""
    gen = torch.Generator().manual_seed(6)
    rows, n_out, k_in = 1280, 2048, 512
    dy = randn(gen, rows, n_out).to(BF16)
    x = randn(gen, rows, k_in).to(BF16)
    out = torch.empty((n_out, k_in), dtype=torch.float32, device="cuda")
    _K().gemm(dy, x, out, a_t=True, b_t=True)
    assert_close("wgrad", out, dy.float().t() @ x.float(), TOL_F32)


@pytest.mark.parametrize("n", [256, 100])  # 100: row pitch not 16-byte aligned -> scalar epilogue path
def test_gemm_fused_epilogue(n):
    gen = torch.Generator().manual_seed(7 + n)
    m, k = 200, 192
    a_s, b_s, a, b = _operands(gen, m, n, k, False, False)
    bias = randn(gen, n)
    gate = torch.tensor([0.7], device="cuda")
    res = randn(gen, m, n).to(BF16)
    mask = randn(gen, m, n).to(BF16)
    alpha = 0.125
    K = _K()

    # bias + alpha + relu
    out = torch.empty((m, n), dtype=BF16, device="cuda")
    K.gemm(a_s, b_s, out, bias=bias, alpha=alpha, relu=True)
    assert_close("bias/alpha/relu", out, torch.relu(alpha * (a @ b.t() + bias)), TOL_BF16)

    # bias, aux (pre-gate), tanh gate, residual
    aux = torch.empty((m, n), dtype=BF16, device="cuda")
    K.gemm(a_s, b_s, out, bias=bias, aux=aux, gate=gate, residual=res)
    pre = a @ b.t() + bias
    assert_close("aux", aux, pre, TOL_BF16)
    assert_close("gate+residual", out, res.float() + torch.tanh(gate) * pre, TOL_BF16)

    # relu_mask then gate (ReLU backward fused into dgrad)
    K.gemm(a_s, b_s, out, relu_mask=mask, gate=gate)
    want = (a @ b.t()) * (mask.float() > 0).float() * torch.tanh(gate)
    assert_close("relu_mask", out, want, TOL_BF16)

    # accumulate into fp32 and bf16 outputs
    base = randn(gen, m, n)
    acc = base.clone()
    K.gemm(a_s, b_s, acc, accumulate=True)
    assert_close("accumulate fp32", acc, base + a @ b.t(), TOL_F32)
    acc16 = base.to(BF16)
    K.gemm(a_s, b_s, acc16, accumulate=True)
    assert_close("accumulate bf16", acc16, base.to(BF16).float() + a @ b.t(), TOL_BF16)


@pytest.mark.parametrize("a_t,b_t", [(False, False), (False, True), (True, True)])
def test_gemm_second_operand_pair(a_t, b_t):
    """acc = A0 B0^T + A1 B1^T in one TMEM tile (LoRA side product, GCN concat, d_bank = dK Wk + dV Wv)."""
    gen = torch.Generator().manual_seed(9)
    m, n, k0, k1 = 264, 328, 200, 64
    a0_s, b0_s, a0, b0 = _operands(gen, m, n, k0, a_t, b_t)
    a1_s, b1_s, a1, b1 = _operands(gen, m, n, k1, a_t, b_t)
    out = torch.empty((m, n), dtype=torch.float32, device="cuda")
    _K().gemm(a0_s, b0_s, out, a_t=a_t, b_t=b_t, a1=a1_s, b1=b1_s)
    assert_close("dual", out, a0 @ b0.t() + a1 @ b1.t(), TOL_F32)


def test_gemm_column_sliced_views():
    """Operands / outputs that are column slices of wider buffers (fused K|V buffer, W[:, :D] halves)."""
    gen = torch.Generator().manual_seed(10)
    m, n, k = 96, 128, 64
    abuf = randn(gen, m, 2 * k).to(BF16)
    wbuf = randn(gen, n, 2 * k).to(BF16)
    obuf = torch.zeros((m, 2 * n), dtype=BF16, device="cuda")
    _K().gemm(abuf[:, k:], wbuf[:, :k], obuf[:, n:])
    assert_close("sliced", obuf[:, n:], abuf[:, k:].float() @ wbuf[:, :k].float().t(), TOL_BF16)
    assert float(obuf[:, :n].abs().max()) == 0.0, "wrote outside the output slice"


def test_gemm_rejects_bad_arguments():
    K = _K()
    a = torch.zeros((8, 12), dtype=BF16, device="cuda")  # row pitch 24 B: not 16-byte aligned rows
    b = torch.zeros((8, 12), dtype=BF16, device="cuda")
    out = torch.zeros((8, 8), dtype=BF16, device="cuda")
    with pytest.raises(RuntimeError, match="leading dims"):
        K.gemm(a, b, out)
    with pytest.raises(RuntimeError, match="CUDA"):
        K.gemm(a.cpu(), b.cpu(), out.cpu())


def test_launch_counter_moves():
    K = _K()
    n0 = K.launch_count()
    a = torch.ones((128, 64), dtype=BF16, device="cuda")
    out = torch.empty((128, 128), dtype=BF16, device="cuda")
    K.gemm(a, torch.ones((128, 64), dtype=BF16, device="cuda"), out)
    torch.cuda.synchronize()
    assert K.launch_count() == n0 + 1
    assert float(out.float().min()) == 64.0 and float(out.float().max()) == 64.0
