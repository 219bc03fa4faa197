"""GPU: the tcgen05 GEMM (dmc_gemm) and the FFMA GEMM (dmc_gemm_simt) through the C ABI, against a
float64 numpy product of the same (already rounded) operands."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _ops():
    import dinomc_b200
    return dinomc_b200.ops, dinomc_b200._lib


def _make(M, N, K, a_mn, b_mn, dtype, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    if dtype == torch.bfloat16:
        A, B = A.bfloat16().float(), B.bfloat16().float()       # exactly representable operands
    A_st = (A.t().contiguous() if a_mn else A).to(dtype).cuda()
    B_st = (B.t().contiguous() if b_mn else B).to(dtype).cuda()
    ref = A.double().numpy() @ B.double().numpy().T
    return A_st, B_st, ref


SHAPES = [(128, 256, 64), (128, 64, 128), (200, 320, 136), (512, 2048, 256), (384, 128, 1000)]
LAYOUTS = [(False, False), (False, True), (True, False), (True, True)]


@pytest.mark.parametrize("a_mn,b_mn", LAYOUTS)
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_bf16(M, N, K, a_mn, b_mn):
    ops, _ = _ops()
    A, B, ref = _make(M, N, K, a_mn, b_mn, torch.bfloat16)
    D = ops.gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert rel_err(D.cpu().numpy(), ref) < 2e-5      # exact products, fp32 accumulation


@pytest.mark.parametrize("a_mn,b_mn", LAYOUTS)
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_3xtf32(M, N, K, a_mn, b_mn):
    ops, _ = _ops()
    A, B, ref = _make(M, N, K, a_mn, b_mn, torch.float32, seed=1)
    Ah, Al = ops.split_tf32(A)
    Bh, Bl = ops.split_tf32(B)
    D = ops.gemm(Ah, Bh, M, N, K, a_mn=a_mn, b_mn=b_mn, A_lo=Al, B_lo=Bl, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert rel_err(D.cpu().numpy(), ref) < 5e-6      # ~fp32 accuracy from three TF32 passes


@pytest.mark.parametrize("a_mn,b_mn", LAYOUTS)
def test_gemm_tf32_single_pass(a_mn, b_mn):
    ops, _ = _ops()
    M, N, K = 256, 256, 256
    A, B, ref = _make(M, N, K, a_mn, b_mn, torch.float32, seed=2)
    D = ops.gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert rel_err(D.cpu().numpy(), ref) < 5e-3      # plain TF32: 10-bit mantissa operands


@pytest.mark.parametrize("a_mn,b_mn", LAYOUTS)
@pytest.mark.parametrize("M,N,K", [(130, 70, 45), (64, 64, 16), (257, 129, 300)])
def test_gemm_simt(M, N, K, a_mn, b_mn):
    ops, _ = _ops()
    A, B, ref = _make(M, N, K, a_mn, b_mn, torch.float32, seed=3)
    D = ops.gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, simt=True)
    torch.cuda.synchronize()
    assert rel_err(D.cpu().numpy(), ref) < 2e-6


@pytest.mark.parametrize("split", [0, 1, 3, 8])
def test_gemm_split_k(split):
    """dgrad-shaped problem: tiny output, long contraction -> split-K partials + deterministic reduce."""
    ops, _ = _ops()
    M, N, K = 256, 256, 8192
    A, B, ref = _make(M, N, K, False, True, torch.bfloat16, seed=4)
    D = ops.gemm(A, B, M, N, K, b_mn=True, split_k=split)
    D2 = ops.gemm(A, B, M, N, K, b_mn=True, split_k=split)
    torch.cuda.synchronize()
    assert rel_err(D.cpu().numpy(), ref) < 2e-5
    assert torch.equal(D, D2)                          # deterministic (no atomics)


def _gelu(x):
    from oracle.np_oracle import gelu
    return gelu(x)


@pytest.mark.parametrize("simt", [False, True])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_gemm_epilogue(simt, out_dtype):
    from oracle.np_oracle import gelu, gelu_grad
    ops, L = _ops()
    M, N, K = 192, 320, 128
    dt = torch.float32 if simt else torch.bfloat16
    A, B, ref = _make(M, N, K, False, False, dt if not simt else torch.bfloat16, seed=5)
    if simt:
        A, B = A.float(), B.float()
    g = torch.Generator().manual_seed(9)
    scale = (torch.rand(N, generator=g) + 0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    alpha_dev = torch.tensor(0.75, device="cuda")
    z_ref = ref * scale.cpu().double().numpy() * (0.5 * 0.75) + bias.cpu().double().numpy()
    tol = 2e-5 if out_dtype == torch.float32 else 1e-2
    # scale/alpha/bias + GELU forward with saved pre-activation
    aux = torch.empty(M, N, dtype=out_dtype, device="cuda")
    D = ops.gemm(A, B, M, N, K, out_dtype=out_dtype, col_scale=scale, bias=bias, alpha=0.5, alpha_dev=alpha_dev,
                 act=L.ACT_GELU, aux=aux, simt=simt)
    torch.cuda.synchronize()
    assert rel_err(aux.float().cpu().numpy(), z_ref) < tol
    assert rel_err(D.float().cpu().numpy(), gelu(z_ref)) < tol
    # GELU backward: D = z * gelu'(aux)
    pre = torch.randn(M, N, generator=g).to(out_dtype).cuda()
    D = ops.gemm(A, B, M, N, K, out_dtype=out_dtype, act=L.ACT_GELU_BWD, aux=pre, simt=simt)
    torch.cuda.synchronize()
    assert rel_err(D.float().cpu().numpy(), ref * gelu_grad(pre.float().cpu().double().numpy())) < tol


def test_gemm_headline_shape():
    """The last-layer forward at its real size (rows 2048, out_dim 65536, bottleneck 256), checked on a
    random sample of output entries and through linearity in A."""
    ops, _ = _ops()
    M, N, K = 2048, 65536, 256
    g = torch.Generator().manual_seed(7)
    A = torch.randn(M, K, generator=g).bfloat16().cuda()
    B = (torch.randn(N, K, generator=g) * 0.06).bfloat16().cuda()
    D = ops.gemm(A, B, M, N, K, out_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    rows = torch.randint(0, M, (64,), generator=g)
    cols = torch.randint(0, N, (64,), generator=g)
    ref = A[rows].double().cpu().numpy() @ B[cols].double().cpu().numpy().T
    got = D[rows][:, cols].float().cpu().numpy()
    assert rel_err(got, ref) < 1e-2                     # bf16 output rounding
    # linearity: (2A) B^T == 2 (A B^T) exactly in floating point (power-of-two scaling)
    D2 = ops.gemm((A.float() * 2).bfloat16(), B, M, N, K, out_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    assert torch.equal(D2[:256].float(), D[:256].float() * 2)


def test_gemm_argument_errors():
    ops, _ = _ops()
    A = torch.zeros(128, 64, dtype=torch.bfloat16, device="cuda")
    B = torch.zeros(128, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(ValueError):
        ops.gemm(A, B, 128, 128, 32)                   # K does not match the operands
    with pytest.raises(RuntimeError):
        ops.gemm(A.cpu(), B.cpu(), 128, 128, 64)       # no CPU path
    A_odd = torch.zeros(128, 20, dtype=torch.bfloat16, device="cuda")[:, :12]   # 40-byte row stride
    B_odd = torch.zeros(128, 20, dtype=torch.bfloat16, device="cuda")[:, :12]
    with pytest.raises(RuntimeError, match="16 bytes"):
        ops.gemm(A_odd, B_odd, 128, 128, 12)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,N,K", [(200, 1000, 64), (512, 4096, 256), (40, 264, 128)])
def test_gemm_fused_statistics(M, N, K, dtype):
    """EPI 2: softmax row partials (+ center) and 32-row column sums of the STORED output, merged by
    dmc_lse_finalize / dmc_teacher_finalize, against numpy on the stored output."""
    ops, _ = _ops()
    g = torch.Generator().manual_seed(11)
    A = torch.nn.functional.normalize(torch.randn(M, K, generator=g), dim=-1)      # unit rows, like the head's zhat
    B = torch.nn.functional.normalize(torch.randn(N, K, generator=g), dim=-1) * 1.5  # rows of norm g = 1.5
    if dtype == torch.bfloat16:
        Ad, Bd, kw = A.bfloat16().cuda(), B.bfloat16().cuda(), {}
    else:
        Ah, Al = ops.split_tf32(A.cuda())
        Bh, Bl = ops.split_tf32(B.cuda())
        Ad, Bd, kw = Ah, Bh, dict(A_lo=Al, B_lo=Bl)
    parts = ops.gemm_stats_parts(N)
    center = (torch.randn(N, generator=g) * 0.2).cuda()
    bound = torch.tensor(1.5, device="cuda")
    for use_center, use_bound in ((False, True), (False, False), (True, False)):
        rp = torch.zeros(M, parts, 2, device="cuda")
        cp = torch.zeros((M + 31) // 32, N, device="cuda")
        scale = 10.0 if not use_center else 25.0
        stats = dict(scale=scale, center=center if use_center else None, row_partials=rp, colsum_partials=cp,
                     bound=bound if use_bound else None)
        D = ops.gemm(Ad, Bd, M, N, K, out_dtype=dtype, stats=stats, **kw)
        torch.cuda.synchronize()
        Dn = D.double().cpu().numpy()                                  # the stored (rounded) values
        y = (Dn - (center.double().cpu().numpy() if use_center else 0.0)) * scale
        ref_lse = np.log(np.exp(y - y.max(-1, keepdims=True)).sum(-1)) + y.max(-1)
        lse = ops.lse_finalize(rp)
        assert rel_err(lse.cpu().numpy(), ref_lse) < 2e-6
        stats_t, colsum = ops.teacher_finalize(rp, cp, M, N)
        assert rel_err(colsum.cpu().numpy(), Dn.sum(0)) < 1e-5
        m2 = y.max(-1) * np.log2(np.e)
        if not use_bound:
            assert rel_err(stats_t[:, 0].cpu().numpy(), m2) < 1e-6
        l_ref = np.exp2(y * np.log2(np.e) - stats_t[:, 0].double().cpu().numpy()[:, None]).sum(-1)
        assert rel_err(stats_t[:, 1].cpu().numpy(), 1.0 / l_ref) < 1e-5


def test_gemm_cta_pair_mode_subprocess():
    """The opt-in CTA-pair (tcgen05 cta_group::2) schedule gives the same results (own process: the mode is read once
    from DMC_GEMM_FLAGS)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, torch, numpy as np
sys.path.insert(0, %r)
import dinomc_b200
ops = dinomc_b200.ops
worst = 0.0
for (M, N, K, a_mn, b_mn) in [(256, 512, 256, False, False), (300, 320, 136, False, True), (512, 256, 1000, True, True),
                              (2048, 256, 4096, False, True), (1024, 2048, 384, True, False)]:
    g = torch.Generator().manual_seed(0)
    A = torch.randn(M, K, generator=g).bfloat16().float(); B = torch.randn(N, K, generator=g).bfloat16().float()
    Ast = (A.t().contiguous() if a_mn else A).bfloat16().cuda(); Bst = (B.t().contiguous() if b_mn else B).bfloat16().cuda()
    D = ops.gemm(Ast, Bst, M, N, K, a_mn=a_mn, b_mn=b_mn)
    torch.cuda.synchronize()
    ref = A.double().numpy() @ B.double().numpy().T
    worst = max(worst, float(np.abs(D.cpu().double().numpy() - ref).max() / np.abs(ref).max()))
print("WORST", worst)
""" % root
    env = dict(os.environ, DMC_GEMM_FLAGS="64")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    worst = float(r.stdout.strip().split("WORST")[-1])
    assert worst < 2e-5


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(256, 512, 128), (200, 320, 136), (128, 64, 64), (384, 1000, 256)])
def test_gemm_lean_epilogue_bias_gelu(M, N, K, out_dtype):
    """bias + GELU (+ saved pre-activation) and GELU' through the lean full-tile epilogue (TMA store, no column scale)
    AND, for the ragged shapes, through the generic chunk code on the edge tiles of the same launch."""
    from oracle.np_oracle import gelu, gelu_grad
    ops, L = _ops()
    A, B, ref = _make(M, N, K, False, False, torch.bfloat16, seed=11)
    g = torch.Generator().manual_seed(12)
    bias = torch.randn(N, generator=g).cuda()
    z_ref = ref * 0.25 + bias.cpu().double().numpy()
    tol = 2e-5 if out_dtype == torch.float32 else 1e-2
    aux = torch.empty(M, N, dtype=out_dtype, device="cuda")
    D = ops.gemm(A, B, M, N, K, out_dtype=out_dtype, bias=bias, alpha=0.25, act=L.ACT_GELU, aux=aux)
    torch.cuda.synchronize()
    assert rel_err(aux.float().cpu().numpy(), z_ref) < tol
    assert rel_err(D.float().cpu().numpy(), gelu(z_ref)) < tol
    Dp = ops.gemm(A, B, M, N, K, out_dtype=out_dtype, bias=bias, alpha=0.25)          # bias only
    torch.cuda.synchronize()
    assert rel_err(Dp.float().cpu().numpy(), z_ref) < tol
    pre = (torch.randn(M, N, generator=g) * 2).to(out_dtype).cuda()
    Db = ops.gemm(A, B, M, N, K, out_dtype=out_dtype, act=L.ACT_GELU_BWD, aux=pre)
    torch.cuda.synchronize()
    assert rel_err(Db.float().cpu().numpy(), ref * gelu_grad(pre.float().cpu().double().numpy())) < tol


def test_gelu_device_function_accuracy():
    """The branch-free GELU / GELU' of the epilogues (erfc by Abramowitz-Stegun 7.1.26) against fp64 erf on a dense
    grid, through a K=1 'GEMM' whose accumulator is the grid itself (fp32 mode parity needs ~1e-6 absolute)."""
    from oracle.np_oracle import gelu, gelu_grad
    ops, L = _ops()
    M, N = 256, 512
    x = torch.linspace(-9.0, 9.0, M * N, dtype=torch.float64).reshape(M, N)
    # D = A (M x 8) . B^T (N x 8) with A = [x_hi, x_lo, 0...] rows ... simpler: aux carries x for GELU', ones product for GELU
    A = torch.zeros(M, 8); A[:, 0] = 1.0
    B = torch.zeros(N, 8); B[:, 0] = 1.0
    pre = x.float().cuda()
    Db = ops.gemm(A.cuda(), B.cuda(), M, N, 8, out_dtype=torch.float32, act=L.ACT_GELU_BWD, aux=pre, simt=True)
    torch.cuda.synchronize()
    assert np.abs(Db.cpu().double().numpy() - gelu_grad(pre.cpu().double().numpy())).max() < 1e-6
    # GELU forward: bias carries one grid row per launch (accumulator 0 + bias)
    Az = torch.zeros(M, 8).cuda()
    bias = torch.linspace(-9.0, 9.0, N).cuda()
    Dg = ops.gemm(Az, B.cuda(), M, N, 8, out_dtype=torch.float32, bias=bias, act=L.ACT_GELU, simt=True)
    torch.cuda.synchronize()
    assert np.abs(Dg[0].cpu().double().numpy() - gelu(bias.cpu().double().numpy())).max() < 1e-6


def test_gemm_long_contraction_uses_pairs_and_matches():
    """K >= 32768 with N <= 256 (the last layer's dgrad shape class) runs on CTA pairs with split-K; same numbers."""
    ops, _ = _ops()
    M, N, K = 384, 256, 32768 + 192
    A, B, ref = _make(M, N, K, False, True, torch.bfloat16, seed=21)
    D = ops.gemm(A, B, M, N, K, b_mn=True, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert rel_err(D.cpu().numpy(), ref) < 2e-5


def test_pdl_toggle_is_bit_identical():
    """dmc_set_pdl only changes HOW kernels are launched (programmatic dependent launch), never what they compute."""
    ops, L = _ops()
    lib = L.load()
    M, N, K = 512, 1024, 384
    A, B, _ = _make(M, N, K, False, False, torch.bfloat16, seed=31)
    bias = torch.randn(N).cuda()
    outs = []
    prev = lib.dmc_set_pdl(1)
    try:
        for on in (1, 0, 1):
            lib.dmc_set_pdl(on)
            h = ops.gemm(A, B, M, N, K, out_dtype=torch.bfloat16, bias=bias, act=L.ACT_GELU)
            z, _, _ = ops.normalize_rows_fwd(h.float(), want_bf16=False)
            cs = ops.colsum(z)
            torch.cuda.synchronize()
            outs.append((h.clone(), z.clone(), cs.clone()))
    finally:
        lib.dmc_set_pdl(prev)
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert torch.equal(a, b)
