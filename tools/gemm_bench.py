"""Stand-alone timing of dmc_gemm on the shapes of the step (CUDA events, L2-sized outputs).  Not a test."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dinomc_b200
ops = dinomc_b200.ops

SHAPES = {
    # name: (M, N, K, a_mn, b_mn, in dtype, out dtype)
    "last_fwd_student": (2048, 65536, 256, False, False, torch.bfloat16, torch.bfloat16),
    "last_fwd_teacher": (512, 65536, 256, False, False, torch.bfloat16, torch.bfloat16),
    "last_fwd_student_bound_stats": (2048, 65536, 256, False, False, torch.bfloat16, torch.bfloat16),
    "last_fwd_student_max_stats": (2048, 65536, 256, False, False, torch.bfloat16, torch.bfloat16),
    "last_dgrad": (2048, 256, 65536, False, True, torch.bfloat16, torch.float32),
    "last_wgrad": (65536, 256, 2048, True, True, torch.bfloat16, torch.float32),
    "big8192": (8192, 8192, 8192, False, False, torch.bfloat16, torch.bfloat16),
    "big4096": (4096, 4096, 4096, False, False, torch.bfloat16, torch.bfloat16),
    "mlp_fwd1": (2048, 2048, 384, False, False, torch.bfloat16, torch.bfloat16),
    "mlp_fwd2": (2048, 2048, 2048, False, False, torch.bfloat16, torch.bfloat16),
    "mlp_fwd3": (2048, 256, 2048, False, False, torch.bfloat16, torch.float32),
    "mlp_wgrad2": (2048, 2048, 2048, True, True, torch.bfloat16, torch.float32),
    "mlp_dgrad2": (2048, 2048, 2048, False, True, torch.bfloat16, torch.bfloat16),
}

def run(name, iters=10):
    M, N, K, a_mn, b_mn, dt, odt = SHAPES[name]
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(dt)
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(dt)
    out = torch.empty(M, N, dtype=odt, device="cuda")
    stats = None
    if name.endswith("_stats"):
        A = torch.nn.functional.normalize(A.float(), dim=-1).to(dt)
        B = torch.nn.functional.normalize(B.float(), dim=-1).to(dt)
        stats = dict(scale=10.0, center=None, row_partials=torch.empty(M, ops.gemm_stats_parts(N), 2, device="cuda"),
                     bound=torch.tensor(1.0, device="cuda") if "bound" in name else None)
    if stats is not None:
        import functools
        ops_gemm = functools.partial(ops.gemm, stats=stats)
    else:
        ops_gemm = ops.gemm
    for _ in range(3):
        ops_gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops_gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # the library GEMM for context (cuBLAS through torch), same operands
    Am = A.t() if a_mn else A
    Bm = B if b_mn else B.t()
    for _ in range(3):
        torch.matmul(Am, Bm)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(Am, Bm)
    e1.record()
    torch.cuda.synchronize()
    ms_lib = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
    print(f"{name:18s} M={M:6d} N={N:6d} K={K:6d}  ours {ms*1e3:8.1f} us ({tf:7.1f} TF/s)   cublas {ms_lib*1e3:8.1f} us", flush=True)

if __name__ == "__main__":
    names = sys.argv[1:] or list(SHAPES)
    for n in names:
        run(n)
