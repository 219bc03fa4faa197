"""Timing of the small (MLP-sized) GEMMs of the step under forced split-K choices.  Not a test.
   python tools/gemm_small.py            # every shape x split_k in (auto, 1, 2, 4, 8)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dinomc_b200
ops = dinomc_b200.ops
bf, f32 = torch.bfloat16, torch.float32
SHAPES = {
    # name: (M, N, K, a_mn, b_mn, out dtype)
    "tiny": (128, 256, 64, False, False, bf),
    "tiny2": (128, 256, 2048, False, False, bf),
    "s_fwd1": (2048, 2048, 384, False, False, bf),
    "s_fwd2": (2048, 2048, 2048, False, False, bf),
    "s_fwd3": (2048, 256, 2048, False, False, f32),
    "t_fwd1": (512, 2048, 384, False, False, bf),
    "t_fwd2": (512, 2048, 2048, False, False, bf),
    "t_fwd3": (512, 256, 2048, False, False, f32),
    "wgrad3": (256, 2048, 2048, True, True, f32),
    "wgrad2": (2048, 2048, 2048, True, True, f32),
    "wgrad1": (2048, 384, 2048, True, True, f32),
    "dgrad3": (2048, 2048, 256, False, True, bf),
    "dgrad2": (2048, 2048, 2048, False, True, bf),
    "dgrad1": (2048, 384, 2048, False, True, f32),
}

def t_us(fn, iters=20):
    """GPU time per call with the host out of the picture: `iters` calls captured in one CUDA graph."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(iters):
                fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def t_us_eager(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

if __name__ == "__main__":
    names = sys.argv[1:] or list(SHAPES)
    for name in names:
        M, N, K, a_mn, b_mn, odt = SHAPES[name]
        A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(bf)
        B = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(bf)
        out = torch.empty(M, N, dtype=odt, device="cuda")
        res = []
        for sk in (0, 1, 2, 4, 8, 16):
            if sk > max(K // 64 // 1, 1):
                continue
            res.append(f"sk{sk}:{t_us(lambda: ops.gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, out=out, split_k=sk)):6.1f}")
        Am, Bm = (A.t() if a_mn else A), (B if b_mn else B.t())
        lib = t_us(lambda: torch.matmul(Am, Bm))
        print(f"{name:8s} M={M:5d} N={N:5d} K={K:5d}  " + "  ".join(res) + f"   cublas {lib:6.1f} us", flush=True)
