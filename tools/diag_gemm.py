"""Diagnostic (not a test): run the tcgen05 GEMM in every operand layout / dtype, each in its own
process so a trap in one variant does not poison the CUDA context of the next.  Prints one line per case."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASE = r"""
import sys, torch, numpy as np
sys.path.insert(0, %(root)r)
import dinomc_b200
ops = dinomc_b200.ops
kind, a_mn, b_mn, M, N, K = %(kind)r, %(a_mn)d, %(b_mn)d, %(M)d, %(N)d, %(K)d
g = torch.Generator().manual_seed(0)
A = torch.randn(M, K, generator=g); B = torch.randn(N, K, generator=g)
if kind == 'bf16':
    A, B = A.bfloat16().float(), B.bfloat16().float()
dt = torch.bfloat16 if kind == 'bf16' else torch.float32
Ast = (A.t().contiguous() if a_mn else A).to(dt).cuda(); Bst = (B.t().contiguous() if b_mn else B).to(dt).cuda()
ref = A.double().numpy() @ B.double().numpy().T
if kind == 'tf32x3':
    Ah, Al = ops.split_tf32(Ast); Bh, Bl = ops.split_tf32(Bst)
    D = ops.gemm(Ah, Bh, M, N, K, a_mn=bool(a_mn), b_mn=bool(b_mn), A_lo=Al, B_lo=Bl)
else:
    D = ops.gemm(Ast, Bst, M, N, K, a_mn=bool(a_mn), b_mn=bool(b_mn))
torch.cuda.synchronize()
d = D.cpu().double().numpy()
err = np.abs(d - ref).max() / np.abs(ref).max()
bad = np.argwhere(np.abs(d - ref) > 1e-2 * np.abs(ref).max())
print('err=%%.3e nbad=%%d first_bad=%%s' %% (err, len(bad), bad[:3].tolist()))
"""


def main():
    shapes = [(128, 256, 64), (256, 128, 256), (200, 320, 136)]
    for kind in ("bf16", "tf32", "tf32x3"):
        for a_mn in (0, 1):
            for b_mn in (0, 1):
                for (M, N, K) in shapes:
                    code = CASE % dict(root=ROOT, kind=kind, a_mn=a_mn, b_mn=b_mn, M=M, N=N, K=K)
                    try:
                        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
                        out = (r.stdout.strip().splitlines() or ["<no output>"])[-1]
                        if r.returncode != 0:
                            out = "FAILED rc=%d %s" % (r.returncode, (r.stderr.strip().splitlines() or [""])[-1][:200])
                    except subprocess.TimeoutExpired:
                        out = "TIMEOUT"
                    print(f"{kind:7s} a_mn={a_mn} b_mn={b_mn} M={M:4d} N={N:4d} K={K:4d}: {out}", flush=True)


if __name__ == "__main__":
    main()
