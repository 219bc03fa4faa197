"""Pipeline trace of CTA 0 of one dmc_gemm launch.  DMC_GEMM_TRACE=1 makes the library print clock64 stamps of the
producer / MMA / epilogue hand-offs; this tool runs one shape in a child process and summarises them.  Debug tool.
   python tools/gemm_trace.py M N K [a_mn b_mn split_k out_f32] [--raw]"""
import os, re, subprocess, sys

if os.environ.get("DMC_GEMM_TRACE_CHILD"):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import dinomc_b200
    ops = dinomc_b200.ops
    M, N, K, a_mn, b_mn, sk, of32 = (int(v) for v in sys.argv[1:8])
    bf = torch.bfloat16
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(bf)
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(bf)
    out = torch.empty(M, N, dtype=torch.float32 if of32 else bf, device="cuda")
    for _ in range(2):
        ops.gemm(A, B, M, N, K, a_mn=bool(a_mn), b_mn=bool(b_mn), out=out, split_k=sk)
    torch.cuda.synchronize()
    sys.exit(0)

args = [a for a in sys.argv[1:] if not a.startswith("--")]
vals = (args + ["0"] * 7)[:7]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
trace_lib = os.path.join(root, "self-supervised-learning-for-aerial-image-segmentation_b200", "libdinomc_trace.so")
if not os.path.isfile(trace_lib):
    sys.exit("build the trace library first: DMC_TRACE=1 ./build.sh")
env = dict(os.environ, DMC_GEMM_TRACE="1", DMC_GEMM_TRACE_CHILD="1", DMC_LIB=trace_lib)
r = subprocess.run([sys.executable, __file__] + vals, env=env, capture_output=True, text=True)
blocks = r.stderr.split("TRACE ")
if len(blocks) < 3:
    print(r.stderr[-2000:]); sys.exit(1)
txt = "TRACE " + blocks[2].split("TRACE ")[0]        # second launch of the main kernel (warm)
if "--raw" in sys.argv:
    print(txt)
print(txt.splitlines()[0])
kb = [[int(v) for v in re.findall(r"=(-?\d+)", l)] for l in txt.splitlines() if l.strip().startswith("kb")]
tl = [[int(v) for v in re.findall(r"=(-?\d+)", l)] for l in txt.splitlines() if l.strip().startswith("tile")]
if kb:
    n = len(kb)
    d = lambda c: (kb[-1][c] - kb[0][c]) / max(n - 1, 1)
    print(f"  k-blocks traced {n}: first P.empty {kb[0][0]}  first M.full {kb[0][2]}  last M.commit {kb[-1][3]}")
    print(f"  per k-block: producer {d(0):.0f} cyc  (issue {sum(k[1]-k[0] for k in kb)/n:.0f})   mma {d(2):.0f} cyc  (issue {sum(k[3]-k[2] for k in kb)/n:.0f})")
for i, t in enumerate(tl[:6]):
    print(f"  tile{i}: M.tmem_empty {t[0]}  E.tmem_full {t[1]}  E.done {t[2]}  (epilogue {t[2]-t[1]})")
if len(tl) > 6:
    print(f"  ... {len(tl)} tiles; last E.done {tl[-1][2]}")
