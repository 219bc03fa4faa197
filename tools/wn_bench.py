import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dinomc_b200
ops = dinomc_b200.ops
K=65536
v=torch.randn(K,256,device="cuda"); g=torch.ones(K,device="cuda")
def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(gr, stream=s):
            for _ in range(reps): fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps*1e3
print("weightnorm_fwd bf16 (incl. absmax): %.1f us" % timed(lambda: ops.weightnorm_fwd(v,g,"bf16")))
w,_,sc,iv = ops.weightnorm_fwd(v,g,"bf16")
ref = (v/ v.norm(dim=1,keepdim=True)).bfloat16()
print("max diff", (w.float()-ref.float()).abs().max().item())
