"""One short GPU call (the round's last ~20 s of box time): smoke() -- parity of the default routes against the oracle in both
modes -- then the cfg2 bench line with the per-op CUPTI pass (PDL off), in ONE process (one torch import)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
t0 = time.time()
import __graft_entry__ as g  # noqa: E402

g.smoke()
print(f"[last shot] smoke done at {time.time() - t0:.1f} s", flush=True)
import bench  # noqa: E402

out = os.open(os.path.join(ROOT, "gpurun_out", "r02x_bench.json"), os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
bench._JSON_FD = out                      # the JSON line goes to the file, everything else stays on stdout / stderr
sys.argv = ["bench.py", "--steps", "20", "--warmup", "3", "--no-cpu-baseline"]
bench.main()
print(f"[last shot] bench done at {time.time() - t0:.1f} s", flush=True)
