"""Launcher: CUPTI timeline of one graph replay of the 2-GPU (or N-GPU) data-parallel step, rank 0.  torchrun inside."""
import os
import subprocess
import sys

n = int(os.environ.get("DMC_PROF_GPUS", "2"))
here = os.path.dirname(os.path.abspath(__file__))
sys.exit(subprocess.call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                          "--master-port", "29547", os.path.join(here, "prof_step_dp_worker.py")] + sys.argv[1:]))
