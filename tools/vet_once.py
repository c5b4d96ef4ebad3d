"""vet() at 2048^2 once after a small warm-up (for ncu captures of vet_eval_kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pysteps_b200
from pysteps_b200 import _synthetic as syn

fr = syn.rain_frames(2048, 2048, 2, 0)
vet = pysteps_b200.motion.get_method("vet")
vet(fr[:, :256, :256], verbose=False, options={"maxiter": 2})
V = vet(fr, verbose=False, options={"maxiter": int(os.environ.get("MAXITER", "3"))})
torch.cuda.synchronize()
