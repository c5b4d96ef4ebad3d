"""dense_lucaskanade a few times on 2 frames of 2048^2 (for ncu launch lists)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pysteps_b200
from pysteps_b200 import _synthetic as syn

m = n = int(os.environ.get("SIZE", "2048"))
frames = torch.from_numpy(syn.rain_frames(m, n, 2, 0, dx=3, dy=-2)).cuda()
lk = pysteps_b200.motion.get_method("lk")
for _ in range(int(os.environ.get("REPS", "3"))):
    V = lk(frames)
torch.cuda.synchronize()
