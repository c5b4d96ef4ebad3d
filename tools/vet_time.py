import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import pysteps_b200
from pysteps_b200 import _synthetic as syn, _lib
fr = syn.rain_frames(2048, 2048, 2, 0)
vet = pysteps_b200.motion.get_method("vet")
V = vet(fr[:, :256, :256], verbose=False)  # warm-up
torch.cuda.synchronize()
calls = {"n": 0}
orig = _lib.call
def counting(name, *a):
    if name in ("b200_vet_cost", "b200_vet_value_and_gradient"): calls["n"] += 1
    return orig(name, *a)
_lib.call = counting
import pysteps_b200.motion.vet as vm
t = time.time(); V = vet(fr, verbose=False); torch.cuda.synchronize(); dt = time.time() - t
wet = fr[1] > 0
print("VET 2048^2: %.3f s, %d fused value+gradient evaluations, mean V in rain (%.4f, %.4f)" % (dt, calls["n"], V[0][wet].mean(), V[1][wet].mean()))
with _lib.Trace() as tr:
    V = vet(fr, verbose=False)
s = tr.summary()
print({k: (len(v), round(sum(v), 2)) for k, v in s.items()})
