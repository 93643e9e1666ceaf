"""Phase breakdown of the cluster recurrence kernel (B2C_RECUR_TRACE=1): clock64() stamps of worker thread 0 of every CTA at the phase
boundaries of every time step -> average cycles per phase.  Analysis aid, not a bench value.  Usage: cluster_trace.py [B]"""
import ctypes
import os
import sys

os.environ["B2C_RECUR_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from imagecaptioner_b200 import _ops
from oracle import kd_oracle as O
from tests.harness import build_student

dev = torch.device("cuda:0")
cfg = bench.CFG
B = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["B"]
T, V, E, H, L = cfg["T"], cfg["V"], cfg["E"], cfg["H"], cfg["L"]
params = O.init_student_params(V, E, H, L, True, seed=0)
model, _ = build_student(params, {}, V, E, H, L, True, E, dev)
model.decoder.compute_dtype = torch.bfloat16
feats = torch.randn(B, 49, E, device=dev)
cap = torch.randint(4, V, (T, B), device=dev)
with torch.no_grad():
    for _ in range(3):
        model.decoder(feats.bfloat16(), cap)
torch.cuda.synchronize()
lib = _ops.load_library()
n = 160 * T * 8
buf = (ctypes.c_uint64 * n)()
g, s_ = ctypes.c_int32(), ctypes.c_int32()
_ops._check(lib.b2c_debug_recur_trace(buf, n, ctypes.byref(g), ctypes.byref(s_)), "trace")
G, Ts = g.value, s_.value
a = np.frombuffer(buf, dtype=np.uint64)[: G * Ts * 8].reshape(G, Ts, 8).astype(np.float64)
step = a[:, 1:, 0] - a[:, :-1, 0]
print(f"B {B}: grid {G} CTAs, {Ts} steps; cycles per step (mean over CTAs, steps 1..): {step.mean():.0f}  (min CTA {step.mean(1).min():.0f}, max {step.mean(1).max():.0f})")
names = ["step start -> u accumulator ready, u rows sent", "wait own u rows, e^{2u}", "scores + softmax", "context, ctx rows stored, multicast issued",
         "layer-0 accumulators + cell + h0 multicast", "layer-1 accumulators + cell + h1 multicast"]
for i, nm in enumerate(names):
    x = (a[:, :, i + 1] - a[:, :, i])[:, 1:]
    print(f"  {nm:52s} mean {x.mean():8.0f}  p10 {np.percentile(x, 10):8.0f}  p90 {np.percentile(x, 90):8.0f}")
