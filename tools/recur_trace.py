"""Phase breakdown of the persistent forward-recurrence kernel (B2C_RECUR_TRACE=1): clock64() stamps of every CTA's epilogue group at
the phase boundaries of every time step -> average cycles per phase.  Analysis aid, not a bench value."""
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
B, T, V, E, H, L = cfg["B"], cfg["T"], cfg["V"], cfg["E"], cfg["H"], cfg["L"]
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
n = 148 * T * 16
buf = (ctypes.c_uint64 * n)()
g, s_ = ctypes.c_int32(), ctypes.c_int32()
_ops._check(lib.b2c_debug_recur_trace(buf, n, ctypes.byref(g), ctypes.byref(s_)), "trace")
G, Ts = g.value, s_.value
a = np.frombuffer(buf, dtype=np.uint64)[: G * Ts * 16].reshape(G, Ts, 16).astype(np.float64)
step = a[:, 1:, 0] - a[:, :-1, 0]
print(f"grid {G} steps {Ts}; cycles per step (mean over CTAs, steps 1..): {step.mean():.0f}  (min CTA {step.mean(1).min():.0f}, max {step.mean(1).max():.0f})")


def show(name, x):
    x = x[:, 1:]
    print(f"  {name:58s} mean {x.mean():8.0f}  p10 {np.percentile(x, 10):8.0f}  p90 {np.percentile(x, 90):8.0f}")


show("epi: step start -> u tile stored, arrived", a[:, :, 1] - a[:, :, 0])
show("epi: wait barrier 0 (all u tiles)", a[:, :, 2] - a[:, :, 1])
show("epi: e^{2u} into smem", a[:, :, 3] - a[:, :, 2])
show("epi: scores (stream e^{2P}, rcp)", a[:, :, 4] - a[:, :, 3])
show("epi: softmax", a[:, :, 5] - a[:, :, 4])
show("epi: context", a[:, :, 6] - a[:, :, 5])
show("epi: attention arrive -> layer 0 cell done", a[:, :, 7] - a[:, :, 6])
show("epi: layer 0 arrive -> layer 1 cell done", a[:, :, 8] - a[:, :, 7])
tail = a[:, 1:, 0] - a[:, :-1, 8]
print(f"  {'epi: layer 1 arrive -> next step start':58s} mean {tail.mean():8.0f}")
show("producer: step-start barrier passed (after epi step start)", a[:, :, 15] - a[:, :, 0])
ut = a[:32]
show("mma (u-tile CTAs): U job issued (after producer start)", ut[:, :, 11] - ut[:, :, 15])
show("mma: early halves issued (after producer start)", a[:, :, 12] - a[:, :, 15])
show("mma: late 0 issued after attention-done stamp", a[:, :, 13] - a[:, :, 6])
show("mma: late 1 issued after layer-0-done stamp", a[:, :, 14] - a[:, :, 7])
