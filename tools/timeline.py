"""Kernel timeline of graph replays of the KD step (torch.profiler / CUPTI): which kernels run on which stream, when.
Analysis aid only -- numbers taken under a profiler are never bench values.  Writes gpurun_out/timeline.csv + a summary."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from imagecaptioner_b200.distillation_utils import DistillationLoss
from imagecaptioner_b200.graph import GraphedKDStep
from imagecaptioner_b200.optim import FlatAdamW, reference_param_groups
from oracle import kd_oracle as O
from tests.harness import build_student

dev = torch.device("cuda:0")
cfg = bench.CFG
B, T, V, E, H, L = cfg["B"], cfg["T"], cfg["V"], cfg["E"], cfg["H"], cfg["L"]
params = O.init_student_params(V, E, H, L, True, seed=0)
pparams = O.init_projector_params(cfg["Et"], E, seed=1)
model, projector = build_student(params, pparams, V, E, H, L, True, cfg["Et"], dev)
model.decoder.compute_dtype = torch.bfloat16
if os.environ.get("B2C_TL_TRAIN"):          # training mode (dropout on), like bench.py's value_dropout
    model.train(); projector.train()
loss_mod = DistillationLoss(0.7, 0.2, 0.1, 4.0, vocab_size=V)
opt = FlatAdamW(reference_param_groups(model, projector, 1e-4), weight_decay=0.01, max_grad_norm=1.0)
host = bench.make_batch(cfg, 1234)
resident = {k: host[k].to(dev) for k in GraphedKDStep.INPUT_KEYS}
kd = GraphedKDStep(model, projector, loss_mod, opt, None, resident, autocast_dtype=torch.bfloat16)
for _ in range(10):
    kd.step()
torch.cuda.synchronize()
NREP = 3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(NREP):
        kd.step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
rows = sorted(((e.time_range.start, e.time_range.end, getattr(e, "device_resource_id", -1), e.name) for e in evs), key=lambda r: r[0])
os.makedirs("gpurun_out", exist_ok=True)
# keep the LAST replay only
n_per = len(rows) // NREP
last = rows[-n_per:]
t0 = last[0][0]
with open("gpurun_out/timeline.csv", "w") as f:
    f.write("start_us,end_us,dur_us,stream,name\n")
    for s, e, st, name in last:
        f.write(f"{s - t0:.2f},{e - t0:.2f},{e - s:.2f},{st},\"{name[:120]}\"\n")
span = max(r[1] for r in last) - t0
print(f"kernels/replay {n_per}  span {span:.1f} us")
streams = {}
for s, e, st, name in last:
    streams.setdefault(st, []).append((s - t0, e - t0, name))
for st, ks in streams.items():
    busy = sum(e - s for s, e, _ in ks)
    print(f"stream {st}: {len(ks)} kernels, busy {busy:.1f} us, first {ks[0][0]:.1f} last-end {ks[-1][1]:.1f}")
# the longest kernels and the gaps on the busiest stream (where does the step time go)
top = sorted(last, key=lambda r: r[0] - r[1])[:12]
for s, e, st, name in top:
    print(f"  {e - s:9.1f} us  start {s - t0:8.1f}  stream {st}  {name[:90]}")
fam = {}
for s, e, st, name in last:
    key = name.split("<")[0].split("(")[0][-48:]
    fam.setdefault(key, [0, 0.0]); fam[key][0] += 1; fam[key][1] += e - s
for key, (n, tot) in sorted(fam.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {tot:9.1f} us  x{n:<4d} {key}")
