"""bf16-mode parity table: per-tensor L2 error of the native kernels and of the reference's own stock-module arithmetic under
torch.autocast(bf16), both against the fp64 oracle, at BASELINE config-1 and config-2 shapes.  -> profiles/r2_bf16_parity.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import kd_oracle as O
from tests.harness import autocast_reference_errors, build_student, run_kd_step, step_errors, to_device

DEV = "cuda:0"
out = []
for name, B in (("config 1 (B=16)", 16), ("config 2 (B=512)", 512)):
    V, E, H, L, T = 5000, 256, 512, 2, 20
    params = O.init_student_params(V, E, H, L, True, seed=0)
    pparams = O.init_projector_params(384, E, seed=1)
    batch = O.synthetic_batch(B, T, V, E, H, seed=1234)
    ref = O.kd_step(to_device(params, DEV), to_device(pparams, DEV), to_device(batch, DEV), dtype=torch.float64)
    model, projector = build_student(params, pparams, V, E, H, L, True, 384, DEV)
    got = step_errors(run_kd_step(model, projector, batch, DEV, torch.bfloat16), ref, "l2")
    got32 = step_errors(run_kd_step(model, projector, batch, DEV, torch.float32), ref, "max")
    auto = autocast_reference_errors(params, pparams, dict(V=V, E=E, H=H, L=L), batch, ref, DEV)      # loss-scaled by 2^16 like the reference's GradScaler
    from oracle import eager_torch as ET
    m_, p_ = ET.build(params, pparams, V, E, H, L, True, 384, 49, DEV)
    unscaled = step_errors(ET.kd_step(m_, p_, batch, torch.bfloat16), ref, "l2")
    out.append(f"   (autocast reference WITHOUT loss scaling: worst gradient error {max(v for k, v in unscaled.items() if 'grad' in k):.3e})")
    out.append(f"== {name}: T=20 V=5000 E256 H512 L2; per-tensor relative L2 error vs the fp64 oracle (fp32 column: max-norm)")
    out.append(f"{'tensor':58s} {'kernel bf16':>12s} {'autocast ref':>12s} {'ratio':>7s} {'kernel fp32':>12s}")
    better = 0
    for k in got:
        r = got[k] / max(auto[k], 1e-30)
        better += r < 1.0
        out.append(f"{k:58s} {got[k]:12.3e} {auto[k]:12.3e} {r:7.2f} {got32[k]:12.3e}")
    out.append(f"kernel more accurate than the autocast reference on {better} of {len(got)} quantities; worst kernel bf16 {max(got.values()):.3e}, worst autocast {max(auto.values()):.3e}")
    out.append("")
txt = "\n".join(out)
print(txt)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/r2_bf16_parity.txt", "w").write(txt)
