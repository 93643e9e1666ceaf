"""Shared by tests/, __graft_entry__.smoke() and bench.py: build the public modules from an oracle parameter
dict and replay the reference's KD step structure (src/train_student_kd.py:262-288) on a synthetic batch."""
from __future__ import annotations

import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def build_student(params, proj_params, V, E, H, L, refinement, Et, device, dropout=0.0):
    """CaptioningStudent (features fed directly) + FeatureProjector loaded from oracle/reference state_dicts."""
    from imagecaptioner_b200.student_model import CaptioningStudent, PrecomputedFeatures
    from imagecaptioner_b200.distillation_utils import FeatureProjector
    model = CaptioningStudent(V, E, H, L, dropout=dropout, use_attention_refinement=refinement, encoder=PrecomputedFeatures(E))
    missing, unexpected = model.load_state_dict({k: v.float() for k, v in params.items()}, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("encoder.") for k in missing), missing
    projector = FeatureProjector(Et, E, 197, 49)
    if proj_params:
        projector.load_state_dict({k: v.float() for k, v in proj_params.items()})
    return model.to(device).eval(), projector.to(device).eval()


def run_kd_step(model, projector, batch, device, dtype=torch.float32, alpha=0.7, beta=0.2, gamma=0.1, temperature=4.0):
    """forward + DistillationLoss + backward; returns what oracle.kd_step returns (on the CPU, fp32)."""
    from imagecaptioner_b200.distillation_utils import DistillationLoss
    model.zero_grad(set_to_none=True)
    projector.zero_grad(set_to_none=True)
    model.decoder.compute_dtype = dtype                     # the whole native path runs in this precision mode
    if getattr(model, "use_attention_refinement", False):
        model.attention_refinement.compute_dtype = dtype
    projector.compute_dtype = dtype
    feats = batch["encoder_features"].to(device).clone().requires_grad_(True)
    cap = batch["captions_input"].to(device)
    tgt = batch["targets"].to(device)
    outputs, enc, hids, atts = model(feats, cap)
    th = batch.get("teacher_hiddens")
    t_out = {"logits": batch["teacher_logits"].to(device),
             "encoder_features": projector(batch["teacher_features"].to(device)),
             "hidden_states": None if th is None else [th[t].to(device) for t in range(th.shape[0])]}
    s_out = {"logits": outputs, "encoder_features": enc, "hidden_states": hids}
    loss_mod = DistillationLoss(alpha, beta, gamma, temperature, vocab_size=outputs.shape[-1])
    total, loss_dict = loss_mod(s_out, t_out, tgt)
    total.backward()
    f32 = lambda t: t.detach().float().cpu()
    return {
        "loss": loss_dict,
        "grads": {k: f32(v.grad) for k, v in model.named_parameters() if v.grad is not None},
        "proj_grads": {k: f32(v.grad) for k, v in projector.named_parameters() if v.grad is not None},
        "d_encoder_features": f32(feats.grad),
        "logits": f32(outputs), "hidden_states": f32(torch.stack(list(hids))),
        "attention_weights": f32(torch.stack(list(atts))),
        "teacher_projected": f32(t_out["encoder_features"]),
    }


def relerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def relerr_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def compare_step(got, ref, tol, verbose=True, metric="max", loosen=None):
    """Relative error of every output / gradient / loss component (max-norm: max|a-b| / max|b|, or per-tensor L2:
    ||a-b|| / ||b||); asserts < tol (or tol * loosen[name] for the listed names); returns the worst."""
    relerr = globals()["relerr"] if metric == "max" else relerr_l2
    loosen = loosen or {}
    rows = []
    for k in ("logits", "hidden_states", "attention_weights", "teacher_projected", "d_encoder_features"):
        rows.append((k, relerr(got[k], ref[k])))
    for k, v in ref["grads"].items():
        rows.append(("grad:" + k, relerr(got["grads"][k], v)))
    for k, v in ref["proj_grads"].items():
        rows.append(("pgrad:" + k, relerr(got["proj_grads"][k], v)))
    for k, v in ref["loss"].items():
        rows.append(("loss:" + k, abs(got["loss"][k] - v) / (abs(v) + 1e-30) if v != 0 else abs(got["loss"][k])))
    worst = max(r[1] for r in rows)
    lim = lambda name: tol * loosen.get(name, 1.0)
    bad = [r for r in rows if not r[1] < lim(r[0])]
    if verbose or bad:
        for name, e in rows:
            print(f"   {name:55s} {e:.3e}{'   <-- FAIL' if not e < lim(name) else ''}")
    assert not bad, f"{len(bad)} quantities exceed rel tol {tol}: {bad[:4]}"
    return worst


def to_device(d, device, dtype=None):
    """dict of tensors (or None) -> same dict on `device` (floating tensors optionally cast)."""
    out = {}
    for k, v in d.items():
        if v is None or not torch.is_tensor(v):
            out[k] = v
        else:
            out[k] = v.to(device=device, dtype=dtype) if (dtype is not None and v.is_floating_point()) else v.to(device)
    return out


def step_errors(got, ref, metric="l2"):
    """name -> relative error of every output / gradient / loss component of a KD step (same names as compare_step)."""
    err = relerr if metric == "max" else relerr_l2
    cpu = lambda t: t.detach().float().cpu() if torch.is_tensor(t) else t
    rows = {}
    for k in ("logits", "hidden_states", "attention_weights", "teacher_projected", "d_encoder_features"):
        rows[k] = err(cpu(got[k]), cpu(ref[k]))
    for k, v in ref["grads"].items():
        rows["grad:" + k] = err(cpu(got["grads"][k]), cpu(v))
    for k, v in ref["proj_grads"].items():
        rows["pgrad:" + k] = err(cpu(got["proj_grads"][k]), cpu(v))
    for k, v in ref["loss"].items():
        rows["loss:" + k] = abs(got["loss"][k] - v) / (abs(v) + 1e-30) if v != 0 else abs(got["loss"][k])
    return rows


def autocast_reference_errors(params, proj_params, meta, batch, ref, device, autocast_dtype=torch.bfloat16, metric="l2", use_refinement=True):
    """What reduced-precision autocast does to the REFERENCE's own arithmetic: the stock torch.nn composition of the reference
    (oracle/eager_torch.py) run on the GPU under torch.autocast(dtype), per-tensor error against `ref` (an fp32 / fp64 result).
    This is the yardstick for the bf16 mode of the native kernels: kernel error <= max(2e-2, 1.2 x this)."""
    from oracle import eager_torch as ET
    Et = batch["teacher_features"].shape[-1]
    model, proj = ET.build(params, proj_params, meta["V"], meta["E"], meta["H"], meta["L"], use_refinement, Et, batch["encoder_features"].shape[1], device)
    # with the reference's own loss scaling (GradScaler's initial 2^16): un-scaled, the stock path drops most of the recurrent
    # gradients at batch 512 (see eager_torch.kd_step), which would make the yardstick meaningless
    got = ET.kd_step(model, proj, batch, autocast_dtype, meta.get("alpha", 0.7), meta.get("beta", 0.2), meta.get("gamma", 0.1), meta.get("temperature", 4.0),
                     loss_scale=65536.0)
    return step_errors(got, ref, metric)


# Gradients of the Linear layers that sit directly behind a ReLU: in reduced precision a few pre-activations change sign and each
# flipped mask element adds or removes a whole term of the gradient sum, so these are the tensors whose bf16 error exceeds 2e-2 --
# in the native kernels AND in the reference's own arithmetic under autocast, by the same amount (profiles/r2_bf16_parity.txt:
# 2.9-4.4e-2 vs 2.9-4.5e-2 at BASELINE config 2).  They are held to the same 1.2x as everything else.
RELU_GATED = ("grad:decoder.output_projection.0.weight", "grad:decoder.output_projection.0.bias",
              "grad:attention_refinement.ffn.0.weight", "grad:attention_refinement.ffn.0.bias",
              "pgrad:feature_projection.0.weight", "pgrad:feature_projection.0.bias")


def compare_step_calibrated(got, ref, ref_err, base_tol=2e-2, factor=1.2, metric="l2", verbose=True):
    """Every quantity within max(base_tol, factor x the (loss-scaled) autocast reference's own error on the same inputs)."""
    rows = step_errors(got, ref, metric)
    bad = []
    for name, e in rows.items():
        lim = max(base_tol, factor * ref_err.get(name, 0.0))
        flag = not e < lim
        if flag:
            bad.append((name, e, lim))
        if verbose or flag:
            print(f"   {name:55s} kernel {e:.3e}   autocast reference {ref_err.get(name, float('nan')):.3e}   limit {lim:.3e}{'   <-- FAIL' if flag else ''}")
    assert not bad, f"{len(bad)} quantities exceed max({base_tol}, {factor} x autocast-reference error): {bad[:4]}"
    return rows
