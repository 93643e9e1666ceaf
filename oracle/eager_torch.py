"""TEST / BASELINE INFRASTRUCTURE ONLY (never imported by imagecaptioner_b200/).

The KD step written with the STOCK torch.nn building blocks the reference composes (nn.LSTM stepped one token at a time,
nn.MultiheadAttention, nn.Linear, F.kl_div, CrossEntropyLoss, MSELoss, CosineEmbeddingLoss): the path a user of the
reference gets on a GPU today ("existing Blackwell path", SURVEY.md section 2.1 / 8d secondary baseline: cuDNN RNN, cuBLAS,
ATen element-wise kernels).  kd_oracle.py restates the same arithmetic with plain tensor operations; this file exists so that

  * bench.py can time the library path on the same B200 (fp32 and under torch.autocast) next to the native step, and
  * tests can measure what reduced-precision autocast does to the REFERENCE's own arithmetic (per-tensor error against
    fp32), which is the yardstick for the bf16 tolerance of the native kernels.

It is pinned on the CPU against kd_oracle (tests/test_oracle.py), which is itself pinned against the real reference modules
(oracle/pin_against_reference.py).  Module / attribute names follow the reference's state_dict keys, the only contract here:
  decoder.{embedding, attention, attention_combine, lstm, output_projection.{0,3}}     src/student_model.py:125-156
  attention_refinement.{attention, ffn.{0,3}, norm1, norm2}                            src/student_model.py:76-101
  feature_projection.{0,3}                                                             src/distillation_utils.py:213-231
Every dropout probability is 0 (the arithmetic of eval mode, like the oracle); the modules stay in train() for cuDNN.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

Tensor = torch.Tensor


class _Decoder(nn.Module):
    def __init__(self, V, E, H, L):
        super().__init__()
        self.H = H
        self.embedding = nn.Embedding(V, E)
        self.attention = nn.Linear(H + E, E)
        self.attention_combine = nn.Linear(2 * E, E)
        self.lstm = nn.LSTM(E, H, num_layers=L, batch_first=True)
        self.output_projection = nn.Sequential(nn.Linear(H, E), nn.ReLU(), nn.Dropout(0.0), nn.Linear(E, V))

    def attend(self, h_top, feats):                                   # src/student_model.py:173-203 (concat form, as shipped)
        B, S, _ = feats.shape
        joined = torch.cat([h_top[:, None, :].expand(B, S, self.H), feats], dim=2)
        scores = torch.tanh(self.attention(joined)).sum(dim=2)
        w = F.softmax(scores, dim=1)
        return torch.bmm(w[:, None, :], feats)[:, 0], w

    def forward(self, feats, captions):                               # src/student_model.py:205-256
        T, B = captions.shape
        emb = self.embedding(captions)                                # (T,B,E)
        L = self.lstm.num_layers
        state = (feats.new_zeros(L, B, self.H), feats.new_zeros(L, B, self.H))
        ys, hs, ws = [], [], []
        for t in range(T):
            ctx, w = self.attend(state[0][-1], feats)
            x = self.attention_combine(torch.cat([emb[t], ctx], dim=1))
            out, state = self.lstm(x[:, None, :], state)
            ys.append(self.output_projection(out[:, 0]))
            hs.append(state[0][-1])
            ws.append(w)
        return torch.stack(ys), hs, ws


class _Refinement(nn.Module):
    def __init__(self, E, heads=4):
        super().__init__()
        self.attention = nn.MultiheadAttention(E, heads, dropout=0.0, batch_first=True)
        self.ffn = nn.Sequential(nn.Linear(E, 2 * E), nn.ReLU(), nn.Dropout(0.0), nn.Linear(2 * E, E))
        self.norm1, self.norm2 = nn.LayerNorm(E), nn.LayerNorm(E)

    def forward(self, x):                                             # src/student_model.py:103-118
        x = self.norm1(x + self.attention(x, x, x, need_weights=False)[0])
        return self.norm2(x + self.ffn(x))


class EagerStudent(nn.Module):
    """decoder (+ refinement) fed with encoder features; forward -> (logits, un-refined features, hiddens, attention)."""

    def __init__(self, V, E, H, L, refinement=True):
        super().__init__()
        self.decoder = _Decoder(V, E, H, L)
        self.attention_refinement = _Refinement(E) if refinement else None

    def forward(self, feats, captions):                               # src/student_model.py:288-312
        refined = self.attention_refinement(feats) if self.attention_refinement is not None else feats
        y, hs, ws = self.decoder(refined, captions)
        return y, feats, hs, ws


class EagerProjector(nn.Module):
    def __init__(self, Et, Es, out_tokens):
        super().__init__()
        self.feature_projection = (nn.Sequential(nn.Linear(Et, Es), nn.ReLU(), nn.Dropout(0.0), nn.LayerNorm(Es))
                                   if Et != Es else nn.Identity())
        self.pool = nn.AdaptiveAvgPool1d(out_tokens)

    def forward(self, x):                                             # src/distillation_utils.py:233-252
        return self.pool(self.feature_projection(x).transpose(1, 2)).transpose(1, 2)


def eager_loss(s_out, t_out, targets, alpha=0.7, beta=0.2, gamma=0.1, temperature=4.0):
    """DistillationLoss.forward with the library calls the reference makes (src/distillation_utils.py:30-54, :56-94, :96-136, :138-200)."""
    y, z = s_out["logits"], t_out["logits"]
    V = y.shape[-1]
    ce = F.cross_entropy(y.reshape(-1, V), targets.reshape(-1), ignore_index=0)
    kd = F.kl_div(F.log_softmax(y.reshape(-1, V) / temperature, dim=-1), F.softmax(z.reshape(-1, V) / temperature, dim=-1),
                  reduction="batchmean") * temperature ** 2

    def pooled(f):
        return (F.softmax(f.sum(dim=2), dim=1)[:, :, None] * f).sum(dim=1)
    fs, ft = s_out["encoder_features"], t_out["encoder_features"]
    feat = 0.6 * F.mse_loss(fs.mean(dim=1), ft.mean(dim=1)) + 0.4 * F.mse_loss(pooled(fs), pooled(ft))
    hid = y.new_zeros(())
    hs, ht = s_out.get("hidden_states"), t_out.get("hidden_states")
    if hs is not None and ht is not None:
        n = min(len(hs), len(ht))
        one = torch.ones(hs[0].shape[0], device=hs[0].device)
        hid = torch.stack([0.7 * F.mse_loss(hs[t], ht[t]) + 0.3 * F.cosine_embedding_loss(hs[t], ht[t], one) for t in range(n)]).mean()
    total = (1 - alpha - beta - gamma) * ce + alpha * kd + beta * feat + gamma * hid
    return total, {"total_loss": total, "ce_loss": ce, "token_kd_loss": kd, "feature_kd_loss": feat, "hidden_kd_loss": hid}


def build(params: Dict[str, Tensor], proj_params: Dict[str, Tensor], V, E, H, L, refinement, Et, S=49, device="cpu"):
    model = EagerStudent(V, E, H, L, refinement)
    missing, unexpected = model.load_state_dict({k: v.float() for k, v in params.items()}, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    proj = EagerProjector(Et, E, S)
    if proj_params:
        proj.load_state_dict({k: v.float() for k, v in proj_params.items()})
    # train(): cuDNN's RNN backward refuses eval mode; every Dropout above has p = 0, so the arithmetic is the eval-mode one
    return model.to(device).train(), proj.to(device).train()


def _forward_loss(model, proj, batch, autocast_dtype, feats, alpha=0.7, beta=0.2, gamma=0.1, temperature=4.0):
    dev = feats.device
    on = autocast_dtype is not None
    with torch.autocast(dev.type, dtype=autocast_dtype if on else torch.bfloat16, enabled=on):
        y, enc, hs, ws = model(feats, batch["captions_input"].to(dev))
        tproj = proj(batch["teacher_features"].to(dev))
        th = batch.get("teacher_hiddens")
        t_out = {"logits": batch["teacher_logits"].to(dev), "encoder_features": tproj,
                 "hidden_states": None if th is None else [th[t].to(dev) for t in range(th.shape[0])]}
        total, parts = eager_loss({"logits": y, "encoder_features": enc, "hidden_states": hs}, t_out, batch["targets"].to(dev),
                                  alpha, beta, gamma, temperature)
    return total, parts, y, hs, ws, tproj


def kd_loss(model, proj, batch, autocast_dtype: Optional[torch.dtype] = None):
    """forward + projector + loss of one batch (src/train_student_kd.py:271-285) -> the loss tensor (the caller runs backward)."""
    dev = next(model.parameters()).device
    feats = batch["encoder_features"].to(dev).clone().requires_grad_(True)
    return _forward_loss(model, proj, batch, autocast_dtype, feats)[0]


def kd_step(model, proj, batch, autocast_dtype: Optional[torch.dtype] = None, alpha=0.7, beta=0.2, gamma=0.1, temperature=4.0,
            loss_scale: float = 1.0):
    """The reference loop body (src/train_student_kd.py:262-288): forward + projector + loss under autocast, then backward.
    Returns what kd_oracle.kd_step returns (fp32 tensors on the model's device).
    `loss_scale`: the loss is multiplied by it before backward and the gradients divided afterwards, like the reference's
    GradScaler (:239, :288-291).  Without it the stock reduced-precision path LOSES most of the recurrent gradients at batch 512
    (per-element gradients of a mean loss are ~1e-8: measured 75-96 % relative error under bf16 autocast, profiles/r2_bf16_parity.txt)."""
    dev = next(model.parameters()).device
    model.zero_grad(set_to_none=True)
    proj.zero_grad(set_to_none=True)
    feats = batch["encoder_features"].to(dev).clone().requires_grad_(True)
    total, parts, y, hs, ws, tproj = _forward_loss(model, proj, batch, autocast_dtype, feats, alpha, beta, gamma, temperature)
    (total * loss_scale).backward()
    f32 = lambda t: t.detach().float() / loss_scale
    return {
        "loss": {k: float(v.detach()) for k, v in parts.items()},
        "grads": {k: f32(v.grad) for k, v in model.named_parameters() if v.grad is not None},
        "proj_grads": {k: f32(v.grad) for k, v in proj.named_parameters() if v.grad is not None},
        "d_encoder_features": f32(feats.grad), "logits": y.detach().float(), "hidden_states": torch.stack(list(hs)).detach().float(),
        "attention_weights": torch.stack(list(ws)).detach().float(), "teacher_projected": tproj.detach().float(),
    }
