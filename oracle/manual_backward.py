"""Dataflow blueprint of the CUDA decoder (forward saves + hand-derived backward) — TEST INFRASTRUCTURE ONLY.

`imagecaptioner_b200/csrc/api.cu` runs exactly this sequence of contractions and pointwise
kernels on the GPU.  Here the same sequence is restated with plain CPU tensor arithmetic and NO
autograd, so that `tests/test_oracle.py` can check the hand-derived backward (BPTT through the
attention-LSTM) against autograd of `oracle/kd_oracle.py` before any GPU
is involved.  Nothing under `imagecaptioner_b200/` imports this file.

Reference math: src/student_model.py:173-256 (forward); the backward is the adjoint of it.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

Tensor = torch.Tensor


def _sig(x):
    return 1.0 / (1.0 + torch.exp(-x))


def decoder_forward_saved(p: Dict[str, Tensor], feats: Tensor, captions: Tensor, prefix: str = "decoder."):
    """Forward with the time-invariant half of the attention projection hoisted out of the loop
    (P = F W_f^T + b_a) and the teacher-forced embedding half of attention_combine time-batched."""
    T, B = captions.shape
    _, S, E = feats.shape
    H = p[prefix + "lstm.weight_hh_l0"].shape[1]
    L = 0
    while (prefix + f"lstm.weight_ih_l{L}") in p:
        L += 1
    Wa, ba = p[prefix + "attention.weight"], p[prefix + "attention.bias"]
    Wh, Wf = Wa[:, :H], Wa[:, H:]
    Wc, bc = p[prefix + "attention_combine.weight"], p[prefix + "attention_combine.bias"]
    Wce, Wcc = Wc[:, :E], Wc[:, E:]
    Wcat = [torch.cat([p[prefix + f"lstm.weight_ih_l{k}"], p[prefix + f"lstm.weight_hh_l{k}"]], dim=1) for k in range(L)]
    bcat = [p[prefix + f"lstm.bias_ih_l{k}"] + p[prefix + f"lstm.bias_hh_l{k}"] for k in range(L)]
    W1, b1 = p[prefix + "output_projection.0.weight"], p[prefix + "output_projection.0.bias"]
    W2, b2 = p[prefix + "output_projection.3.weight"], p[prefix + "output_projection.3.bias"]
    ins = [E] + [H] * (L - 1)

    P = (feats.reshape(B * S, E) @ Wf.t() + ba).reshape(B, S, E)
    emb = p[prefix + "embedding.weight"][captions.reshape(-1)]                 # (T*B,E)
    xh = [feats.new_zeros(T + 1, B, ins[k] + H) for k in range(L)]              # [input ; h_prev]
    xh[0][:T, :, :E] = (emb @ Wce.t() + bc).reshape(T, B, E)
    c = [feats.new_zeros(T + 1, B, H) for _ in range(L)]
    gates = [feats.new_zeros(T, B, 4 * H) for _ in range(L)]
    u = feats.new_zeros(T, B, E)
    ctx = feats.new_zeros(T, B, E)
    w = feats.new_zeros(T, B, S)
    hid = feats.new_zeros(T, B, H)
    for t in range(T):
        q = xh[L - 1][t][:, ins[L - 1]:]
        u[t] = q @ Wh.t()
        s = torch.tanh(P + u[t].unsqueeze(1)).sum(-1)
        s = s - s.max(dim=1, keepdim=True).values
        e = torch.exp(s)
        w[t] = e / e.sum(dim=1, keepdim=True)
        ctx[t] = (w[t].unsqueeze(-1) * feats).sum(1)
        xh[0][t][:, :E] += ctx[t] @ Wcc.t()
        for k in range(L):
            pre = xh[k][t] @ Wcat[k].t() + bcat[k]
            i, f, g, o = _sig(pre[:, :H]), _sig(pre[:, H:2 * H]), torch.tanh(pre[:, 2 * H:3 * H]), _sig(pre[:, 3 * H:])
            gates[k][t] = torch.cat([i, f, g, o], dim=1)
            c[k][t + 1] = f * c[k][t] + i * g
            h = o * torch.tanh(c[k][t + 1])
            xh[k][t + 1][:, ins[k]:] = h
            if k + 1 < L:
                xh[k + 1][t][:, :H] = h
            else:
                hid[t] = h
    o1 = (hid.reshape(T * B, H) @ W1.t() + b1).clamp_min(0)
    logits = (o1 @ W2.t() + b2).reshape(T, B, -1)
    saved = dict(P=P, emb=emb, xh=xh, c=c, gates=gates, u=u, ctx=ctx, w=w, hid=hid, o1=o1,
                 Wh=Wh, Wf=Wf, Wce=Wce, Wcc=Wcc, Wcat=Wcat, W1=W1, W2=W2, ins=ins, L=L, H=H)
    return logits, hid, w, saved


def decoder_backward_manual(p: Dict[str, Tensor], feats: Tensor, captions: Tensor, saved: dict,
                            dlogits: Tensor, dhid: Optional[Tensor], prefix: str = "decoder."):
    """Hand-derived adjoint; returns (grads by state_dict key, dfeats)."""
    T, B = captions.shape
    _, S, E = feats.shape
    sv = saved
    H, L, ins = sv["H"], sv["L"], sv["ins"]
    V = dlogits.shape[-1]
    G: Dict[str, Tensor] = {}
    dl = dlogits.reshape(T * B, V)
    # ---- output head, time-batched
    do1 = dl @ sv["W2"]
    do1 = torch.where(sv["o1"] > 0, do1, torch.zeros_like(do1))
    G[prefix + "output_projection.3.weight"] = dl.t() @ sv["o1"]
    G[prefix + "output_projection.3.bias"] = dl.sum(0)
    dH_ext = (do1 @ sv["W1"]).reshape(T, B, H)
    G[prefix + "output_projection.0.weight"] = do1.t() @ sv["hid"].reshape(T * B, H)
    G[prefix + "output_projection.0.bias"] = do1.sum(0)
    # ---- BPTT
    dgates = [feats.new_zeros(T, B, 4 * H) for _ in range(L)]
    dxh0 = feats.new_zeros(T, B, E + H)
    dxh = [None] + [feats.new_zeros(B, 2 * H) for _ in range(1, L)]
    dc = [feats.new_zeros(B, H) for _ in range(L)]
    dctx = feats.new_zeros(T, B, E)
    ds = feats.new_zeros(T, B, S)
    du = feats.new_zeros(T, B, E)
    dq = None
    for t in range(T - 1, -1, -1):
        for k in range(L - 1, -1, -1):
            dh = feats.new_zeros(B, H)
            if t < T - 1:                                    # recurrent carry from step t+1
                dh = dh + (dxh0[t + 1][:, E:] if k == 0 else dxh[k][:, H:])
            if k == L - 1:
                dh = dh + dH_ext[t]
                if dhid is not None:
                    dh = dh + dhid[t]
                if dq is not None:
                    dh = dh + dq
            else:
                dh = dh + dxh[k + 1][:, :H]                  # input gradient of the layer above, same step
            gt = sv["gates"][k][t]
            i, f, g, o = gt[:, :H], gt[:, H:2 * H], gt[:, 2 * H:3 * H], gt[:, 3 * H:]
            tc = torch.tanh(sv["c"][k][t + 1])
            dcc = dc[k] + dh * o * (1 - tc * tc)
            dgates[k][t] = torch.cat([dcc * g * i * (1 - i), dcc * sv["c"][k][t] * f * (1 - f),
                                      dcc * i * (1 - g * g), dh * tc * o * (1 - o)], dim=1)
            dc[k] = dcc * f
            d = dgates[k][t] @ sv["Wcat"][k]                 # (B, in+H)
            if k == 0:
                dxh0[t] = d
            else:
                dxh[k] = d
        dx = dxh0[t][:, :E]
        dctx[t] = dx @ sv["Wcc"]
        dw = (dctx[t].unsqueeze(1) * feats).sum(-1)          # (B,S)
        wt = sv["w"][t]
        ds[t] = wt * (dw - (wt * dw).sum(1, keepdim=True))
        th = torch.tanh(sv["P"] + sv["u"][t].unsqueeze(1))
        du[t] = (ds[t].unsqueeze(-1) * (1 - th * th)).sum(1)
        dq = du[t] @ sv["Wh"]
    # ---- post-loop, time-batched
    dP = feats.new_zeros(B, S, E)
    dF = feats.new_zeros(B, S, E)
    for t in range(T):
        th = torch.tanh(sv["P"] + sv["u"][t].unsqueeze(1))
        dP += ds[t].unsqueeze(-1) * (1 - th * th)
        dF += sv["w"][t].unsqueeze(-1) * dctx[t].unsqueeze(1)
    for k in range(L):
        dg = dgates[k].reshape(T * B, 4 * H)
        xk = sv["xh"][k][:T].reshape(T * B, -1)
        G[prefix + f"lstm.weight_ih_l{k}"] = dg.t() @ xk[:, :ins[k]]
        G[prefix + f"lstm.weight_hh_l{k}"] = dg.t() @ xk[:, ins[k]:]
        G[prefix + f"lstm.bias_ih_l{k}"] = dg.sum(0)
        G[prefix + f"lstm.bias_hh_l{k}"] = dg.sum(0)
    qall = sv["xh"][L - 1][:T].reshape(T * B, -1)[:, ins[L - 1]:]
    dWh = du.reshape(T * B, E).t() @ qall
    dWf = dP.reshape(B * S, E).t() @ feats.reshape(B * S, E)
    G[prefix + "attention.weight"] = torch.cat([dWh, dWf], dim=1)
    G[prefix + "attention.bias"] = dP.reshape(B * S, E).sum(0)
    dF = dF + (dP.reshape(B * S, E) @ sv["Wf"]).reshape(B, S, E)
    dxa = dxh0[:, :, :E].reshape(T * B, E)
    G[prefix + "attention_combine.weight"] = torch.cat([dxa.t() @ sv["emb"], dxa.t() @ sv["ctx"].reshape(T * B, E)], dim=1)
    G[prefix + "attention_combine.bias"] = dxa.sum(0)
    demb = dxa @ sv["Wce"]
    dE = torch.zeros_like(p[prefix + "embedding.weight"])
    dE.index_add_(0, captions.reshape(-1), demb)
    G[prefix + "embedding.weight"] = dE
    return G, dF


def kd_token_grad(y: Tensor, z: Tensor, targets: Tensor, temperature: float, alpha: float, w_ce: float):
    """Closed-form d(alpha*KD + w_ce*CE)/dlogits — what kernel (3) writes."""
    V = y.shape[-1]
    yy, zz, tt = y.reshape(-1, V), z.reshape(-1, V), targets.reshape(-1)
    N = yy.shape[0]
    pS = torch.softmax(yy / temperature, dim=1)
    pT = torch.softmax(zz / temperature, dim=1)
    valid = (tt != 0).to(y.dtype)
    nv = valid.sum()
    p1 = torch.softmax(yy, dim=1)
    onehot = torch.zeros_like(yy)
    onehot[torch.arange(N), tt] = 1.0
    g = alpha * temperature / N * (pS - pT) + (w_ce / nv) * valid.unsqueeze(1) * (p1 - onehot)
    return g.reshape(y.shape)


# --------------------------------------------------------------------------------------------------------------
# v2 dataflow: attention_combine folded into layer 0's gate contraction (what csrc/api.cu runs since round 1, step 12)
#   pre0_t = emb_t (W_ih0 W_ce)^T + ctx_t (W_ih0 W_cc)^T + h0_{t-1} W_hh0^T + (W_ih0 b_c + b_ih0 + b_hh0)
# so x_t = W_c [emb_t ; ctx_t] + b_c is never materialised, the context GEMM leaves the serial chain, and layer 0's
# [input ; h] operand is [ctx_t ; h0_{t-1}].  The gate-row interleave used on the GPU is a pure relabelling and is not modelled.
# --------------------------------------------------------------------------------------------------------------
def decoder_forward_saved_v2(p: Dict[str, Tensor], feats: Tensor, captions: Tensor, prefix: str = "decoder."):
    T, B = captions.shape
    _, S, E = feats.shape
    H = p[prefix + "lstm.weight_hh_l0"].shape[1]
    L = 0
    while (prefix + f"lstm.weight_ih_l{L}") in p:
        L += 1
    Wa, ba = p[prefix + "attention.weight"], p[prefix + "attention.bias"]
    Wh, Wf = Wa[:, :H], Wa[:, H:]
    Wc, bc = p[prefix + "attention_combine.weight"], p[prefix + "attention_combine.bias"]
    Wce, Wcc = Wc[:, :E], Wc[:, E:]
    Wih0 = p[prefix + "lstm.weight_ih_l0"]
    Wx, We = Wih0 @ Wcc, Wih0 @ Wce                                     # (4H,E) each
    bx = Wih0 @ bc + p[prefix + "lstm.bias_ih_l0"] + p[prefix + "lstm.bias_hh_l0"]
    Wcat = [torch.cat([Wx, p[prefix + "lstm.weight_hh_l0"]], dim=1)]
    Wcat += [torch.cat([p[prefix + f"lstm.weight_ih_l{k}"], p[prefix + f"lstm.weight_hh_l{k}"]], dim=1) for k in range(1, L)]
    bcat = [None] + [p[prefix + f"lstm.bias_ih_l{k}"] + p[prefix + f"lstm.bias_hh_l{k}"] for k in range(1, L)]
    W1, b1 = p[prefix + "output_projection.0.weight"], p[prefix + "output_projection.0.bias"]
    W2, b2 = p[prefix + "output_projection.3.weight"], p[prefix + "output_projection.3.bias"]
    ins = [E] + [H] * (L - 1)
    P = (feats.reshape(B * S, E) @ Wf.t() + ba).reshape(B, S, E)
    emb = p[prefix + "embedding.weight"][captions.reshape(-1)]
    G0 = (emb @ We.t() + bx).reshape(T, B, 4 * H)                        # time-batched addend of layer 0
    xh = [feats.new_zeros(T + 1, B, ins[k] + H) for k in range(L)]       # layer 0: [ctx_t ; h0_{t-1}]
    c = [feats.new_zeros(T + 1, B, H) for _ in range(L)]
    gates = [feats.new_zeros(T, B, 4 * H) for _ in range(L)]
    u = feats.new_zeros(T, B, E); w = feats.new_zeros(T, B, S); hid = feats.new_zeros(T, B, H)
    for t in range(T):
        q = xh[L - 1][t][:, ins[L - 1]:]
        u[t] = q @ Wh.t()
        s = torch.tanh(P + u[t].unsqueeze(1)).sum(-1)
        e = torch.exp(s - s.max(dim=1, keepdim=True).values)
        w[t] = e / e.sum(dim=1, keepdim=True)
        xh[0][t][:, :E] = (w[t].unsqueeze(-1) * feats).sum(1)            # ctx_t goes straight into layer 0's operand
        for k in range(L):
            pre = xh[k][t] @ Wcat[k].t() + (G0[t] if k == 0 else bcat[k])
            i, f, g, o = _sig(pre[:, :H]), _sig(pre[:, H:2 * H]), torch.tanh(pre[:, 2 * H:3 * H]), _sig(pre[:, 3 * H:])
            gates[k][t] = torch.cat([i, f, g, o], dim=1)
            c[k][t + 1] = f * c[k][t] + i * g
            h = o * torch.tanh(c[k][t + 1])
            xh[k][t + 1][:, ins[k]:] = h
            if k + 1 < L:
                xh[k + 1][t][:, :H] = h
            else:
                hid[t] = h
    o1 = (hid.reshape(T * B, H) @ W1.t() + b1).clamp_min(0)
    logits = (o1 @ W2.t() + b2).reshape(T, B, -1)
    saved = dict(P=P, emb=emb, xh=xh, c=c, gates=gates, u=u, w=w, hid=hid, o1=o1, Wh=Wh, Wf=Wf, Wce=Wce, Wcc=Wcc, Wih0=Wih0,
                 Wx=Wx, We=We, Wcat=Wcat, W1=W1, W2=W2, ins=ins, L=L, H=H)
    return logits, hid, w, saved


def decoder_backward_manual_v2(p: Dict[str, Tensor], feats: Tensor, captions: Tensor, saved: dict,
                               dlogits: Tensor, dhid: Optional[Tensor], prefix: str = "decoder."):
    T, B = captions.shape
    _, S, E = feats.shape
    sv = saved
    H, L, ins = sv["H"], sv["L"], sv["ins"]
    V = dlogits.shape[-1]
    G: Dict[str, Tensor] = {}
    dl = dlogits.reshape(T * B, V)
    do1 = dl @ sv["W2"]
    do1 = torch.where(sv["o1"] > 0, do1, torch.zeros_like(do1))
    G[prefix + "output_projection.3.weight"] = dl.t() @ sv["o1"]
    G[prefix + "output_projection.3.bias"] = dl.sum(0)
    dH_ext = (do1 @ sv["W1"]).reshape(T, B, H)
    G[prefix + "output_projection.0.weight"] = do1.t() @ sv["hid"].reshape(T * B, H)
    G[prefix + "output_projection.0.bias"] = do1.sum(0)
    dgates = [feats.new_zeros(T, B, 4 * H) for _ in range(L)]
    dxh0 = feats.new_zeros(T, B, E + H)                                  # [dctx_t ; dh0 carry]
    dxh = [None] + [feats.new_zeros(B, 2 * H) for _ in range(1, L)]
    dc = [feats.new_zeros(B, H) for _ in range(L)]
    ds = feats.new_zeros(T, B, S); du = feats.new_zeros(T, B, E)
    dq = None
    for t in range(T - 1, -1, -1):
        for k in range(L - 1, -1, -1):
            dh = feats.new_zeros(B, H)
            if t < T - 1:
                dh = dh + (dxh0[t + 1][:, E:] if k == 0 else dxh[k][:, H:])
            if k == L - 1:
                dh = dh + dH_ext[t]
                if dhid is not None:
                    dh = dh + dhid[t]
                if dq is not None:
                    dh = dh + dq
            else:
                dh = dh + dxh[k + 1][:, :H]
            gt = sv["gates"][k][t]
            i, f, g, o = gt[:, :H], gt[:, H:2 * H], gt[:, 2 * H:3 * H], gt[:, 3 * H:]
            tc = torch.tanh(sv["c"][k][t + 1])
            dcc = dc[k] + dh * o * (1 - tc * tc)
            dgates[k][t] = torch.cat([dcc * g * i * (1 - i), dcc * sv["c"][k][t] * f * (1 - f),
                                      dcc * i * (1 - g * g), dh * tc * o * (1 - o)], dim=1)
            dc[k] = dcc * f
            d = dgates[k][t] @ sv["Wcat"][k]
            if k == 0:
                dxh0[t] = d
            else:
                dxh[k] = d
        dctx = dxh0[t][:, :E]                                            # no context GEMM any more
        dw = (dctx.unsqueeze(1) * feats).sum(-1)
        wt = sv["w"][t]
        ds[t] = wt * (dw - (wt * dw).sum(1, keepdim=True))
        th = torch.tanh(sv["P"] + sv["u"][t].unsqueeze(1))
        du[t] = (ds[t].unsqueeze(-1) * (1 - th * th)).sum(1)
        dq = du[t] @ sv["Wh"]
    dP = feats.new_zeros(B, S, E); dF = feats.new_zeros(B, S, E)
    for t in range(T):
        th = torch.tanh(sv["P"] + sv["u"][t].unsqueeze(1))
        dP += ds[t].unsqueeze(-1) * (1 - th * th)
        dF += sv["w"][t].unsqueeze(-1) * dxh0[t][:, :E].unsqueeze(1)
    dg0 = dgates[0].reshape(T * B, 4 * H)
    x0 = sv["xh"][0][:T].reshape(T * B, -1)
    dWx = dg0.t() @ x0[:, :E]                                            # (4H,E)
    dWe = dg0.t() @ sv["emb"]                                            # (4H,E)
    dbx = dg0.sum(0)
    G[prefix + "lstm.weight_hh_l0"] = dg0.t() @ x0[:, E:]
    bc = p[prefix + "attention_combine.bias"]
    G[prefix + "lstm.weight_ih_l0"] = dWx @ sv["Wcc"].t() + dWe @ sv["Wce"].t() + torch.outer(dbx, bc)     # via W_x, W_e and b_x
    G[prefix + "lstm.bias_ih_l0"] = dbx
    G[prefix + "lstm.bias_hh_l0"] = dbx
    G[prefix + "attention_combine.weight"] = torch.cat([sv["Wih0"].t() @ dWe, sv["Wih0"].t() @ dWx], dim=1)
    G[prefix + "attention_combine.bias"] = sv["Wih0"].t() @ dbx
    for k in range(1, L):
        dg = dgates[k].reshape(T * B, 4 * H)
        xk = sv["xh"][k][:T].reshape(T * B, -1)
        G[prefix + f"lstm.weight_ih_l{k}"] = dg.t() @ xk[:, :ins[k]]
        G[prefix + f"lstm.weight_hh_l{k}"] = dg.t() @ xk[:, ins[k]:]
        G[prefix + f"lstm.bias_ih_l{k}"] = dg.sum(0)
        G[prefix + f"lstm.bias_hh_l{k}"] = dg.sum(0)
    qall = sv["xh"][L - 1][:T].reshape(T * B, -1)[:, ins[L - 1]:]
    dWh = du.reshape(T * B, E).t() @ qall
    dWf = dP.reshape(B * S, E).t() @ feats.reshape(B * S, E)
    G[prefix + "attention.weight"] = torch.cat([dWh, dWf], dim=1)
    G[prefix + "attention.bias"] = dP.reshape(B * S, E).sum(0)
    dF = dF + (dP.reshape(B * S, E) @ sv["Wf"]).reshape(B, S, E)
    demb = dg0 @ sv["We"]
    dE = torch.zeros_like(p[prefix + "embedding.weight"])
    dE.index_add_(0, captions.reshape(-1), demb)
    G[prefix + "embedding.weight"] = dE
    return G, dF
