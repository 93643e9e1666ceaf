"""CPU oracle for the KD hot path — TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement (plain tensor arithmetic on the CPU, no
nn.Module, no nn.LSTM, no F.cross_entropy / F.kl_div) of the algorithm that the
reference implements on its hot path.  It exists so that the CUDA path can be
checked against it; nothing under ``imagecaptioner_b200/`` may import it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / CPU baseline.

Parity pin: ``oracle/pin_against_reference.py`` imports the real reference
modules from ``/root/reference/src`` (possible only in the build container) and
checks every function below against them on seeded inputs; it also writes the
golden vectors in ``tests/golden/``.  The reference's own tests pin no values
(SURVEY.md §8c), so those generated vectors are the pin.

Parameter dictionaries use the reference ``state_dict`` key names
(``decoder.lstm.weight_ih_l0`` ...), so a reference checkpoint is directly an
oracle parameter set.

Reference citations are ``file:line`` into ``/root/reference``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor
PAD, START, END, UNK = 0, 1, 2, 3  # src/data_loader.py:22-23


# --------------------------------------------------------------------------
# small numerics helpers (written out so nothing hides in a library call)
# --------------------------------------------------------------------------
def _f(x: Tensor) -> Tensor:
    """Losses run in >= fp32 (autocast promotes them in the reference)."""
    return x if x.dtype in (torch.float32, torch.float64) else x.float()


def _sigmoid(x: Tensor) -> Tensor:
    return 1.0 / (1.0 + torch.exp(-x))


def _softmax(x: Tensor, dim: int) -> Tensor:
    m = x.max(dim=dim, keepdim=True).values
    e = torch.exp(x - m)
    return e / e.sum(dim=dim, keepdim=True)


def _log_softmax(x: Tensor, dim: int) -> Tensor:
    m = x.max(dim=dim, keepdim=True).values
    z = x - m
    return z - torch.log(torch.exp(z).sum(dim=dim, keepdim=True))


def _layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    y = x @ w.t()
    return y if b is None else y + b


# --------------------------------------------------------------------------
# AttentionRefinement  (src/student_model.py:72-118)  — eval mode (dropout off)
# --------------------------------------------------------------------------
def refinement_forward(p: Dict[str, Tensor], x: Tensor, num_heads: int = 4,
                       prefix: str = "attention_refinement.") -> Tensor:
    """Post-norm block: x = LN(x + MHA(x)); x = LN(x + FFN(x)).  (B,S,E)->(B,S,E).

    nn.MultiheadAttention(batch_first=True) packs q/k/v as in_proj_weight (3E,E)
    (src/student_model.py:83-88); heads split the embedding dim contiguously,
    scores scaled by 1/sqrt(head_dim).
    """
    B, S, E = x.shape
    hd = E // num_heads
    w_in, b_in = p[prefix + "attention.in_proj_weight"], p[prefix + "attention.in_proj_bias"]
    qkv = _linear(x, w_in, b_in)                       # (B,S,3E)
    q, k, v = qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:]

    def heads(t):
        return t.reshape(B, S, num_heads, hd).permute(0, 2, 1, 3)   # (B,h,S,hd)

    q, k, v = heads(q), heads(k), heads(v)
    att = _softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (att @ v).permute(0, 2, 1, 3).reshape(B, S, E)
    o = _linear(o, p[prefix + "attention.out_proj.weight"], p[prefix + "attention.out_proj.bias"])
    x = _layer_norm(x + o, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"])       # :112
    f = _linear(x, p[prefix + "ffn.0.weight"], p[prefix + "ffn.0.bias"]).clamp_min(0)    # :91-96
    f = _linear(f, p[prefix + "ffn.3.weight"], p[prefix + "ffn.3.bias"])
    return _layer_norm(x + f, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"])     # :116


# --------------------------------------------------------------------------
# LSTMDecoder  (src/student_model.py:121-256)
# --------------------------------------------------------------------------
def attention_step(q: Tensor, feats: Tensor, w_a: Tensor, b_a: Tensor) -> Tuple[Tensor, Tensor]:
    """Spatial attention of one decode step (src/student_model.py:173-203).

    a_l = W_a [q ; F_l] + b_a (hidden FIRST, :189), s_l = sum_e tanh(a_le),
    w = softmax_l(s), ctx = sum_l w_l F_l.   q (B,H), feats (B,S,E).
    """
    H = q.shape[1]
    a = (q @ w_a[:, :H].t()).unsqueeze(1) + feats @ w_a[:, H:].t() + b_a   # (B,S,E)
    s = torch.tanh(a).sum(dim=2)                                           # (B,S)
    w = _softmax(s, dim=1)
    ctx = (w.unsqueeze(2) * feats).sum(dim=1)                              # (B,E)
    return ctx, w


def lstm_cell(x: Tensor, h: Tensor, c: Tensor, w_ih: Tensor, w_hh: Tensor,
              b_ih: Tensor, b_hh: Tensor) -> Tuple[Tensor, Tensor]:
    """One nn.LSTM layer step; gate order i,f,g,o; both biases (src/student_model.py:142-148,244)."""
    Hn = h.shape[1]
    g = x @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
    i, f, gg, o = g[:, :Hn], g[:, Hn:2 * Hn], g[:, 2 * Hn:3 * Hn], g[:, 3 * Hn:]
    c2 = _sigmoid(f) * c + _sigmoid(i) * torch.tanh(gg)
    h2 = _sigmoid(o) * torch.tanh(c2)
    return h2, c2


def num_lstm_layers(p: Dict[str, Tensor], prefix: str = "decoder.") -> int:
    n = 0
    while (prefix + f"lstm.weight_ih_l{n}") in p:
        n += 1
    return n


def decoder_step(p: Dict[str, Tensor], feats: Tensor, emb_t: Tensor,
                 h: List[Tensor], c: List[Tensor], prefix: str = "decoder."):
    """One full decode step (src/student_model.py:232-251 / :348-363), eval mode.

    Returns logits (B,V), attention weights (B,S); updates h, c lists in place.
    """
    L = len(h)
    ctx, w = attention_step(h[L - 1], feats, p[prefix + "attention.weight"], p[prefix + "attention.bias"])
    x = _linear(torch.cat([emb_t, ctx], dim=1),                           # embedding FIRST (:240)
                p[prefix + "attention_combine.weight"], p[prefix + "attention_combine.bias"])
    inp = x
    for k in range(L):
        h[k], c[k] = lstm_cell(inp, h[k], c[k],
                               p[prefix + f"lstm.weight_ih_l{k}"], p[prefix + f"lstm.weight_hh_l{k}"],
                               p[prefix + f"lstm.bias_ih_l{k}"], p[prefix + f"lstm.bias_hh_l{k}"])
        inp = h[k]
    y = _linear(h[L - 1], p[prefix + "output_projection.0.weight"], p[prefix + "output_projection.0.bias"])
    y = _linear(y.clamp_min(0), p[prefix + "output_projection.3.weight"], p[prefix + "output_projection.3.bias"])
    return y, w


def decoder_forward(p: Dict[str, Tensor], feats: Tensor, captions: Tensor, prefix: str = "decoder.", hidden=None):
    """LSTMDecoder.forward (src/student_model.py:205-256), eval mode.

    feats (B,S,E) refined features, captions (T,B) int64, hidden = None (zero state, :219-220) or (h0, c0) each (L,B,H) ->
    outputs (T,B,V), hidden_states list of T (B,H) (top layer), attention list of T (B,S).
    """
    T, B = captions.shape
    L = num_lstm_layers(p, prefix)
    H = p[prefix + "lstm.weight_hh_l0"].shape[1]
    emb = p[prefix + "embedding.weight"][captions]                         # (T,B,E)  :224
    if hidden is None:
        h = [feats.new_zeros(B, H) for _ in range(L)]                      # :167-171
        c = [feats.new_zeros(B, H) for _ in range(L)]
    else:                                                                   # :205: the caller's (h0, c0) seeds nn.LSTM's state (:243)
        h = [hidden[0][k].to(feats.dtype) for k in range(L)]
        c = [hidden[1][k].to(feats.dtype) for k in range(L)]
    outs, hids, atts = [], [], []
    for t in range(T):
        y, w = decoder_step(p, feats, emb[t], h, c, prefix)
        outs.append(y)
        hids.append(h[L - 1])
        atts.append(w)
    return torch.stack(outs, dim=0), hids, atts


def student_forward(p: Dict[str, Tensor], encoder_features: Tensor, captions: Tensor,
                    use_refinement: bool = True):
    """CaptioningStudent.forward with the encoder outside the path (src/student_model.py:288-312).

    Returns the UN-refined encoder features for KD but decodes from the refined ones (:301-312).
    """
    refined = refinement_forward(p, encoder_features) if use_refinement else encoder_features
    outputs, hids, atts = decoder_forward(p, refined, captions)
    return outputs, encoder_features, hids, atts


def greedy_decode(p: Dict[str, Tensor], feats: Tensor, max_len: int,
                  start_id: int = START, end_id: int = END, prefix: str = "decoder."):
    """Batched restatement of caption_image's loop (src/student_model.py:339-381).

    feats are the (already refined) features (B,S,E).  Returns tokens (max_len,B) int64,
    lengths (B,) = number of words emitted before <END> (caption_image's result length),
    and the per-step top1-top2 logit margin (max_len,B) so a fixture can prove it is
    well conditioned.  Samples keep stepping after <END> (tokens past `lengths` are
    don't-care), exactly like running caption_image per sample and truncating.
    """
    B = feats.shape[0]
    L = num_lstm_layers(p, prefix)
    H = p[prefix + "lstm.weight_hh_l0"].shape[1]
    h = [feats.new_zeros(B, H) for _ in range(L)]
    c = [feats.new_zeros(B, H) for _ in range(L)]
    tok = torch.full((B,), start_id, dtype=torch.long)
    toks = torch.zeros(max_len, B, dtype=torch.long)
    margins = torch.zeros(max_len, B, dtype=feats.dtype)
    lengths = torch.full((B,), max_len, dtype=torch.long)
    done = torch.zeros(B, dtype=torch.bool)
    for t in range(max_len):
        y, _ = decoder_step(p, feats, p[prefix + "embedding.weight"][tok], h, c, prefix)
        top2 = y.topk(2, dim=1).values
        margins[t] = top2[:, 0] - top2[:, 1]
        tok = y.argmax(dim=1)                                              # :369
        toks[t] = tok
        newly = (tok == end_id) & ~done                                    # :372
        lengths[newly] = t
        done |= newly
    return toks, lengths, margins


# --------------------------------------------------------------------------
# FeatureProjector  (src/distillation_utils.py:203-252) — eval mode
# --------------------------------------------------------------------------
def adaptive_avg_pool_tokens(x: Tensor, out_len: int) -> Tensor:
    """AdaptiveAvgPool1d over the token axis of (B,L,E): window i = [floor(i*L/o), ceil((i+1)*L/o))."""
    B, L, E = x.shape
    rows = []
    for i in range(out_len):
        lo = (i * L) // out_len
        hi = -((-(i + 1) * L) // out_len)
        rows.append(x[:, lo:hi, :].mean(dim=1))
    return torch.stack(rows, dim=1)


def feature_projector(p: Dict[str, Tensor], x: Tensor, student_seq_len: int, prefix: str = "") -> Tensor:
    """Linear -> ReLU -> (Dropout) -> LayerNorm, then token pooling; identity projection if no weights."""
    if (prefix + "feature_projection.0.weight") in p:
        x = _linear(x, p[prefix + "feature_projection.0.weight"], p[prefix + "feature_projection.0.bias"]).clamp_min(0)
        x = _layer_norm(x, p[prefix + "feature_projection.3.weight"], p[prefix + "feature_projection.3.bias"])
    if x.shape[1] != student_seq_len:
        x = adaptive_avg_pool_tokens(x, student_seq_len)
    return x


# --------------------------------------------------------------------------
# DistillationLoss  (src/distillation_utils.py:8-200)
# --------------------------------------------------------------------------
def cross_entropy_ignore_pad(logits: Tensor, targets: Tensor) -> Tensor:
    """CrossEntropyLoss(ignore_index=0): mean over non-PAD rows (src/distillation_utils.py:22,154)."""
    V = logits.shape[-1]
    y = _f(logits.reshape(-1, V))
    t = targets.reshape(-1)
    lse = torch.logsumexp(y, dim=1)
    picked = y.gather(1, t.unsqueeze(1)).squeeze(1)
    valid = (t != PAD)
    n = valid.sum()
    return ((lse - picked) * valid).sum() / n


def token_kd(student_logits: Tensor, teacher_logits: Tensor, temperature: float) -> Tensor:
    """T^2 * KL(softmax(z/T) || softmax(y/T)), 'batchmean' over ALL N rows (src/distillation_utils.py:30-54)."""
    V = student_logits.shape[-1]
    y = _f(student_logits.reshape(-1, V)) / temperature
    z = _f(teacher_logits.reshape(-1, V)) / temperature
    log_ps = _log_softmax(y, dim=1)
    log_pt = _log_softmax(z, dim=1)
    pt = torch.exp(log_pt)
    kl = torch.where(pt > 0, pt * (log_pt - log_ps), torch.zeros_like(pt))   # xlogy convention
    return kl.sum() / y.shape[0] * (temperature ** 2)


def feature_kd(s: Tensor, t: Tensor) -> Tensor:
    """0.6*MSE(mean_l) + 0.4*MSE(softmax-pooled) (src/distillation_utils.py:56-94)."""
    if s.shape[-1] != t.shape[-1]:
        raise ValueError(f"Feature dimensions don't match: student {s.shape[-1]}, teacher {t.shape[-1]}")
    s, t = _f(s), _f(t)
    g = ((s.mean(dim=1) - t.mean(dim=1)) ** 2).mean()
    sa = _softmax(s.sum(dim=-1), dim=1)
    ta = _softmax(t.sum(dim=-1), dim=1)
    a = (((s * sa.unsqueeze(-1)).sum(dim=1) - (t * ta.unsqueeze(-1)).sum(dim=1)) ** 2).mean()
    return 0.6 * g + 0.4 * a


def hidden_kd(s_h: Optional[Sequence[Tensor]], t_h: Optional[Sequence[Tensor]]) -> Tensor:
    """mean_t[0.7*MSE + 0.3*mean_b(1-cos)], cos eps 1e-12 inside the sqrt (src/distillation_utils.py:96-136)."""
    if s_h is None or t_h is None:
        return torch.tensor(0.0)
    n = min(len(s_h), len(t_h))
    per_t = []
    for s, t in zip(s_h[:n], t_h[:n]):
        if s.shape[-1] != t.shape[-1]:
            raise ValueError(f"Hidden dimensions don't match: student {s.shape[-1]}, teacher {t.shape[-1]}")
        s, t = _f(s), _f(t)
        mse = ((s - t) ** 2).mean()
        eps = 1e-12                                    # ATen cosine_embedding_loss EPSILON
        cos = (s * t).sum(1) / torch.sqrt(((s * s).sum(1) + eps) * ((t * t).sum(1) + eps))
        per_t.append(0.7 * mse + 0.3 * (1.0 - cos).mean())
    return torch.stack(per_t).mean()


def distillation_loss(student_outputs: dict, teacher_outputs: dict, targets: Tensor,
                      alpha: float = 0.7, beta: float = 0.2, gamma: float = 0.1,
                      temperature: float = 4.0):
    """DistillationLoss.forward (src/distillation_utils.py:138-200) -> (total, dict of 0-dim tensors)."""
    y, z = student_outputs["logits"], teacher_outputs["logits"]
    ce = cross_entropy_ignore_pad(y, targets)
    kd = token_kd(y, z, temperature)
    feat = torch.tensor(0.0)
    if "encoder_features" in student_outputs and "encoder_features" in teacher_outputs:
        feat = feature_kd(student_outputs["encoder_features"], teacher_outputs["encoder_features"])
    hid = torch.tensor(0.0)
    if "hidden_states" in student_outputs and "hidden_states" in teacher_outputs:
        hid = hidden_kd(student_outputs["hidden_states"], teacher_outputs["hidden_states"])
    total = (1 - alpha - beta - gamma) * ce + alpha * kd + beta * feat + gamma * hid     # :184-189
    parts = {"total_loss": total, "ce_loss": ce, "token_kd_loss": kd,
             "feature_kd_loss": feat, "hidden_kd_loss": hid}
    return total, parts


# --------------------------------------------------------------------------
# One whole KD step (the unit of work of BASELINE.json's metric): fwd + loss + bwd
# --------------------------------------------------------------------------
def kd_step(params: Dict[str, Tensor], proj_params: Dict[str, Tensor], batch: dict,
            alpha=0.7, beta=0.2, gamma=0.1, temperature=4.0, use_refinement=True,
            dtype=torch.float32):
    """Replays src/train_student_kd.py:262-288 with the encoders outside the path.

    batch: encoder_features (B,S,E), captions_input (T,B), targets (T,B),
           teacher_logits (T,B,V), teacher_features (B,St,Et), teacher_hiddens (T,B,H) or None.
    Returns dict(loss parts as floats), grads for every parameter / projector parameter and
    for encoder_features, plus the forward outputs.
    """
    P = {k: v.detach().to(dtype).requires_grad_(v.is_floating_point()) for k, v in params.items()}
    Q = {k: v.detach().to(dtype).requires_grad_(True) for k, v in proj_params.items()}
    feats = batch["encoder_features"].detach().to(dtype).requires_grad_(True)
    outputs, enc, hids, atts = student_forward(P, feats, batch["captions_input"], use_refinement)
    S = feats.shape[1]
    tproj = feature_projector(Q, batch["teacher_features"].to(dtype), S)
    th = batch.get("teacher_hiddens")
    t_out = {"logits": batch["teacher_logits"].to(dtype), "encoder_features": tproj,
             "hidden_states": None if th is None else [th[t].to(dtype) for t in range(th.shape[0])]}
    s_out = {"logits": outputs, "encoder_features": enc, "hidden_states": hids}
    total, parts = distillation_loss(s_out, t_out, batch["targets"], alpha, beta, gamma, temperature)
    total.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in P.items() if v.requires_grad}
    pgrads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in Q.items()}
    return {
        "loss": {k: float(v.detach()) for k, v in parts.items()},
        "grads": grads, "proj_grads": pgrads, "d_encoder_features": feats.grad,
        "logits": outputs.detach(), "hidden_states": torch.stack([h.detach() for h in hids]),
        "attention_weights": torch.stack([a.detach() for a in atts]),
        "teacher_projected": tproj.detach(),
    }


# --------------------------------------------------------------------------
# Synthetic data / parameter initialisation shared by tests, smoke and bench
# --------------------------------------------------------------------------
def init_student_params(V: int, E: int = 256, H: int = 512, L: int = 2, refinement: bool = True,
                        seed: int = 0, logit_scale: float = 1.0) -> Dict[str, Tensor]:
    """Random parameters with the reference's initialisers' distributions
    (src/student_model.py:135,159-165; nn.Linear / nn.MultiheadAttention defaults).
    `logit_scale` multiplies output_projection.3.weight so greedy-decode fixtures have
    realistic argmax margins."""
    g = torch.Generator().manual_seed(seed)

    def U(shape, a):
        return (torch.rand(shape, generator=g) * 2 - 1) * a

    def lin(prefix, out_f, in_f, p):
        a = 1.0 / math.sqrt(in_f)
        p[prefix + ".weight"] = U((out_f, in_f), a)
        p[prefix + ".bias"] = U((out_f,), a)

    p: Dict[str, Tensor] = {}
    p["decoder.embedding.weight"] = U((V, E), 0.1)
    lin("decoder.attention", E, H + E, p)
    lin("decoder.attention_combine", E, 2 * E, p)
    for k in range(L):
        in_k = E if k == 0 else H
        a = math.sqrt(6.0 / (4 * H + in_k))                       # xavier_uniform
        p[f"decoder.lstm.weight_ih_l{k}"] = U((4 * H, in_k), a)
        q, _ = torch.linalg.qr(torch.randn(4 * H, H, generator=g))  # orthogonal columns
        p[f"decoder.lstm.weight_hh_l{k}"] = q.contiguous()
        p[f"decoder.lstm.bias_ih_l{k}"] = torch.zeros(4 * H)
        p[f"decoder.lstm.bias_hh_l{k}"] = torch.zeros(4 * H)
    lin("decoder.output_projection.0", E, H, p)
    lin("decoder.output_projection.3", V, E, p)
    p["decoder.output_projection.3.weight"] *= logit_scale
    if refinement:
        a = math.sqrt(6.0 / (3 * E + E))
        p["attention_refinement.attention.in_proj_weight"] = U((3 * E, E), a)
        p["attention_refinement.attention.in_proj_bias"] = torch.zeros(3 * E)
        lin("attention_refinement.attention.out_proj", E, E, p)
        p["attention_refinement.attention.out_proj.bias"].zero_()
        lin("attention_refinement.ffn.0", 2 * E, E, p)
        lin("attention_refinement.ffn.3", E, 2 * E, p)
        for n in ("norm1", "norm2"):
            p[f"attention_refinement.{n}.weight"] = torch.ones(E)
            p[f"attention_refinement.{n}.bias"] = torch.zeros(E)
    return p


def init_projector_params(Et: int, Es: int, seed: int = 1) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    if Et == Es:
        return {}
    a = 1.0 / math.sqrt(Et)
    return {
        "feature_projection.0.weight": (torch.rand((Es, Et), generator=g) * 2 - 1) * a,
        "feature_projection.0.bias": (torch.rand((Es,), generator=g) * 2 - 1) * a,
        "feature_projection.3.weight": torch.ones(Es),
        "feature_projection.3.bias": torch.zeros(Es),
    }


def bleu1(predicted: Tensor, targets: Tensor) -> Tensor:
    """compute_bleu_score (src/distillation_utils.py:398-409) for every column of (T,B) token matrices, on token ids:
    |set(pred) & set(target)| / |set(target)| with PAD / START / END (0, 1, 2) removed; 0 when the target set is empty."""
    out = []
    for b in range(predicted.shape[1]):
        ps = {int(t) for t in predicted[:, b].tolist() if int(t) not in (PAD, START, END)}
        ts = {int(t) for t in targets[:, b].tolist() if int(t) not in (PAD, START, END)}
        out.append(len(ps & ts) / len(ts) if ts else 0.0)
    return torch.tensor(out, dtype=torch.float32)


def eval_step(params: Dict[str, Tensor], proj_params: Dict[str, Tensor], batch: dict, alpha=0.7, beta=0.2, gamma=0.1, temperature=4.0) -> dict:
    """One batch of validate_student_model (src/train_student_kd.py:43-80): forward, the distillation loss without gradients,
    teacher-forced predictions logits.argmax(-1) (:74) and the BLEU-1 of EVERY sample (the reference scores the first two)."""
    got = kd_step(params, proj_params, batch, alpha, beta, gamma, temperature)      # the gradients it also returns are ignored here
    pred = got["logits"].argmax(dim=-1)
    return {"loss": got["loss"], "predicted_tokens": pred, "bleu": bleu1(pred, batch["targets"])}


def synthetic_batch(B: int, T: int, V: int, E: int = 256, H: int = 512, S: int = 49,
                    St: int = 197, Et: int = 384, seed: int = 1234, teacher_hiddens: bool = True) -> dict:
    """Synthetic KD batch of SURVEY.md §8d: N(0,1) encoder features, START-first captions,
    ~15 % PAD suffix in the targets with END as the last real token, N(0,2^2) fp32 teacher logits,
    N(0,1) teacher ViT features and (synthetic) teacher hidden states."""
    g = torch.Generator().manual_seed(seed)
    caps = torch.randint(4, V, (T + 1, B), generator=g)
    caps[0] = START
    n_pad = torch.randint(0, max(1, int(0.3 * T) + 1), (B,), generator=g)     # mean ~15 % of T
    for b in range(B):
        k = int(n_pad[b])
        last = T - k                                  # index (in the T+1 long caption) of <END>
        caps[last, b] = END
        if k > 0:
            caps[last + 1:, b] = PAD
    batch = {
        "encoder_features": torch.randn(B, S, E, generator=g),
        "captions_input": caps[:-1].contiguous(),
        "targets": caps[1:].contiguous(),
        "teacher_logits": torch.randn(T, B, V, generator=g) * 2.0,
        "teacher_features": torch.randn(B, St, Et, generator=g),
        "teacher_hiddens": torch.randn(T, B, H, generator=g) if teacher_hiddens else None,
    }
    return batch
