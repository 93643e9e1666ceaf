"""Pins the validation-path oracle (oracle/kd_oracle.py: bleu1, eval_step) against the REAL reference and writes
tests/golden/validation_case.pt.  Runs only where /root/reference exists (the build container); the committed fixture is the pin
on the GPU box.  Reference: validate_student_model (src/train_student_kd.py:29-86) and compute_bleu_score
(src/distillation_utils.py:398-409).

    python oracle/pin_validation.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import kd_oracle as O  # noqa: E402
from oracle.pin_against_reference import load_reference  # noqa: E402


class IdVocab:
    """itos[i] -> a distinct word per id, like the reference Vocabulary (src/data_loader.py:22-23)."""
    def __init__(self, V):
        self.itos = {i: f"w{i}" for i in range(V)}
        self.itos.update({0: "<PAD>", 1: "<START>", 2: "<END>", 3: "<UNK>"})


def main():
    ref_sm, ref_du = load_reference()
    ok = True
    # (1) compute_bleu_score on random id sequences with PAD / START / END, repeats, empty targets
    g = torch.Generator().manual_seed(11)
    T, B, V = 12, 64, 40
    pred = torch.randint(0, V, (T, B), generator=g)
    tgt = torch.randint(0, V, (T, B), generator=g)
    tgt[:, 0] = 0                                   # empty target set -> 0.0
    tgt[:, 1] = torch.tensor([1, 5, 5, 5, 2, 0, 0, 0, 0, 0, 0, 0])
    pred[:, 1] = torch.tensor([5, 5, 1, 2, 0, 9, 9, 9, 9, 9, 9, 9])
    tgt[6:, 2] = 0
    vocab = IdVocab(V)
    bleu_ref = torch.tensor([ref_du.compute_bleu_score(pred[:, b].numpy(), tgt[:, b].numpy(), vocab) for b in range(B)], dtype=torch.float64)
    bleu_orc = O.bleu1(pred, tgt)
    err = float((bleu_ref - bleu_orc.double()).abs().max())
    print(f"bleu1 oracle vs reference compute_bleu_score: max abs err {err:.2e}")
    ok &= err < 1e-6                                # the oracle returns fp32
    # (2) the validation step on a golden KD case: the reference's loss (already pinned) + logits.argmax(-1) + BLEU of every sample
    case = torch.load(os.path.join(ROOT, "tests", "golden", "kd_small_default.pt"), weights_only=False)
    ref_logits = case["reference"]["logits"]
    pred_ref = ref_logits.argmax(dim=-1)            # train_student_kd.py:74
    tg = case["batch"]["targets"]
    vocab2 = IdVocab(case["meta"]["V"])
    bleu_ref2 = torch.tensor([ref_du.compute_bleu_score(pred_ref[:, b].numpy(), tg[:, b].numpy(), vocab2) for b in range(tg.shape[1])])
    ev = O.eval_step(case["params"], case["proj_params"], case["batch"])
    tok_ok = bool(torch.equal(ev["predicted_tokens"], pred_ref))
    bleu_ok = float((ev["bleu"].double() - bleu_ref2.double()).abs().max()) < 1e-6
    loss_ok = all(abs(ev["loss"][k] - v) <= 2e-6 * max(1.0, abs(v)) for k, v in case["reference"]["loss"].items())
    print(f"eval_step oracle vs reference: tokens equal {tok_ok}, bleu equal {bleu_ok}, loss parts equal {loss_ok}")
    ok &= tok_ok and bleu_ok and loss_ok
    torch.save({"bleu": {"pred": pred, "targets": tgt, "reference": bleu_ref.float()},
                "kd_small_default": {"predicted_tokens": pred_ref, "bleu": bleu_ref2.float()},
                "generator": "oracle/pin_validation.py (reference compute_bleu_score / logits.argmax)"},
               os.path.join(ROOT, "tests", "golden", "validation_case.pt"))
    print("ALL PINNED" if ok else "PIN FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
