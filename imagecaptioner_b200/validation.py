"""validate_student_model (reference src/train_student_kd.py:29-86) on the native path.

Same signature and return value as the reference function.  Per batch: teacher forward (stock module, out of the path),
student forward through the C ABI, ``DistillationLoss.evaluate`` (the loss WITHOUT gradients and the teacher-forced
``logits.argmax(-1)`` from one pass over the logits, ``b2c_kd_token_eval``), and the set-overlap BLEU-1 of every sample of the
batch on the device (``b2c_bleu1``).  The reference scores only the first two samples of the first five batches on the host
(:72-80); that is the default here too so the returned number is the same metric, ``bleu_all_samples=True`` averages all of them.
"""
from __future__ import annotations

import torch

from . import _ops
from .distillation_utils import TeacherWrapper


@torch.no_grad()
def validate_student_model(student_model, teacher_model, data_loader, distill_loss, projectors, device, vocab=None, max_batches=50,
                           bleu_all_samples=False, fused=True):
    """-> (average loss per sample, average BLEU-1).  `vocab` is accepted for signature compatibility: the metric is computed
    on token ids (the vocabulary maps ids to words one-to-one)."""
    was_training = student_model.training
    student_model.eval()
    teacher_wrapper = teacher_model if isinstance(teacher_model, TeacherWrapper) else TeacherWrapper(teacher_model)
    total_loss, total_samples, bleu_scores = 0.0, 0, []
    for batch_idx, (imgs, captions) in enumerate(data_loader):
        if batch_idx >= max_batches:
            break
        imgs, captions = imgs.to(device), captions.to(device)
        captions_input, captions_target = captions[:-1, :], captions[1:, :]
        teacher_outputs = teacher_wrapper(imgs.float(), captions_input.long())
        teacher_outputs["encoder_features"] = projectors["encoder"](teacher_outputs["encoder_features"])
        if fused and student_model.supports_fused_validation():
            # bf16 mode: the vocabulary head is reduced to the loss partials and the argmax in the GEMM's epilogue -- the (T,B,V)
            # student logits are never written
            _, loss_dict, predicted = distill_loss.evaluate_fused(student_model, imgs, captions_input, teacher_outputs, captions_target)
        else:
            student_logits, student_encoder_features, student_hidden_states, _ = student_model(imgs, captions_input)
            student_outputs = {"logits": student_logits, "encoder_features": student_encoder_features,
                               "hidden_states": student_hidden_states}
            _, loss_dict, predicted = distill_loss.evaluate(student_outputs, teacher_outputs, captions_target)
        n = imgs.size(0)
        total_loss += loss_dict["total_loss"] * n
        total_samples += n
        if batch_idx < 5:
            bleu = _ops.bleu1(predicted, captions_target)                  # every sample, on the device
            k = bleu.numel() if bleu_all_samples else min(2, bleu.numel())
            bleu_scores.extend(bleu[:k].tolist())
    if was_training:
        student_model.train()
    avg_loss = total_loss / max(total_samples, 1)
    avg_bleu = float(sum(bleu_scores) / len(bleu_scores)) if bleu_scores else 0.0
    return avg_loss, avg_bleu
