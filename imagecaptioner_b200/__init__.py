"""imagecaptioner_b200 — B200-native (sm_100a) KD hot path of VeeraKarthick609/ImageCaptioner.

Only what the hot path needs: the C-ABI CUDA library (csrc/, include/b2c.h), its ctypes
binding and autograd Functions (_ops), the drop-in mirrors of the reference's two modules on the path
(student_model, distillation_utils) and the data-parallel glue (ddp).
"""
from . import _ops  # noqa: F401
from .student_model import CNNEncoder, AttentionRefinement, LSTMDecoder, CaptioningStudent, count_parameters  # noqa: F401
from .distillation_utils import (DistillationLoss, FeatureProjector, TeacherWrapper, create_feature_projectors,  # noqa: F401
                                 validate_distillation_setup, compute_bleu_score, log_training_progress)

__all__ = ["CNNEncoder", "AttentionRefinement", "LSTMDecoder", "CaptioningStudent", "count_parameters",
           "DistillationLoss", "FeatureProjector", "TeacherWrapper", "create_feature_projectors",
           "validate_distillation_setup", "compute_bleu_score", "log_training_progress"]
