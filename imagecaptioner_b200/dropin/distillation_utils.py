"""Shadow of the reference's ``src/distillation_utils.py`` (imported at train_student_kd.py:17-24 and
evaluate_student.py:19); see ``dropin/student_model.py`` and INTEGRATION.md."""
from imagecaptioner_b200.distillation_utils import (DistillationLoss, FeatureProjector, TeacherWrapper,  # noqa: F401
                                                    create_feature_projectors, validate_distillation_setup,
                                                    compute_bleu_score, log_training_progress)
