"""Shadow of the reference's ``src/student_model.py``: put this directory on sys.path AHEAD of the
reference's ``src/`` and ``from student_model import CaptioningStudent`` (train_student_kd.py:14,
evaluate_student.py:17) resolves to the B200-native implementation.  See INTEGRATION.md."""
from imagecaptioner_b200.student_model import *  # noqa: F401,F403
from imagecaptioner_b200.student_model import CNNEncoder, AttentionRefinement, LSTMDecoder, CaptioningStudent, count_parameters  # noqa: F401
