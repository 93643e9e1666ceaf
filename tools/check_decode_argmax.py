"""Greedy decode in bf16 mode on a seeded random student; writes tokens + lengths to argv[1].
Run once as is (argmax fused into the vocabulary-head GEMM epilogue) and once with B2C_DECODE_LOGITS=1 (logits + argmax kernel):
the two files must be identical (tests/test_gpu_parity.py::test_fused_argmax_decode_equals_logits_path)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import kd_oracle as O
from tests.harness import build_student

V, E, H, L, B, S, max_len = 5000, 256, 512, 2, 300, 49, 12
dev = torch.device("cuda:0")
params = O.init_student_params(V, E, H, L, True, seed=21)
params["decoder.output_projection.3.weight"] = params["decoder.output_projection.3.weight"] * 8     # spread the logits
model, _ = build_student(params, O.init_projector_params(384, E, seed=1), V, E, H, L, True, 384, dev)
model.decoder.compute_dtype = torch.bfloat16
model.attention_refinement.compute_dtype = torch.bfloat16
feats = torch.randn(B, S, E, generator=torch.Generator().manual_seed(5)).to(dev)
with torch.no_grad():
    refined = model.attention_refinement(feats)
    tokens, lengths = model.decoder.greedy(refined, max_len, 1, 2)
torch.cuda.synchronize()
torch.save({"tokens": tokens.cpu(), "lengths": lengths.cpu()}, sys.argv[1])
print("distinct tokens", int(tokens.unique().numel()), "mean length", float(lengths.float().mean()))
