"""Small end-to-end run of every entry point for compute-sanitizer (memcheck): KD step fp32 + bf16, greedy decode, attention accessor."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import kd_oracle as O
from tests.harness import build_student, run_kd_step
dev = torch.device("cuda:0")
for (B, T, V, E, H, L, refine, Et) in [(5, 4, 104, 32, 64, 2, True, 40), (3, 3, 203, 48, 96, 3, True, 48), (130, 2, 200, 64, 128, 1, False, 64)]:
    params = O.init_student_params(V, E, H, L, refine, seed=0)
    pparams = O.init_projector_params(Et, E, seed=1)
    batch = O.synthetic_batch(B, T, V, E, H, Et=Et, seed=7)
    model, projector = build_student(params, pparams, V, E, H, L, refine, Et, dev)
    for dt in (torch.float32, torch.bfloat16):
        out = run_kd_step(model, projector, batch, dev, dt)
        assert all(v == v for v in out["loss"].values())
    toks, lens = model.decoder.greedy(torch.randn(B, 49, E, device=dev), 4)
    ctx, w = model.decoder.attention_mechanism(torch.randn(B, H, device=dev), torch.randn(B, 49, E, device=dev))
    torch.cuda.synchronize()
print("sanitize run ok")
