"""Runs the named kernels of the hot path alone at BASELINE config-2 shapes (for `ncu --set full -k regex:...`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecaptioner_b200 import _ops

lib = _ops.load_library()
dev = torch.device("cuda:0")
B, T, V, E, H = 512, 20, 5000, 256, 512
N = T * B
st = torch.cuda.current_stream().cuda_stream
y = torch.randn(N, V, device=dev).bfloat16(); z = torch.randn(N, V, device=dev) * 2
tgt = torch.randint(1, V, (N,), device=dev); nval = torch.tensor([N], dtype=torch.int32, device=dev)
dy = torch.empty_like(y); rows = torch.empty(2, N, device=dev)
A = torch.randn(N, E, device=dev).bfloat16(); W = torch.randn(V, E, device=dev).bfloat16(); C = torch.empty(N, V, device=dev, dtype=torch.bfloat16)
A2 = torch.randn(B, E + H, device=dev).bfloat16(); W2 = torch.randn(4 * H, E + H, device=dev).bfloat16(); C2 = torch.empty(B, 4 * H, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
# AttentionRefinement at the bench shape (tensor-core MHA forward / backward) and the native optimizer step on 7.3 M parameters
from imagecaptioner_b200.student_model import AttentionRefinement
from imagecaptioner_b200.optim import FlatAdamW
ref = AttentionRefinement(E, 4).to(dev).eval(); ref.compute_dtype = torch.bfloat16
xr = torch.randn(B, 49, E, device=dev, requires_grad=True)
big = [torch.nn.Parameter(torch.randn(7_330_000 // 2, device=dev)), torch.nn.Parameter(torch.randn(7_330_000 // 2, device=dev))]
opt = FlatAdamW([{"params": [big[0]], "clip_group": 0}, {"params": [big[1]], "clip_group": 1}], lr=1e-4)
opt.reducer.flat.normal_()
for _ in range(reps):
    out = ref(xr); out.float().sum().backward()
    opt.step()
    assert lib.b2c_kd_token_loss(y.data_ptr(), z.data_ptr(), tgt.data_ptr(), N, V, 4.0, 0.7, 0.0, 1.0, nval.data_ptr(), dy.data_ptr(),
                                 rows[0].data_ptr(), rows[1].data_ptr(), _ops.B2C_BF16, st) == 0
    _ops.gemm(A, W, N, V, E, C=C)
    _ops.gemm(A2, W2, B, 4 * H, E + H, C=C2)
torch.cuda.synchronize()
print("ok")
