"""Do small kernels on forked streams overlap inside a CUDA graph?  Times chains of gate GEMMs / tiny kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecaptioner_b200 import _ops
lib = _ops.load_library()
dev = torch.device("cuda:0")
E, H = 256, 512
W = torch.randn(4 * H, E + H, device=dev).bfloat16()
def mk(M): return torch.randn(M, E + H, device=dev).bfloat16(), torch.empty(M, 4 * H, device=dev)
A512, C512 = mk(512)
subs = [mk(128) for _ in range(4)]
tiny = [torch.ones(8, device=dev) for _ in range(4)]
one = torch.ones(1, device=dev)
streams = [torch.cuda.Stream() for _ in range(4)]

def gemm(A, C, M): _ops.gemm(A, W, M, 4 * H, E + H, C=C)
def tinyk(t):
    lib.b2c_scale_inplace(t.data_ptr(), 8, 0, one.data_ptr(), torch.cuda.current_stream().cuda_stream)

def serial(fn, n):
    for _ in range(n): fn()
def forked(fns, n):
    cur = torch.cuda.current_stream()
    for s, f in zip(streams, fns):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            for _ in range(n): f()
    for s in streams: cur.wait_stream(s)

def timed(build, reps=20):
    build(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): build()
    for _ in range(3): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

print("A  80 x gemm M=128 serial        : %8.1f us" % timed(lambda: serial(lambda: gemm(*subs[0], 128), 80)))
print("B  4 branches x 20 gemm M=128    : %8.1f us" % timed(lambda: forked([lambda i=i: gemm(*subs[i], 128) for i in range(4)], 20)))
print("C  20 x gemm M=512 serial        : %8.1f us" % timed(lambda: serial(lambda: gemm(A512, C512, 512), 20)))
print("D  80 x tiny kernel serial       : %8.1f us" % timed(lambda: serial(lambda: tinyk(tiny[0]), 80)))
print("E  4 branches x 20 tiny kernels  : %8.1f us" % timed(lambda: forked([lambda i=i: tinyk(tiny[i]) for i in range(4)], 20)))
print("F  1 x gemm M=512                : %8.1f us" % timed(lambda: serial(lambda: gemm(A512, C512, 512), 1)))
