"""Launch-boundary micro-benchmark: a chain of N tiny dependent kernels (RMSNorm over [64, 4096] bf16), as a CUDA graph and as
plain stream launches; run with RSTNET_PDL=0 and =1."""
import os, sys, torch
sys.path.insert(0, ".")
from rstnet_b200 import _lib, ops
L = _lib.lib(); dev = "cuda"
N = 200
a = torch.randn(64, 4096, device=dev).to(torch.bfloat16); b = torch.empty_like(a)
w = torch.ones(4096, device=dev, dtype=torch.bfloat16)

def chain():
    st = ops._stream()
    x, y = a, b
    for _ in range(N):
        _lib.check(L.rstnet_lm_rms_norm_bf16(x.data_ptr(), w.data_ptr(), y.data_ptr(), 64, 4096, 1e-5, 0, st))
        x, y = y, x

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3 / N

chain(); torch.cuda.synchronize()
t_stream = timeit(chain, 5)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    chain()
t_graph = timeit(g.replay)
print(f"PDL={os.environ.get('RSTNET_PDL','0')}: stream {t_stream:.2f} us/launch, graph {t_graph:.2f} us/launch")
