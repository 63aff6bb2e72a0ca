"""Micro-benchmark of the LM ring decode attention kernel (B=64, 32 heads, 2047 keys, hs=128)."""
import sys, torch
sys.path.insert(0, ".")
from rstnet_b200 import _lib, ops
B, nh, hs, cap = 64, 32, 128, 2048
dev = "cuda"
kvs = [torch.randn(2, B, nh, cap, hs, device=dev, dtype=torch.bfloat16) for _ in range(8)]  # 8 x 2.1 GB > L2
q = torch.randn(B, nh * hs, device=dev, dtype=torch.bfloat16)
out = torch.empty_like(q)
off = torch.full((1,), cap + 8, dtype=torch.int64, device=dev)
L = _lib.lib(); st = ops._stream()
def run(kv):
    _lib.check(L.rstnet_lm_ring_decode_attention_bf16(q.data_ptr(), kv.data_ptr(), off.data_ptr(), out.data_ptr(), B, nh, hs, cap, cap, st))
for kv in kvs: run(kv)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    for kv in kvs: run(kv)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 24
gb = 2 * B * nh * (cap - 1) * hs * 2 / 1e9
print(f"attention: {ms*1e3:.1f} us per launch, {gb/ms*1e3/1e3:.3f} TB/s ({gb/ms/6.5716*100:.1f}% of measured HBM peak)")
