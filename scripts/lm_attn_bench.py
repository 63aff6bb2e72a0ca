"""Micro-benchmark of the LM ring decode attention kernel (B=64, 32 heads, 2047 keys, hs=128): one CTA per (stream, head)
vs the persistent key-split form."""
import sys, torch
sys.path.insert(0, ".")
from rstnet_b200 import _lib, ops
import json
B, nh, hs, cap = 64, 32, 128, 2048
dev = "cuda"
kvs = [torch.randn(2, B, nh, cap, hs, device=dev, dtype=torch.bfloat16) for _ in range(8)]  # 8 x 2.1 GB > L2
q = torch.randn(B, nh * hs, device=dev, dtype=torch.bfloat16)
out = torch.empty_like(q)
off = torch.full((B,), cap + 8, dtype=torch.int64, device=dev)
L = _lib.lib(); st = ops._stream()
ws = torch.zeros(L.rstnet_lm_attention_split_workspace(B, nh, hs), dtype=torch.uint8, device=dev)

res = {}
for name, w in (("one_cta_per_job", None), ("key_split", ws.data_ptr())):
    def run(kv):
        _lib.check(L.rstnet_lm_ring_decode_attention_bf16(q.data_ptr(), kv.data_ptr(), off.data_ptr(), 1, out.data_ptr(), B, B, nh, nh, hs,
                                                          cap, cap, w, st))
    for kv in kvs: run(kv)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        for kv in kvs: run(kv)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 24
    gb = 2 * B * nh * (cap - 1) * hs * 2 / 1e9
    res[name] = {"us": ms * 1e3, "gbps": gb / ms * 1e3}
    print(f"{name}: {ms*1e3:.1f} us per launch, {gb/ms:.3f} TB/s")
    res[name]["out"] = out.float().clone()
d = (res["one_cta_per_job"].pop("out") - res["key_split"].pop("out")).abs().max().item()
res["max_abs_diff"] = d
print(json.dumps(res))
