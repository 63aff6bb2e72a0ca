"""Same-box A/B of MimiCodec knobs on the 256-stream streaming step (graph replay, ms per encode+decode frame step).
usage: codec_ab.py knob=value[,value...]   e.g.  codec_ab.py fused_rope_attention=1,0"""
import sys, torch
sys.path.insert(0, ".")
import bench
from specs import mimi_spec as S

dev = torch.device("cuda", 0)
knob, vals = sys.argv[1].split("=")
vals = [int(v) for v in vals.split(",")]
x = torch.randn(256, 1, 1920 * 60, device=dev)


def run(v):
    m = bench._mimi(dev, S)
    setattr(m, knob, bool(v) if isinstance(getattr(m, knob), bool) else v)
    with torch.no_grad(), m.streaming(256):
        def step(i):
            c = m.encode(x[..., (i % 60) * 1920:(i % 60 + 1) * 1920]); m.decode(c)
        for i in range(8):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(50):
            step(8 + i)
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 50


for rep in range(2):
    for v in vals:
        print(f"{knob}={v}: {run(v):.4f} ms per frame step", flush=True)
