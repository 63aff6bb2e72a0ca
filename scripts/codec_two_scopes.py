"""Experiment: one 256-stream streaming scope vs two 128-stream scopes replayed on two CUDA streams (small-M transformer GEMMs
of one half overlap the other half's)."""
import sys, torch
sys.path.insert(0, ".")
import bench
from specs import mimi_spec as S
from rstnet_b200.codec import _StreamState

dev = torch.device("cuda", 0)
m = bench._mimi(dev, S)
eng = m._eng()
NF = 40
x = torch.randn(256, 1, 1920 * NF, device=dev)


def timeit(fn, n=30):
    for i in range(6):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(6 + i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    one = _StreamState(eng, 256)
    def f1(i):
        c = one.encode(x[..., i * 1920:(i + 1) * 1920]); one.decode(c)
    t1 = timeit(f1)
    del one
    parts = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    h = 256 // parts
    sts = [_StreamState(eng, h) for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    def f2(i):
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(cur)
        for k in range(parts):
            streams[k].wait_event(ev)
            with torch.cuda.stream(streams[k]):
                c = sts[k].encode(x[k * h:(k + 1) * h, :, i * 1920:(i + 1) * 1920]); sts[k].decode(c)
                e = torch.cuda.Event(); e.record(streams[k]); cur.wait_event(e)
    t2 = timeit(f2)
print(f"one scope of 256: {t1:.3f} ms/step ({256/t1:.1f} k frames/s); {parts} scopes of {h}: {t2:.3f} ms/step ({256/t2:.1f} k frames/s)")
