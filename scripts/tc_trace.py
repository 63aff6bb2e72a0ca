"""Profiling aid: per-k-iteration clock64 stamps of one tcgen05 GEMM CTA and per-CTA lifetimes."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from rstnet_b200 import ops, _lib
mode = sys.argv[1]
prec = 0
if mode == "lin":
    M, K, N = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda"); out = torch.empty(M, N, device="cuda")
    plan = ops.TcGemm(a, 0, K, M * K, K, M, 1, w, K, out, 0, N, M * N, M, 1, precision=prec)
    nk = K // 32
else:  # conv in time-major layout: B Cin Cout k s T pre
    B, Cin, Cout, k, s, T, pre = [int(v) for v in sys.argv[2:9]]
    use_r = len(sys.argv) > 9 and sys.argv[9] == "1"
    buf = torch.randn(k - s + T, B, Cin, device="cuda"); w = torch.randn(Cout, k * Cin, device="cuda")
    out = torch.empty(T // s, B, Cout, device="cuda")
    res = torch.randn(T // s, B, Cout, device="cuda") if use_r else None
    plan = ops.TcGemm(buf, 0, Cin, B * Cin, Cin, B, k - s + T, w, Cin, out, 0, Cout, B * Cout, B, T // s, taps=k, tap_do=1, o_mul=s,
                      pre_act=0, post_act=pre, precision=prec, R=res, r_i_stride=Cout, r_o_stride=B * Cout)
    nk = k * Cin // 32
gx, gy, bn = C.c_int32(), C.c_int32(), C.c_int32()
_lib.lib().rstnet_tc_gemm_grid(plan._h, C.byref(gx), C.byref(gy), C.byref(bn))
trace = torch.zeros(max(nk, 8), 8, dtype=torch.int64, device="cuda")
ct = torch.zeros(max(gx.value, 148), 4, dtype=torch.int64, device="cuda")
for _ in range(3): plan.run()
_lib.lib().rstnet_tc_gemm_set_trace(plan._h, trace.data_ptr(), ct.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); plan.run(); e1.record(); torch.cuda.synchronize()
t = trace.cpu(); c = ct.cpu(); c = c[c[:, 0] > 0]
t0 = int(t[0, 0])
print(f"{sys.argv[1:]}: kernel {e0.elapsed_time(e1)*1e3:.1f} us, grid ({gx.value},{gy.value}) BN={bn.value}, k-iters {nk}")
g0 = int(c[:, 0].min())
life = (c[:, 3] - c[:, 0]).float()
print(f"CTA lifetime ns: mean {life.mean():.0f} min {life.min():.0f} max {life.max():.0f}; setup {(c[:,1]-c[:,0]).float().mean():.0f} ns; "
      f"main {(c[:,2]-c[:,1]).float().mean():.0f} ns; all CTAs span {int(c[:,3].max()) - g0} ns")
order = torch.argsort(c[:, 0])
print("first CTA starts (ns):", [int(c[i, 0]) - g0 for i in order[:4]], " 149th..:", "")
print("epilogue stamps (end, pre-bar1, post-bar1, boxes free / residual landed, staged, fenced, post-bar3, next tile's drain start):", [int(t[k, 7]) - t0 if int(t[k, 7]) else -1 for k in range(8)])
print("kit  prod  landed  xformed  mma_start mma_issued | drain_b drain_e")
for k in range(min(max(nk, 4), 16)):
    r = [int(v) - t0 if int(v) else -1 for v in t[k]]
    print(f"{k:3d} {r[0]:6d} {r[1]:7d} {r[2]:8d} {r[3]:9d} {r[4]:9d} | {r[5]:7d} {r[6]:7d}")
