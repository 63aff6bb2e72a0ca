"""Tuning aid: time the LM weight-streaming GEMM (+ its finalize) for each split-K factor, per layer shape.
Each (shape, splits) runs over 16..64 distinct weight matrices back to back (working set >> L2) as one replayed CUDA graph."""
import os, sys, torch
sys.path.insert(0, ".")
from rstnet_b200.lm import SkinnyGemm
dev = "cuda"
M = 64
shapes = {"qkv": (12288, 4096, "plain"), "proj": (4096, 4096, "norm"), "fc": (22016, 4096, "silu"), "mlp_proj": (4096, 11008, "norm"),
          "d_in": (3072, 1024, "plain"), "d_out": (1024, 1024, "norm"), "d_gate": (8448, 1024, "silu"), "d_ffout": (1024, 4224, "norm"),
          "d_head": (2052, 1024, "plain")}
only = sys.argv[1:] or list(shapes)
for name in only:
    N, K, mode = shapes[name]
    reps = 16 if N * K > 8e6 else 64
    Ws = [torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.02 for _ in range(reps)]
    X = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    R = torch.randn(M, N, device=dev, dtype=torch.bfloat16)
    ws = torch.empty(8 * M * N, dtype=torch.float32, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    aux = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    nw = torch.ones(N, device=dev, dtype=torch.bfloat16)
    so = torch.empty(M, N // 2, device=dev, dtype=torch.bfloat16)
    res = {}
    for s in (1, 2, 3, 4, 6, 8):
        os.environ["RSTNET_SKINNY_SPLITS"] = str(s)
        try:
            if mode == "plain":
                plans = [SkinnyGemm(X, W, out, None, ws) for W in Ws]
            elif mode == "norm":
                plans = [SkinnyGemm(X, W, out, R, ws, norm_w=nw, aux=aux, eps=1e-5) for W in Ws]
            else:
                plans = [SkinnyGemm(X, W, None, None, ws, silu_out=so) for W in Ws]
        except Exception as e:
            res[s] = None
            continue
        for p in plans[:2]:
            p.run()
        torch.cuda.synchronize()
        # timed as a replayed CUDA graph (how the decode step runs them): kernel + finalize launch boundaries included
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for p in plans:
                p.run()
        g.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
        res[s] = round(best, 1)
        del g
        del plans
    os.environ.pop("RSTNET_SKINNY_SPLITS", None)
    print(name, (N, K), mode, res, flush=True)
    del Ws
