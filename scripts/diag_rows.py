"""Diagnostic: identical streams must give identical tokens / audio (batch rows are independent).
Run as the first process on a fresh box, several times: the rare mismatch was only ever seen there."""
import sys, torch
sys.path.insert(0, ".")
from oracle import mimi_spec as S
from rstnet_b200.codec import MimiCodec
w = S.synthetic_weights(S.OFFICIAL, seed=41)
m = MimiCodec(encoder_rates=[8, 6, 5, 4], codebook_size=2048, codebook_dim=256, rvq_layers=8)
m.load_state_dict(w, strict=True); m = m.to("cuda").eval()
B = 256
x = S.synthetic_audio(4, 1920 * 2, seed=33).repeat(B // 4, 1, 1).cuda()
bad = 0
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    with torch.no_grad(), m.streaming(B):
        for i in range(2):
            c = m.encode(x[..., i * 1920:(i + 1) * 1920]); wv = m.decode(c)
            c4 = c.view(B // 4, 4, 8, -1); w4 = wv.view(B // 4, 4, -1)
            cb = (c4 != c4[:1]).any(-1).any(-1); wb = (w4 != w4[:1]).any(-1)
            if cb.any() or wb.any():
                bad += 1
                d = (w4 - w4[:1]).abs()
                cols = d.amax(dim=(0, 1)).nonzero().flatten()
                st = m._stream_state
                for plan in st.dec.values():
                    for name in ("qup", "X", "a", "yd", "yda", "hd"):
                        for li, bf in enumerate(plan.debug_bufs[name]):
                            t = bf.t                                   # [rows, B, C]
                            v = t.view(t.shape[0], B // 4, 4, t.shape[2])
                            dd = (v != v[:, :1]).any(-1)              # [rows, groups, 4]
                            if dd.any():
                                rows_t = dd.any(-1).any(-1).nonzero().flatten().tolist()
                                streams = dd.any(0).nonzero().tolist()
                                ch = (v != v[:, :1]).any(0).any(0).any(0).nonzero().flatten()
                                print(f"   buf {name}[{li}] C={t.shape[2]} ctx={bf.ctx}: time rows {rows_t[:12]} ({len(rows_t)}), streams {[g*4+k for g,k in streams][:10]} ({len(streams)}), channels {int(ch.min())}..{int(ch.max())} ({len(ch)})")
                try:
                    plan = list(st.dec.values())[0]
                    done = False
                    for name in ("X", "a", "yd", "yda", "hd"):
                        for li, bf in enumerate(plan.debug_bufs[name]):
                            t = bf.t
                            v = t.view(t.shape[0], B // 4, 4, t.shape[2])
                            dd = (v != v[:, :1]).any(-1)
                            if dd.any() and not done:
                                done = True
                                idx = dd.nonzero()
                                print("   FIRST corrupted buffer", name, li, "all corrupted streams:", sorted(set((int(g) * 4 + int(k)) for _, g, k in idx)))
                                r, g, k = [int(z) for z in idx[0]]
                                bad_v, good_v = v[r, g, k], v[r, 0, k]
                                print("   row", r, "stream", g * 4 + k, "bad[:6]", bad_v[:6].tolist(), "good[:6]", good_v[:6].tolist())
                                same_t = [(int(tt)) for tt in range(t.shape[0]) if torch.equal(v[tt, 0, k], bad_v)]
                                print("   bad vector equals the correct output of time rows:", same_t[:8])
                                for kk in range(4):
                                    same_k = [int(tt) for tt in range(max(0, r - 3), min(t.shape[0], r + 4)) if torch.equal(v[tt, 0, kk], bad_v)]
                                    if same_k: print("   equals content", kk, "at rows", same_k)
                                if name == "a" and li >= 1:
                                    R = plan.debug_bufs["yd"][li - 1]
                                    rv = R.t[r - bf.ctx + R.ctx, g * 4 + k]
                                    elu = torch.where(rv > 0, rv, torch.exp(rv) - 1)
                                    print("   residual[:6]", rv[:6].tolist(), " bad==elu(residual):", bool(torch.allclose(bad_v, elu, atol=1e-6)),
                                          " bad-good[:6]", (bad_v - good_v)[:6].tolist())
                                    nz = (bad_v != good_v).nonzero().flatten()
                                    print("   differing channels:", int(nz.min()), "..", int(nz.max()), "count", len(nz))
                except Exception as e:
                    print("   diag error", repr(e))
                print(f"rep {rep} frame {i}: code-mismatch rows {cb.nonzero().tolist()[:8]} ({int(cb.sum())}), wav-mismatch rows {wb.nonzero().tolist()[:8]} ({int(wb.sum())})"
                      f" max wav diff {float(d.max()):.3e}; wav sample range {int(cols.min()) if len(cols) else -1}..{int(cols.max()) if len(cols) else -1} ({len(cols)} samples)")
print("mismatching frames:", bad)
