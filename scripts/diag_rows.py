"""Diagnostic: identical streams must give identical tokens / audio (batch rows are independent)."""
import sys, torch
sys.path.insert(0, ".")
from oracle import mimi_spec as S
from rstnet_b200.codec import MimiCodec
w = S.synthetic_weights(S.OFFICIAL, seed=41)
m = MimiCodec(encoder_rates=[8, 6, 5, 4], codebook_size=2048, codebook_dim=256, rvq_layers=8)
m.load_state_dict(w, strict=True); m = m.to("cuda").eval()
B = 256
x = S.synthetic_audio(4, 1920 * 3, seed=33).repeat(B // 4, 1, 1).cuda()
for rep in range(3):
    with torch.no_grad(), m.streaming(B):
        for i in range(3):
            c = m.encode(x[..., i * 1920:(i + 1) * 1920]); wv = m.decode(c)
            c4 = c.view(B // 4, 4, 8, -1); w4 = wv.view(B // 4, 4, -1)
            cb = (c4 != c4[:1]).any(-1).any(-1); wb = (w4 != w4[:1]).any(-1)
            print(f"rep {rep} frame {i}: code-mismatch rows {cb.nonzero().tolist()[:6]} ({int(cb.sum())}), wav-mismatch rows {wb.nonzero().tolist()[:6]} ({int(wb.sum())})"
                  f" max wav diff {float((w4 - w4[:1]).abs().max()):.3e}")
