"""Generate tests/golden/lm_small.npz by running the UNMODIFIED reference GPT (streaming decode) here.

Run:  python -m oracle.gen_golden_lm      (needs /root/reference)

The reference `models.llama_streaming.GPT` (small config, seeded synthetic weights from
oracle/lm_oracle.synthetic_weights) is stepped for a few frames under `with gpt.streaming(B)` with
greedy sampling, in fp32 and in bf16; the oracle restatement must reproduce it bit for bit, and the
reference outputs are stored.
"""
from __future__ import annotations

import os
import sys

os.environ.setdefault("NO_TORCH_COMPILE", "1")
os.environ.setdefault("NO_CUDA_GRAPH", "1")
sys.dont_write_bytecode = True
REF = "/root/reference/MLLM_v2"

import numpy as np
import torch

from . import lm_oracle as L
from .gen_golden import weights_digest

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def reference_frames(w, cfg, seqs, dtype):
    sys.path.insert(0, REF)
    from models.llama_streaming import GPT, Config
    m = GPT(Config(**cfg.reference_kwargs())).eval()
    sd = m.state_dict()
    assert set(sd.keys()) == set(w.keys()), set(sd.keys()) ^ set(w.keys())
    m.load_state_dict({k: v.float() for k, v in w.items()}, strict=True)
    m = m.to(dtype)
    B = seqs[0].shape[0]
    outs = []
    with torch.no_grad(), m.streaming(B):
        for seq in seqs:
            out, tl = m.forward_global(seq)
            tt = torch.argmax(tl.float(), dim=-1)
            toks, prev, al = [tt[:, 0]], tt[:, :, None], []
            with m.codecformer.streaming(B):
                for k in range(cfg.dep_q):
                    lg = m.forward_codecformer(k, prev, out)
                    al.append(lg[:, 0, 0])
                    prev = torch.argmax(lg.float(), dim=-1)
                    toks.append(prev[:, 0, 0])
            outs.append((out, tl, torch.stack(al, 1), torch.stack(toks, 1)))
    return outs


def main():
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    os.makedirs(GOLDEN, exist_ok=True)
    cfg = L.SMALL
    B, frames = 3, 20  # 20 > context (16): the ring wraps
    g = torch.Generator().manual_seed(99)
    seqs = []
    for f in range(frames):
        s = torch.randint(0, 2048, (B, cfg.n_q + 1, 1), generator=g)
        s[:, 0] = torch.randint(0, 150000, (B, 1), generator=g)
        if f == 0:
            s[:, 1:] = cfg.audio_card          # initial audio token (llama_streaming.py:609)
            s[:, 0] = 151655                    # hard-coded text_initial_token_id (:598-604)
        if f == 3:
            s[0, 2] = -1                        # zero_token_id row (ScaledEmbedding, :505-517)
        seqs.append(s)
    save = {}
    for dtype, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        w = L.synthetic_weights(cfg, seed=7, dtype=torch.float32, std=0.05)
        wd = {k: v.to(dtype) for k, v in w.items()}
        ref = reference_frames(w, cfg, seqs, dtype)
        gs = L.GPTStream(wd, cfg, B)
        with torch.no_grad():
            for f, seq in enumerate(seqs):
                o = L.greedy_frame(gs, seq)
                for a, b, name in zip(o, ref[f], ("out", "text_logits", "audio_logits", "tokens")):
                    assert torch.equal(a, b), f"oracle != reference: {tag} frame {f} {name}: {(a.float() - b.float()).abs().max()}"
        print(f"{tag}: oracle == reference on {frames} frames (bit for bit)")
        keep = (0, 1, 3, 15, 16, 19)
        save[f"{tag}_out"] = torch.stack([ref[f][0].float() for f in keep]).numpy()
        save[f"{tag}_text_top"] = torch.stack([ref[f][1].float().topk(8, dim=-1).values for f in keep]).numpy()
        save[f"{tag}_audio_logits"] = torch.stack([ref[f][2].float() for f in keep]).numpy()
        save[f"{tag}_tokens"] = torch.stack([ref[f][3] for f in range(frames)]).numpy()
        save[f"{tag}_keep"] = np.array(keep)
        if tag == "f32":
            save["weights_sha256"] = np.array(weights_digest(w))
    save["seqs"] = torch.stack(seqs).numpy()
    np.savez_compressed(os.path.join(GOLDEN, "lm_small.npz"), **save)
    print("wrote", os.path.join(GOLDEN, "lm_small.npz"))


if __name__ == "__main__":
    main()
