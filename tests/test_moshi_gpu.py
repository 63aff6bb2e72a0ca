"""GPU parity tests (-m gpu) of the Moshi-style LMModel / LMGen twin (rstnet_b200/moshi.py) against the golden vectors of
the unmodified reference (tests/golden/moshi_small.npz, oracle/gen_golden_moshi.py) and the oracle restatement."""
import os

import numpy as np
import pytest
import torch

from oracle import moshi_oracle as M
from rstnet_b200.moshi import LMGen, LMModel

pytestmark = pytest.mark.gpu
DEV, BF = "cuda", torch.bfloat16


def _cos(a, b):
    a, b = a.float().flatten().cpu(), b.float().flatten().cpu()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp(min=1e-12))


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-6))


@pytest.fixture(scope="module")
def moshi(golden_dir):
    from oracle.gen_golden import weights_digest
    cfg = M.SMALL
    w = M.synthetic_weights(cfg, seed=5)
    gold = np.load(os.path.join(golden_dir, "moshi_small.npz"))
    assert weights_digest(w) == str(gold["weights_sha256"])
    m = LMModel(**cfg.reference_kwargs())
    assert set(m.state_dict().keys()) == set(w.keys())
    m.load_state_dict(w, strict=True)
    return m.to(DEV, BF).eval(), w, cfg, gold


def test_forward_text_and_depformer_vs_reference_golden(moshi):
    """Teacher-forced with the frames the reference's own LMGen fed its model (20 steps, the ring wraps at 16):
    transformer_out and the 8 depth logits vs the reference's bf16 values; greedy tokens vs the reference's."""
    m, w, cfg, gold = moshi
    feeds = torch.from_numpy(gold["bf16_feed"])                # [steps, B, 17, 1]
    outs = torch.from_numpy(gold["bf16_out"])                  # [steps - 1, B, 9, 1] (delayed view of the tokens)
    keep = list(gold["keep"])
    B = feeds.shape[1]
    m.use_cuda_graphs = True
    agree = total = 0
    with m.streaming(B):
        for t in range(feeds.shape[0]):
            out, tl = m.forward_text(feeds[t].to(DEV))
            assert out.shape == (B, 1, cfg.dim) and tl.shape == (B, 1, 1, cfg.text_card)
            # the tokens the reference generated at step t are what it fed at step t + 1 (own streams, delays 0 / 1)
            lgs = []
            with m.depformer.streaming(B):
                prev = tl.float().argmax(-1)[:, 0, :, None] if t + 1 >= feeds.shape[0] else feeds[t + 1][:, 0:1].to(DEV)
                for k in range(cfg.dep_q):
                    lg = m.forward_depformer(k, prev, out)
                    assert lg.shape == (B, 1, 1, cfg.card)
                    lgs.append(lg[:, 0, 0])
                    prev = lg.float().argmax(-1)
            al = torch.stack(lgs, 1)
            if t in keep:
                i = keep.index(t)
                assert _cos(out, torch.from_numpy(gold["bf16_transformer_out"][i])) >= 0.999, t
                assert _rel(out, torch.from_numpy(gold["bf16_transformer_out"][i])) <= 5e-2, t
            if t + 1 < feeds.shape[0]:
                ref_text = feeds[t + 1][:, 0, 0]
                agree += int((tl.float().argmax(-1)[:, 0, 0].cpu() == ref_text).sum()); total += B
    print(f"moshi text-token agreement with the reference under teacher forcing: {agree}/{total}")
    assert agree >= 0.85 * total


def test_lmgen_step_closed_loop(moshi):
    """LMGen.step (models/model.py:490-562): None during the max_delay warm-up, then [B, dep_q + 1, 1] in the delayed
    layout; greedy closed loop vs the reference's tokens and, decision by decision, vs the oracle teacher-forced with
    OUR tokens (a bf16 near-tie flip changes everything after it, so the golden comparison alone would be brittle)."""
    m, w, cfg, gold = moshi
    inputs = torch.from_numpy(gold["inputs"])
    ref = torch.from_numpy(gold["bf16_out"])
    B = inputs.shape[1]
    gen = LMGen(m, use_sampling=False)
    outs = []
    with gen.streaming(B):
        for t in range(inputs.shape[0]):
            o = gen.step(inputs[t].to(DEV))
            if t < max(cfg.delays):
                assert o is None
            else:
                assert o.shape == (B, cfg.dep_q + 1, 1) and o.dtype == torch.int64
                outs.append(o.cpu())
    mine = torch.stack(outs)
    same = (mine == ref).all(dim=(1, 2, 3))
    first_diff = int((~same).nonzero()[0]) if bool((~same).any()) else len(same)
    print(f"moshi LMGen closed loop: identical to the reference for the first {first_diff}/{len(same)} output frames")
    # decision-level check: every (text, audio) decision of OUR closed loop under the oracle teacher-forced with our tokens
    gen2 = LMGen(m, use_sampling=False)
    ora = M.LMGenOracle({k: v.to(BF) for k, v in w.items()}, cfg, B)
    exact = n = 0
    worst = 0.0
    with gen2.streaming(B), torch.no_grad():
        for t in range(inputs.shape[0]):
            gen2.step(inputs[t].to(DEV))
            pos = gen2._st.offset % gen2._st.cache.shape[2]
            ours = gen2._st.cache[:, :cfg.dep_q + 1, pos].cpu()            # [B, 9] tokens we generated at this step
            ora.step(inputs[t], force=ours)
            _, _, text_logits, alog = ora.last
            lt = text_logits.float()[:, 0, 0]
            d = [lt.max(-1).values - lt.gather(1, ours[:, :1])[:, 0]]
            la = alog.float()                                               # [B, dep_q, card]
            d.append((la.max(-1).values - la.gather(2, ours[:, 1:, None])[:, :, 0]).flatten())
            d = torch.cat(d)
            worst = max(worst, float(d.max()))
            exact += int((d == 0).sum()); n += d.numel()
    print(f"moshi LMGen: {exact}/{n} decisions are the oracle's exact argmax; worst deficit {worst:.3f}")
    assert worst <= 0.1 and exact >= 0.8 * n
    # reset restarts the generator
    with gen.streaming(B):
        a = [gen.step(inputs[t].to(DEV)) for t in range(3)]
        gen.reset_streaming()
        b = [gen.step(inputs[t].to(DEV)) for t in range(3)]
    assert a[0] is None and b[0] is None and torch.equal(a[2], b[2])
    # sampling mode runs (moshi's default top_k 250 goes through the threshold-select sampler)
    gs = LMGen(m)
    with gs.streaming(B):
        for t in range(3):
            o = gs.step(inputs[t].to(DEV))
    assert o.shape == (B, cfg.dep_q + 1, 1) and int(o.min()) >= 0
