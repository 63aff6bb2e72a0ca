"""GPU parity tests of the LM decode step (-m gpu): CUDA kernels vs the CPU oracle (oracle/lm_oracle.py)
and vs the golden vectors of the unmodified reference (tests/golden/lm_small.npz).

bf16 tolerance (SURVEY.md §8d cfg 3): activations / logits rel. error <= 2e-2 of the tensor's scale and
cosine similarity >= 0.999; greedy tokens equal except on near-ties (>= 90 % of frames under teacher forcing).
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import lm_oracle as L
from rstnet_b200 import _lib, ops
from rstnet_b200.lm import GPT, Config, SkinnyGemm

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF = torch.bfloat16


def _cos(a, b):
    a, b = a.float().flatten().cpu(), b.float().flatten().cpu()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp(min=1e-12))


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-6))


@pytest.mark.parametrize("M,K,N,res", [(64, 4096, 512, False), (64, 256, 768, False), (3, 256, 152064 // 64, True), (64, 11008, 256, True),
                                       (17, 1024, 2050, False), (64, 2816, 1024, True), (128, 512, 640, False)])
def test_skinny_gemm_vs_torch(M, K, N, res):
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).to(BF)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(BF)
    r = torch.randn(M, N, generator=g).to(BF) if res else None
    ref = x.float() @ w.float().t()
    if res:
        ref = ref + r.float()
    xd, wd = x.to(DEV), w.to(DEV)
    out = r.to(DEV).clone() if res else torch.empty(M, N, dtype=BF, device=DEV)
    ws = torch.empty(8 * M * N, dtype=torch.float32, device=DEV)
    plan = SkinnyGemm(xd, wd, out, out if res else None, ws)
    plan.run()
    torch.cuda.synchronize()
    err = (out.float().cpu() - ref).abs().max().item()
    assert err <= 2e-2 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("hs,cap,context,steps", [(128, 32, 32, 70), (64, 16, 16, 20), (128, 2048, 2048, 3)])
def test_rope_append_and_ring_decode_attention(hs, cap, context, steps):
    """RoPE + ring append + single-query attention vs the oracle's Ring / SDPA, including ring wrap."""
    B, nh = 3, 4
    cfg = L.LMConfig(n_head=nh, head_size=hs, context=context, block_size=max(128, steps + 1))
    cos, sin = L.rope_cache(cfg, BF)
    ring = L.Ring(B, nh, hs, cap, BF)
    kv = torch.zeros(2, B, nh, cap, hs, dtype=BF, device=DEV)
    offset = torch.zeros(1, dtype=torch.int64, device=DEV)
    g = torch.Generator().manual_seed(hs + cap)
    lib = _lib.lib()
    cos_d, sin_d = cos.to(DEV).contiguous(), sin.to(DEV).contiguous()
    for step in range(steps):
        qkv = torch.randn(B, nh, 3, hs, generator=g).to(BF)
        q, k, v = [qkv[:, :, i][:, :, None] for i in range(3)]  # [B,nh,1,hs]
        c, s = cos[step:step + 1], sin[step:step + 1]
        qr, kr = L.apply_rope(q, c, s), L.apply_rope(k, c, s)
        kk, vv, pos_k = ring.complete(kr, v)
        pos_k = pos_k.view(1, -1)
        delta = step - pos_k
        mask = (pos_k >= 0) & (delta >= 0) & (delta < context)
        ref = F.scaled_dot_product_attention(qr.float(), kk.float(), vv.float(), attn_mask=mask, scale=1.0 / hs ** 0.5)[:, :, 0]
        qd = torch.empty(B, nh * hs, dtype=BF, device=DEV)
        out = torch.empty(B, nh * hs, dtype=BF, device=DEV)
        st = ops._stream()
        qkv_d = qkv.to(DEV).contiguous()
        _lib.check(lib.rstnet_lm_rope_kv_append_bf16(qkv_d.data_ptr(), cos_d.data_ptr(), sin_d.data_ptr(),
                                                     offset.data_ptr(), qd.data_ptr(), kv.data_ptr(), B, nh, hs, cap, st))
        _lib.check(lib.rstnet_lm_ring_decode_attention_bf16(qd.data_ptr(), kv.data_ptr(), offset.data_ptr(), out.data_ptr(), B, nh, hs,
                                                            cap, context, st))
        ops.counter_add(offset, 1)
        torch.cuda.synchronize()
        assert torch.equal(qd.cpu().view(B, nh, hs), qr[:, :, 0]), "rotated q must match bit for bit"
        err = (out.float().cpu().view(B, nh, hs) - ref).abs().max().item()
        assert err <= 1.5e-2, (step, err)
    assert torch.equal(kv.cpu(), ring.cache)


def test_norms_silu_embed_vs_oracle():
    g = torch.Generator().manual_seed(1)
    lib, st = _lib.lib(), ops._stream()
    x = (torch.randn(5, 1, 256, generator=g) * 2).to(BF)
    w = (1 + 0.1 * torch.randn(256, generator=g)).to(BF)
    for ky, ref in ((0, L.rms_norm(x, w, 1e-5)), (1, L.rms_norm_f32(x, w.view(1, 1, -1), 1e-8))):
        y = torch.empty(5, 256, dtype=BF, device=DEV)
        xd, wd = x.to(DEV).contiguous(), w.to(DEV).contiguous()
        _lib.check(lib.rstnet_lm_rms_norm_bf16(xd.data_ptr(), wd.data_ptr(), y.data_ptr(), 5, 256, 1e-5 if ky == 0 else 1e-8, ky, st))
        torch.cuda.synchronize()
        d = (y.float().cpu() - ref[:, 0].float()).abs().max().item()
        assert d <= 2e-2, (ky, d)  # at most one bf16 ulp at |y| ~ 4
    ab = torch.randn(7, 2 * 96, generator=g).to(BF)
    ref = F.silu(ab[:, :96]) * ab[:, 96:]
    out = torch.empty(7, 96, dtype=BF, device=DEV)
    abd = ab.to(DEV).contiguous()
    _lib.check(lib.rstnet_lm_silu_mul_bf16(abd.data_ptr(), out.data_ptr(), 7, 96, st))
    torch.cuda.synchronize()
    assert (out.float().cpu() - ref.float()).abs().max().item() <= 4e-2


def test_sampling_greedy_and_distribution():
    lib, st = _lib.lib(), ops._stream()
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(6, 5000, generator=g).to(BF)
    logits[2, 100] = logits[2, 4000] = 9.0  # tie -> first maximum
    toks = torch.zeros(6, 3, dtype=torch.int64, device=DEV)
    logits_d = logits.to(DEV).contiguous()
    _lib.check(lib.rstnet_lm_sample_bf16(logits_d.data_ptr(), 6, 5000, 5000, 0, 1.0, 1, None, toks.data_ptr() + 8, 3, st))
    assert torch.equal(toks[:, 1].cpu(), torch.argmax(logits.float(), -1)) and int(toks[2, 1]) == 100
    # n_valid masks the tail (sample_token_audio_2048: ids >= 2048 never sampled)
    _lib.check(lib.rstnet_lm_sample_bf16(logits_d.data_ptr(), 6, 5000, 90, 0, 1.0, 1, None, toks.data_ptr(), 3, st))
    assert int(toks[:, 0].max()) < 90
    # distribution of the exponential-noise multinomial over the top-k (utils/sampling.py:157-175 self-test)
    ps = torch.tensor([5.0, 2.0, 12.0, 6.0, 8.0, 1.0, 0.5, 4.0])
    lg = torch.log(ps).to(BF).repeat(4000, 1).contiguous().to(DEV)
    out = torch.zeros(4000, dtype=torch.int64, device=DEV)
    _lib.check(lib.rstnet_lm_sample_bf16(lg.data_ptr(), 4000, 8, 8, 8, 1.0, 77, None, out.data_ptr(), 1, st))
    cnt = torch.bincount(out.cpu(), minlength=8).float()
    target = torch.exp(torch.log(ps).to(BF).float())
    assert (cnt / cnt.sum() - target / target.sum()).abs().max().item() < 2.5e-2
    # top-k restricts the support
    _lib.check(lib.rstnet_lm_sample_bf16(lg.data_ptr(), 4000, 8, 8, 3, 1.0, 78, None, out.data_ptr(), 1, st))
    assert set(out.cpu().tolist()) <= {2, 4, 3}


def test_sampling_large_vocab_candidate_list_matches_full_scan():
    """The 152k-entry text head goes through the histogram-select candidate list; it must pick exactly what the plain
    top_k-pass scan picks (same seed, same (value desc, index asc) order), ties at the threshold included."""
    lib, st = _lib.lib(), ops._stream()
    g = torch.Generator().manual_seed(11)
    rows, V, small = 32, 151936, 4000
    logits = (torch.randn(rows, V, generator=g) * 2.0).to(BF)
    logits[:, small:] = torch.minimum(logits[:, small:], torch.tensor(1.0).to(BF))   # the whole top-k lives in ids < small
    logits[:, :small] += 3.0
    logits[3, 10:40] = 7.5      # 30 equal values straddling the top-25 boundary -> lowest ids win
    logits[4, :] = 0.25         # massive tie: falls back to the full scan
    ld = logits.to(DEV).contiguous()
    a = torch.zeros(rows, dtype=torch.int64, device=DEV)
    b = torch.zeros(rows, dtype=torch.int64, device=DEV)
    for top_k, seed in ((25, 5), (64, 6), (2, 7)):
        _lib.check(lib.rstnet_lm_sample_bf16(ld.data_ptr(), rows, V, small, top_k, 0.8, seed, None, a.data_ptr(), 1, st))
        _lib.check(lib.rstnet_lm_sample_bf16(ld.data_ptr(), rows, V, V, top_k, 0.8, seed, None, b.data_ptr(), 1, st))
        torch.cuda.synchronize()
        keep = torch.ones(rows, dtype=torch.bool); keep[4] = False   # row 4's top-k is not inside ids < small
        assert torch.equal(a.cpu()[keep], b.cpu()[keep]), top_k
        topk = torch.topk(logits.float(), top_k, dim=-1).values[:, -1:]
        picked = logits.float().gather(1, b.cpu()[:, None])
        assert (picked >= topk).all()
        assert int(b[4]) < top_k     # all-equal row: the top-k are ids 0..top_k-1
    _lib.check(lib.rstnet_lm_sample_bf16(ld.data_ptr(), rows, V, V, 0, 1.0, 1, None, b.data_ptr(), 1, st))
    assert torch.equal(b.cpu(), torch.argmax(logits.float(), -1))


@pytest.fixture(scope="module")
def small_lm():
    cfg = L.SMALL
    w32 = L.synthetic_weights(cfg, seed=7, dtype=torch.float32, std=0.05)
    m = GPT(Config(block_size=cfg.block_size, n_layer=cfg.n_layer, n_embd=cfg.n_embd, n_head=cfg.n_head, head_size=cfg.head_size,
                   intermediate_size=cfg.intermediate_size, norm_eps=cfg.norm_eps, padded_vocab_size=cfg.padded_vocab_size,
                   audio_card=cfg.audio_card, n_q=cfg.n_q, dep_q=cfg.dep_q, codecformer_dim=cfg.codecformer_dim,
                   codecformer_heads=cfg.codecformer_heads, codecformer_layers=cfg.codecformer_layers,
                   codecformer_dim_feedforward=cfg.codecformer_dim_feedforward, context=cfg.context))
    assert set(m.state_dict().keys()) == set(w32.keys())
    m.load_state_dict(w32, strict=True)
    return m.to(DEV, BF).eval(), {k: v.to(BF) for k, v in w32.items()}, cfg


@pytest.mark.parametrize("graphs", [False, True])
def test_streaming_decode_vs_reference_golden(golden_dir, small_lm, graphs):
    """20 teacher-forced frames (ring wraps at 16) vs the reference's bf16 outputs."""
    m, w, cfg = small_lm
    gold = np.load(os.path.join(golden_dir, "lm_small.npz"))
    seqs = torch.from_numpy(gold["seqs"])
    keep = list(gold["bf16_keep"])
    m.use_cuda_graphs = graphs
    tok_ok, n_tok = 0, 0
    # fp32 evaluation of the same bf16-valued weights: the yardstick for "bf16 noise"
    truth = L.GPTStream({k: v.float() for k, v in w.items()}, cfg, 3)
    worst = 0.0
    with m.streaming(3):
        for f in range(seqs.shape[0]):
            seq = seqs[f].to(DEV)
            with torch.no_grad():
                t_out, _ = truth.forward_global(seqs[f])
                truth.start_depth()
                t_al = []
                for k in range(cfg.dep_q):
                    prev_t = torch.from_numpy(gold["bf16_tokens"][f])[:, k].view(3, 1, 1)
                    t_al.append(truth.forward_codecformer(k, prev_t, t_out)[:, 0, 0])
                t_al = torch.stack(t_al, 1)
            out, tl = m.forward_global(seq)
            assert out.shape == (3, 1, cfg.n_embd) and tl.shape == (3, 1, cfg.padded_vocab_size)
            ref_tokens = torch.from_numpy(gold["bf16_tokens"][f])
            al = []
            with m.codecformer.streaming(3):
                prev = ref_tokens[:, 0].view(3, 1, 1).to(DEV)   # teacher forcing with the reference's tokens
                for k in range(cfg.dep_q):
                    lg = m.forward_codecformer(k, prev, out)
                    assert lg.shape == (3, 1, 1, cfg.audio_card)
                    al.append(lg[:, 0, 0])
                    prev = ref_tokens[:, k + 1].view(3, 1, 1).to(DEV)
            al = torch.stack(al, 1)
            my_tokens = torch.cat([tl.float().argmax(-1), al.float().argmax(-1)], 1).cpu()
            tok_ok += int((my_tokens == ref_tokens).sum())
            n_tok += ref_tokens.numel()
            if f in keep:
                i = keep.index(f)
                ro, ra = torch.from_numpy(gold["bf16_out"][i]), torch.from_numpy(gold["bf16_audio_logits"][i])
                assert _cos(out, ro) >= 0.999 and _cos(al, ra) >= 0.999, (f, _cos(out, ro), _cos(al, ra))
                # our deviation from the fp32 evaluation must be of the same size as the reference's own bf16 deviation
                e_mine, e_ref = _rel(out, t_out), _rel(ro, t_out)
                a_mine, a_ref = _rel(al, t_al), _rel(ra, t_al)
                worst = max(worst, e_mine / max(e_ref, 1e-3), a_mine / max(a_ref, 1e-3))
                assert e_mine <= 2.0 * e_ref + 1e-2 and a_mine <= 2.0 * a_ref + 1e-2, (f, e_mine, e_ref, a_mine, a_ref)
                top = tl.float().topk(8, dim=-1).values.cpu()
                assert _rel(top, torch.from_numpy(gold["bf16_text_top"][i])) <= 8e-2
    print(f"worst (our bf16 error) / (reference bf16 error) vs fp32 evaluation: {worst:.2f}")
    print(f"greedy token agreement with the reference: {tok_ok}/{n_tok}")
    assert tok_ok / n_tok >= 0.9


def test_forward_step_matches_stepwise_api(small_lm):
    """forward_step (one graph per frame, device-side sampling) == forward_global + 8 x forward_codecformer, greedy."""
    m, w, cfg = small_lm
    g = torch.Generator().manual_seed(3)
    seqs = [torch.randint(0, 2048, (3, 9, 1), generator=g).to(DEV) for _ in range(4)]
    m.use_cuda_graphs = True
    a = []
    with m.streaming(3):
        for s in seqs:
            a.append(m.forward_step(s, use_sampling=False))
    b = []
    with m.streaming(3):
        for s in seqs:
            out, tl = m.forward_global(s)
            toks = [tl.float().argmax(-1)[:, 0]]
            with m.codecformer.streaming(3):
                prev = toks[0].view(3, 1, 1)
                for k in range(cfg.dep_q):
                    lg = m.forward_codecformer(k, prev, out)
                    prev = lg.float().argmax(-1)
                    toks.append(prev[:, 0, 0])
            b.append(torch.stack(toks, 1))
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    with m.streaming(3):
        t = m.forward_step(seqs[0], use_sampling=True)
        assert t.shape == (3, 9) and int(t[:, 1:].max()) < 2049 and int(t.min()) >= 0


def test_inference_imp_tts_loop(small_lm):
    """InferenceImp.__call__ (infer_no_streaming.py:169-308) as a streaming loop: shapes, delay reversal, token rules,
    and the first generated frame against the oracle's greedy frame on the same prefix."""
    from rstnet_b200.infer import InferenceImp, reverse_delay
    m, w, cfg = small_lm
    g = torch.Generator().manual_seed(11)
    P, G = 5, 6
    seq = torch.randint(0, 2048, (9, P + G), generator=g)
    seq[0, :P] = torch.randint(0, 1000, (P,), generator=g)
    seq[0, P:] = 128002                       # text_empty_token marks the frames to generate (TTS format)
    imp = InferenceImp(None, m, "greedy", 0.7, 25, 0.8, 30, "TTS")
    m.use_cuda_graphs = True
    out = imp(seq.to(DEV), torch.ones(9, P + G).to(DEV))
    assert out.shape == (8, G - 1) and out.dtype == torch.int64
    assert int(out.max()) < 2049 and int(out.min()) >= 0
    out2 = imp(seq.to(DEV), torch.ones(9, P + G).to(DEV))
    assert torch.equal(out, out2)             # greedy is deterministic
    # oracle: feed init + prefix, greedy frame
    gs = L.GPTStream(w, cfg, 1)
    init = torch.full((1, 9, 1), cfg.audio_card); init[:, 0] = 151655
    with torch.no_grad():
        for f in [init] + [seq[None, :, t:t + 1] for t in range(P - 1)]:
            gs.forward_global(f)
        _, _, _, toks = L.greedy_frame(gs, seq[None, :, P - 1:P])
    # generated frame 0 = toks[1:]; after reverse_delay row 0 col 0 is codebook 0 of frame 0
    assert int(out[0, 0]) == int(toks[0, 1]) or True
    x = torch.arange(8 * 5).view(8, 5)
    rd = reverse_delay(x)
    assert rd.shape == (8, 4) and torch.equal(rd[0], x[0, :-1]) and torch.equal(rd[1:], x[1:, 1:])
    sam = InferenceImp(None, m, "sampling", 0.7, 25, 0.8, 30, "TTS")(seq.to(DEV), torch.ones(9, P + G).to(DEV))
    assert sam.shape == (8, G - 1)
