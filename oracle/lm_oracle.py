"""CPU oracle: functional restatement of the speech-text LM's streaming decode step.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): never imported by rstnet_b200/.

Reference: MLLM_v2/models/llama_streaming.py (GPT, Block, CausalSelfAttention, LLaMAMLP),
MLLM_v2/models/lit_model.py (RMSNorm, build_rope_cache, apply_rope, RingKVCache),
MLLM_v2/modules/transformer.py (depth "codecformer": StreamingTransformer with per-step weights),
MLLM_v2/modules/gating.py, MLLM_v2/utils/sampling.py.  Pinned bit-for-bit (fp32 and bf16, CPU)
against the unmodified reference by oracle/gen_golden_lm.py.  The functions are dtype-agnostic:
weights in fp32 or bf16 give the reference's arithmetic in that dtype.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

W = Dict[str, torch.Tensor]


@dataclass(frozen=True)
class LMConfig:
    """Fields of models.llama_streaming.Config that the decode step reads (config.py:20-60,
    llama_streaming.py:447-486)."""
    n_layer: int = 32
    n_embd: int = 4096
    n_head: int = 32
    head_size: int = 128
    intermediate_size: int = 11008
    norm_eps: float = 1e-5
    rope_base: int = 10000
    block_size: int = 4096
    padded_vocab_size: int = 152064   # >= 151656: text_initial_token_id is hard-coded 151655 (llama_streaming.py:598-604)
    audio_card: int = 2050
    n_q: int = 8
    dep_q: int = 8
    codecformer_dim: int = 1024
    codecformer_heads: int = 16
    codecformer_layers: int = 6
    codecformer_dim_feedforward: int = 4224
    context: int = 3000
    n_query_groups: Optional[int] = None        # None -> n_head (MHA); config.py:104-109
    rotary_percentage: float = 1.0              # rope_n_elem = int(rotary_percentage * head_size), config.py:113
    rope_condense_ratio: int = 1
    rope_adjustments: Optional[dict] = None     # Llama 3.1 frequency scaling (lit_model.py:110-144)

    @property
    def n_kv(self) -> int:
        return self.n_head if self.n_query_groups is None else self.n_query_groups

    @property
    def rope_n_elem(self) -> int:
        return int(self.rotary_percentage * self.head_size)

    @property
    def ff_hidden(self) -> int:
        """ActivationGating hidden size (modules/gating.py:40-43)."""
        d, ff = self.codecformer_dim, self.codecformer_dim_feedforward
        return (21 * d) // 8 if ff == 4 * d else (2 * ff) // 3

    def reference_kwargs(self) -> dict:
        return dict(block_size=self.block_size, n_layer=self.n_layer, n_embd=self.n_embd, n_head=self.n_head,
                    n_query_groups=self.n_kv, head_size=self.head_size, intermediate_size=self.intermediate_size,
                    norm_class_name="RMSNorm", norm_eps=self.norm_eps, rotary_percentage=self.rotary_percentage,
                    rope_base=self.rope_base, rope_condense_ratio=self.rope_condense_ratio, rope_adjustments=self.rope_adjustments,
                    parallel_residual=False, bias=False, mlp_class_name="LLaMAMLP", padded_vocab_size=self.padded_vocab_size,
                    audio_card=self.audio_card, n_q=self.n_q, dep_q=self.dep_q, codecformer_dim=self.codecformer_dim,
                    codecformer_heads=self.codecformer_heads, codecformer_layers=self.codecformer_layers,
                    codecformer_dim_feedforward=self.codecformer_dim_feedforward, context=self.context)


LM7B = LMConfig()
# small config with the 7B topology (same vocab: the hard-coded initial text token needs it)
SMALL = LMConfig(n_layer=2, n_embd=256, n_head=4, head_size=64, intermediate_size=512, block_size=64, codecformer_dim=128,
                 codecformer_heads=4, codecformer_layers=2, codecformer_dim_feedforward=192, context=16)


def param_spec(cfg: LMConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """state_dict of llama_streaming.GPT (lora_r == 0): (name, shape, kind)."""
    E, V, I, D = cfg.n_embd, cfg.padded_vocab_size, cfg.intermediate_size, cfg.codecformer_dim
    hs, nh = cfg.head_size, cfg.n_head
    spec = [("lm_head.linear.weight", (V, E), "w"), ("transformer.wte.weight", (V, E), "w")]
    for l in range(cfg.n_layer):
        p = f"transformer.h.{l}"
        spec += [(f"{p}.norm_1.weight", (E,), "norm"), (f"{p}.attn.attn.linear.weight", ((nh + 2 * cfg.n_kv) * hs, E), "w"),
                 (f"{p}.attn.proj.linear.weight", (E, nh * hs), "w"), (f"{p}.norm_2.weight", (E,), "norm"),
                 (f"{p}.mlp.fc_1.linear.weight", (I, E), "w"), (f"{p}.mlp.fc_2.linear.weight", (I, E), "w"),
                 (f"{p}.mlp.proj.linear.weight", (E, I), "w")]
    spec.append(("transformer.ln_f.weight", (E,), "norm"))
    spec += [(f"input_emb.{i}.weight", (cfg.audio_card + 1, E), "w") for i in range(cfg.n_q)]
    spec += [(f"codecformer_in.{i}.weight", (D, E), "w") for i in range(cfg.dep_q)]
    spec += [(f"codecformer_emb.{i}.weight", (cfg.audio_card + 1, D), "w") for i in range(cfg.dep_q - 1)]
    spec.append(("codecformer_text_emb.weight", (V, D), "w"))
    H = cfg.ff_hidden
    for l in range(cfg.codecformer_layers):
        p = f"codecformer.layers.{l}"
        spec += [(f"{p}.self_attn.in_proj_weight", (cfg.dep_q * 3 * D, D), "w"),
                 (f"{p}.self_attn.out_proj.weight", (cfg.dep_q * D, D), "w"),
                 (f"{p}.norm1.alpha", (1, 1, D), "norm"), (f"{p}.norm2.alpha", (1, 1, D), "norm")]
        for k in range(cfg.dep_q):
            spec += [(f"{p}.gating.{k}.linear_in.weight", (2 * H, D), "w"), (f"{p}.gating.{k}.linear_out.weight", (D, H), "w")]
    spec += [(f"audio_linears.{i}.weight", (cfg.audio_card, D), "w") for i in range(cfg.dep_q)]
    return spec


def synthetic_weights(cfg: LMConfig, seed: int = 7, dtype=torch.float32, std: float = 0.02, device="cpu") -> W:
    """normal(0, std) for matrices / embeddings (lit_model.py:65-72 uses 0.02), 1 + 0.1 N(0,1) for norms."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    w: W = {}
    for name, shape, kind in param_spec(cfg):
        if kind == "w":
            # keep tiny models lively: scale std so that activations stay O(1) through the stack
            t = torch.empty(shape, dtype=torch.float32).normal_(0.0, std, generator=g)
        else:
            t = torch.empty(shape, dtype=torch.float32).normal_(0.0, 0.1, generator=g).add_(1.0)
        w[name] = t.to(dtype=dtype, device=device)
    return w


def synthetic_lora(cfg: LMConfig, seed: int, r: int, enable=(True, True, True)) -> W:
    """Seeded lora_A / lora_B for every LoRA-wrapped linear of llama_streaming.GPT with lora_r = r on query/key/value
    (per `enable`), projection, mlp and head (shapes: llama_streaming.py:95-97, 203-215)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    E, V, I, hs, nh, nkv = cfg.n_embd, cfg.padded_vocab_size, cfg.intermediate_size, cfg.head_size, cfg.n_head, cfg.n_kv
    rnd = lambda *shape: torch.empty(shape).normal_(0.0, 0.05, generator=g)
    out: W = {"lm_head.lora_A": rnd(r, E), "lm_head.lora_B": rnd(V, r)}
    qkv_rows = sum(s for s, e in zip((hs * nh, hs * nkv, hs * nkv), enable) if e)
    for l in range(cfg.n_layer):
        p = f"transformer.h.{l}"
        out[f"{p}.attn.attn.lora_A"] = rnd(r * sum(enable), E)
        out[f"{p}.attn.attn.lora_B"] = rnd(qkv_rows, r)
        out[f"{p}.attn.proj.lora_A"], out[f"{p}.attn.proj.lora_B"] = rnd(r, nh * hs), rnd(E, r)
        for n, (o, i) in (("fc_1", (I, E)), ("fc_2", (I, E)), ("proj", (E, I))):
            out[f"{p}.mlp.{n}.lora_A"], out[f"{p}.mlp.{n}.lora_B"] = rnd(r, i), rnd(o, r)
    return out


# --------------------------------------------------------------------------- pieces
def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """lit_model.RMSNorm.forward (lit_model.py:707-714): fp32 inside, cast back."""
    dtype = x.dtype
    x = x.float()
    norm_x = torch.mean(x * x, dim=-1, keepdim=True)
    return (x * torch.rsqrt(norm_x + eps) * weight.float()).to(dtype)


def rms_norm_f32(x: torch.Tensor, alpha: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """Kyutai _rms_norm with dtype=float (modules/transformer.py:34-48, create_norm_fn 'rms_norm_f32')."""
    x_dtype = x.dtype
    xf = x.to(torch.float)
    var = eps + torch.mean(xf ** 2, dim=2, keepdim=True)
    return (xf * (alpha.to(var) * torch.rsqrt(var))).to(x_dtype)


def rope_cache(cfg: LMConfig, dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """build_rope_cache (lit_model.py:441-488) over rope_n_elem dims incl. the Llama-3.1 rope_adjustments; the buffers
    follow the model dtype (GPT(...).to(bfloat16) casts them)."""
    n = cfg.rope_n_elem
    theta = 1.0 / (cfg.rope_base ** (torch.arange(0, n, 2).float() / n))
    if cfg.rope_adjustments is not None:
        ec = cfg.rope_adjustments
        wavelen = 2 * torch.pi / theta
        ratio = ec["original_max_seq_len"] / wavelen
        smooth = torch.clamp((ratio - ec["low_freq_factor"]) / (ec["high_freq_factor"] - ec["low_freq_factor"]), min=0.0, max=1.0)
        theta = (1 - smooth) * (theta / ec["factor"]) + smooth * theta
    idx_theta = torch.outer(torch.arange(cfg.block_size) / cfg.rope_condense_ratio, theta).repeat(1, 2)
    return torch.cos(idx_theta).to(dtype), torch.sin(idx_theta).to(dtype)


def rope_partial(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, n: int) -> torch.Tensor:
    """llama_streaming.py:979-982: the first rope_n_elem dims rotate, the rest pass through."""
    return torch.cat((apply_rope(x[..., :n], cos, sin), x[..., n:]), dim=-1)


def split_qkv(qkv: torch.Tensor, cfg: "LMConfig") -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """llama_streaming.py:952-970: per-group interleave [q..q, k, v]; k/v expanded to n_head heads for GQA."""
    B, T, _ = qkv.shape
    nh, nkv, hs = cfg.n_head, cfg.n_kv, cfg.head_size
    q_per_kv = nh // nkv
    qkv = qkv.view(B, T, nkv, q_per_kv + 2, hs).permute(0, 2, 3, 1, 4)
    q, k, v = qkv.split((q_per_kv, 1, 1), dim=2)
    if nkv != nh and nkv != 1:
        k = k.expand(B, nkv, q_per_kv, T, hs)
        v = v.expand(B, nkv, q_per_kv, T, hs)
    return q.reshape(B, -1, T, hs), k.reshape(B, -1, T, hs), v.reshape(B, -1, T, hs)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """lit_model.apply_rope (lit_model.py:560-573), rotate-half."""
    hs = x.size(-1)
    x1, x2 = x[..., : hs // 2], x[..., hs // 2:]
    rotated = torch.cat((-x2, x1), dim=-1)
    return ((x * cos) + (rotated * sin)).to(dtype=x.dtype)


class Ring:
    """lit_model.RingKVCache (lit_model.py:589-658) with its `delta <= 0` labelling."""

    def __init__(self, B, H, D, capacity, dtype, device="cpu"):
        self.capacity = capacity
        self.cache = torch.zeros(2, B, H, capacity, D, dtype=dtype, device=device)
        self.end_offset = 0

    def complete(self, k, v):
        T = k.shape[2]
        dev = self.cache.device
        idx = (torch.arange(T, device=dev) + self.end_offset) % self.capacity
        self.cache[0].index_copy_(2, idx, k)
        self.cache[1].index_copy_(2, idx, v)
        self.end_offset += T
        slots = torch.arange(self.capacity, device=dev)
        end_index = self.end_offset % self.capacity
        delta = slots - end_index
        pos = torch.where(delta <= 0, self.end_offset + delta, self.end_offset + delta - self.capacity)
        pos = torch.where(slots >= self.end_offset, torch.full_like(pos, -1), pos)
        return self.cache[0], self.cache[1], pos


def embed_sum(seq: torch.Tensor, w: W, cfg: LMConfig) -> torch.Tensor:
    """GPT.forward_global embedding part (llama_streaming.py:680-687): ScaledEmbedding rows (zero for id -1,
    :505-517) of the n_q audio streams summed in order, then + wte(text)."""
    x = None
    for cb in range(cfg.n_q):
        ids = seq[:, cb + 1]
        e = F.embedding(ids.clamp(min=0), w[f"input_emb.{cb}.weight"])
        e = torch.where((ids == -1)[..., None], torch.zeros(1, dtype=e.dtype, device=e.device), e)
        x = e if x is None else x + e
    return x + F.embedding(seq[:, 0], w["transformer.wte.weight"])


class GPTStream:
    """`with gpt.streaming(B):` decode state + the two step functions."""

    def __init__(self, w: W, cfg: LMConfig, B: int):
        self.w, self.cfg, self.B = w, cfg, B
        self.dtype = w["transformer.wte.weight"].dtype
        self.device = w["transformer.wte.weight"].device    # the same ATen calls on a GPU = "the reference eager on that GPU"
        self.cos, self.sin = [t.to(self.device) for t in rope_cache(cfg, self.dtype)]
        self.rings = [Ring(B, cfg.n_head, cfg.head_size, cfg.context, self.dtype, self.device) for _ in range(cfg.n_layer)]
        self.offset = 0
        self.dep_rings: Optional[List[Ring]] = None
        self.dep_offset = 0

    # ---- temporal transformer -------------------------------------------------------------
    def attn(self, x: torch.Tensor, l: int) -> torch.Tensor:
        """CausalSelfAttention.forward, streaming, T == 1 (llama_streaming.py:935-998)."""
        cfg, w = self.cfg, self.w
        B, T, _ = x.shape
        nh, hs = cfg.n_head, cfg.head_size
        q, k, v = split_qkv(F.linear(x, w[f"transformer.h.{l}.attn.attn.linear.weight"]), cfg)
        cos = self.cos.index_select(0, torch.tensor([self.offset], device=self.device))
        sin = self.sin.index_select(0, torch.tensor([self.offset], device=self.device))
        q, k = rope_partial(q, cos, sin, cfg.rope_n_elem), rope_partial(k, cos, sin, cfg.rope_n_elem)
        kk, vv, pos_k = self.rings[l].complete(k, v)
        pos_k = pos_k.view(1, -1)
        delta = (self.offset + torch.arange(T, device=self.device).view(-1, 1)) - pos_k
        mask = (pos_k >= 0) & (delta >= 0) & (delta < cfg.context)
        y = F.scaled_dot_product_attention(q, kk, vv, attn_mask=mask, dropout_p=0.0, scale=1.0 / math.sqrt(hs))
        y = y.transpose(1, 2).reshape(B, T, hs * nh)
        return F.linear(y, w[f"transformer.h.{l}.attn.proj.linear.weight"])

    def forward_global(self, seq: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """GPT.forward_global on one frame seq[B, n_q+1, 1] (llama_streaming.py:665-692) through
        LLAMAStreamingTransformer.forward / Block.forward (:788-800, :834-853)."""
        cfg, w = self.cfg, self.w
        x = embed_sum(seq, w, cfg)
        for l in range(cfg.n_layer):
            p = f"transformer.h.{l}"
            x = self.attn(rms_norm(x, w[f"{p}.norm_1.weight"], cfg.norm_eps), l) + x
            h = rms_norm(x, w[f"{p}.norm_2.weight"], cfg.norm_eps)
            h = F.linear(F.silu(F.linear(h, w[f"{p}.mlp.fc_1.linear.weight"])) * F.linear(h, w[f"{p}.mlp.fc_2.linear.weight"]),
                         w[f"{p}.mlp.proj.linear.weight"])  # LLaMAMLP (lit_model.py:399-403)
            x = h + x
        x = rms_norm(x, w["transformer.ln_f.weight"], cfg.norm_eps)
        self.offset += seq.shape[-1]
        return x, F.linear(x, w["lm_head.linear.weight"])

    # ---- depth transformer ----------------------------------------------------------------
    def start_depth(self):
        """`with gpt.codecformer.streaming(B):` (fresh per frame, llama_streaming.py:581)."""
        cfg = self.cfg
        hd = cfg.codecformer_dim // cfg.codecformer_heads
        self.dep_rings = [Ring(self.B, cfg.codecformer_heads, hd, cfg.dep_q, self.dtype, self.device) for _ in range(cfg.codecformer_layers)]
        self.dep_offset = 0

    def forward_codecformer(self, k: int, prev: torch.Tensor, transformer_out: torch.Tensor) -> torch.Tensor:
        """GPT.forward_codecformer (llama_streaming.py:727-749): prev [B,1,1] -> logits [B,1,1,card]."""
        cfg, w = self.cfg, self.w
        assert k == self.dep_offset
        D, H = cfg.codecformer_dim, cfg.codecformer_heads
        x = F.linear(transformer_out, w[f"codecformer_in.{k}.weight"])
        ids = prev[:, 0]
        table = w["codecformer_text_emb.weight"] if k == 0 else w[f"codecformer_emb.{k - 1}.weight"]
        e = F.embedding(ids.clamp(min=0), table)
        e = torch.where((ids == -1)[..., None], torch.zeros(1, dtype=e.dtype, device=e.device), e)
        x = x + e
        B, T, _ = x.shape
        for l in range(cfg.codecformer_layers):
            p = f"codecformer.layers.{l}"
            # StreamingTransformerLayer._sa_block with weights_per_step (modules/transformer.py:375-419, 568-577)
            h = rms_norm_f32(x, w[f"{p}.norm1.alpha"])
            proj = F.linear(h[:, 0], w[f"{p}.self_attn.in_proj_weight"].view(cfg.dep_q, -1, D)[k])[:, None]
            q, kk, v = proj.view(B, T, 3, H, D // H).permute(2, 0, 3, 1, 4)
            kk, v, pos_k = self.dep_rings[l].complete(kk, v)
            pos_k = pos_k.view(1, -1)
            delta = (self.dep_offset + torch.arange(T, device=self.device).view(-1, 1)) - pos_k
            bias = (pos_k >= 0) & (delta >= 0)
            a = F.scaled_dot_product_attention(q, kk, v, bias, dropout_p=0.0).permute(0, 2, 1, 3).reshape(B, T, D)
            a = F.linear(a[:, 0], w[f"{p}.self_attn.out_proj.weight"].view(cfg.dep_q, -1, D)[k])[:, None]
            x = x + a
            # _ff_block with per-step ActivationGating (transformer.py:550-567; gating.py:12-21)
            h = rms_norm_f32(x, w[f"{p}.norm2.alpha"])
            g = F.linear(h, w[f"{p}.gating.{k}.linear_in.weight"]).view(B, T, 2, -1)
            g = F.linear(F.silu(g[..., 0, :]) * g[..., 1, :], w[f"{p}.gating.{k}.linear_out.weight"])
            x = x + g
        self.dep_offset += 1
        return F.linear(x, w[f"audio_linears.{k}.weight"])[:, None]


def greedy_frame(gs: GPTStream, seq: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """One generated frame with argmax sampling (sample_token(use_sampling=False), utils/sampling.py:85-105):
    temporal step -> text token -> 8 depth steps.  Returns (transformer_out, text_logits, audio_logits[B,8,card], tokens[B,9])."""
    cfg = gs.cfg
    out, text_logits = gs.forward_global(seq)
    text_tok = torch.argmax(text_logits.float(), dim=-1)  # [B,1]
    toks = [text_tok[:, 0]]
    prev = text_tok[:, :, None]
    gs.start_depth()
    alog = []
    for k in range(cfg.dep_q):
        lg = gs.forward_codecformer(k, prev, out)
        alog.append(lg[:, 0, 0])
        nxt = torch.argmax(lg.float(), dim=-1)  # [B,1,1]
        toks.append(nxt[:, 0, 0])
        prev = nxt
    return out, text_logits, torch.stack(alog, 1), torch.stack(toks, 1)


# ----------------------------------------------------------------------------- non-streaming forms
def embed_sum_T(seq: torch.Tensor, w: W, cfg: LMConfig) -> torch.Tensor:
    """Embedding part of GPT.forward_global on seq[B, n_q+1, T] (llama_streaming.py:680-687)."""
    x = None
    for cb in range(cfg.n_q):
        ids = seq[:, cb + 1]
        e = F.embedding(ids.clamp(min=0), w[f"input_emb.{cb}.weight"])
        e = torch.where((ids == -1)[..., None], torch.zeros(1, dtype=e.dtype), e)
        x = e if x is None else x + e
    return x + F.embedding(seq[:, 0], w["transformer.wte.weight"])


def forward_global_full(w: W, cfg: LMConfig, seq: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """GPT.forward_global OUTSIDE a streaming scope (llama_streaming.py:665-692 with state None in
    CausalSelfAttention.forward, :935-998): positions 0..T-1, cos[:T]/sin[:T], KVCacheResult.from_kv (all T keys,
    lit_model.py:582-587), mask (delta >= 0) & (delta < context).  This is what infer_no_streaming.py:239 calls."""
    B, K, T = seq.shape
    hs = cfg.head_size
    dtype = w["transformer.wte.weight"].dtype
    cos, sin = rope_cache(cfg, dtype)
    cos, sin = cos[:T], sin[:T]
    x = embed_sum_T(seq, w, cfg)
    pos = torch.arange(T)
    delta = pos.view(-1, 1) - pos.view(1, -1)
    mask = (pos.view(1, -1) >= 0) & (delta >= 0) & (delta < cfg.context)
    for l in range(cfg.n_layer):
        p = f"transformer.h.{l}"
        h = rms_norm(x, w[f"{p}.norm_1.weight"], cfg.norm_eps)
        q, k, v = split_qkv(F.linear(h, w[f"{p}.attn.attn.linear.weight"]), cfg)
        q, k = rope_partial(q, cos, sin, cfg.rope_n_elem), rope_partial(k, cos, sin, cfg.rope_n_elem)
        y = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.0, scale=1.0 / math.sqrt(hs))
        y = y.transpose(1, 2).reshape(B, T, hs * cfg.n_head)
        x = F.linear(y, w[f"{p}.attn.proj.linear.weight"]) + x
        h = rms_norm(x, w[f"{p}.norm_2.weight"], cfg.norm_eps)
        h = F.linear(F.silu(F.linear(h, w[f"{p}.mlp.fc_1.linear.weight"])) * F.linear(h, w[f"{p}.mlp.fc_2.linear.weight"]),
                     w[f"{p}.mlp.proj.linear.weight"])
        x = h + x
    x = rms_norm(x, w["transformer.ln_f.weight"], cfg.norm_eps)
    return x, F.linear(x, w["lm_head.linear.weight"])


def scaled_embedding(ids: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    """ScaledEmbedding.forward (llama_streaming.py:505-517): exact zero row for id -1."""
    e = F.embedding(ids.clamp(min=0), table)
    return torch.where((ids == -1)[..., None], torch.zeros(1, dtype=e.dtype), e)


def forward_local(w: W, cfg: LMConfig, local_start_token: torch.Tensor, sequence: torch.Tensor,
                  transformer_out: torch.Tensor) -> torch.Tensor:
    """GPT.forward_local (llama_streaming.py:694-725): the depth transformer over B*T rows x dep_q steps, teacher-forced,
    non-streaming (multi_linear with offset 0, modules/transformer.py:155-179; all dep_q keys visible causally --
    KVCacheResult.from_kv has no ring, so step dep_q-1 sees key 0, unlike the streaming form).
    local_start_token [B,T,D] features, sequence [B,dep_q,T] tokens, transformer_out [B,T,E] -> logits [B,T,dep_q,card]."""
    B, K, S = sequence.shape
    D, H, Q = cfg.codecformer_dim, cfg.codecformer_heads, cfg.dep_q
    local_inputs = [local_start_token.reshape(-1, local_start_token.shape[-1])]
    for cb in range(Q - 1):
        e = scaled_embedding(sequence[:, cb:cb + 1, :], w[f"codecformer_emb.{cb}.weight"])
        local_inputs.append(e.reshape(-1, e.shape[-1]))
    views = []
    for cb in range(Q):
        t = F.linear(transformer_out, w[f"codecformer_in.{cb}.weight"])
        t = t.reshape(-1, t.shape[-1])
        views.append((t + local_inputs[cb]).unsqueeze(1))
    x = torch.cat(views, dim=1)                      # [B*T, Q, D]
    N = x.shape[0]
    pos = torch.arange(Q)
    bias = (pos.view(1, -1) >= 0) & ((pos.view(-1, 1) - pos.view(1, -1)) >= 0)
    for l in range(cfg.codecformer_layers):
        p = f"codecformer.layers.{l}"
        h = rms_norm_f32(x, w[f"{p}.norm1.alpha"])
        w_in = w[f"{p}.self_attn.in_proj_weight"].view(Q, -1, D)
        proj = torch.stack([F.linear(h[:, t], w_in[t]) for t in range(Q)], 1)
        q, kk, v = proj.view(N, Q, 3, H, D // H).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, kk, v, bias, dropout_p=0.0).permute(0, 2, 1, 3).reshape(N, Q, D)
        w_out = w[f"{p}.self_attn.out_proj.weight"].view(Q, -1, D)
        x = x + torch.stack([F.linear(a[:, t], w_out[t]) for t in range(Q)], 1)
        h = rms_norm_f32(x, w[f"{p}.norm2.alpha"])
        ys = []
        for t in range(Q):
            g = F.linear(h[:, t:t + 1], w[f"{p}.gating.{t}.linear_in.weight"]).view(N, 1, 2, -1)
            ys.append(F.linear(F.silu(g[..., 0, :]) * g[..., 1, :], w[f"{p}.gating.{t}.linear_out.weight"]))
        x = x + torch.cat(ys, dim=1)
    logits = []
    for cb in range(Q):
        lg = F.linear(x[:, cb:cb + 1, :], w[f"audio_linears.{cb}.weight"])
        logits.append(lg.reshape(B, -1, 1, lg.shape[-1]))
    return torch.cat(logits, dim=2)
