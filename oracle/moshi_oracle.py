"""CPU oracle: restatement of the Moshi-style LMModel streaming step and LMGen.step.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): never imported by rstnet_b200/.

Reference: MLLM_v2/models/model.py (LMModel.forward_text :364-389, forward_depformer :392-428, LMGen.step :490-562,
depformer_step :564-597) over the Kyutai StreamingTransformer (modules/transformer.py:375-419, 550-588, 669-691),
apply_rope (modules/rope.py:11-68), ActivationGating (modules/gating.py:12-21).  Greedy decoding only (use_sampling False).
Pinned bit for bit against the unmodified reference by oracle/gen_golden_moshi.py (fp32 and bf16).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import lm_oracle as L

W = Dict[str, torch.Tensor]


@dataclass(frozen=True)
class MoshiConfig:
    dim: int = 256
    num_heads: int = 4
    num_layers: int = 2
    hidden_scale: float = 4.125
    n_q: int = 16
    dep_q: int = 8
    card: int = 2048
    text_card: int = 32000
    existing_text_padding_id: Optional[int] = 3
    context: int = 16
    max_period: float = 10000.0
    depformer_dim: int = 128
    depformer_dim_feedforward: int = 528
    depformer_num_heads: int = 4
    depformer_num_layers: int = 2
    delays: Tuple[int, ...] = (0, 0, 1, 1, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 1)

    @property
    def hidden(self) -> int:
        ff = int(self.hidden_scale * self.dim)
        return (21 * self.dim) // 8 if ff == 4 * self.dim else (2 * ff) // 3

    @property
    def dep_hidden(self) -> int:
        d, ff = self.depformer_dim, self.depformer_dim_feedforward
        return (21 * d) // 8 if ff == 4 * d else (2 * ff) // 3

    def reference_kwargs(self) -> dict:
        return dict(dim=self.dim, text_card=self.text_card, existing_text_padding_id=self.existing_text_padding_id, n_q=self.n_q,
                    dep_q=self.dep_q, card=self.card, num_heads=self.num_heads, num_layers=self.num_layers, hidden_scale=self.hidden_scale,
                    causal=True, layer_scale=None, context=self.context, max_period=self.max_period, gating="silu", norm="rms_norm_f32",
                    positional_embedding="rope", depformer_dim=self.depformer_dim, depformer_dim_feedforward=self.depformer_dim_feedforward,
                    depformer_num_heads=self.depformer_num_heads, depformer_num_layers=self.depformer_num_layers, depformer_causal=True,
                    depformer_layer_scale=None, depformer_multi_linear=True, depformer_context=8, depformer_max_period=10000,
                    depformer_gating="silu", depformer_pos_emb="none", depformer_weights_per_step=True, delays=list(self.delays))


SMALL = MoshiConfig()


def param_spec(cfg: MoshiConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    d, D = cfg.dim, cfg.depformer_dim
    extra = 1 if cfg.existing_text_padding_id is None else 0
    spec = [(f"emb.{i}.weight", (cfg.card + 1, d), "w") for i in range(cfg.n_q)]
    spec += [("text_emb.weight", (cfg.text_card + 1, d), "w"), ("text_linear.weight", (cfg.text_card + extra, d), "w")]
    for l in range(cfg.num_layers):
        p = f"transformer.layers.{l}"
        spec += [(f"{p}.self_attn.in_proj_weight", (3 * d, d), "w"), (f"{p}.self_attn.out_proj.weight", (d, d), "w"),
                 (f"{p}.norm1.alpha", (1, 1, d), "norm"), (f"{p}.norm2.alpha", (1, 1, d), "norm"),
                 (f"{p}.gating.linear_in.weight", (2 * cfg.hidden, d), "w"), (f"{p}.gating.linear_out.weight", (d, cfg.hidden), "w")]
    spec.append(("out_norm.alpha", (1, 1, d), "norm"))
    spec += [(f"depformer_in.{i}.weight", (D, d), "w") for i in range(cfg.dep_q)]
    spec += [(f"depformer_emb.{i}.weight", (cfg.card + 1, D), "w") for i in range(cfg.dep_q - 1)]
    spec.append(("depformer_text_emb.weight", (cfg.text_card + 1, D), "w"))
    for l in range(cfg.depformer_num_layers):
        p = f"depformer.layers.{l}"
        spec += [(f"{p}.self_attn.in_proj_weight", (cfg.dep_q * 3 * D, D), "w"), (f"{p}.self_attn.out_proj.weight", (cfg.dep_q * D, D), "w"),
                 (f"{p}.norm1.alpha", (1, 1, D), "norm"), (f"{p}.norm2.alpha", (1, 1, D), "norm")]
        for k in range(cfg.dep_q):
            spec += [(f"{p}.gating.{k}.linear_in.weight", (2 * cfg.dep_hidden, D), "w"),
                     (f"{p}.gating.{k}.linear_out.weight", (D, cfg.dep_hidden), "w")]
    spec += [(f"linears.{i}.weight", (cfg.card, D), "w") for i in range(cfg.dep_q)]
    return spec


def synthetic_weights(cfg: MoshiConfig, seed: int = 5, std: float = 0.05) -> W:
    g = torch.Generator().manual_seed(seed)
    w: W = {}
    for name, shape, kind in param_spec(cfg):
        if kind == "w":
            w[name] = torch.empty(shape).normal_(0.0, std, generator=g)
        else:
            w[name] = torch.empty(shape).normal_(0.0, 0.1, generator=g).add_(1.0)
    return w


def rope_pairs(q: torch.Tensor, k: torch.Tensor, offset: int, max_period: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """apply_rope, time_before_heads=False (modules/rope.py:11-68): q, k [B,H,T,D]; fp32 inside, cast back."""
    B, H, T, D = q.shape
    ds = torch.arange(D // 2, dtype=torch.float32)
    freqs = torch.exp(ds * (-math.log(max_period) * 2 / D))
    ts = (torch.tensor([offset]).float() + torch.arange(T, dtype=torch.float32)).view(1, -1, 1)
    q2, k2 = q.view(B, H, T, D // 2, 2), k.view(B, H, T, D // 2, 2)
    qr, qi, kr, ki = q2[..., 0].float(), q2[..., 1].float(), k2[..., 0].float(), k2[..., 1].float()
    rotr, roti = torch.cos(freqs * ts), torch.sin(freqs * ts)
    dt = q.dtype
    qo = torch.stack([(qr * rotr - qi * roti).to(dt), (qr * roti + qi * rotr).to(dt)], dim=-1)
    ko = torch.stack([(kr * rotr - ki * roti).to(dt), (kr * roti + ki * rotr).to(dt)], dim=-1)
    return qo.view(B, H, T, D), ko.view(B, H, T, D)


class MoshiStream:
    """`with lm.streaming(B):` state + forward_text / forward_depformer on one frame."""

    def __init__(self, w: W, cfg: MoshiConfig, B: int):
        self.w, self.cfg, self.B = w, cfg, B
        self.dtype = w["text_emb.weight"].dtype
        hd = cfg.dim // cfg.num_heads
        self.rings = [L.Ring(B, cfg.num_heads, hd, cfg.context, self.dtype) for _ in range(cfg.num_layers)]
        self.offset = 0
        # the depth transformer is the GPT oracle's (same module upstream), under its weight names
        ren = {}
        for k, v in w.items():
            k2 = (k.replace("depformer_in.", "codecformer_in.").replace("depformer_emb.", "codecformer_emb.")
                  .replace("depformer_text_emb.", "codecformer_text_emb.").replace("depformer.layers.", "codecformer.layers.")
                  .replace("linears.", "audio_linears.") if not k.startswith("text_linear") else k)
            ren[k2] = v
        ren["transformer.wte.weight"] = w["text_emb.weight"]
        dcfg = L.LMConfig(n_layer=0, n_embd=cfg.dim, n_head=cfg.num_heads, head_size=hd, audio_card=cfg.card, n_q=cfg.n_q, dep_q=cfg.dep_q,
                          codecformer_dim=cfg.depformer_dim, codecformer_heads=cfg.depformer_num_heads,
                          codecformer_layers=cfg.depformer_num_layers, codecformer_dim_feedforward=cfg.depformer_dim_feedforward,
                          context=cfg.context, block_size=8)
        self.depth = L.GPTStream(ren, dcfg, B)

    def forward_text(self, seq: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """LMModel.forward_text on seq [B, K, 1] (models/model.py:364-389)."""
        cfg, w = self.cfg, self.w
        x = None
        for cb in range(cfg.n_q):
            e = L.scaled_embedding(seq[:, cb + 1], w[f"emb.{cb}.weight"])
            x = e if x is None else x + e
        x = x + L.scaled_embedding(seq[:, 0], w["text_emb.weight"])
        B, T, d = x.shape
        H = cfg.num_heads
        for l in range(cfg.num_layers):
            p = f"transformer.layers.{l}"
            h = L.rms_norm_f32(x, w[f"{p}.norm1.alpha"])
            proj = F.linear(h, w[f"{p}.self_attn.in_proj_weight"])
            q, k, v = proj.view(B, T, 3, H, d // H).permute(2, 0, 3, 1, 4)
            q, k = rope_pairs(q, k, self.offset, cfg.max_period)
            k, v, pos_k = self.rings[l].complete(k, v)
            pos_k = pos_k.view(1, -1)
            delta = (self.offset + torch.arange(T).view(-1, 1)) - pos_k
            bias = (pos_k >= 0) & (delta >= 0) & (delta < cfg.context)
            a = F.scaled_dot_product_attention(q, k, v, bias, dropout_p=0.0).permute(0, 2, 1, 3).reshape(B, T, d)
            x = x + F.linear(a, w[f"{p}.self_attn.out_proj.weight"])
            h = L.rms_norm_f32(x, w[f"{p}.norm2.alpha"])
            g = F.linear(h, w[f"{p}.gating.linear_in.weight"]).view(B, T, 2, -1)
            x = x + F.linear(F.silu(g[..., 0, :]) * g[..., 1, :], w[f"{p}.gating.linear_out.weight"])
        self.offset += T
        out = L.rms_norm_f32(x, w["out_norm.alpha"])
        return out, F.linear(out, w["text_linear.weight"])[:, None]

    def depformer_step(self, text_token: torch.Tensor, transformer_out: torch.Tensor,
                       force: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """LMGen.depformer_step, greedy (:564-597): tokens [B, dep_q], logits [B, dep_q, card].
        force [B, dep_q]: teacher-force these audio tokens (the logits are then those a candidate's decisions are judged on)."""
        self.depth.start_depth()
        prev, toks, lgs = text_token, [], []
        for k in range(self.cfg.dep_q):
            lg = self.depth.forward_codecformer(k, prev[:, None, None], transformer_out)
            nxt = torch.argmax(lg.float(), dim=-1)[:, 0, 0] if force is None else force[:, k]
            toks.append(nxt); lgs.append(lg[:, 0, 0])
            prev = nxt
        return torch.stack(toks, 1), torch.stack(lgs, 1)


class LMGenOracle:
    """LMGen.step with use_sampling False (models/model.py:490-562)."""

    def __init__(self, w: W, cfg: MoshiConfig, B: int):
        self.cfg, self.B = cfg, B
        self.lm = MoshiStream(w, cfg, B)
        self.max_delay = max(cfg.delays)
        K = cfg.n_q + 1
        self.cache = torch.full((B, K, self.max_delay + 2), -2, dtype=torch.long)
        self.initial = torch.full((1, K, 1), cfg.card, dtype=torch.long)
        self.initial[:, 0] = cfg.text_card
        self.offset = 0
        self.last = None

    def step(self, input_tokens: torch.Tensor, force: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """force [B, dep_q + 1]: the (text, audio) tokens to take at this step instead of the oracle's own argmaxes."""
        cfg = self.cfg
        CT = self.cache.shape[2]
        for q_other in range(input_tokens.shape[1]):
            k = cfg.dep_q + 1 + q_other
            wp = (self.offset + cfg.delays[k]) % CT
            self.cache[:, k, wp:wp + 1] = input_tokens[:, q_other]
        position = self.offset % CT
        for k, delay in enumerate(cfg.delays):
            if self.offset <= delay:
                self.cache[:, k, position] = self.initial[:, k, 0]
        input_ = self.cache[:, :, position:position + 1]
        out, text_logits = self.lm.forward_text(input_)
        text_token = torch.argmax(text_logits.float(), dim=-1)[:, 0, 0] if force is None else force[:, 0]
        audio, alog = self.lm.depformer_step(text_token, out, None if force is None else force[:, 1:])
        self.last = (input_.clone(), out, text_logits, alog)
        self.offset += 1
        position = self.offset % CT
        self.cache[:, 0, position] = text_token
        self.cache[:, 1:cfg.dep_q + 1, position] = audio
        if self.offset <= self.max_delay:
            return None
        gd = torch.tensor(cfg.delays[:cfg.dep_q + 1])
        index = ((self.offset - self.max_delay + gd) % CT).view(1, -1, 1).expand(self.B, -1, 1)
        return self.cache.gather(dim=2, index=index)
