"""CPU oracle: functional restatement of the reference Mimi codec (torch CPU fp32).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): never imported by rstnet_b200/.

Reference root: /root/reference/MLLM_v2/tools/tokenizer/MimiCodec/model  (abbreviated R/ below;
byte-identical twins live in MLLM_v2/modules, MLLM_v2/moshi/modules, AudioCodec/MimiCodec/modules).
Pinned against the unmodified reference by oracle/gen_golden.py (bit-for-bit on CPU).

Layouts follow the reference: activations [B, C, T] fp32, codes [B, K, T] int64.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .mimi_spec import MimiConfig, OFFICIAL, codebook_prefixes

W = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------- convs
def _extra_padding(length: int, k_eff: int, stride: int, padding_total: int) -> int:
    """R/modules/conv.py:50-58 get_extra_padding_for_conv1d."""
    n_frames = (length - k_eff + padding_total) / stride + 1
    ideal = (math.ceil(n_frames) - 1) * stride + (k_eff - padding_total)
    return ideal - length


def causal_conv1d(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int = 1,
                  pad_mode: str = "constant") -> torch.Tensor:
    """Non-streaming causal StreamingConv1d.forward (R/modules/conv.py:232-254): left pad
    k_eff - stride, right pad so the last window is full, then nn.Conv1d (dilation 1 everywhere in
    Mimi because n_residual_layers == 1, R/modules/seanet.py:190-196)."""
    k = weight.shape[-1]
    pt = k - stride
    extra = _extra_padding(x.shape[-1], k, stride, pt)
    x = F.pad(x, (pt, extra), mode=pad_mode)
    return F.conv1d(x, weight, bias, stride=stride)


def causal_convtr1d(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int,
                    groups: int = 1) -> torch.Tensor:
    """Non-streaming causal StreamingConvTranspose1d.forward (R/modules/conv.py:306-329):
    ConvTranspose1d then trim k - stride samples on the right (trim_right_ratio == 1)."""
    k = weight.shape[-1]
    y = F.conv_transpose1d(x, weight, bias, stride=stride, groups=groups)
    return y[..., : y.shape[-1] - (k - stride)]


def resblock(x: torch.Tensor, w: W, prefix: str) -> torch.Tensor:
    """SEANetResnetBlock.forward (R/modules/seanet.py:54-94): x + conv_k1(ELU(conv_k3(ELU(x))))."""
    h = causal_conv1d(F.elu(x), w[f"{prefix}.block.1.conv.conv.weight"], w[f"{prefix}.block.1.conv.conv.bias"])
    h = causal_conv1d(F.elu(h), w[f"{prefix}.block.3.conv.conv.weight"], w[f"{prefix}.block.3.conv.conv.bias"])
    return x + h


def seanet_encoder(x: torch.Tensor, w: W, cfg: MimiConfig = OFFICIAL) -> torch.Tensor:
    """SEANetEncoder.forward (R/modules/seanet.py:177-241): [B,1,L] -> [B,D,L/hop]."""
    idx = 0
    y = causal_conv1d(x, w[f"encoder.model.{idx}.conv.conv.weight"], w[f"encoder.model.{idx}.conv.conv.bias"])
    idx += 1
    for ratio in reversed(cfg.ratios):
        y = resblock(y, w, f"encoder.model.{idx}")
        idx += 2
        y = causal_conv1d(F.elu(y), w[f"encoder.model.{idx}.conv.conv.weight"],
                          w[f"encoder.model.{idx}.conv.conv.bias"], stride=ratio)
        idx += 1
    idx += 1
    return causal_conv1d(F.elu(y), w[f"encoder.model.{idx}.conv.conv.weight"], w[f"encoder.model.{idx}.conv.conv.bias"])


def seanet_decoder(z: torch.Tensor, w: W, cfg: MimiConfig = OFFICIAL) -> torch.Tensor:
    """SEANetDecoder.forward (R/modules/seanet.py:327-395): [B,D,F] -> [B,1,F*hop]."""
    idx = 0
    y = causal_conv1d(z, w[f"decoder.model.{idx}.conv.conv.weight"], w[f"decoder.model.{idx}.conv.conv.bias"])
    idx += 1
    for ratio in cfg.ratios:
        idx += 1
        y = causal_convtr1d(F.elu(y), w[f"decoder.model.{idx}.convtr.convtr.weight"],
                            w[f"decoder.model.{idx}.convtr.convtr.bias"], stride=ratio)
        idx += 1
        y = resblock(y, w, f"decoder.model.{idx}")
        idx += 1
    idx += 1
    return causal_conv1d(F.elu(y), w[f"decoder.model.{idx}.conv.conv.weight"], w[f"decoder.model.{idx}.conv.conv.bias"])


def downsample(x: torch.Tensor, w: W, cfg: MimiConfig = OFFICIAL) -> torch.Tensor:
    """ConvDownsample1d learnt=True (R/modules/resample.py:39-65): dense Conv1d k=2s stride s,
    no bias, pad_mode 'replicate'."""
    return causal_conv1d(x, w["downsample.conv.conv.conv.weight"], None, stride=cfg.resample_stride,
                         pad_mode="replicate")


def upsample(x: torch.Tensor, w: W, cfg: MimiConfig = OFFICIAL) -> torch.Tensor:
    """ConvTrUpsample1d learnt=True channel_wise=True (R/modules/resample.py:86-119):
    depthwise ConvTranspose1d k=2s stride s, no bias."""
    return causal_convtr1d(x, w["upsample.convtr.convtr.convtr.weight"], None, stride=cfg.resample_stride,
                           groups=x.shape[1])


# --------------------------------------------------------------------------- codec transformer
def rope_pairs(q: torch.Tensor, k: torch.Tensor, offset: int, max_period: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """Kyutai apply_rope, time_before_heads=False (R/modules/rope.py:11-68): rotation of
    (even, odd) pairs, angles computed on the fly in fp32. q,k: [B,H,T,D]."""
    B, H, T, D = q.shape
    ds = torch.arange(D // 2, dtype=torch.float32, device=q.device)
    freqs = torch.exp(ds * (-math.log(max_period) * 2 / D))
    ts = (float(offset) + torch.arange(T, dtype=torch.float32, device=q.device)).view(1, -1, 1)
    rotr, roti = torch.cos(freqs * ts), torch.sin(freqs * ts)
    q2, k2 = q.view(B, H, T, D // 2, 2), k.view(B, H, T, D // 2, 2)
    qr, qi, kr, ki = q2[..., 0].float(), q2[..., 1].float(), k2[..., 0].float(), k2[..., 1].float()
    qo = torch.stack([(qr * rotr - qi * roti).to(q.dtype), (qr * roti + qi * rotr).to(q.dtype)], dim=-1)
    ko = torch.stack([(kr * rotr - ki * roti).to(q.dtype), (kr * roti + ki * rotr).to(q.dtype)], dim=-1)
    return qo.view(B, H, T, D), ko.view(B, H, T, D)


class KVRing:
    """RingKVCache (R/modules/transformer.py:211-278): fixed-capacity ring written at
    (end_offset + t) % capacity; positions[-1] mark never-written slots."""

    def __init__(self, batch: int, heads: int, dim: int, capacity: int, dtype=torch.float32, device="cpu"):
        self.capacity = capacity
        self.cache = torch.zeros(2, batch, heads, capacity, dim, dtype=dtype, device=device)
        self.end_offset = 0

    def complete(self, k: torch.Tensor, v: torch.Tensor):
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


def mha(x: torch.Tensor, w: W, p: str, cfg: MimiConfig, offset: int, ring: Optional[KVRing]) -> torch.Tensor:
    """StreamingMultiheadAttention.forward (R/modules/transformer.py:375-419): in_proj laid out
    (p h d); pair-RoPE on q,k; causal mask with finite `context`; SDPA; out_proj."""
    B, T, D = x.shape
    H = cfg.num_heads
    proj = F.linear(x, w[f"{p}.self_attn.in_proj_weight"])
    q, k, v = proj.view(B, T, 3, H, D // H).permute(2, 0, 3, 1, 4)
    q, k = rope_pairs(q, k, offset, cfg.max_period)
    if ring is None:
        pos_k = torch.arange(T, device=x.device)
    else:
        k, v, pos_k = ring.complete(k, v)
    pos_k = pos_k.view(1, -1)
    pos_q = offset + torch.arange(T, device=x.device).view(-1, 1)
    delta = pos_q - pos_k
    bias = (pos_k >= 0) & (delta >= 0) & (delta < cfg.context)
    y = F.scaled_dot_product_attention(q, k, v, bias, dropout_p=0.0)
    y = y.permute(0, 2, 1, 3).reshape(B, T, D)
    return F.linear(y, w[f"{p}.self_attn.out_proj.weight"])


def transformer_layer(x: torch.Tensor, w: W, p: str, cfg: MimiConfig, offset: int, ring: Optional[KVRing]) -> torch.Tensor:
    """StreamingTransformerLayer.forward (R/modules/transformer.py:550-588): pre-LayerNorm(eps 1e-5),
    LayerScale on both residual branches, GELU FFN without biases."""
    D = x.shape[-1]
    h = F.layer_norm(x, (D,), w[f"{p}.norm1.weight"], w[f"{p}.norm1.bias"], 1e-5)
    x = x + w[f"{p}.layer_scale_1.scale"] * mha(h, w, p, cfg, offset, ring)
    h = F.layer_norm(x, (D,), w[f"{p}.norm2.weight"], w[f"{p}.norm2.bias"], 1e-5)
    h = F.linear(F.gelu(F.linear(h, w[f"{p}.linear1.weight"])), w[f"{p}.linear2.weight"])
    return x + w[f"{p}.layer_scale_2.scale"] * h


def codec_transformer(x_bct: torch.Tensor, w: W, side: str, cfg: MimiConfig = OFFICIAL, offset: int = 0,
                      rings: Optional[List[KVRing]] = None) -> torch.Tensor:
    """ProjectedTransformer.forward, conv_layout=True, no projections since d_model == dimension
    (R/modules/transformer.py:699-750) around StreamingTransformer.forward (:669-691, rope only)."""
    x = x_bct.transpose(1, 2)
    for l in range(cfg.num_layers):
        x = transformer_layer(x, w, f"{side}.transformer.layers.{l}", cfg, offset,
                              None if rings is None else rings[l])
    return x.transpose(1, 2)


# --------------------------------------------------------------------------- RVQ
def codebook_embedding(w: W, prefix: str, eps: float = 1e-5) -> torch.Tensor:
    """EuclideanCodebook.embedding (R/quantization/core_vq.py:142-150)."""
    return w[f"{prefix}.embedding_sum"] / w[f"{prefix}.cluster_usage"].clamp(min=eps)[:, None]


def codebooks(w: W, cfg: MimiConfig = OFFICIAL) -> torch.Tensor:
    """[n_q, bins, dim] centroids in code order."""
    return torch.stack([codebook_embedding(w, p) for p in codebook_prefixes(cfg)])


def quantize_nearest(x_nd: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """EuclideanCodebook._quantize (R/quantization/core_vq.py:179-185): cdist(p=2) + argmin
    (first minimum wins)."""
    return torch.cdist(x_nd[None], emb[None], p=2)[0].argmin(dim=-1)


def rvq_levels_encode(x_bdt: torch.Tensor, embs: torch.Tensor) -> torch.Tensor:
    """ResidualVectorQuantization.encode (R/quantization/core_vq.py:365-376) with
    VectorQuantization.encode/decode (:285-297; project_in/out are Identity because
    codebook_dim == dim). x [B,d,T] -> codes [n, B, T]."""
    residual = x_bdt
    out = []
    for emb in embs:
        xr = residual.permute(0, 2, 1)
        shape = xr.shape
        codes = quantize_nearest(xr.reshape(-1, shape[-1]), emb).view(*shape[:-1])
        quantized = F.embedding(codes, emb).permute(0, 2, 1)
        residual = residual - quantized
        out.append(codes)
    return torch.stack(out)


def rvq_encode(z: torch.Tensor, w: W, cfg: MimiConfig = OFFICIAL) -> torch.Tensor:
    """SplitResidualVectorQuantizer.encode (R/quantization/vq.py:305-315) over
    ResidualVectorQuantizer.encode (:134-147): rvq_first and rvq_rest each project the SAME
    latent with their own 1x1 input_proj. z [B,D,T] -> codes [B,n_q,T] int64."""
    if z.shape[-1] == 0:
        return torch.empty((z.shape[0], cfg.n_q, 0), dtype=torch.int64, device=z.device)
    embs = codebooks(w, cfg)
    ns = cfg.n_q_semantic
    x1 = F.conv1d(z, w["quantizer.rvq_first.input_proj.weight"])
    c1 = rvq_levels_encode(x1, embs[:ns]).transpose(0, 1)
    x2 = F.conv1d(z, w["quantizer.rvq_rest.input_proj.weight"])
    c2 = rvq_levels_encode(x2, embs[ns:]).transpose(0, 1)
    return torch.cat([c1, c2], dim=1)


def rvq_decode(codes: torch.Tensor, w: W, cfg: MimiConfig = OFFICIAL) -> torch.Tensor:
    """SplitResidualVectorQuantizer.decode (R/quantization/vq.py:317-323): sum of centroid
    gathers per group -> 1x1 output_proj per group -> sum. codes [B,K,T] -> [B,D,T]."""
    embs = codebooks(w, cfg)
    ns = cfg.n_q_semantic

    def group(cs, es, proj):
        q = torch.zeros((), device=cs.device)
        for lvl in range(cs.shape[1]):
            q = q + F.embedding(cs[:, lvl], es[lvl])
        return F.conv1d(q.permute(0, 2, 1), proj)

    out = group(codes[:, :ns], embs[:ns], w["quantizer.rvq_first.output_proj.weight"])
    if codes.shape[1] > ns:
        out = out + group(codes[:, ns:], embs[ns:], w["quantizer.rvq_rest.output_proj.weight"])
    return out


# --------------------------------------------------------------------------- whole codec
def encode_latent(audio: torch.Tensor, w: W, cfg: MimiConfig = OFFICIAL) -> torch.Tensor:
    z = seanet_encoder(audio, w, cfg)
    z = codec_transformer(z, w, "encoder_transformer", cfg)
    return downsample(z, w, cfg)


def encode(audio: torch.Tensor, w: W, cfg: MimiConfig = OFFICIAL) -> torch.Tensor:
    """MimiCodec.encode (R/models/MimiCodec.py:93-101): [B,1,L] -> codes [B,n_q,ceil(L/frame)]."""
    return rvq_encode(encode_latent(audio, w, cfg), w, cfg)


def decode(codes: torch.Tensor, w: W, cfg: MimiConfig = OFFICIAL) -> torch.Tensor:
    """MimiCodec.decode (R/models/MimiCodec.py:103-110): codes [B,n_q,T] -> wav [B,1,T*frame]."""
    z = rvq_decode(codes, w, cfg)
    z = upsample(z, w, cfg)
    z = codec_transformer(z, w, "decoder_transformer", cfg)
    return seanet_decoder(z, w, cfg)


# --------------------------------------------------------------------------- streaming
class _ConvStream:
    """RawStreamingConv1d.forward + StreamingConv1d first-call padding
    (R/modules/streaming.py:216-244; R/modules/conv.py:245-254)."""

    def __init__(self, weight, bias, stride=1, pad_mode="constant"):
        self.w, self.b, self.s, self.mode = weight, bias, stride, pad_mode
        self.k = weight.shape[-1]
        self.previous: Optional[torch.Tensor] = None
        self.padding_to_add = self.k - stride

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if self.padding_to_add > 0 and x.shape[-1] > 0:
            x = F.pad(x, (self.padding_to_add, 0), mode=self.mode)
            self.padding_to_add = 0
        if self.previous is not None:
            x = torch.cat([self.previous, x], dim=-1)
        T = x.shape[-1]
        n = max(0, (T - self.k) // self.s + 1)
        self.previous = x[..., n * self.s:]
        if n == 0:
            return x.new_empty(x.shape[0], self.w.shape[0], 0)
        return F.conv1d(x[..., : (n - 1) * self.s + self.k], self.w, self.b, stride=self.s)


class _ConvTrStream:
    """RawStreamingConvTranspose1d.forward (R/modules/streaming.py:270-303): overlap-add the
    k - stride tail (`partial`, minus bias) into the next call."""

    def __init__(self, weight, bias, stride, groups=1):
        self.w, self.b, self.s, self.g = weight, bias, stride, groups
        self.k = weight.shape[-1]
        self.partial: Optional[torch.Tensor] = None

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if x.shape[-1] == 0:
            return x.new_empty(x.shape[0], self.w.shape[1] * self.g, 0)
        out = F.conv_transpose1d(x, self.w, self.b, stride=self.s, groups=self.g)
        if self.partial is not None:
            PT = self.partial.shape[-1]
            out[..., :PT] += self.partial - self.b[:, None] if self.b is not None else self.partial
        inv = self.k - self.s
        OT = out.shape[-1]
        self.partial = out[..., OT - inv:]
        return out[..., : OT - inv]


class _ResStream:
    """SEANetResnetBlock in streaming mode; StreamingAdd (R/modules/streaming.py:177-194) never has
    to realign here because both branches are k_eff - stride == ctx causal with stride 1."""

    def __init__(self, w: W, prefix: str):
        self.c1 = _ConvStream(w[f"{prefix}.block.1.conv.conv.weight"], w[f"{prefix}.block.1.conv.conv.bias"])
        self.c2 = _ConvStream(w[f"{prefix}.block.3.conv.conv.weight"], w[f"{prefix}.block.3.conv.conv.bias"])

    def __call__(self, x):
        h = self.c2(F.elu(self.c1(F.elu(x))))
        assert h.shape[-1] == x.shape[-1]
        return x + h


class StreamingCodec:
    """Chunked encode/decode with the reference's StreamingModule carry semantics
    (R/modules/streaming.py:33-151; MimiModel.encode/decode under `streaming(B)`,
    MLLM_v2/moshi/models/compression.py:368-423).  One instance == one `with streaming(B)` scope."""

    def __init__(self, w: W, batch: int, cfg: MimiConfig = OFFICIAL):
        self.w, self.cfg, self.B = w, cfg, batch
        D = cfg.dimension
        # encoder
        self.enc: list = []
        idx = 0
        self.enc.append(("conv", _ConvStream(w[f"encoder.model.{idx}.conv.conv.weight"], w[f"encoder.model.{idx}.conv.conv.bias"])))
        idx += 1
        for ratio in reversed(cfg.ratios):
            self.enc.append(("res", _ResStream(w, f"encoder.model.{idx}")))
            idx += 2
            self.enc.append(("elu_conv", _ConvStream(w[f"encoder.model.{idx}.conv.conv.weight"],
                                                     w[f"encoder.model.{idx}.conv.conv.bias"], ratio)))
            idx += 1
        idx += 1
        self.enc.append(("elu_conv", _ConvStream(w[f"encoder.model.{idx}.conv.conv.weight"], w[f"encoder.model.{idx}.conv.conv.bias"])))
        self.down = _ConvStream(w["downsample.conv.conv.conv.weight"], None, cfg.resample_stride, "replicate")
        # decoder
        self.up = _ConvTrStream(w["upsample.convtr.convtr.convtr.weight"], None, cfg.resample_stride, groups=D)
        self.dec: list = []
        idx = 0
        self.dec.append(("conv", _ConvStream(w[f"decoder.model.{idx}.conv.conv.weight"], w[f"decoder.model.{idx}.conv.conv.bias"])))
        idx += 1
        for ratio in cfg.ratios:
            idx += 1
            self.dec.append(("elu_convtr", _ConvTrStream(w[f"decoder.model.{idx}.convtr.convtr.weight"],
                                                         w[f"decoder.model.{idx}.convtr.convtr.bias"], ratio)))
            idx += 1
            self.dec.append(("res", _ResStream(w, f"decoder.model.{idx}")))
            idx += 1
        idx += 1
        self.dec.append(("elu_conv", _ConvStream(w[f"decoder.model.{idx}.conv.conv.weight"], w[f"decoder.model.{idx}.conv.conv.bias"])))
        hd = D // cfg.num_heads
        dev = w["downsample.conv.conv.conv.weight"].device   # weights on a GPU = the same ATen calls the reference eager makes there
        self.enc_rings = [KVRing(batch, cfg.num_heads, hd, cfg.context, device=dev) for _ in range(cfg.num_layers)]
        self.dec_rings = [KVRing(batch, cfg.num_heads, hd, cfg.context, device=dev) for _ in range(cfg.num_layers)]
        self.enc_offset = 0
        self.dec_offset = 0

    @staticmethod
    def _run(stack, x):
        for kind, m in stack:
            if kind.startswith("elu_"):
                x = F.elu(x)
            x = m(x)
        return x

    def encode(self, chunk: torch.Tensor) -> torch.Tensor:
        z = self._run(self.enc, chunk)
        if z.shape[-1] > 0:
            T = z.shape[-1]
            z = codec_transformer(z, self.w, "encoder_transformer", self.cfg, self.enc_offset, self.enc_rings)
            self.enc_offset += T
        z = self.down(z)
        self.last_latent = z          # for rvq_margins: which frames may legitimately flip an index
        return rvq_encode(z, self.w, self.cfg)

    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        z = rvq_decode(codes, self.w, self.cfg)
        z = self.up(z)
        T = z.shape[-1]
        z = codec_transformer(z, self.w, "decoder_transformer", self.cfg, self.dec_offset, self.dec_rings)
        self.dec_offset += T
        return self._run(self.dec, z)


# --------------------------------------------------------------------------- diagnostics
def rvq_margins(z: torch.Tensor, w: W, cfg: MimiConfig = OFFICIAL) -> torch.Tensor:
    """Relative top-1/top-2 distance margin per (level, frame) in float64 along the oracle's own
    decision path: frames whose margin is below ~1e-5 can legitimately flip under any change of
    fp32 summation order. Returns [n_q, B*T]."""
    embs = codebooks(w, cfg).double()
    ns = cfg.n_q_semantic
    out = []
    for part, es in (("rvq_first", embs[:ns]), ("rvq_rest", embs[ns:])):
        x = F.conv1d(z, w[f"quantizer.{part}.input_proj.weight"]).double().permute(0, 2, 1).reshape(-1, cfg.codebook_dim)
        for emb in es:
            d = torch.cdist(x[None], emb[None])[0]
            top2 = d.topk(2, dim=-1, largest=False)
            out.append((top2.values[:, 1] - top2.values[:, 0]) / top2.values[:, 1].clamp(min=1e-30))
            x = x - emb[top2.indices[:, 0]]
    return torch.stack(out)
