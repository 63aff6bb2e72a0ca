"""Shapes, state_dict key names and seeded synthetic weights for the Mimi codec.

Neutral specification data (see specs/__init__.py): imported by the oracle, the tests and bench.py's product arm.

The key names and shapes below are the ``state_dict`` of the reference's
tokenizer-variant ``MimiCodec`` (MLLM_v2/tools/tokenizer/MimiCodec/model/models/MimiCodec.py:25-74)
built with ``mimi_config.yaml`` (encoder_rates [8,6,5,4], codebook_size 2048,
codebook_dim 256, rvq_layers 8); ``oracle/gen_golden.py`` asserts that they match
the real module key-for-key, so real Kyutai checkpoints load into the same layout.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch


@dataclass(frozen=True)
class MimiConfig:
    """Hyper-parameters (MimiCodec.py:27-57; loaders.py:24-66)."""
    sample_rate: int = 24000
    n_filters: int = 64
    ratios: Tuple[int, ...] = (8, 6, 5, 4)       # decoder order; encoder uses reversed
    dimension: int = 512                          # latent_dim
    kernel_size: int = 7
    last_kernel_size: int = 3
    residual_kernel_size: int = 3
    compress: int = 2
    codebook_size: int = 2048
    codebook_dim: int = 256
    n_q: int = 8
    n_q_semantic: int = 1
    num_heads: int = 8
    num_layers: int = 8
    dim_feedforward: int = 2048
    context: int = 250
    max_period: float = 10000.0
    layer_scale: float = 0.01
    resample_stride: int = 2                      # encoder_frame_rate 25 Hz -> 12.5 Hz
    semantic_feature_dim: int = 1024              # unused at inference (MimiCodec.py:69)

    @property
    def hop_length(self) -> int:
        return int(math.prod(self.ratios))

    @property
    def frame_size(self) -> int:
        """Samples per 12.5 Hz code frame (1920 for the official config)."""
        return self.hop_length * self.resample_stride


OFFICIAL = MimiConfig()
# std of the projected latents under synthetic_weights(seed=41) (measured once, oracle/gen_golden.py)
CODEBOOK_SCALE = 1.3
# A small config with the same topology, for fast CPU tests of host logic.
TINY = MimiConfig(n_filters=4, ratios=(4, 3, 2, 2), dimension=32, codebook_size=64, codebook_dim=16,
                  n_q=4, num_heads=2, num_layers=2, dim_feedforward=64, context=10,
                  semantic_feature_dim=8)


def param_spec(cfg: MimiConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """Ordered (name, shape, kind) list. kind in {weight,bias,norm_w,norm_b,scale}.

    Mirrors SEANetEncoder/SEANetDecoder construction (modules/seanet.py:160-237, 313-390),
    ConvDownsample1d/ConvTrUpsample1d (modules/resample.py:39-57, 86-104),
    StreamingTransformerLayer (modules/transformer.py:468-548) and
    ResidualVectorQuantizer projections (quantization/vq.py:80-91).
    """
    spec: List[Tuple[str, Tuple[int, ...], str]] = []
    nf, D = cfg.n_filters, cfg.dimension

    def conv(prefix, cout, cin, k, bias=True):
        spec.append((f"{prefix}.weight", (cout, cin, k), "weight"))
        if bias:
            spec.append((f"{prefix}.bias", (cout,), "bias"))

    def resblock(prefix, dim):
        hidden = dim // cfg.compress
        conv(f"{prefix}.block.1.conv.conv", hidden, dim, cfg.residual_kernel_size)
        conv(f"{prefix}.block.3.conv.conv", dim, hidden, 1)

    # ---- encoder (seanet.py:177-237)
    idx = 0
    mult = 1
    conv(f"encoder.model.{idx}.conv.conv", mult * nf, 1, cfg.kernel_size)
    idx += 1
    for ratio in reversed(cfg.ratios):
        resblock(f"encoder.model.{idx}", mult * nf)
        idx += 2  # resblock, ELU
        conv(f"encoder.model.{idx}.conv.conv", mult * nf * 2, mult * nf, 2 * ratio)
        idx += 1
        mult *= 2
    idx += 1  # ELU
    conv(f"encoder.model.{idx}.conv.conv", D, mult * nf, cfg.last_kernel_size)

    # ---- decoder (seanet.py:327-390)
    idx = 0
    mult = 2 ** len(cfg.ratios)
    conv(f"decoder.model.{idx}.conv.conv", mult * nf, D, cfg.kernel_size)
    idx += 1
    for ratio in cfg.ratios:
        idx += 1  # ELU
        # ConvTranspose1d weight is [Cin, Cout, k]
        spec.append((f"decoder.model.{idx}.convtr.convtr.weight", (mult * nf, mult * nf // 2, 2 * ratio), "weight"))
        spec.append((f"decoder.model.{idx}.convtr.convtr.bias", (mult * nf // 2,), "bias"))
        idx += 1
        resblock(f"decoder.model.{idx}", mult * nf // 2)
        idx += 1
        mult //= 2
    idx += 1  # ELU
    conv(f"decoder.model.{idx}.conv.conv", 1, nf, cfg.last_kernel_size)

    # ---- resampling (resample.py; MimiCodec.py:65-66)
    s = cfg.resample_stride
    spec.append(("downsample.conv.conv.conv.weight", (D, D, 2 * s), "weight"))
    spec.append(("upsample.convtr.convtr.convtr.weight", (D, 1, 2 * s), "weight"))
    # semantic distillation head: present in the state_dict, unused by encode/decode
    spec.append(("semantic_mapping_layer.ln_layer.weight", (D, cfg.semantic_feature_dim), "weight"))
    spec.append(("semantic_mapping_layer.ln_layer.bias", (D,), "bias"))

    # ---- transformers (transformer.py:468-548)
    for side in ("encoder_transformer", "decoder_transformer"):
        for l in range(cfg.num_layers):
            p = f"{side}.transformer.layers.{l}"
            spec.append((f"{p}.self_attn.in_proj_weight", (3 * D, D), "weight"))
            spec.append((f"{p}.self_attn.out_proj.weight", (D, D), "weight"))
            spec.append((f"{p}.norm1.weight", (D,), "norm_w"))
            spec.append((f"{p}.norm1.bias", (D,), "norm_b"))
            spec.append((f"{p}.norm2.weight", (D,), "norm_w"))
            spec.append((f"{p}.norm2.bias", (D,), "norm_b"))
            spec.append((f"{p}.linear1.weight", (cfg.dim_feedforward, D), "weight"))
            spec.append((f"{p}.linear2.weight", (D, cfg.dim_feedforward), "weight"))
            spec.append((f"{p}.layer_scale_1.scale", (D,), "scale"))
            spec.append((f"{p}.layer_scale_2.scale", (D,), "scale"))

    # ---- quantizer projections (vq.py:80-91, 213-223)
    for part in ("rvq_first", "rvq_rest"):
        spec.append((f"quantizer.{part}.input_proj.weight", (cfg.codebook_dim, D, 1), "weight"))
        spec.append((f"quantizer.{part}.output_proj.weight", (D, cfg.codebook_dim, 1), "weight"))
    return spec


def buffer_spec(cfg: MimiConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """EuclideanCodebook buffers (quantization/core_vq.py:120-124)."""
    out = []
    parts = [("rvq_first", cfg.n_q_semantic), ("rvq_rest", cfg.n_q - cfg.n_q_semantic)]
    for part, n in parts:
        for i in range(n):
            p = f"quantizer.{part}.vq.layers.{i}._codebook"
            out.append((f"{p}._initialized", (1,)))
            out.append((f"{p}.cluster_usage", (cfg.codebook_size,)))
            out.append((f"{p}.embedding_sum", (cfg.codebook_size, cfg.codebook_dim)))
    return out


def codebook_prefixes(cfg: MimiConfig) -> List[str]:
    """state_dict prefixes of the n_q codebooks in code order (vq.py:305-315)."""
    out = [f"quantizer.rvq_first.vq.layers.{i}._codebook" for i in range(cfg.n_q_semantic)]
    out += [f"quantizer.rvq_rest.vq.layers.{i}._codebook" for i in range(cfg.n_q - cfg.n_q_semantic)]
    return out


def synthetic_weights(cfg: MimiConfig = OFFICIAL, seed: int = 41,
                      codebook_scale: float | None = None) -> Dict[str, torch.Tensor]:
    """Deterministic fp32 state_dict (CPU).

    Recipe (ours; mirrors the spirit of moshi/modules/conv_test.py:53-60 which uses
    xavier-uniform seed 41): >=2-D weights xavier-uniform; biases U(-0.05,0.05) (non-zero so
    bias handling is exercised); norm weights 1+0.1 N(0,1); norm biases 0.05 N(0,1);
    LayerScale U(0.25,0.75) (larger than the trained 0.01 so transformer errors are visible);
    codebooks: embedding_sum = usage * N(0, scale^2) with usage U(0.5, 2) so that the
    `embedding_sum / cluster_usage.clamp(min=eps)` division (core_vq.py:142-150) matters.
    """
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, torch.Tensor] = {}
    for name, shape, kind in param_spec(cfg):
        t = torch.empty(shape, dtype=torch.float32)
        if kind == "weight":
            torch.nn.init.xavier_uniform_(t, generator=g)
        elif kind == "bias":
            t.uniform_(-0.05, 0.05, generator=g)
        elif kind == "norm_w":
            t.normal_(0.0, 0.1, generator=g).add_(1.0)
        elif kind == "norm_b":
            t.normal_(0.0, 0.05, generator=g)
        elif kind == "scale":
            t.uniform_(0.25, 0.75, generator=g)
        w[name] = t
    scale = codebook_scale if codebook_scale is not None else CODEBOOK_SCALE
    for lvl, p in enumerate(codebook_prefixes(cfg)):
        usage = torch.empty(cfg.codebook_size).uniform_(0.5, 2.0, generator=g)
        emb = torch.empty(cfg.codebook_size, cfg.codebook_dim).normal_(0.0, 1.0, generator=g)
        # residual energy shrinks level after level; keep centroids on the residual's scale
        emb *= scale * (0.75 ** lvl if lvl > 0 else 1.0)
        w[f"{p}._initialized"] = torch.ones(1)
        w[f"{p}.cluster_usage"] = usage
        w[f"{p}.embedding_sum"] = emb * usage[:, None]
    return w


def synthetic_audio(batch: int, length: int, seed: int = 0) -> torch.Tensor:
    """Seeded synthetic 24 kHz mono clip [B,1,L]: 5 sinusoids 100-4000 Hz + noise, peak ~0.3
    (SURVEY.md §8d cfg 1)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(length, dtype=torch.float32) / 24000.0
    x = torch.zeros(batch, 1, length)
    for b in range(batch):
        freqs = torch.empty(5).uniform_(100.0, 4000.0, generator=g)
        amps = torch.empty(5).uniform_(0.02, 0.06, generator=g)
        phases = torch.empty(5).uniform_(0.0, 6.2831853, generator=g)
        sig = (amps[:, None] * torch.sin(6.2831853 * freqs[:, None] * t[None] + phases[:, None])).sum(0)
        x[b, 0] = sig
    x += 0.02 * torch.empty(batch, 1, length).normal_(0.0, 1.0, generator=g)
    return x
