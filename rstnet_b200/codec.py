"""B200-native Mimi codec behind the reference's Python API.

Mirrors (names, signatures, tensor layouts, state_dict keys):
  * ``MimiCodec.encode(audio[B,1,L]) -> codes[B,n_q,ceil(L/1920)] int64`` and
    ``MimiCodec.decode(codes[B,n_q,T]) -> wav[B,1,1920*T]``
    (MLLM_v2/tools/tokenizer/MimiCodec/model/models/MimiCodec.py:25-110);
  * the ``StreamingModule`` protocol used by ``MimiModel`` under ``with m.streaming(B):`` /
    ``streaming_forever`` / ``reset_streaming`` (MLLM_v2/modules/streaming.py:33-151,
    MLLM_v2/moshi/models/compression.py:368-423): chunked encode/decode with conv carry rows and
    ring KV caches kept in device buffers;
  * ``MimiTokenizer.tokenize / detokenize`` (tools/tokenizer/MimiCodec/mimi_tokenizer.py:56-82).

All arithmetic runs in librstnet_b200.so (hand-written sm_100a CUDA); this file only lays out
buffers in HBM and sequences launches.  Internal activation layout is [B, T, C] (channels last)
so that every conv is a GEMM over overlapping contiguous rows; each layer's input buffer carries
its own left context rows (`k - stride`), which double as the streaming carry.
"""
from __future__ import annotations

import math
from contextlib import contextmanager
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import ops
from ._lib import ACT_ELU, ACT_GELU, ACT_NONE, RstnetError


class _Node(nn.Module):
    """Anonymous container so that parameters get the reference's dotted state_dict names."""


def _register(root: nn.Module, dotted: str, tensor: torch.Tensor, buffer: bool = False) -> None:
    *path, leaf = dotted.split(".")
    m = root
    for p in path:
        if p not in m._modules:
            m.add_module(p, _Node())
        m = m._modules[p]
    if buffer:
        m.register_buffer(leaf, tensor)
    else:
        m.register_parameter(leaf, nn.Parameter(tensor, requires_grad=False))


_OLD_CODEBOOK_NAMES = {"inited": "_initialized", "cluster_size": "cluster_usage", "embed_avg": "embedding_sum",
                       "embed_sum": "embedding_sum"}  # quantization/core_vq.py:126-140


class _Buf:
    """[B, ctx + T + extra, C] fp32 activation buffer; rows [ctx, ctx+T) are the live rows."""

    def __init__(self, B: int, ctx: int, T: int, extra: int, C: int, device):
        self.B, self.ctx, self.T, self.extra, self.C = B, ctx, T, extra, C
        self.rows = ctx + T + extra
        self.t = torch.zeros(B, self.rows, C, device=device, dtype=torch.float32)
        self.bs = self.rows * C

    def off(self, row: int) -> int:
        return row * self.C


class MimiCodec(nn.Module):
    """Drop-in for the reference ``MimiCodec`` (same constructor arguments and defaults)."""

    def __init__(self, sample_rate=24000, n_filters=64, encoder_rates=[4, 5, 6, 8], compress=2, causal=True,
                 latent_dim=512, codebook_size=4096, codebook_dim=32, rvq_layers=8, num_heads=8, num_layers=8,
                 layer_scale=0.01, context=250, dim_feedforward=2048, semantic_feature_dim=1024,
                 target_frame_rate=12.5):
        super().__init__()
        if not causal:
            raise NotImplementedError("only the causal codec (the streaming hot path) is implemented")
        self.sample_rate = sample_rate
        self.n_filters = n_filters
        self.ratios = list(encoder_rates)
        self.compress = compress
        self.latent_dim = latent_dim
        self.codebook_size, self.codebook_dim, self.n_q, self.n_q_semantic = codebook_size, codebook_dim, rvq_layers, 1
        self.num_heads, self.num_layers = num_heads, num_layers
        self.context, self.dim_feedforward = context, 2048  # the reference hard-codes 2048 (MimiCodec.py:55)
        self.max_period = 10000.0
        self.kernel_size, self.last_kernel_size, self.residual_kernel_size = 7, 3, 3
        self.hop_length = int(math.prod(self.ratios))
        self.encoder_frame_rate = sample_rate / self.hop_length
        self.target_frame_rate = target_frame_rate
        self.resample_stride = int(self.encoder_frame_rate / self.target_frame_rate)
        self.frame_size = self.hop_length * self.resample_stride
        self.codebook_eps = 1e-5
        assert latent_dim % num_heads == 0 and n_filters % (4 * compress) == 0 and codebook_dim % 16 == 0

        D, nf = latent_dim, n_filters
        g = torch.Generator().manual_seed(0)

        def w_(shape):
            t = torch.empty(shape)
            if len(shape) >= 2:
                nn.init.xavier_uniform_(t, generator=g)
            else:
                t.zero_()
            return t

        def conv(prefix, cout, cin, k, bias=True):
            _register(self, f"{prefix}.weight", w_((cout, cin, k)))
            if bias:
                _register(self, f"{prefix}.bias", w_((cout,)))

        def resblock(prefix, dim):
            conv(f"{prefix}.block.1.conv.conv", dim // compress, dim, self.residual_kernel_size)
            conv(f"{prefix}.block.3.conv.conv", dim, dim // compress, 1)

        # encoder (modules/seanet.py:177-237)
        idx, mult = 0, 1
        conv(f"encoder.model.{idx}.conv.conv", nf, 1, self.kernel_size)
        idx += 1
        for r in reversed(self.ratios):
            resblock(f"encoder.model.{idx}", mult * nf)
            idx += 2
            conv(f"encoder.model.{idx}.conv.conv", mult * nf * 2, mult * nf, 2 * r)
            idx += 1
            mult *= 2
        idx += 1
        conv(f"encoder.model.{idx}.conv.conv", D, mult * nf, self.last_kernel_size)
        # decoder (modules/seanet.py:327-390)
        idx, mult = 0, 2 ** len(self.ratios)
        conv(f"decoder.model.{idx}.conv.conv", mult * nf, D, self.kernel_size)
        idx += 1
        for r in self.ratios:
            idx += 1
            _register(self, f"decoder.model.{idx}.convtr.convtr.weight", w_((mult * nf, mult * nf // 2, 2 * r)))
            _register(self, f"decoder.model.{idx}.convtr.convtr.bias", w_((mult * nf // 2,)))
            idx += 1
            resblock(f"decoder.model.{idx}", mult * nf // 2)
            idx += 1
            mult //= 2
        idx += 1
        conv(f"decoder.model.{idx}.conv.conv", 1, nf, self.last_kernel_size)
        s = self.resample_stride
        _register(self, "downsample.conv.conv.conv.weight", w_((D, D, 2 * s)))
        _register(self, "upsample.convtr.convtr.convtr.weight", w_((D, 1, 2 * s)))
        _register(self, "semantic_mapping_layer.ln_layer.weight", w_((D, semantic_feature_dim)))
        _register(self, "semantic_mapping_layer.ln_layer.bias", w_((D,)))
        for side in ("encoder_transformer", "decoder_transformer"):
            for l in range(num_layers):
                p = f"{side}.transformer.layers.{l}"
                _register(self, f"{p}.self_attn.in_proj_weight", w_((3 * D, D)))
                _register(self, f"{p}.self_attn.out_proj.weight", w_((D, D)))
                for n in ("norm1", "norm2"):
                    _register(self, f"{p}.{n}.weight", torch.ones(D))
                    _register(self, f"{p}.{n}.bias", torch.zeros(D))
                _register(self, f"{p}.linear1.weight", w_((self.dim_feedforward, D)))
                _register(self, f"{p}.linear2.weight", w_((D, self.dim_feedforward)))
                _register(self, f"{p}.layer_scale_1.scale", torch.full((D,), float(layer_scale)))
                _register(self, f"{p}.layer_scale_2.scale", torch.full((D,), float(layer_scale)))
        for part, n in (("rvq_first", self.n_q_semantic), ("rvq_rest", self.n_q - self.n_q_semantic)):
            _register(self, f"quantizer.{part}.input_proj.weight", w_((codebook_dim, D, 1)))
            _register(self, f"quantizer.{part}.output_proj.weight", w_((D, codebook_dim, 1)))
            for i in range(n):
                p = f"quantizer.{part}.vq.layers.{i}._codebook"
                _register(self, f"{p}._initialized", torch.tensor([False], dtype=torch.float), buffer=True)
                _register(self, f"{p}.cluster_usage", torch.ones(codebook_size), buffer=True)
                _register(self, f"{p}.embedding_sum", torch.zeros(codebook_size, codebook_dim), buffer=True)
        self._engine: Optional["_Engine"] = None
        self._stream_state: Optional["_StreamState"] = None
        self.use_cuda_graphs = True

    # ------------------------------------------------------------------ parameters
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        sd = {}
        for k, v in state_dict.items():
            head, _, leaf = k.rpartition(".")
            if head.endswith("_codebook") and leaf in _OLD_CODEBOOK_NAMES:
                k = f"{head}.{_OLD_CODEBOOK_NAMES[leaf]}"
            sd[k] = v
        self._engine = None
        return super().load_state_dict(sd, strict=strict, **kw)

    def _apply(self, fn, *a, **kw):
        self._engine = None
        return super()._apply(fn, *a, **kw)

    @classmethod
    def from_config(cls, config_path):
        import json
        with open(config_path, "r") as f:
            return cls(**json.load(f))

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    def _eng(self) -> "_Engine":
        dev = self.device
        if dev.type != "cuda":
            raise RstnetError("MimiCodec runs on CUDA only (sm_100a kernels; the CPU path is the reference itself)")
        if self._engine is None or self._engine.device != dev:
            self._engine = _Engine(self, dev)
        return self._engine

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def encode(self, audio_data: torch.Tensor) -> torch.Tensor:
        """[B,1,L] float -> codes [B,n_q,T] int64 (MimiCodec.py:93-101; MimiModel.encode when streaming)."""
        if audio_data.dim() != 3 or audio_data.shape[1] != 1:
            raise ValueError(f"expected audio of shape [B,1,L], got {tuple(audio_data.shape)}")
        eng = self._eng()
        x = audio_data.to(device=eng.device, dtype=torch.float32)
        if self._stream_state is not None:
            return self._stream_state.encode(x)
        return eng.encode_batch(x)

    @torch.no_grad()
    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """codes [B,K,T] int -> wav [B,1,T*frame_size] float32 (MimiCodec.py:103-110)."""
        if codes.dim() != 3:
            raise ValueError(f"expected codes of shape [B,K,T], got {tuple(codes.shape)}")
        if codes.dtype.is_floating_point:
            raise ValueError("codes must be integers")  # core_vq.py:202-204
        eng = self._eng()
        c = codes.to(device=eng.device, dtype=torch.int64).contiguous()
        if self._stream_state is not None:
            return self._stream_state.decode(c)
        return eng.decode_batch(c)

    def forward(self, *a, **kw):
        raise NotImplementedError("training forward (with semantic distillation) is out of scope; use encode/decode")

    # ------------------------------------------------------------------ StreamingModule protocol
    @property
    def is_streaming(self) -> bool:
        return self._stream_state is not None

    def streaming_forever(self, batch_size: int):
        self._stream_state = _StreamState(self._eng(), batch_size)

    @contextmanager
    def streaming(self, batch_size: int):
        self.streaming_forever(batch_size)
        try:
            yield
        finally:
            self._stream_state = None

    def reset_streaming(self):
        if self._stream_state is None:
            raise ValueError("Trying to reset streaming, but the codec wasn't streaming.")
        self._stream_state.reset()


class MimiTokenizer:
    """tokenize / detokenize wrapper (tools/tokenizer/MimiCodec/mimi_tokenizer.py:14-82).  The
    checkpoint download is left to the caller: pass a MimiCodec with weights loaded."""

    def __init__(self, model: MimiCodec, device=torch.device("cuda")):
        self.model = model.to(device).eval()
        self.device = device
        self.sr = 24000

    def find_length(self, x):
        return x.shape[1]

    def tokenize(self, wav, sample_rate: int = 24000):
        if not isinstance(wav, torch.Tensor):
            raise NotImplementedError
        if wav.dim() == 1:
            return wav
        if wav.dim() == 2:
            if wav.numel() == 0:
                return None
            if sample_rate != self.sr:
                raise NotImplementedError("resample to 24 kHz before tokenizing")
            wav = wav.unsqueeze(1)
        codes = self.model.encode(wav.to(self.device))
        return codes.squeeze(0).detach().cpu().to(torch.int16)

    def detokenize(self, codes):
        assert codes.shape[0] == 8
        wav = self.model.decode(codes.unsqueeze(0).to(self.device).long())
        return wav.squeeze(1).detach().cpu()


# ====================================================================== engine
class _Engine:
    """Device-resident packed weights + launch sequences."""

    BATCH_MODE_BYTES = 12 << 30  # activation budget per non-streaming pass; larger batches are split

    def __init__(self, m: MimiCodec, device: torch.device):
        self.m, self.device = m, device
        sd = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in m.state_dict().items()}
        self.D, self.nf = m.latent_dim, m.n_filters
        self.ratios = list(m.ratios)
        self.enc_ratios = list(reversed(m.ratios))
        D = self.D

        def conv_w(prefix):  # [Cout,Cin,k] -> Wt[(tap,ci), co]
            w = sd[f"{prefix}.weight"]
            cout, cin, k = w.shape
            return w.permute(2, 1, 0).reshape(k * cin, cout).contiguous(), sd.get(f"{prefix}.bias")

        def convtr_w(prefix, s):  # [Cin,Cout,2s] -> Wt[(half,ci), (j,co)], half 0 multiplies x[t-1]
            w = sd[f"{prefix}.weight"]
            cin, cout, k = w.shape
            assert k == 2 * s
            w_prev = w[:, :, s:].permute(0, 2, 1).reshape(cin, s * cout)
            w_cur = w[:, :, :s].permute(0, 2, 1).reshape(cin, s * cout)
            return torch.cat([w_prev, w_cur], 0).contiguous(), sd[f"{prefix}.bias"].repeat(s).contiguous()

        # ---- encoder
        self.e_conv0_w = sd["encoder.model.0.conv.conv.weight"].reshape(self.nf, -1).contiguous()
        self.e_conv0_b = sd["encoder.model.0.conv.conv.bias"]
        self.e_res, self.e_down = [], []
        idx = 1
        for r in self.enc_ratios:
            self.e_res.append((conv_w(f"encoder.model.{idx}.block.1.conv.conv"), conv_w(f"encoder.model.{idx}.block.3.conv.conv")))
            idx += 2
            self.e_down.append(conv_w(f"encoder.model.{idx}.conv.conv"))
            idx += 1
        idx += 1
        self.e_final = conv_w(f"encoder.model.{idx}.conv.conv")
        self.down_w = conv_w("downsample.conv.conv.conv")[0]
        # ---- decoder
        self.d_conv0 = conv_w("decoder.model.0.conv.conv")
        self.d_tr, self.d_res = [], []
        idx = 1
        for r in self.ratios:
            idx += 1
            self.d_tr.append(convtr_w(f"decoder.model.{idx}.convtr.convtr", r))
            idx += 1
            self.d_res.append((conv_w(f"decoder.model.{idx}.block.1.conv.conv"), conv_w(f"decoder.model.{idx}.block.3.conv.conv")))
            idx += 1
        idx += 1
        wf = sd[f"decoder.model.{idx}.conv.conv.weight"]  # [1, nf, k]
        self.d_final_w = wf[0].t().contiguous().reshape(-1)  # (tap, ci)
        self.d_final_b = sd[f"decoder.model.{idx}.conv.conv.bias"]
        self.up_w = sd["upsample.convtr.convtr.convtr.weight"].reshape(D, -1).contiguous()
        # ---- transformers
        self.tr = {}
        for side in ("encoder_transformer", "decoder_transformer"):
            layers = []
            for l in range(m.num_layers):
                p = f"{side}.transformer.layers.{l}"
                layers.append(dict(
                    in_w=sd[f"{p}.self_attn.in_proj_weight"].t().contiguous(),
                    out_w=sd[f"{p}.self_attn.out_proj.weight"].t().contiguous(),
                    n1w=sd[f"{p}.norm1.weight"], n1b=sd[f"{p}.norm1.bias"],
                    n2w=sd[f"{p}.norm2.weight"], n2b=sd[f"{p}.norm2.bias"],
                    w1=sd[f"{p}.linear1.weight"].t().contiguous(), w2=sd[f"{p}.linear2.weight"].t().contiguous(),
                    ls1=sd[f"{p}.layer_scale_1.scale"], ls2=sd[f"{p}.layer_scale_2.scale"]))
            self.tr[side] = layers
        hd = D // m.num_heads
        # freqs exactly as rope.py:36-37 evaluates them (fp32 tensor * python scalar, then exp)
        ds = torch.arange(hd // 2, dtype=torch.float32)
        self.freqs = torch.exp(ds * (-math.log(m.max_period) * 2 / hd)).to(device)
        # ---- quantizer
        cd = m.codebook_dim
        w1 = sd["quantizer.rvq_first.input_proj.weight"].reshape(cd, D)
        w2 = sd["quantizer.rvq_rest.input_proj.weight"].reshape(cd, D)
        self.q_in_w = torch.cat([w1.t(), w2.t()], 1).contiguous()            # [D, 2cd]
        o1 = sd["quantizer.rvq_first.output_proj.weight"].reshape(D, cd)
        o2 = sd["quantizer.rvq_rest.output_proj.weight"].reshape(D, cd)
        self.q_out_w = torch.cat([o1.t(), o2.t()], 0).contiguous()           # [2cd, D]
        prefixes = [f"quantizer.rvq_first.vq.layers.{i}._codebook" for i in range(m.n_q_semantic)]
        prefixes += [f"quantizer.rvq_rest.vq.layers.{i}._codebook" for i in range(m.n_q - m.n_q_semantic)]
        # centroids = embedding_sum / cluster_usage.clamp(min=eps) (core_vq.py:142-150); the squared
        # norms are taken on the host with the same ATen ops the reference's cdist uses
        embs = [sd[f"{p}.embedding_sum"].cpu() / sd[f"{p}.cluster_usage"].cpu().clamp(min=m.codebook_eps)[:, None] for p in prefixes]
        E = torch.stack(embs)
        self.E = E.contiguous().to(device)
        self.Et = E.transpose(1, 2).contiguous().to(device)
        self.enorm = E.pow(2).sum(dim=-1).contiguous().to(device)
        self.zero_counter = torch.zeros(1, dtype=torch.int64, device=device)
        self._plans: Dict[tuple, object] = {}

    # ------------------------------------------------------------------ shared launch sequences
    def _transformer(self, side: str, X: _Buf, x_row: int, B: int, F: int, kv: List[torch.Tensor], cap: int,
                     offset: torch.Tensor, scratch) -> None:
        m = self.m
        D, H = self.D, m.num_heads
        hd = D // H
        ln, qkv, att, ff = scratch
        xo = X.off(x_row)
        for l, w in enumerate(self.tr[side]):
            kvl = kv[l] if len(kv) > 1 else kv[0]
            ops.layer_norm(X.t, xo, X.bs, w["n1w"], w["n1b"], ln, B, F, D, 1e-5)
            ops.gemm_rows(ln, 0, F * D, D, w["in_w"], qkv, 0, F * 3 * D, 3 * D, B, F)
            ops.rope_kv_append(qkv, kvl, offset, self.freqs, B, F, H, hd, cap)
            ops.ring_attention(qkv, kvl, offset, att, B, F, H, hd, cap, m.context, len(kv) == 1)
            ops.gemm_rows(att, 0, F * D, D, w["out_w"], X.t, xo, X.bs, D, B, F, scale=w["ls1"], R=X.t, r_off=xo,
                          r_bs=X.bs, r_rs=D)
            ops.layer_norm(X.t, xo, X.bs, w["n2w"], w["n2b"], ln, B, F, D, 1e-5)
            ops.gemm_rows(ln, 0, F * D, D, w["w1"], ff, 0, F * m.dim_feedforward, m.dim_feedforward, B, F, post_act=ACT_GELU)
            ops.gemm_rows(ff, 0, F * m.dim_feedforward, m.dim_feedforward, w["w2"], X.t, xo, X.bs, D, B, F, scale=w["ls2"],
                          R=X.t, r_off=xo, r_bs=X.bs, r_rs=D)

    # ------------------------------------------------------------------ plans
    def enc_plan(self, B: int, L: int, streaming: bool) -> "_EncPlan":
        key = ("enc", B, L, streaming)
        if key not in self._plans:
            if not streaming:  # keep a single batch-mode plan per kind alive
                for k in [k for k in self._plans if k[0] == "enc" and not k[3]]:
                    del self._plans[k]
            self._plans[key] = _EncPlan(self, B, L, streaming)
        return self._plans[key]

    def dec_plan(self, B: int, T: int, streaming: bool) -> "_DecPlan":
        key = ("dec", B, T, streaming)
        if key not in self._plans:
            if not streaming:
                for k in [k for k in self._plans if k[0] == "dec" and not k[3]]:
                    del self._plans[k]
            self._plans[key] = _DecPlan(self, B, T, streaming)
        return self._plans[key]

    # ------------------------------------------------------------------ non-streaming entry points
    def encode_batch(self, x: torch.Tensor) -> torch.Tensor:
        B, _, L = x.shape
        m = self.m
        if L == 0:
            return torch.empty((B, m.n_q, 0), dtype=torch.int64, device=self.device)
        per_stream = 4 * L * (self.nf * 3.6) + (1 << 20)
        sub = max(1, min(B, int(self.BATCH_MODE_BYTES // per_stream)))
        outs = []
        for b0 in range(0, B, sub):
            xb = x[b0:b0 + sub]
            plan = self.enc_plan(xb.shape[0], L, False)
            outs.append(plan.run(xb, None).clone())
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)

    def decode_batch(self, codes: torch.Tensor) -> torch.Tensor:
        B, K, T = codes.shape
        m = self.m
        if T == 0:
            return torch.empty((B, 1, 0), dtype=torch.float32, device=self.device)
        per_stream = 4 * T * m.frame_size * (self.nf * 3.6) + (1 << 20)
        sub = max(1, min(B, int(self.BATCH_MODE_BYTES // per_stream)))
        outs = []
        for b0 in range(0, B, sub):
            cb = codes[b0:b0 + sub].contiguous()
            plan = self.dec_plan(cb.shape[0], T, False)
            outs.append(plan.run(cb, None).clone())
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)


def _ceil_div(a: int, b: int) -> int:
    return -(-a // b)


class _EncPlan:
    """Buffers + launch order of one encode pass (whole clip, or one streaming chunk)."""

    def __init__(self, eng: _Engine, B: int, L: int, streaming: bool):
        m, dev = eng.m, eng.device
        self.eng, self.B, self.L, self.streaming = eng, B, L, streaming
        if streaming and L % m.frame_size != 0:
            raise RstnetError(f"streaming chunks must be multiples of {m.frame_size} samples, got {L}")
        nf, D = eng.nf, eng.D
        k0 = m.kernel_size
        self.Ts = [L]
        for r in eng.enc_ratios:
            self.Ts.append(_ceil_div(self.Ts[-1], r))
        T = self.Ts
        self.F = T[-1]
        s = m.resample_stride
        self.T5 = _ceil_div(self.F, s)
        self.xin = _Buf(B, k0 - 1, L, 0, 1, dev)
        self.y, self.h, self.r = [], [], []
        C = nf
        for i, ratio in enumerate(eng.enc_ratios):
            self.y.append(_Buf(B, m.residual_kernel_size - 1, T[i], 0, C, dev))
            self.h.append(_Buf(B, 0, T[i], 0, C // m.compress, dev))
            self.r.append(_Buf(B, ratio, T[i], T[i + 1] * ratio - T[i], C, dev))
            C *= 2
        self.y4 = _Buf(B, m.last_kernel_size - 1, self.F, 0, C, dev)
        self.xtr = _Buf(B, s, self.F, self.T5 * s - self.F, D, dev)
        F = self.F
        self.scratch = (torch.empty(B * F, D, device=dev), torch.empty(B * F, 3 * D, device=dev),
                        torch.empty(B * F, D, device=dev), torch.empty(B * F, m.dim_feedforward, device=dev))
        self.lat = torch.empty(B * self.T5, D, device=dev)
        self.xproj = torch.empty(B * self.T5, 2 * m.codebook_dim, device=dev)
        self.codes = torch.zeros(B, m.n_q, self.T5, dtype=torch.int64, device=dev)
        self.work = torch.empty(ops.rvq_encode_workspace(B * self.T5, m.n_q, m.codebook_dim, m.codebook_size),
                                dtype=torch.uint8, device=dev)
        hd = D // m.num_heads
        if streaming:
            self.cap = m.context
            self.kv = [torch.zeros(2, B, m.num_heads, self.cap, hd, device=dev) for _ in range(m.num_layers)]
            self.offset = torch.zeros(1, dtype=torch.int64, device=dev)
            carries = [self.xin] + self.y + self.r + [self.y4, self.xtr]
            self.n_copy = len(carries)
            self.copy_table = ops.make_copy_table([(b.t, b.bs, b.C, b.T, 0, b.ctx) for b in carries], dev)
        else:
            self.cap = F
            self.kv = [torch.zeros(2, B, m.num_heads, self.cap, hd, device=dev)]
            self.offset = eng.zero_counter
        self.graph: Optional[torch.cuda.CUDAGraph] = None

    def reset(self):
        assert self.streaming
        for b in [self.xin] + self.y + self.r + [self.y4, self.xtr]:
            b.t[:, :b.ctx].zero_()
        self.offset.zero_()

    def launch(self):
        eng, m, B = self.eng, self.eng.m, self.B
        T, D = self.Ts, eng.D
        k0 = m.kernel_size
        # conv0: 1 -> nf, k7
        ops.conv1d_cin1(self.xin.t, self.xin.bs, eng.e_conv0_w, eng.e_conv0_b, self.y[0].t, self.y[0].off(self.y[0].ctx),
                        self.y[0].bs, B, T[0], eng.nf, k0, ACT_NONE)
        C = eng.nf
        for i, ratio in enumerate(eng.enc_ratios):
            y, h, r = self.y[i], self.h[i], self.r[i]
            (w1, b1), (w2, b2) = eng.e_res[i]
            # SEANetResnetBlock: ELU -> k3 -> ELU -> k1, + skip; the ELU that follows is fused as post_act
            ops.gemm_rows(y.t, 0, y.bs, C, w1, h.t, 0, h.bs, h.C, B, T[i], bias=b1, pre_act=ACT_ELU, post_act=ACT_ELU)
            ops.gemm_rows(h.t, 0, h.bs, h.C, w2, r.t, r.off(r.ctx), r.bs, C, B, T[i], bias=b2, R=y.t, r_off=y.off(y.ctx),
                          r_bs=y.bs, r_rs=C, post_act=ACT_ELU)
            wd, bd = eng.e_down[i]
            nxt = self.y[i + 1] if i + 1 < len(self.y) else self.y4
            ops.gemm_rows(r.t, 0, r.bs, ratio * C, wd, nxt.t, nxt.off(nxt.ctx), nxt.bs, 2 * C, B, T[i + 1], bias=bd,
                          post_act=ACT_NONE if nxt is not self.y4 else ACT_ELU)
            C *= 2
        wf, bf = eng.e_final
        X = self.xtr
        ops.gemm_rows(self.y4.t, 0, self.y4.bs, C, wf, X.t, X.off(X.ctx), X.bs, D, B, self.F, bias=bf)
        eng._transformer("encoder_transformer", X, X.ctx, B, self.F, self.kv, self.cap, self.offset, self.scratch)
        # ConvDownsample1d: replicate padding (left on the first call only when streaming)
        ops.rows_fill(X.t, X.bs, B, D, 0, X.ctx, mode=1, src_row=X.ctx, only_if_zero=self.offset if self.streaming else None)
        if X.extra:
            ops.rows_fill(X.t, X.bs, B, D, X.ctx + X.T, X.extra, mode=1, src_row=X.ctx + X.T - 1)
        s = m.resample_stride
        ops.gemm_rows(X.t, 0, X.bs, s * D, eng.down_w, self.lat, 0, self.T5 * D, D, B, self.T5)
        cd = m.codebook_dim
        ops.gemm_rows(self.lat, 0, self.T5 * D, D, eng.q_in_w, self.xproj, 0, self.T5 * 2 * cd, 2 * cd, B, self.T5)
        ops.rvq_encode(self.xproj, 2 * cd, eng.E, eng.Et, eng.enorm, self.codes, self.work, B * self.T5, self.T5, m.n_q,
                       m.n_q_semantic, cd, m.codebook_size)
        if self.streaming:
            ops.rows_copy_table(self.copy_table, self.n_copy, B)
            ops.counter_add(self.offset, self.F)

    def run(self, x: torch.Tensor, graphs: Optional[bool]) -> torch.Tensor:
        self.xin.t[:, self.xin.ctx:self.xin.ctx + self.L, 0].copy_(x[:, 0, :])
        _run_plan(self, graphs)
        return self.codes


class _DecPlan:
    """Buffers + launch order of one decode pass."""

    def __init__(self, eng: _Engine, B: int, T: int, streaming: bool):
        m, dev = eng.m, eng.device
        self.eng, self.B, self.T, self.streaming = eng, B, T, streaming
        D, nf = eng.D, eng.nf
        s = m.resample_stride
        self.F = F = T * s
        cd = m.codebook_dim
        self.q = torch.empty(B * T, 2 * cd, device=dev)
        self.qup = _Buf(B, 1, T, 0, D, dev)
        self.xdec = _Buf(B, m.kernel_size - 1, F, 0, D, dev)
        C = nf * 2 ** len(eng.ratios)
        self.a = [_Buf(B, 1, F, 0, C, dev)]
        self.yd, self.hd = [], []
        Tin = F
        for i, r in enumerate(eng.ratios):
            Tout = Tin * r
            self.yd.append(_Buf(B, m.residual_kernel_size - 1, Tout, 0, C // 2, dev))
            self.hd.append(_Buf(B, 0, Tout, 0, C // 2 // m.compress, dev))
            last = i == len(eng.ratios) - 1
            self.a.append(_Buf(B, (m.last_kernel_size - 1) if last else 1, Tout, 0, C // 2, dev))
            C //= 2
            Tin = Tout
        self.Lout = Tin
        self.wav = torch.empty(B, 1, self.Lout, device=dev)
        self.scratch = (torch.empty(B * F, D, device=dev), torch.empty(B * F, 3 * D, device=dev),
                        torch.empty(B * F, D, device=dev), torch.empty(B * F, m.dim_feedforward, device=dev))
        self.codes_in = torch.zeros(B, m.n_q, T, dtype=torch.int64, device=dev)
        hd = D // m.num_heads
        if streaming:
            self.cap = m.context
            self.kv = [torch.zeros(2, B, m.num_heads, self.cap, hd, device=dev) for _ in range(m.num_layers)]
            self.offset = torch.zeros(1, dtype=torch.int64, device=dev)
            carries = [self.qup, self.xdec] + self.a + self.yd
            self.n_copy = len(carries)
            self.copy_table = ops.make_copy_table([(b.t, b.bs, b.C, b.T, 0, b.ctx) for b in carries], dev)
        else:
            self.cap = F
            self.kv = [torch.zeros(2, B, m.num_heads, self.cap, hd, device=dev)]
            self.offset = eng.zero_counter
        self.graph: Optional[torch.cuda.CUDAGraph] = None

    def reset(self):
        assert self.streaming
        for b in [self.qup, self.xdec] + self.a + self.yd:
            b.t[:, :b.ctx].zero_()
        self.offset.zero_()

    def launch(self):
        eng, m, B, T, F = self.eng, self.eng.m, self.B, self.T, self.F
        D, cd = eng.D, m.codebook_dim
        K = self.codes_in.shape[1]
        ops.rvq_decode_gather(self.codes_in, eng.E, self.q, B * T, T, K, m.n_q_semantic, cd, m.codebook_size)
        qup = self.qup
        ops.gemm_rows(self.q, 0, T * 2 * cd, 2 * cd, eng.q_out_w, qup.t, qup.off(1), qup.bs, D, B, T)
        X = self.xdec
        ops.convtr1d_depthwise(qup.t, qup.bs, eng.up_w, X.t, X.off(X.ctx), X.bs, B, T, D, m.resample_stride)
        eng._transformer("decoder_transformer", X, X.ctx, B, F, self.kv, self.cap, self.offset, self.scratch)
        w0, b0 = eng.d_conv0
        a = self.a[0]
        ops.gemm_rows(X.t, 0, X.bs, D, w0, a.t, a.off(1), a.bs, a.C, B, F, bias=b0, post_act=ACT_ELU)
        Tin = F
        for i, r in enumerate(eng.ratios):
            a, y, h, nxt = self.a[i], self.yd[i], self.hd[i], self.a[i + 1]
            wt, bt = eng.d_tr[i]
            Cin, Cout = a.C, y.C
            # ConvTranspose1d k=2r stride r as a GEMM over [x[t-1], x[t]]: one output row = r time steps
            ops.gemm_rows(a.t, 0, a.bs, Cin, wt, y.t, y.off(y.ctx), y.bs, r * Cout, B, Tin, bias=bt)
            Tout = Tin * r
            (w1, b1), (w2, b2) = eng.d_res[i]
            ops.gemm_rows(y.t, 0, y.bs, Cout, w1, h.t, 0, h.bs, h.C, B, Tout, bias=b1, pre_act=ACT_ELU, post_act=ACT_ELU)
            ops.gemm_rows(h.t, 0, h.bs, h.C, w2, nxt.t, nxt.off(nxt.ctx), nxt.bs, Cout, B, Tout, bias=b2, R=y.t,
                          r_off=y.off(y.ctx), r_bs=y.bs, r_rs=Cout, post_act=ACT_ELU)
            Tin = Tout
        last = self.a[-1]
        ops.conv1d_cout1(last.t, last.bs, eng.d_final_w, eng.d_final_b, self.wav, self.Lout, B, self.Lout, last.C,
                         m.last_kernel_size)
        if self.streaming:
            ops.rows_copy_table(self.copy_table, self.n_copy, B)
            ops.counter_add(self.offset, F)

    def run(self, codes: torch.Tensor, graphs: Optional[bool]) -> torch.Tensor:
        if codes.shape[1] != self.codes_in.shape[1]:
            self.codes_in = torch.zeros(self.B, codes.shape[1], self.T, dtype=torch.int64, device=self.eng.device)
            self.graph = None
        self.codes_in.copy_(codes)
        _run_plan(self, graphs)
        return self.wav


def _run_plan(plan, graphs: Optional[bool]) -> None:
    """Launch eagerly, or (streaming steps) replay a CUDA graph of the whole step: all buffers are
    static and the stream position lives in a device counter, so one capture serves every step."""
    if not graphs:
        plan.launch()
        return
    if plan.graph is None:
        plan.warmups = getattr(plan, "warmups", 0) + 1
        if plan.warmups <= 2:
            plan.launch()  # eager warm-up steps (also set the kernels' smem attributes)
            return
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            plan.launch()  # capture only; nothing executes until replay
        plan.graph = g
    plan.graph.replay()


class _StreamState:
    """Per-`streaming(B)` scope state: one encode plan and one decode plan per chunk size, sharing
    nothing with other scopes (mirrors `_MimiState`, compression.py:37-60)."""

    def __init__(self, eng: _Engine, batch_size: int):
        self.eng, self.B = eng, batch_size
        self.enc: Dict[int, _EncPlan] = {}
        self.dec: Dict[int, _DecPlan] = {}

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        B, _, L = x.shape
        if B != self.B:
            raise RstnetError(f"streaming batch size is {self.B}, got {B}")
        if L == 0:
            return torch.empty((B, self.eng.m.n_q, 0), dtype=torch.int64, device=self.eng.device)
        if L not in self.enc:
            if self.enc:
                raise RstnetError("the chunk size must stay constant within one streaming scope")
            self.enc[L] = _EncPlan(self.eng, B, L, True)
        return self.enc[L].run(x, self.eng.m.use_cuda_graphs).clone()

    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        B, K, T = codes.shape
        if B != self.B:
            raise RstnetError(f"streaming batch size is {self.B}, got {B}")
        if T not in self.dec:
            if self.dec:
                raise RstnetError("the chunk size must stay constant within one streaming scope")
            self.dec[T] = _DecPlan(self.eng, B, T, True)
        return self.dec[T].run(codes, self.eng.m.use_cuda_graphs).clone()

    def reset(self):
        for p in list(self.enc.values()) + list(self.dec.values()):
            p.reset()
