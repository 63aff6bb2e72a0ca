"""B200-native streaming decode step of the speech-text LM behind the reference's Python API.

Mirrors ``models.llama_streaming.GPT`` (MLLM_v2/models/llama_streaming.py:520-766) for the streaming
path (`with gpt.streaming(B):`, one frame = one position per call):
  * ``forward_global(seq[B,9,1]) -> (transformer_out[B,1,E], text_logits[B,1,V])``  (:665-692)
  * ``with gpt.codecformer.streaming(B): forward_codecformer(k, prev[B,1,1], transformer_out)`` (:727-749)
  * ``_get_initial_token``, ``codecformer_text_emb``, token-id properties, identical state_dict keys
    (LoRA merged / r == 0: ``...attn.attn.linear.weight`` etc., so reference checkpoints load).
plus ``forward_step`` -- the whole frame (temporal step, text sampling, 8 depth steps with sampling) as
one CUDA-graph replay; BASELINE.json names it although no such symbol exists upstream.

All arithmetic runs in librstnet_b200.so: tcgen05 weight-streaming GEMMs, ring decode attention, fused
norm / RoPE / gating / sampling kernels.  bf16 weights and activations, fp32 accumulation, exactly the
dtype recipe of `GPT(config).to(device, bfloat16)` (infer_no_streaming.py:104-105).
"""
from __future__ import annotations

import ctypes as C
from contextlib import contextmanager
from dataclasses import dataclass
from typing import Dict, Optional

import torch
from torch import nn

from . import _lib, ops
from ._lib import RstnetError
from .codec import _register


@dataclass
class Config:
    """The fields of models.llama_streaming.Config the decode path reads (same names / defaults)."""
    block_size: int = 4096
    n_layer: int = 16
    n_embd: int = 4096
    n_head: int = 32
    n_query_groups: Optional[int] = None
    head_size: Optional[int] = None
    intermediate_size: int = 11008
    norm_eps: float = 1e-5
    rope_base: int = 10000
    rotary_percentage: float = 1.0
    padded_vocab_size: int = 152064
    audio_card: int = 2048
    n_q: int = 9
    dep_q: int = 8
    codecformer_dim: int = 1024
    codecformer_heads: int = 32
    codecformer_layers: int = 6
    codecformer_dim_feedforward: int = 1024
    context: int = 3000

    def __post_init__(self):
        if self.head_size is None:
            self.head_size = self.n_embd // self.n_head
        if self.n_query_groups is None:
            self.n_query_groups = self.n_head

    @property
    def ff_hidden(self) -> int:  # modules/gating.py:40-43
        d, ff = self.codecformer_dim, self.codecformer_dim_feedforward
        return (21 * d) // 8 if ff == 4 * d else (2 * ff) // 3


class SkinnyGemm:
    """rstnet_skinny_gemm_* plan: out[m,n] = sum_k X[m,k] W[n,k] (+ R[m,n]), bf16, optionally with a fused finalize:
    norm_w/aux -> aux = RMSNorm(out) * norm_w (the next GEMM's pre-norm); silu_out -> silu_out = silu(a) * b."""

    def __init__(self, X: torch.Tensor, W: torch.Tensor, out: Optional[torch.Tensor], R: Optional[torch.Tensor],
                 ws: Optional[torch.Tensor], max_splits: int = 8, norm_w: Optional[torch.Tensor] = None,
                 aux: Optional[torch.Tensor] = None, eps: float = 0.0, kyutai: bool = False, silu_out: Optional[torch.Tensor] = None):
        M, K = X.shape
        N = W.shape[0]
        assert W.shape[1] == K and X.dtype == W.dtype == torch.bfloat16 and X.is_contiguous() and W.is_contiguous()
        if N % 4 != 0:
            ws = None  # split-K / fused finalize work on 4-element groups
        mode = 2 if silu_out is not None else (1 if norm_w is not None else 0)
        if mode and ws is None:
            raise RstnetError("fused finalize needs a workspace and N % 4 == 0")
        aux_t = silu_out if mode == 2 else aux
        self._keep = (X, W, out, R, ws, norm_w, aux_t)
        self._h = C.c_void_p()
        self.flops = 2.0 * M * N * K
        self.bytes = 2.0 * (N * K + M * K + M * N)
        p = lambda t: None if t is None else t.data_ptr()
        _lib.check(_lib.lib().rstnet_skinny_gemm_create_fused(p(X), p(W), p(R), p(out), p(ws), M, N, K,
                                                              max_splits if ws is not None else 1, mode, p(norm_w), p(aux_t),
                                                              float(eps), int(kyutai), C.byref(self._h)), "skinny_gemm_create")

    def run(self):
        _lib.check(_lib.lib().rstnet_skinny_gemm_run(self._h, ops._stream()), "skinny_gemm_run")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().rstnet_skinny_gemm_destroy(h)
            except Exception:
                pass
            self._h = None


class _DepthScope:
    """Stand-in for `gpt.codecformer` (a StreamingTransformer upstream): `with gpt.codecformer.streaming(B):`
    starts the per-frame depth state (llama_streaming.py:581: the subtree is fenced from the outer scope)."""

    def __init__(self, gpt: "GPT"):
        self._gpt = gpt

    @contextmanager
    def streaming(self, batch_size: int):
        st = self._gpt._state
        if st is None or st.B != batch_size:
            raise RstnetError("enter gpt.streaming(B) with the same batch size first")
        st.depth_step = 0
        try:
            yield
        finally:
            st.depth_step = None


class GPT(nn.Module):
    def __init__(self, config: Config, device=None, dtype=None):
        """device/dtype: create the (random-init) parameters directly there (a 7B model in bf16 on the GPU
        without a 28 GB fp32 host copy); default = CPU fp32 like the reference constructor."""
        super().__init__()
        c = self.config = config
        if c.n_query_groups != c.n_head:
            raise NotImplementedError("grouped-query configurations are not implemented yet (the 7B backbone is MHA)")
        if c.rotary_percentage != 1.0:
            raise NotImplementedError("partial rotary embeddings are not implemented")
        E, V, I, D, H = c.n_embd, c.padded_vocab_size, c.intermediate_size, c.codecformer_dim, c.ff_hidden
        g = torch.Generator(device=device if device is not None else "cpu").manual_seed(0)
        fk = dict(device=device, dtype=dtype)

        def w_(*shape):
            return torch.empty(*shape, **fk).normal_(0.0, 0.02, generator=g)

        def ones(*shape):
            return torch.ones(*shape, **fk)

        _register(self, "lm_head.linear.weight", w_(V, E))
        _register(self, "transformer.wte.weight", w_(V, E))
        for l in range(c.n_layer):
            p = f"transformer.h.{l}"
            _register(self, f"{p}.norm_1.weight", ones(E))
            _register(self, f"{p}.attn.attn.linear.weight", w_(3 * c.n_head * c.head_size, E))
            _register(self, f"{p}.attn.proj.linear.weight", w_(E, c.n_head * c.head_size))
            _register(self, f"{p}.norm_2.weight", ones(E))
            _register(self, f"{p}.mlp.fc_1.linear.weight", w_(I, E))
            _register(self, f"{p}.mlp.fc_2.linear.weight", w_(I, E))
            _register(self, f"{p}.mlp.proj.linear.weight", w_(E, I))
        _register(self, "transformer.ln_f.weight", ones(E))
        for i in range(c.n_q):
            _register(self, f"input_emb.{i}.weight", w_(c.audio_card + 1, E))
        for i in range(c.dep_q):
            _register(self, f"codecformer_in.{i}.weight", w_(D, E))
        for i in range(c.dep_q - 1):
            _register(self, f"codecformer_emb.{i}.weight", w_(c.audio_card + 1, D))
        _register(self, "codecformer_text_emb_.weight", w_(V, D))
        for l in range(c.codecformer_layers):
            p = f"codecformer_.layers.{l}"
            _register(self, f"{p}.self_attn.in_proj_weight", w_(c.dep_q * 3 * D, D))
            _register(self, f"{p}.self_attn.out_proj.weight", w_(c.dep_q * D, D))
            _register(self, f"{p}.norm1.alpha", ones(1, 1, D))
            _register(self, f"{p}.norm2.alpha", ones(1, 1, D))
            for k in range(c.dep_q):
                _register(self, f"{p}.gating.{k}.linear_in.weight", w_(2 * H, D))
                _register(self, f"{p}.gating.{k}.linear_out.weight", w_(D, H))
        for i in range(c.dep_q):
            _register(self, f"audio_linears.{i}.weight", w_(c.audio_card, D))
        self.max_seq_length = c.block_size
        self.codecformer = _DepthScope(self)
        self._state: Optional["_LMState"] = None
        self._packed = None
        self.use_cuda_graphs = True

    # ---- state_dict keys identical to the reference (`codecformer.` / `codecformer_text_emb.` subtrees are
    # stored under private attribute names because `codecformer` / `codecformer_text_emb` are API objects here)
    _RENAME = (("codecformer_.", "codecformer."), ("codecformer_text_emb_.", "codecformer_text_emb."))

    def state_dict(self, *a, **kw):
        sd = super().state_dict(*a, **kw)
        out = type(sd)()
        for k, v in sd.items():
            for src, dst in self._RENAME:
                if k.startswith(src):
                    k = dst + k[len(src):]
            out[k] = v
        return out

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        sd = {}
        old = {"lm_head.weight": "lm_head.linear.weight"}  # llama_streaming.py:762-766 compatibility mapping
        for k, v in state_dict.items():
            k = old.get(k, k)
            for src, dst in self._RENAME:
                if k.startswith(dst):
                    k = src + k[len(dst):]
            for a, b in ((".attn.weight", ".attn.linear.weight"), (".proj.weight", ".proj.linear.weight"),
                         (".fc_1.weight", ".fc_1.linear.weight"), (".fc_2.weight", ".fc_2.linear.weight")):
                if k.endswith(a) and k.startswith("transformer.h."):
                    k = k[: -len(a)] + b  # base-checkpoint names (llama_streaming.py:1000-1009, 1034-1043)
            sd[k] = v
        self._packed = None
        return super().load_state_dict(sd, strict=strict, **kw)

    def _apply(self, fn, *a, **kw):
        self._packed = None
        return super()._apply(fn, *a, **kw)

    # ---- token-id conventions (llama_streaming.py:590-634)
    @property
    def zero_token_id(self) -> int:
        return -1

    @property
    def text_initial_token_id(self) -> int:
        return 151655

    @property
    def initial_token_id(self) -> int:
        return self.config.audio_card

    @property
    def num_codebooks(self) -> int:
        return self.config.n_q + 1

    @property
    def num_audio_codebooks(self) -> int:
        return self.config.n_q

    @property
    def audio_offset(self) -> int:
        return 1

    @property
    def ungenerated_token_id(self) -> int:
        return -2

    @property
    def device(self):
        return next(iter(self.parameters())).device

    def _get_initial_token(self) -> torch.Tensor:
        tok = torch.full([1, self.num_codebooks, 1], self.initial_token_id, device=self.device, dtype=torch.long)
        tok[:, 0] = self.text_initial_token_id
        return tok

    def codecformer_text_emb(self, ids: torch.Tensor) -> torch.Tensor:
        w = dict(self.named_parameters())["codecformer_text_emb_.weight"]
        y = torch.nn.functional.embedding(ids.clamp(min=0), w)
        return torch.where((ids == self.zero_token_id)[..., None], torch.zeros(1, dtype=y.dtype, device=y.device), y)

    # ---- StreamingModule protocol
    def streaming_forever(self, batch_size: int):
        dev = self.device
        if dev.type != "cuda":
            raise RstnetError("GPT decode runs on CUDA only (sm_100a kernels; the CPU path is the reference itself)")
        if next(self.parameters()).dtype != torch.bfloat16:
            raise RstnetError("GPT decode runs in bfloat16: call .to(device, torch.bfloat16) as infer_no_streaming.py:104-105 does")
        self._state = _LMState(self, batch_size)

    @contextmanager
    def streaming(self, batch_size: int):
        self.streaming_forever(batch_size)
        try:
            yield
        finally:
            self._state = None

    def reset_streaming(self):
        if self._state is None:
            raise ValueError("Trying to reset streaming, but the model wasn't streaming.")
        self._state.reset()

    def _st(self) -> "_LMState":
        if self._state is None:
            raise RstnetError("only the streaming decode path is implemented: call inside `with gpt.streaming(B):` "
                              "(the full-sequence forward is the reference's own training/teacher-forcing path)")
        return self._state

    # ---- reference API
    @torch.no_grad()
    def forward_global(self, sequence: torch.Tensor):
        B, K, T = sequence.shape
        assert K == self.num_codebooks, f"Sequence shape {sequence.shape} must match the number of codebooks."
        if T != 1:
            raise RstnetError("streaming forward_global takes one frame per call (the reference's RoPE row select is only "
                              "correct for T == 1 too, llama_streaming.py:972-975)")
        return self._st().forward_global(sequence)

    @torch.no_grad()
    def forward_codecformer(self, codecformer_cb_index: int, sequence: torch.Tensor, transformer_out: torch.Tensor):
        B, K, S = sequence.shape
        assert K == 1, f"Codebooks for Depformer streaming should be passed 1 by 1, got {K}."
        assert S == 1, f"Steps for Depformer streaming should be passed 1 by 1, got {S}."
        assert transformer_out.shape[1] == 1, "Transformer out should be a for a single step."
        return self._st().forward_codecformer(codecformer_cb_index, sequence, transformer_out)

    @torch.no_grad()
    def forward_step(self, sequence: torch.Tensor, *, use_sampling: bool = True, temp_text: float = 0.7, top_k_text: int = 25,
                     temp: float = 0.8, top_k: int = 30, audio_valid=2049) -> torch.Tensor:
        """One generated frame: temporal step on sequence[B,9,1], text token, then the 8 depth steps, each sampled
        on the device (sample_token / sample_token_audio, utils/sampling.py:85-154).  Returns tokens [B, 9]
        (text, audio_0..7).  With use_cuda_graphs the whole frame is a single graph replay."""
        return self._st().forward_step(sequence, use_sampling, temp_text, top_k_text, temp, top_k, audio_valid)

    def forward(self, *a, **kw):
        raise NotImplementedError("training / teacher-forcing forward is out of scope; use the streaming decode API")


class _LMState:
    """Buffers, KV rings, GEMM plans of one `streaming(B)` scope."""

    def __init__(self, m: GPT, B: int):
        c, dev = m.config, m.device
        self.m, self.B, self.c = m, B, c
        bf = torch.bfloat16
        P = {k: v for k, v in m.named_parameters()}
        E, V, I, D, H = c.n_embd, c.padded_vocab_size, c.intermediate_size, c.codecformer_dim, c.ff_hidden
        nh, hs = c.n_head, c.head_size
        self.cap = c.context

        def z(*shape, dtype=bf):
            return torch.zeros(*shape, dtype=dtype, device=dev)

        # activations
        self.seq = z(B, c.n_q + 1, dtype=torch.int64)
        self.x, self.xn, self.q, self.att = z(B, E), z(B, E), z(B, nh * hs), z(B, nh * hs)
        self.qkv, self.ab, self.hmid = z(B, 3 * nh * hs), z(B, 2 * I), z(B, I)
        self.out, self.logits = z(B, E), z(B, V)
        self.tout = z(B, E)
        self.demb, self.dx, self.dn, self.datt = z(B, D), z(B, D), z(B, D), z(B, D)
        self.dqkv, self.dab, self.dh, self.dlogits = z(B, 3 * D), z(B, 2 * H), z(B, H), z(B, c.audio_card)
        self.tokens = z(B, c.dep_q + 1, dtype=torch.int64)
        self._idbuf = z(B, dtype=torch.int64)
        self.offset = z(1, dtype=torch.int64)
        self.frame_counter = z(1, dtype=torch.int64)
        # KV rings (k/v stored per head like the reference: [2,B,nh,cap,hs], lit_model.py:607-615)
        self.kv = [z(2, B, nh, self.cap, hs) for _ in range(c.n_layer)]
        hd = D // c.codecformer_heads
        self.dkv = [z(2, B, c.codecformer_heads, c.dep_q, hd) for _ in range(c.codecformer_layers)]
        # RoPE tables in the model dtype (the reference's buffers are cast by .to(bfloat16))
        theta = 1.0 / (c.rope_base ** (torch.arange(0, hs, 2).float() / hs))
        idx_theta = torch.outer(torch.arange(c.block_size) / 1, theta).repeat(1, 2)
        self.cos, self.sin = torch.cos(idx_theta).to(bf).to(dev).contiguous(), torch.sin(idx_theta).to(bf).to(dev).contiguous()
        # embedding table pointer array
        self.tables = [P[f"input_emb.{i}.weight"] for i in range(c.n_q)]
        self.table_ptrs = torch.tensor([t.data_ptr() for t in self.tables], dtype=torch.int64, device=dev)
        self.wte = P["transformer.wte.weight"]
        # split-K workspace shared by all GEMM plans (launches are stream-ordered)
        self.ws = torch.empty(8 * B * max(3 * E, 2 * I, 4096), dtype=torch.float32, device=dev)
        if m._packed is None:
            m._packed = {f"fc12.{l}": torch.cat([P[f"transformer.h.{l}.mlp.fc_1.linear.weight"],
                                                 P[f"transformer.h.{l}.mlp.fc_2.linear.weight"]], 0).contiguous()
                         for l in range(c.n_layer)}
        wsmax = max(3 * E, 2 * I, 4096)
        G = lambda X, W, out, R=None, **kw: SkinnyGemm(X, W, out, R, self.ws if W.shape[0] <= wsmax else None, **kw)
        L_ = c.n_layer
        n1 = [P[f"transformer.h.{l}.norm_1.weight"] for l in range(L_)]
        n2 = [P[f"transformer.h.{l}.norm_2.weight"] for l in range(L_)]
        self.ln_f = P["transformer.ln_f.weight"]
        self.n1_first = n1[0]
        self.layers = []
        for l in range(L_):
            p = f"transformer.h.{l}"
            last = l == L_ - 1
            self.layers.append(dict(
                qkv=G(self.xn, P[f"{p}.attn.attn.linear.weight"], self.qkv),
                # x = attn + x ; xn = norm_2(x)   (Block.forward, llama_streaming.py:846-849) in the GEMM's finalize
                proj=G(self.att, P[f"{p}.attn.proj.linear.weight"], self.x, self.x, norm_w=n2[l], aux=self.xn, eps=c.norm_eps),
                # hmid = silu(fc_1 x) * fc_2 x   (LLaMAMLP, lit_model.py:399-403)
                fc=G(self.xn, m._packed[f"fc12.{l}"], None, silu_out=self.hmid),
                # x = mlp + x ; xn = norm_1 of the next block (or ln_f -> transformer_out)
                down=G(self.hmid, P[f"{p}.mlp.proj.linear.weight"], self.x, self.x, norm_w=self.ln_f if last else n1[l + 1],
                       aux=self.out if last else self.xn, eps=c.norm_eps)))
        self.head = G(self.out, P["lm_head.linear.weight"], self.logits)
        # depth transformer: per-codebook-step weight slabs
        self.text_emb = P["codecformer_text_emb_.weight"]
        self.dep_emb = [P[f"codecformer_emb.{i}.weight"] for i in range(c.dep_q - 1)]
        Ld = c.codecformer_layers
        a1 = [P[f"codecformer_.layers.{l}.norm1.alpha"].view(-1) for l in range(Ld)]
        a2 = [P[f"codecformer_.layers.{l}.norm2.alpha"].view(-1) for l in range(Ld)]
        self.dsteps = []
        for k in range(c.dep_q):
            layers = []
            for l in range(Ld):
                p = f"codecformer_.layers.{l}"
                w_in = P[f"{p}.self_attn.in_proj_weight"].view(c.dep_q, 3 * D, D)[k]
                w_out = P[f"{p}.self_attn.out_proj.weight"].view(c.dep_q, D, D)[k]
                nxt = dict(norm_w=a1[l + 1], aux=self.dn, eps=1e-8, kyutai=True) if l + 1 < Ld else {}
                layers.append(dict(
                    qkv=G(self.dn, w_in, self.dqkv),
                    out=G(self.datt, w_out, self.dx, self.dx, norm_w=a2[l], aux=self.dn, eps=1e-8, kyutai=True),
                    gin=G(self.dn, P[f"{p}.gating.{k}.linear_in.weight"], None, silu_out=self.dh),
                    gout=G(self.dh, P[f"{p}.gating.{k}.linear_out.weight"], self.dx, self.dx, **nxt)))
            self.dsteps.append(dict(
                inp=G(self.tout, P[f"codecformer_in.{k}.weight"], self.dx, self.demb, norm_w=a1[0], aux=self.dn, eps=1e-8, kyutai=True),
                layers=layers, head=G(self.dx, P[f"audio_linears.{k}.weight"], self.dlogits)))
        self.depth_step: Optional[int] = None
        self.graphs: Dict[tuple, torch.cuda.CUDAGraph] = {}
        self.warm: Dict[tuple, int] = {}
        self.seed = 1234

    def reset(self):
        self.offset.zero_()
        self.frame_counter.zero_()
        self.depth_step = None

    # ---- launch sequences -------------------------------------------------------------------
    def _temporal(self):
        c, B, L = self.c, self.B, _lib.lib()
        st = ops._stream()
        E = c.n_embd
        _lib.check(L.rstnet_lm_embed_sum_bf16(self.seq.data_ptr(), c.n_q + 1, self.wte.data_ptr(), self.table_ptrs.data_ptr(),
                                              c.n_q, E, self.x.data_ptr(), B, st), "lm_embed_sum")
        _lib.check(L.rstnet_lm_rms_norm_bf16(self.x.data_ptr(), self.n1_first.data_ptr(), self.xn.data_ptr(), B, E, c.norm_eps, 0, st), "rms")
        for l, ly in enumerate(self.layers):
            ly["qkv"].run()
            _lib.check(L.rstnet_lm_rope_kv_append_bf16(self.qkv.data_ptr(), self.cos.data_ptr(), self.sin.data_ptr(),
                                                       self.offset.data_ptr(), self.q.data_ptr(), self.kv[l].data_ptr(), B,
                                                       c.n_head, c.head_size, self.cap, st), "rope_kv")
            _lib.check(L.rstnet_lm_ring_decode_attention_bf16(self.q.data_ptr(), self.kv[l].data_ptr(), self.offset.data_ptr(),
                                                              self.att.data_ptr(), B, c.n_head, c.head_size, self.cap, c.context, st),
                       "attention")
            ly["proj"].run()   # + residual + norm_2 -> xn
            ly["fc"].run()     # + SiLU gating -> hmid
            ly["down"].run()   # + residual + next pre-norm -> xn (last layer: ln_f -> transformer_out)
        self.head.run()
        ops.counter_add(self.offset, 1)

    def _depth(self, k: int, ids: torch.Tensor, id_stride: int):
        c, B, L = self.c, self.B, _lib.lib()
        st = ops._stream()
        D = c.codecformer_dim
        table = self.text_emb if k == 0 else self.dep_emb[k - 1]
        _lib.check(L.rstnet_lm_embed_rows_bf16(ids.data_ptr(), id_stride, table.data_ptr(), D, self.demb.data_ptr(), B, st), "embed_rows")
        ds = self.dsteps[k]
        ds["inp"].run()        # dx = in_k(transformer_out) + emb ; dn = norm1_0(dx)
        hd = D // c.codecformer_heads
        for l, ly in enumerate(ds["layers"]):
            ly["qkv"].run()
            _lib.check(L.rstnet_lm_depth_attention_bf16(self.dqkv.data_ptr(), self.dkv[l].data_ptr(), self.datt.data_ptr(), B,
                                                        c.codecformer_heads, hd, c.dep_q, k, st), "depth_attention")
            ly["out"].run()    # dx += out_k(att) ; dn = norm2(dx)
            ly["gin"].run()    # dh = silu(a) * b
            ly["gout"].run()   # dx += out(dh) ; dn = norm1 of the next layer
        ds["head"].run()

    def _sample(self, logits: torch.Tensor, V: int, n_valid: int, top_k: int, temp: float, col: int, salt: int):
        _lib.check(_lib.lib().rstnet_lm_sample_bf16(logits.data_ptr(), self.B, V, n_valid, top_k, float(temp), self.seed + salt,
                                                    self.frame_counter.data_ptr(), self.tokens.data_ptr() + 8 * col,
                                                    self.c.dep_q + 1, ops._stream()), "sample")

    def _replay(self, key, fn):
        if not self.m.use_cuda_graphs:
            fn()
            return
        g = self.graphs.get(key)
        if g is None:
            self.warm[key] = self.warm.get(key, 0) + 1
            if self.warm[key] <= 1:
                fn()
                return
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            self.graphs[key] = g
        g.replay()

    # ---- API ----------------------------------------------------------------------------------
    def forward_global(self, sequence: torch.Tensor):
        self.seq.copy_(sequence[:, :, 0])
        self._replay(("temporal",), self._temporal)
        c = self.c
        return self.out.view(self.B, 1, c.n_embd).clone(), self.logits.view(self.B, 1, c.padded_vocab_size).clone()

    def forward_codecformer(self, k: int, sequence: torch.Tensor, transformer_out: torch.Tensor):
        if self.depth_step is None:
            raise RstnetError("call inside `with gpt.codecformer.streaming(B):`")
        if k != self.depth_step:
            raise RstnetError(f"depth steps must run in order: expected {self.depth_step}, got {k}")
        self.tout.copy_(transformer_out[:, 0])
        self._idbuf.copy_(sequence[:, 0, 0])
        self._replay(("depth", k), lambda: self._depth(k, self._idbuf, 1))
        self.depth_step += 1
        return self.dlogits.view(self.B, 1, 1, self.c.audio_card).clone()

    def forward_step(self, sequence, use_sampling, temp_text, top_k_text, temp, top_k, audio_valid):
        c = self.c
        self.seq.copy_(sequence[:, :, 0])
        tk_text = top_k_text if use_sampling else 0
        tk = top_k if use_sampling else 0

        def frame():
            self._temporal()
            self._sample(self.logits, c.padded_vocab_size, c.padded_vocab_size, tk_text, temp_text, 0, 0)
            ops_t = self.out
            self.tout.copy_(ops_t)
            for k in range(c.dep_q):
                self._depth(k, self.tokens[:, k], c.dep_q + 1)
                av = audio_valid[k] if isinstance(audio_valid, (tuple, list)) else audio_valid
                self._sample(self.dlogits, c.audio_card, min(av, c.audio_card), tk, temp, k + 1, k + 1)
            ops.counter_add(self.frame_counter, 1)

        self._replay(("frame", tk_text, float(temp_text), tk, float(temp), tuple(audio_valid) if isinstance(audio_valid, (tuple, list)) else audio_valid), frame)
        return self.tokens.clone()
