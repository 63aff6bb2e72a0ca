"""CPU oracle: restatement of the offline TTS generation loop `InferenceImp.__call__` + `reverse_delay`.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): never imported by rstnet_b200/.

Reference: MLLM_v2/infer_no_streaming.py:149-323.  The loop is restated exactly as upstream runs it -- the whole
prefix goes through the NON-streaming `forward_global` every generated frame and `forward_local` is evaluated 8 times
per frame (O(T^2)); oracle/lm_oracle.forward_global_full / forward_local are those two functions.  Only deterministic
token rules are restated (torch's Philox stream is not reproduced):
  * use_sampling False          -> argmax over the whole card (the *_2048 / 2049 masks only exist on the sampling path,
                                   utils/sampling.py:107-154);
  * use_sampling True, top_k 1  -> softmax, mask ids >= 2048 / 2049 (-inf), topk(1): deterministic, exercises the masks.
Pinned bit for bit against the unmodified reference by oracle/gen_golden_lm.py (fp32 and bf16).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import lm_oracle as L

TEXT_PAD, ACOUSTIC_PAD, SEMANTIC_PAD, TEXT_EMPTY = 128003, 2049, 2049, 128002


def reverse_delay(x: torch.Tensor) -> torch.Tensor:
    """infer_no_streaming.py:311-323."""
    if x.shape[0] != 8:
        x = x.transpose(0, 1)
    x_new = torch.ones_like(x)
    x_new[0, :-1] = x[0, :-1]
    x_new[1:, :-1] = x[1:, 1:]
    return x_new[:, :-1]


def _pick(logits: torch.Tensor, use_sampling: bool, n_valid: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """sample_token / sample_token_audio[_2048] in their deterministic forms.  Returns (token [...], margin [...]):
    margin = top-1 minus top-2 of the quantity the decision is made on (fp32 logits)."""
    lg = logits.float()
    if use_sampling:       # top_k == 1: topk(softmax(l / temp) with ids >= n_valid at -inf, 1)
        probs = torch.softmax(lg, dim=-1)
        probs[..., n_valid:] = float("-inf")
        tok = torch.topk(probs, 1, dim=-1).indices[..., 0]
        lg = lg.clone()
        lg[..., n_valid:] = float("-inf")
    else:
        tok = torch.argmax(lg, dim=-1)
    top2 = torch.topk(lg, 2, dim=-1).values
    return tok, top2[..., 0] - top2[..., 1]


def _deficit(logits: torch.Tensor, tok: torch.Tensor, use_sampling: bool, n_valid: int) -> torch.Tensor:
    """How far below the best allowed candidate the forced token's logit is (0 = it is an argmax; inf = not allowed)."""
    lg = logits.float().reshape(-1).clone()
    if use_sampling:
        lg[n_valid:] = float("-inf")
    return lg.max() - lg[int(tok)]


def inference_imp_tts(w: L.W, cfg: L.LMConfig, seq: torch.Tensor, use_sampling: bool,
                      force: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """InferenceImp.__call__ for task 'TTS', n_samples == 1 (infer_no_streaming.py:169-308).  seq [9, L] int64.
    Returns {"codes": [8, G-1] (after reverse_delay), "frames": [G, 9] raw tokens per generated frame (text, audio 0..7),
    "margins": [G, 9]}.
    force [G, 9]: teacher-force these decisions instead of the oracle's own (the loop is closed, so one near-tie flip
    changes everything after it: a candidate implementation's tokens are checked decision by decision); the result then
    also holds "deficit" [G, 9] = best allowed logit minus the forced token's logit under the oracle."""
    seq = seq.unsqueeze(0)
    pad_len = int(seq[0, 1:2, :].eq(SEMANTIC_PAD).int().sum().item())
    seq = seq[:, :, : seq.shape[2] - pad_len]
    prefix_len = seq.shape[2] - int(seq[0, 0, :].eq(TEXT_EMPTY).int().sum().item())
    prefix = seq[:, :, :prefix_len]
    maxlen = minlen = seq.shape[2] - prefix_len
    init = torch.full((1, cfg.n_q + 1, 1), cfg.audio_card, dtype=torch.long)
    init[:, 0] = 151655
    pre_gen_len = prefix.shape[2]
    frames: List[torch.Tensor] = []
    margins: List[torch.Tensor] = []
    deficits: List[torch.Tensor] = []
    final = []
    for g_idx in range(maxlen):
        g_len = prefix.shape[2]
        global_prefix = torch.cat([init.expand(prefix.shape[0], -1, -1), prefix], dim=-1)
        transformer_out, text_logits = L.forward_global_full(w, cfg, global_prefix)
        local_pad = torch.ones_like(prefix[:, :, 0:1]) * cfg.audio_card
        prefix = torch.cat([prefix, local_pad], dim=-1)
        text_tok, text_m = _pick(text_logits[:, -1:, :], use_sampling, text_logits.shape[-1])
        defs = []
        if force is not None:
            text_tok = force[g_idx, 0].reshape(1, 1)
            defs.append(_deficit(text_logits[:, -1:, :], text_tok, use_sampling, text_logits.shape[-1]))
        prefix[:, 0, -1] = text_tok.squeeze()
        toks, ms = [text_tok.reshape(())], [text_m.reshape(())]
        audio_seq = []
        for l_idx in range(8):
            local_start = L.scaled_embedding(prefix[:, 0, :], w["codecformer_text_emb.weight"])
            logits = L.forward_local(w, cfg, local_start, prefix[:, 1:, :], transformer_out)
            valid = logits[:, -1:, l_idx:l_idx + 1, :]
            if g_len == pre_gen_len or (l_idx > 0 and g_len > minlen):
                n_valid = 2049
            else:
                n_valid = 2048
            nxt, m = _pick(valid, use_sampling, n_valid)
            if force is not None:
                nxt = force[g_idx, l_idx + 1].reshape(1, 1, 1)
                defs.append(_deficit(valid, nxt, use_sampling, n_valid))
            audio_seq.append(nxt.squeeze())
            prefix[:, l_idx + 1, g_len] = nxt.squeeze()
            toks.append(nxt.reshape(()))
            ms.append(m.reshape(()))
        final.append(torch.stack(audio_seq))
        frames.append(torch.stack(toks))
        margins.append(torch.stack(ms))
        if force is not None:
            deficits.append(torch.stack(defs))
    final_t = torch.stack(final, dim=0)
    out = {"codes": reverse_delay(final_t), "frames": torch.stack(frames), "margins": torch.stack(margins)}
    if force is not None:
        out["deficit"] = torch.stack(deficits)
    return out
