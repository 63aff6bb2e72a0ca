"""Streaming O(1)-per-frame version of the reference's offline generation loop.

Mirrors ``InferenceImp`` and ``reverse_delay`` of MLLM_v2/infer_no_streaming.py:149-323 (same constructor and call
signature, same special token ids, same per-codebook sampling rules), but instead of re-running the whole prefix
through ``forward_global`` and re-running ``forward_local`` 8 times per generated frame (O(T^2) per utterance,
SURVEY.md §3.1) it streams: the prompt goes through ``GPT.prefill`` (multi-position chunks writing the KV rings in one
pass), and each generated frame is one ``GPT.forward_step`` (temporal step + 8 depth steps + device-side sampling, one
CUDA-graph replay).  The arithmetic per frame is the reference's: upstream evaluates the depth transformer through the
NON-streaming ``forward_local`` (every depth step sees all earlier keys), so ``forward_step`` runs with
``depth_ring_quirk=False`` here; the temporal transformer's non-streaming form equals the streamed one while the
sequence is shorter than ``config.context`` (tests/test_lm_gpu.py checks the loop against the reference's own tokens).
Only the 'TTS' task is runnable upstream (the other branches reference undefined variables); same here.
"""
from __future__ import annotations

import torch

from ._lib import RstnetError
from .lm import GPT


def reverse_delay(x: torch.Tensor) -> torch.Tensor:
    """Undo the one-frame acoustic delay (infer_no_streaming.py:311-323): x [8, L] (or [L, 8]) -> [8, L-1] with
    row 0 kept and rows 1..7 shifted left by one frame."""
    if x.shape[0] != 8:
        x = x.transpose(0, 1)
    out = torch.empty_like(x[:, :-1])
    out[0] = x[0, :-1]
    out[1:] = x[1:, 1:]
    return out


class InferenceImp(object):
    def __init__(self, args, model: GPT, mode, temp_text, top_k_text, temp, top_k, task_name):
        self.model, self.args = model, args
        self.n_samples = 1
        self.task_name = task_name
        self.text_pad_token = 128003
        self.acoustic_pad_token = 2049
        self.semantic_pad_token = 2049
        self.text_empty_token = 128002
        self.mode = mode
        self.use_sampling = True          # upstream hard-codes True (:162); set the attribute to False for argmax decoding
        self.temp_text, self.top_k_text, self.temp, self.top_k = temp_text, top_k_text, temp, top_k

    @torch.no_grad()
    def __call__(self, seq: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """seq [9, L] (one utterance, as upstream) -> codes [8, T'] after reverse_delay."""
        return self.generate(seq.unsqueeze(0).expand(self.n_samples, -1, -1))[0]

    @torch.no_grad()
    def generate(self, seq: torch.Tensor, return_frames: bool = False):
        """Batched form: seq [B, 9, L], all rows in the same TTS layout (same prompt length and number of frames to
        generate; row 0 defines them, as upstream reads `seq[0]`).  -> codes [B, 8, T'] (and the raw frames [B, G, 9])."""
        if self.task_name != "TTS":
            raise NotImplementedError("only task 'TTS' is runnable in the reference loop (infer_no_streaming.py:184-226)")
        if self.mode == "teacher-force":
            raise NotImplementedError("teacher-force mode is the training forward (out of scope)")
        m = self.model
        dev = seq.device
        pad_len = int(seq[0, 1:2, :].eq(self.semantic_pad_token).int().sum().item())
        seq = seq[:, :, : seq.shape[2] - pad_len]
        prefix_len = seq.shape[2] - int(seq[0, 0, :].eq(self.text_empty_token).int().sum().item())
        prefix = seq[:, :, :prefix_len]
        maxlen = minlen = seq.shape[2] - prefix_len
        if maxlen <= 0:
            raise RstnetError("nothing to generate: the sequence has no text-empty frames")
        if prefix_len <= 0:
            raise RstnetError("the sequence has no prompt frames")
        B = prefix.shape[0]
        pre_gen_len = prefix.shape[2]
        frames = []
        with m.streaming(B):
            # the init token + all prompt frames but the last only feed the KV rings (their outputs are never sampled);
            # the step on the last prompt frame yields generated frame 0
            init = m._get_initial_token().expand(B, -1, -1).to(dev)
            feed = torch.cat([init, prefix], dim=2)
            m.prefill(feed[:, :, :-1].contiguous())
            cur = feed[:, :, -1:].contiguous()
            for g_idx in range(maxlen):
                g_len = pre_gen_len + g_idx
                # per-codebook candidate sets (infer_no_streaming.py:264-283): 2049 ids on the first generated frame
                # and for codebooks > 0 once g_len > minlen, otherwise 2048
                valid = tuple(2049 if (g_len == pre_gen_len or (l > 0 and g_len > minlen)) else 2048 for l in range(8))
                toks = m.forward_step(cur, use_sampling=self.use_sampling, temp_text=self.temp_text, top_k_text=self.top_k_text,
                                      temp=self.temp, top_k=self.top_k, audio_valid=valid, depth_ring_quirk=False)
                frames.append(toks)
                cur = toks[:, :, None]
            m.check_device_errors()
        raw = torch.stack(frames, dim=1).to(dev)                          # [B, G, 9]
        codes = torch.stack([reverse_delay(raw[b, :, 1:]) for b in range(B)], 0)
        return (codes, raw) if return_frames else codes
