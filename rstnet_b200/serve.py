"""Batched real-time frame scheduler: the serving loop of MLLM_v2/moshi/server.py:44-166 (one websocket session,
batch 1: `mimi.encode(chunk)` -> `lm_gen.step(codes)` -> `mimi.decode(tokens)` per 80 ms of audio, :108-144) generalised to
a BATCH of sessions that share one streaming scope on one GPU (BASELINE cfg 5).

What the batch form needs beyond the reference's all-or-nothing streaming state, and where it lives:
  * admission of a new session into a free batch row of a LIVE scope   -> `reset_streaming(streams=[row])`
  * a row that delivered no audio this tick must keep its exact state  -> `set_active_streams(mask)` (the carry copy and
    the position counters of held rows do not advance; codec.py / lm.py)
  * no re-capture of CUDA graphs when sessions come and go            -> all of the above are device-side flags/counters
The websocket / Opus transport of server.py (:98-103, 141-153, 163) is out of scope (networking); sessions
push raw 24 kHz PCM chunks and receive (tokens, PCM) per tick.

`FrameScheduler` is pure host logic over an engine object with `reset_rows(rows)` and `step(pcm_rows, active) ->
{row: (tokens, pcm)}`; `DuplexEngine` is that engine for a MimiCodec + GPT pair on a GPU.
"""
from __future__ import annotations

import time
from collections import deque
from typing import Deque, Dict, Hashable, List, Optional, Tuple

import torch

from ._lib import RstnetError

FRAME_SAMPLES = 1920       # 80 ms at 24 kHz = one 12.5 Hz frame (moshi/server.py:57: sample_rate / frame_rate)
FRAME_SECONDS = 0.08


class FrameScheduler:
    """Rows of a fixed-capacity batch are leased to sessions.  Every `tick()` steps the engine once for all rows that
    have a full frame of audio queued; the others are held."""

    def __init__(self, engine, capacity: int):
        self.engine, self.capacity = engine, capacity
        self._row_of: Dict[Hashable, int] = {}
        self._free: List[int] = list(range(capacity))
        self._queue: Dict[Hashable, Deque] = {}
        self.ticks = 0

    # ---- session lifecycle
    def admit(self, session: Hashable) -> int:
        """Lease the lowest free row to `session` and restart that row's streaming state (server.py:156-158 does
        `mimi.reset_streaming(); lm_gen.reset_streaming()` for its single session)."""
        if session in self._row_of:
            raise RuntimeError(f"session {session!r} is already admitted")
        if not self._free:
            raise RuntimeError("no free row: the batch is full")
        self._free.sort()
        row = self._free.pop(0)
        self._row_of[session] = row
        self._queue[session] = deque()
        self.engine.reset_rows([row])
        return row

    def release(self, session: Hashable) -> None:
        row = self._row_of.pop(session)
        self._queue.pop(session, None)
        self._free.append(row)

    def sessions(self) -> Dict[Hashable, int]:
        return dict(self._row_of)

    def free_rows(self) -> int:
        return len(self._free)

    # ---- data path
    def push(self, session: Hashable, frame) -> None:
        """Queue one 80 ms frame (1920 samples) of the session's input audio."""
        self._queue[session].append(frame)

    def tick(self) -> Dict[Hashable, Tuple]:
        """One scheduler period: step every session that has a frame queued; returns {session: (tokens, pcm)}."""
        ready = {s: r for s, r in self._row_of.items() if self._queue[s]}
        self.ticks += 1
        if not ready:
            return {}
        pcm_rows = {r: self._queue[s].popleft() for s, r in ready.items()}
        out = self.engine.step(pcm_rows, sorted(pcm_rows))
        return {s: out[r] for s, r in ready.items()}


class DuplexEngine:
    """One streaming scope of a MimiCodec and a GPT for `capacity` sessions: per tick, for all rows at once,
    encode the sessions' 80 ms chunks -> one LM frame (temporal step + 8 depth steps + sampling) -> decode the generated
    codes (the three calls of server.py:128-136).  The LM input frame of a row is [its previous text token, the 8 codes of
    its input audio]; the generated audio codes are restricted to ids < 2048 (decodable)."""

    def __init__(self, codec, gpt, capacity: int, *, use_sampling: bool = True, temp_text: float = 0.7, top_k_text: int = 25,
                 temp: float = 0.8, top_k: int = 30):
        if capacity > 128:
            raise RstnetError("the LM step takes at most 128 streams per scope (one weight-streaming GEMM pass)")
        self.codec, self.gpt, self.B = codec, gpt, capacity
        self.dev = gpt.device
        self.sampling = dict(use_sampling=use_sampling, temp_text=temp_text, top_k_text=top_k_text, temp=temp, top_k=top_k)
        codec.streaming_forever(capacity)
        gpt.streaming_forever(capacity)
        self.pcm_in = torch.zeros(capacity, 1, FRAME_SAMPLES, dtype=torch.float32).pin_memory()
        self.pcm_dev = torch.zeros(capacity, 1, FRAME_SAMPLES, dtype=torch.float32, device=self.dev)
        self.prev_text = torch.full((capacity, 1, 1), gpt.text_initial_token_id, dtype=torch.int64, device=self.dev)
        self.tok_host = torch.zeros(capacity, gpt.config.dep_q + 1, dtype=torch.int64).pin_memory()
        self.pcm_host = torch.zeros(capacity, 1, FRAME_SAMPLES, dtype=torch.float32).pin_memory()
        self.mask_host = torch.zeros(capacity, dtype=torch.int64).pin_memory()
        self.latencies_ms: List[float] = []

    def reset_rows(self, rows) -> None:
        self.codec.reset_streaming(streams=list(rows))
        self.gpt.reset_streaming(streams=list(rows))
        self.prev_text[list(rows)] = self.gpt.text_initial_token_id

    @torch.no_grad()
    def step(self, pcm_rows: Dict[int, torch.Tensor], active: List[int]):
        t0 = time.perf_counter()
        self.mask_host.zero_()
        for r, chunk in pcm_rows.items():
            self.pcm_in[r, 0].copy_(torch.as_tensor(chunk, dtype=torch.float32).reshape(FRAME_SAMPLES))
            self.mask_host[r] = 1
        self.codec.set_active_streams(self.mask_host)
        self.gpt.set_active_streams(self.mask_host)
        self.pcm_dev.copy_(self.pcm_in, non_blocking=True)
        codes = self.codec.encode(self.pcm_dev)                                   # [B, 8, 1]
        frame = torch.cat([self.prev_text, codes], dim=1)                        # [B, 9, 1]
        toks = self.gpt.forward_step(frame, audio_valid=2048, **self.sampling)    # [B, 9]
        held = (self.mask_host == 0).to(self.dev)
        self.prev_text.copy_(torch.where(held[:, None, None], self.prev_text, toks[:, :1, None]))
        pcm = self.codec.decode(toks[:, 1:, None].clamp(max=self.codec.codebook_size - 1))   # [B, 1, 1920]
        self.tok_host.copy_(toks, non_blocking=True)
        self.pcm_host.copy_(pcm, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.latencies_ms.append(1e3 * (time.perf_counter() - t0))
        return {r: (self.tok_host[r].clone(), self.pcm_host[r, 0].clone()) for r in active}
