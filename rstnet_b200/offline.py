"""Offline codec tokenization and reconstruction drivers.

  * `tokenize_utterances` / `python -m rstnet_b200.offline tokenize`: the Mimi branch of
    MLLM_v2/egs/pretraining/local/offline_codec_tokenization.py (and tools/data_scripts/offline_tokenization.py): every
    utterance -> int16 codes [8, T], collected in a dict `utt_id -> tensor` and written with `torch.save` -- the on-disk
    format the reference's data loader reads (tools/tokenizer/MimiCodec/mimi_tokenizer.py:44,72).  Unlike the reference
    (one clip per call) clips of EQUAL length are encoded as one batch (>= 96 of them: on the tensor cores); clips are
    never padded to a common length, because the codec's convs zero-pad each LAYER's input at the end of a clip
    (modules/conv.py:245-254), so audio padding would change a clip's last frame.
  * `reconstruct_directory` / `python -m rstnet_b200.offline reconstruct`: AudioCodec/MimiCodec/inference.py:111-148 -- every
    wav of a directory through encode -> decode, written under the same name.

24 kHz mono PCM wav in / out through scipy.io.wavfile (this image has neither torchaudio nor soundfile); other sample rates
are rejected rather than resampled (the reference resamples with torchaudio / julius).
"""
from __future__ import annotations

import argparse
import os
import sys
from collections import defaultdict
from typing import Dict, Iterable, Tuple

import numpy as np
import torch

from .codec import MimiCodec


def _as_row(wav: torch.Tensor) -> torch.Tensor:
    wav = torch.as_tensor(wav, dtype=torch.float32)
    if wav.dim() == 2:
        if wav.shape[0] != 1:
            raise ValueError(f"mono audio expected, got {tuple(wav.shape)}")
        wav = wav[0]
    if wav.dim() != 1:
        raise ValueError(f"expected [L] or [1, L] audio, got {tuple(wav.shape)}")
    return wav


@torch.no_grad()
def tokenize_utterances(codec: MimiCodec, items: Iterable[Tuple[str, torch.Tensor]], batch_size: int = 256) -> Dict[str, torch.Tensor]:
    """{utt_id: int16 [n_q, ceil(L / 1920)]} -- identical to MimiTokenizer.tokenize on every clip."""
    dev = codec.device
    by_len = defaultdict(list)
    for utt, wav in items:
        w = _as_row(wav)
        if w.numel():
            by_len[w.numel()].append((utt, w))
    out: Dict[str, torch.Tensor] = {}
    for L, group in by_len.items():
        for i in range(0, len(group), batch_size):
            part = group[i:i + batch_size]
            x = torch.stack([w for _, w in part])[:, None].to(dev)            # [B, 1, L]
            codes = codec.encode(x).to(torch.int16).cpu()                    # [B, n_q, T]
            for (utt, _), c in zip(part, codes):
                out[utt] = c.clone()
    return out


def save_tokens(tokens: Dict[str, torch.Tensor], path: str) -> None:
    torch.save(tokens, path)


def read_wav(path: str) -> Tuple[torch.Tensor, int]:
    from scipy.io import wavfile
    sr, data = wavfile.read(path)
    if data.ndim == 2:
        data = data.mean(axis=1)
    if np.issubdtype(data.dtype, np.integer):
        data = data.astype(np.float32) / float(np.iinfo(data.dtype).max)
    return torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)), int(sr)


def write_wav(path: str, wav: torch.Tensor, sr: int = 24000) -> None:
    from scipy.io import wavfile
    w = wav.detach().cpu().float().clamp(-1.0, 1.0).numpy()
    wavfile.write(path, sr, (w * 32767.0).astype(np.int16))


@torch.no_grad()
def reconstruct_directory(codec: MimiCodec, src: str, dst: str) -> int:
    """inference.py:test_batch -- wav -> codes -> wav for every file of `src`."""
    os.makedirs(dst, exist_ok=True)
    n = 0
    for name in sorted(os.listdir(src)):
        if not name.lower().endswith(".wav"):
            continue
        wav, sr = read_wav(os.path.join(src, name))
        if sr != codec.sample_rate:
            raise ValueError(f"{name}: {sr} Hz; resample to {codec.sample_rate} Hz first")
        codes = codec.encode(wav[None, None].to(codec.device))
        rec = codec.decode(codes)[0, 0, : wav.numel()]
        if float(rec.abs().max()) > 0.99:
            print(f"Clipping!! {name}: max scale {float(rec.abs().max()):.3f}", file=sys.stderr)   # inference.py:check_clipping2
        write_wav(os.path.join(dst, name), rec, codec.sample_rate)
        n += 1
    return n


def _load_codec(args) -> MimiCodec:
    import json
    cfg = json.load(open(args.config)) if args.config else dict(encoder_rates=[8, 6, 5, 4], codebook_size=2048, codebook_dim=256, rvq_layers=8)
    m = MimiCodec(**cfg)
    sd = torch.load(args.weights, map_location="cpu")
    m.load_state_dict(sd.get("codec_model", sd), strict=False)
    return m.to(args.device).eval()


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    for name in ("tokenize", "reconstruct"):
        p = sub.add_parser(name)
        p.add_argument("--weights", required=True, help="checkpoint (state_dict, or {'codec_model': state_dict})")
        p.add_argument("--config", default=None, help="json with the MimiCodec constructor arguments")
        p.add_argument("--device", default="cuda")
    sub.choices["tokenize"].add_argument("--wav-scp", required=True, help="kaldi wav.scp: <utt_id> <path>")
    sub.choices["tokenize"].add_argument("--output-file", required=True)
    sub.choices["tokenize"].add_argument("--batch-size", type=int, default=256)
    sub.choices["reconstruct"].add_argument("--input", required=True)
    sub.choices["reconstruct"].add_argument("--output", required=True)
    args = ap.parse_args(argv)
    codec = _load_codec(args)
    if args.cmd == "tokenize":
        def items():
            for line in open(args.wav_scp):
                utt, path = line.strip().split(None, 1)
                wav, sr = read_wav(path)
                if sr != codec.sample_rate:
                    raise ValueError(f"{utt}: {sr} Hz; resample to {codec.sample_rate} Hz first")
                yield utt, wav
        toks = tokenize_utterances(codec, items(), args.batch_size)
        save_tokens(toks, args.output_file)
        print(f"tokenized {len(toks)} utterances -> {args.output_file}")
    else:
        print(f"reconstructed {reconstruct_directory(codec, args.input, args.output)} files -> {args.output}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
