"""Vocabulary byte table for the device detokeniser (`t1_detokenize`, csrc/answer_scan.cu).

The reference turns completion ids into text with `processing_class.batch_decode(completion_ids,
skip_special_tokens=True)` (src/time_r1/rl/timer1_trainer.py:642-644).  For the Qwen2 tokenizer (byte-level BPE) that is:
drop the special added tokens, map every remaining token's characters back to bytes through the GPT-2 byte<->unicode
table (a token with a character outside that table contributes its own UTF-8 bytes instead), concatenate, and decode the
bytes as UTF-8 with replacement.  The first three steps are a table lookup per token — this module builds the table once
on the host and keeps it in HBM; the kernel does the lookups and the concatenation; the scanner (`t1_parse_answers`) then
works on the bytes, treating invalid UTF-8 the way the replacement decoding would.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Optional, Sequence

import numpy as np


def byte_level_alphabet() -> dict:
    """char -> byte of the GPT-2 / tokenizers `ByteLevel` alphabet: printable latin-1 bytes stand for themselves, the
    other 68 byte values are mapped to U+0100.. in increasing byte order."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    table, extra = {}, 0
    for b in range(256):
        if b in keep:
            table[chr(b)] = b
        else:
            table[chr(256 + extra)] = b
            extra += 1
    return table


_ALPHABET = byte_level_alphabet()


def token_bytes(token: str) -> bytes:
    """Bytes one vocabulary entry contributes to the decoded string (tokenizers' ByteLevel decoder, per token)."""
    out = bytearray()
    for ch in token:
        b = _ALPHABET.get(ch)
        if b is None:
            return token.encode("utf-8")
        out.append(b)
    return bytes(out)


@dataclass
class DetokTable:
    """offsets [V + 1] int32, data [total] uint8, skip [V] uint8 — device tensors (or numpy on the host before .to())."""
    offsets: object
    data: object
    skip: object
    vocab_size: int
    max_token_bytes: int

    @staticmethod
    def from_byte_strings(entries: Sequence[bytes], special_ids: Iterable[int] = (), device=None) -> "DetokTable":
        V = len(entries)
        lens = np.fromiter((len(e) for e in entries), np.int64, V)
        off = np.zeros(V + 1, np.int64)
        np.cumsum(lens, out=off[1:])
        if off[-1] >= 2 ** 31:
            raise ValueError("vocabulary byte table exceeds int32 offsets")
        data = np.frombuffer(b"".join(entries) + b"\0", np.uint8).copy()      # never empty
        skip = np.zeros(V, np.uint8)
        for i in special_ids:
            if 0 <= int(i) < V:
                skip[int(i)] = 1
        t = DetokTable(off.astype(np.int32), data, skip, V, int(lens.max()) if V else 0)
        return t.to(device) if device is not None else t

    @staticmethod
    def supports(tokenizer) -> bool:
        """True when batch_decode of this tokenizer IS the table walk: a byte-level BPE decoder (tokenizers `ByteLevel`, or a
        tokenizer that declares `is_byte_level_bpe = True`), a readable vocabulary and no post-hoc clean-up of the decoded
        string (`clean_up_tokenization_spaces`, off for Qwen2 / Qwen2.5-VL)."""
        tok = getattr(tokenizer, "tokenizer", tokenizer)
        if tok is None or not (hasattr(tok, "get_vocab") and hasattr(tok, "all_special_ids")):
            return False
        if getattr(tok, "clean_up_tokenization_spaces", False):
            return False
        if getattr(tok, "is_byte_level_bpe", False):
            return True
        backend = getattr(tok, "backend_tokenizer", None)
        decoder = getattr(backend, "decoder", None)
        return type(decoder).__name__ == "ByteLevel"

    @staticmethod
    def from_tokenizer(tokenizer, device=None, vocab_size: Optional[int] = None) -> "DetokTable":
        """Build the table from an HF tokenizer (or a processor holding one) whose decoder is byte-level BPE.
        ids without a vocabulary entry (the padding of the embedding matrix) decode to nothing."""
        tok = getattr(tokenizer, "tokenizer", tokenizer)
        vocab = tok.get_vocab()                                   # includes the added tokens
        V = max(max(vocab.values()) + 1, int(vocab_size or 0))
        entries = [b""] * V
        for token, i in vocab.items():
            entries[i] = token_bytes(token)
        return DetokTable.from_byte_strings(entries, tok.all_special_ids, device)

    def to(self, device) -> "DetokTable":
        import torch

        conv = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(device)
        return DetokTable(conv(self.offsets), conv(self.data), conv(self.skip), self.vocab_size, self.max_token_bytes)

    # host restatement of the kernel's table walk — used by the CPU tests to pin the table against `tokenizers`
    def decode_bytes_host(self, ids: Sequence[int], skip_special_tokens: bool = True) -> bytes:
        off, data, skip = (np.asarray(a.cpu() if hasattr(a, "cpu") else a) for a in (self.offsets, self.data, self.skip))
        out = bytearray()
        for i in ids:
            i = int(i)
            if 0 <= i < self.vocab_size and not (skip_special_tokens and skip[i]):
                out += data[off[i]:off[i + 1]].tobytes()
        return bytes(out)
