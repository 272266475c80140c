"""The vocabulary byte table behind t1_detokenize, pinned against the real `tokenizers` ByteLevel-BPE decoder that
processing_class.batch_decode(..., skip_special_tokens=True) runs in the reference (timer1_trainer.py:642-644)."""
import random

import numpy as np
import pytest


def _byte_level_tokenizer():
    from tokenizers import AddedToken, Tokenizer, decoders, models, pre_tokenizers

    alphabet = sorted(pre_tokenizers.ByteLevel.alphabet())
    vocab = {ch: i for i, ch in enumerate(alphabet)}
    merges = []
    # a few multi-byte entries: ASCII words, a space-prefixed word, a complete and a partial UTF-8 sequence
    from time_r1_b200.detok import byte_level_alphabet

    inv = {b: ch for ch, b in byte_level_alphabet().items()}
    words = [b"answer", b" to", b" and", b"think", b"</", b"12", b".5", "é".encode(), "٣".encode(), b"\xe2\x80", b"\xf0\x9f\x98",
             " ".encode(), b"\n\n", b"<answer>"[:4]]
    for w in words:
        chars = [inv[b] for b in w]
        cur = chars[0]
        for ch in chars[1:]:
            merges.append((cur, ch))
            cur += ch
            vocab.setdefault(cur, len(vocab))
    tok = Tokenizer(models.BPE(vocab=vocab, merges=merges))
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tok.decoder = decoders.ByteLevel()
    tok.add_special_tokens([AddedToken("<|endoftext|>", special=True), AddedToken("<|im_end|>", special=True)])
    tok.add_tokens([AddedToken("<answer>", special=False), AddedToken("</answer>", special=False), AddedToken("<tool_call>", special=False),
                    AddedToken("<日本>", special=False)])
    return tok


def test_alphabet_is_the_tokenizers_byte_level_alphabet():
    from tokenizers import decoders, pre_tokenizers

    from time_r1_b200.detok import byte_level_alphabet

    table = byte_level_alphabet()
    assert set(table) == set(pre_tokenizers.ByteLevel.alphabet()) and sorted(table.values()) == list(range(256))
    dec = decoders.ByteLevel()
    for ch, b in table.items():
        if b < 0x80:
            assert dec.decode([ch]) == chr(b), (ch, b)


def test_table_decodes_like_hf_batch_decode():
    from transformers import PreTrainedTokenizerFast

    from time_r1_b200.detok import DetokTable

    tok = _byte_level_tokenizer()
    hf = PreTrainedTokenizerFast(tokenizer_object=tok, eos_token="<|im_end|>", pad_token="<|endoftext|>", clean_up_tokenization_spaces=False)
    table = DetokTable.from_tokenizer(hf, vocab_size=hf.vocab_size + 40)      # padded embedding rows decode to nothing
    V = len(hf.get_vocab())
    assert table.vocab_size >= V and table.max_token_bytes >= len("<tool_call>")
    assert set(np.nonzero(table.skip)[0]) == set(hf.all_special_ids)
    rng = random.Random(3)
    rows = [[rng.randrange(V) for _ in range(rng.randint(0, 40))] for _ in range(400)]
    rows.append(hf.encode("<think>ok</think>\n<answer>12.5 to 17 é٣</answer><|im_end|><|endoftext|>"))
    for skip in (True, False):
        want = hf.batch_decode(rows, skip_special_tokens=skip)
        for ids, w in zip(rows, want):
            got = table.decode_bytes_host(ids, skip_special_tokens=skip).decode("utf-8", errors="replace")
            assert got == w, (ids, got, w)
    assert want[-1].startswith("<think>ok</think>\n<answer>12.5 to 17")


def test_token_bytes_falls_back_to_utf8_for_chars_outside_the_alphabet():
    from time_r1_b200.detok import token_bytes

    assert token_bytes("Ġto") == b" to" and token_bytes("<日本>") == "<日本>".encode() and token_bytes("Ċ") == b"\n"


def test_supports_only_byte_level_tokenizers_without_cleanup():
    from transformers import PreTrainedTokenizerFast

    from time_r1_b200.detok import DetokTable

    tok = _byte_level_tokenizer()
    hf = PreTrainedTokenizerFast(tokenizer_object=tok, eos_token="<|im_end|>", clean_up_tokenization_spaces=False)
    assert DetokTable.supports(hf)
    assert DetokTable.supports(type("Processor", (), {"tokenizer": hf})())            # an AutoProcessor holds the tokenizer
    cleaned = PreTrainedTokenizerFast(tokenizer_object=tok, eos_token="<|im_end|>", clean_up_tokenization_spaces=True)
    assert not DetokTable.supports(cleaned)

    from tokenizers import Tokenizer, decoders, models

    wp = Tokenizer(models.WordPiece(vocab={"[UNK]": 0, "a": 1, "##b": 2}, unk_token="[UNK]"))
    wp.decoder = decoders.WordPiece()
    assert not DetokTable.supports(PreTrainedTokenizerFast(tokenizer_object=wp, clean_up_tokenization_spaces=False))

    class Stub:                                   # the trainer tests' stub: decodes by template, no vocabulary
        def batch_decode(self, ids, skip_special_tokens=True):
            return []

    assert not DetokTable.supports(Stub()) and not DetokTable.supports(None)
