"""The device answer scanner's code (time_r1_b200/csrc/answer_scan.cuh), compiled for the host by a test-only harness
(tests/native/answer_scan_host.cpp), against CPython's `re` / `float()` running the reference's own patterns
(main.py:122-142, 234-239; evaluate.py:125-151; calc_difficulty.py:40-46).  Bit-exact: every parsed double must have the
same 64 bits.  The same corpora are replayed through the CUDA kernel in tests/test_answer_scan_gpu.py."""
import ctypes
import os
import re
import subprocess
import time

import numpy as np
import pytest

import scan_cases as sc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("as_host") / "answer_scan_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "time_r1_b200", "csrc"),
                    os.path.join(ROOT, "tests", "native", "answer_scan_host.cpp"), "-o", so], check=True)
    return ctypes.CDLL(so)


def _pack(rows):
    off = np.zeros(len(rows) + 1, np.int64)
    off[1:] = np.cumsum([len(r) for r in rows])
    buf = np.frombuffer(b"".join(rows) + b"\0", np.uint8).copy()
    return buf, off


def scan(host, rows, mode):
    buf, off = _pack(rows)
    n = len(rows)
    s, e = np.zeros(n), np.zeros(n)
    v, f = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    host.as_scan_batch(P(buf), P(off), n, mode, P(s), P(e), P(v), P(f))
    return s, e, v, f


def floats(host, rows, slow=False):
    buf, off = _pack(rows)
    out = np.zeros(len(rows))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    (host.as_float_slow_batch if slow else host.as_float_batch)(P(buf), P(off), len(rows), P(out))
    return out


def check_rows(host, rows, mode):
    rows = [r.encode("utf-8") if isinstance(r, str) else r for r in rows]
    sc.assert_rows_match(rows, mode, *scan(host, rows, mode))


# ---- character classes ----------------------------------------------------------------------------------------------
def test_digit_and_space_tables_match_python(host):
    host.as_digit_value.argtypes = [ctypes.c_uint32]
    host.as_is_space.argtypes = [ctypes.c_uint32]
    d, sp = re.compile(r"\d"), re.compile(r"\s")
    for cp in range(0x110000):
        if 0xD800 <= cp <= 0xDFFF:
            continue
        c = chr(cp)
        want = int(float(c)) if d.fullmatch(c) else -1
        assert host.as_digit_value(cp) == want, hex(cp)
        ws = bool(sp.fullmatch(c))
        assert ws == (("a" + c).strip() == "a")
        assert bool(host.as_is_space(cp)) == ws, hex(cp)


def test_ignorecase_only_matches_ascii_letters():
    for ch in "toand":
        r = re.compile(ch, re.IGNORECASE)
        hits = [cp for cp in range(0x110000) if not 0xD800 <= cp <= 0xDFFF and r.fullmatch(chr(cp))]
        assert hits == [ord(ch.upper()), ord(ch)]


# ---- float() ----------------------------------------------------------------------------------------------------------
def test_float_matches_cpython(host):
    cases = sc.float_cases()
    got_fast = floats(host, [c.encode() for c in cases])
    got_slow = floats(host, [c.encode() for c in cases], slow=True)
    for c, a, b in zip(cases, got_fast, got_slow):
        want = float(c)
        assert sc.bits(a) == sc.bits(want), (c[:60], len(c), a, want)
        assert sc.bits(b) == sc.bits(want), ("big-integer path", c[:60], len(c), b, want)


def test_float_unicode_digits(host):
    cases = ["٣.٥", "1٣", "\U0001D7D8\U0001D7D9.5", "１２３", "०" * 30 + "." + "৪" * 30]
    got = floats(host, [c.encode("utf-8") for c in cases])
    for c, a in zip(cases, got):
        assert sc.bits(a) == sc.bits(float(c)), c


# ---- the patterns -----------------------------------------------------------------------------------------------------
def test_hand_written_cases(host):
    for mode in (0, 1, 2):
        check_rows(host, sc.HAND, mode)


def test_reference_demo_probe(host):
    # SURVEY §8c probe of the unmodified reference: iou 0.529 for this completion vs GT (10, 20)
    s, e, v, f = scan(host, [b"<think>...</think>\n<answer>12.54 to 17.83</answer>"], 0)
    assert (s[0], e[0], v[0], f[0]) == (12.54, 17.83, 1, 1)


def test_fuzz_against_re(host):
    texts = sc.fuzz_texts()
    for mode in (0, 1, 2):
        check_rows(host, texts, mode)


def test_fuzz_invalid_utf8(host):
    rows = sc.fuzz_byte_rows()
    for mode in (0, 1, 2):
        check_rows(host, rows, mode)


def test_long_digit_runs_stay_linear(host):
    short = "<answer>" + "1" * 300 + ".5" + "</answer>"           # CPython's backtracking matcher is cubic in the run length
    check_rows(host, [short, short.replace(".5", " to 2")], 0)
    t = ("<answer>" + "1" * 2000000 + ".5" + "</answer>").encode()
    t0 = time.time()
    s, e, v, f = scan(host, [t, t.replace(b".5", b" to 2")], 0)
    assert time.time() - t0 < 5.0
    assert list(v) == [0, 1] and s[1] == float("inf") and e[1] == 2.0


def test_property_random_unicode_text(host):
    """hypothesis: arbitrary Unicode text interleaved with the tokens the patterns care about — the scanner and CPython's
    `re` must agree on every example (all three modes)."""
    hyp = pytest.importorskip("hypothesis")
    st = hyp.strategies

    piece = st.one_of(st.sampled_from(sc._PIECES), st.text(max_size=6),
                      st.text(alphabet=st.characters(whitelist_categories=("Nd", "Zs", "Cc")), max_size=4),
                      st.integers(0, 10 ** 25).map(str), st.floats(0, 1e6, allow_nan=False).map(lambda f: f"{f:.3f}"))
    texts = st.lists(piece, max_size=14).map("".join)

    @hyp.settings(max_examples=400, deadline=None, suppress_health_check=list(hyp.HealthCheck))
    @hyp.given(st.lists(texts, min_size=1, max_size=8))
    def run(batch):
        for mode in (0, 1, 2):
            check_rows(host, batch, mode)

    run()
