"""Shared case material for the answer-scanner tests (CPU harness and GPU kernel): the reference's own patterns run by
CPython (main.py:122-142, 234-239; evaluate.py:125-151; calc_difficulty.py:40-46) and the corpora both tests replay."""
import random
import re
import struct
from fractions import Fraction

# ---- the reference's patterns, verbatim semantics -------------------------------------------------------------------
_ANSWER = re.compile(r"<answer>(.*?)</answer>", re.DOTALL)
_SPAN_I = re.compile(r"(\d+\.?\d*) (to|and) (\d+\.?\d*)", re.IGNORECASE)
_SPAN = re.compile(r"(\d+\.?\d*) (to|and) (\d+\.?\d*)")
_FORMAT = re.compile(r"<think>.*?</think>\s*<answer>.*?</answer>", re.DOTALL)
_NUMBER = re.compile(r"\d+(?:\.\d+)?")


def ref_train(text):
    blocks = _ANSWER.findall(text)
    parsed = None
    if blocks:
        m = _SPAN_I.findall(blocks[-1])
        if m:
            parsed = (float(m[-1][0]), float(m[-1][2]))
    return parsed, bool(_FORMAT.fullmatch(text.strip()))


def ref_eval(text, force):
    from time_r1_b200.filtering import extract_answer, extract_answer_force

    p = extract_answer(text)
    if None in p and force:
        p = extract_answer_force(text)
    return None if None in p else (p[0], p[1])


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def _decimal_expansion(fr: Fraction, max_frac=1200):
    """Exact decimal string of a dyadic rational."""
    ip = fr.numerator // fr.denominator
    rem = fr - ip
    digs = []
    while rem and len(digs) < max_frac:
        rem *= 10
        d = rem.numerator // rem.denominator
        digs.append(str(d))
        rem -= d
    assert rem == 0
    return str(ip) + ("." + "".join(digs) if digs else "")


HAND = [
    "<think>frames 3-7 show it</think>\n<answer>12.54 to 17.83</answer>",
    "<think>x</think><answer>12.54 and 17.83</answer>",
    "<think>x</think> <answer>3 TO 4</answer>", "<answer>3 And 4</answer>", "<answer>3  to 4</answer>",
    "<answer>1 to 2</answer><answer>5 to 6</answer>", "<answer>1 to 2</answer><answer>nothing</answer>",
    "<answer>1 to 2 to 3</answer>", "<answer>1 to 22 to 3 and 4</answer>", "<answer>1.2.3 to 4</answer>",
    "<answer>1. to 3.</answer>", "<answer>1.to 3</answer>", "<answer>.5 to .7</answer>", "<answer>5 to .7</answer>",
    "<answer>1 to 2</answer", "<answer><answer>1 to 2</answer></answer>", "<answer>\n1 to\n2 10 to 11\n</answer>",
    "<answer>1 to 2</answer> trailing <answer>", "</answer><answer>7 to 8</answer>", "<answer>1 to2</answer>",
    "<ANSWER>1 to 2</ANSWER>", "", " ", "<answer></answer>", "<answer>12 to 13</answer><answer>",
    "<answer>from 12.5 to 17 seconds, or 20 and 30.</answer>", "<answer>1 tO 2</answer>", "<answer>1 tö 2</answer>",
    "<answer>٣ to ٤.٥</answer>", "<answer>1 to 2</answer>", "<answer>12 to 3٠</answer>",
    "  <think>a</think> \n<answer>b</answer>　 ", "<think></think><answer></answer>",
    "<think></think><answer></answer>x", "x<think></think><answer></answer>", "<think></think>x<answer></answer>",
    "<think></think><answer>", "<think></think><answer></answer", "<think></think></think> <answer></answer></answer>",
    "<think><answer></answer></think><answer>1 to 2</answer>", "<think></think><answer></answer>\x1f\x85",
    "<think></think><answer></answer>\xc2", "\x1c<think>\n\n</think>\t\x0b<answer>\n</answer>\r\n",
    "<think>a</think><answer>b</answer><think>c</think><answer>d</answer>", "<think></think><answe></answer>",
    "The event happens 3 to 4 and then 10.5 and 11", "no numbers", "12", "12 13", "first 1.5, then 2.", "a 1.b 2.5.6",
    "<answer>3 to 4</answer>\n<answer>x</answer> 7 and 9", "1" * 50 + " to " + "2" * 50, "١٢ to ٣٤ then 5 and 6.5",
]



_PIECES = ["<answer>", "</answer>", "<think>", "</think>", " to ", " and ", " TO ", " And ", " ", "  ", "\n", "\t", ".", "..",
           "1", "2", "30", "4.5", "007", "12.54", "17.83", "x", "to", "and", "<", ">", "</", "answer", "٣", " ",
           "　", " ", "é", "\U0001F600", "\U0001D7D8", "�", "9" * 20, "0.000", "think"]



def float_cases(seed=5, n_random=3000, n_half=300):
    rng = random.Random(seed)
    cases = ["0", "0.0", "000", "1", "1.", "007.50", "12.54", "17.83", "0.1", "0.30000000000000004",
             "9007199254740993", "9007199254740992", "9007199254740991", "18446744073709551615", "18446744073709551616",
             "123456789012345678901234567890", "0.000000000000000000000000000001", "1" + "0" * 308, "1" + "0" * 309,
             "17976931348623157" + "0" * 292, "17976931348623158" + "0" * 292, "17976931348623159" + "0" * 292,
             "0." + "0" * 323 + "4", "0." + "0" * 323 + "2", "0." + "0" * 323 + "3", "0." + "0" * 400 + "1",
             "0." + "0" * 307 + "22250738585072011", "0." + "0" * 307 + "22250738585072014",
             "1" * 400, "9" * 800 + "." + "9" * 800, "0." + "0" * 1100 + "7", "5" * 1000]
    for _ in range(n_random):
        n_int, n_frac = rng.choice([(1, 0), (2, 2), (3, 1), (5, 5), (17, 0), (1, 25), (20, 20), (40, 3), (300, 10), (1, 330)])
        s = "".join(rng.choice("0123456789") for _ in range(rng.randint(1, n_int)))
        if n_frac and rng.random() < 0.8:
            s += "." + "".join(rng.choice("0123456789") for _ in range(rng.randint(0, n_frac)))
        cases.append(s)
    # exact halfway points between adjacent doubles (and one digit either side): the hard cases of decimal->binary
    for _ in range(n_half):
        m = rng.getrandbits(52) | (1 << 52)
        e = rng.choice([-1074, -1073, -1060, -1022, -600, -80, -53, -10, 0, 5, 60, 300, 900]) if rng.random() < 0.7 else rng.randint(-1074, 960)
        if e == -1074:
            m = rng.getrandbits(rng.randint(1, 52)) | 1
        half = Fraction(2 * m + 1, 2) * Fraction(2) ** e
        if half >= Fraction(10) ** 309:
            continue
        d = _decimal_expansion(half)
        cases += [d, d + "0000", d + "0" * 30 + "1", d[:-1] + str(int(d[-1]) - 1) + "9" * 25] if d[-1] != "0" else [d]
    return cases


def fuzz_texts(seed=11, n=6000):
    rng = random.Random(seed)
    texts = []
    for _ in range(n):
        k = rng.choice([1, 2, 3, 5, 8, 13, 30])
        t = "".join(rng.choice(_PIECES) for _ in range(rng.randint(0, k)))
        if rng.random() < 0.4:
            t = "<think>" + t + rng.choice(["</think>", "</think>\n", "</think>  "]) + "<answer>" + \
                "".join(rng.choice(_PIECES) for _ in range(rng.randint(0, 6))) + rng.choice(["</answer>", "</answer>\n ", "</answer>."])
        if rng.random() < 0.3:
            t = rng.choice([" ", "\n", "\u2003", "\x1d", ""]) + t + rng.choice([" ", "\n\n", "\u3000", "\x85", ""])
        texts.append(t)
    return texts


def fuzz_byte_rows(seed=13, n=5000):
    """Rows that are not valid UTF-8 (partial multi-byte tokens of a byte-level BPE): the reference sees
    bytes.decode('utf-8', errors='replace'); the scanner sees the raw bytes."""
    rng = random.Random(seed)
    frag = [p.encode("utf-8") for p in _PIECES] + [b"\xe2", b"\xe2\x80", b"\x80", b"\xbf", b"\xc2", b"\xf0\x9f", b"\xf0\x9f\x98",
                                                    b"\xed\xa0\x80", b"\xc0\xaf", b"\xe0\x80\xaf", b"\xf4\x90\x80\x80", b"\xff", b"\xd9", b"\xa3",
                                                    b"\xe3\x80", b"\xc2\x85", b"\xe1\x9a\x80", b"\xe1\x9a"]
    rows = []
    for _ in range(n):
        r = b"".join(rng.choice(frag) for _ in range(rng.randint(0, rng.choice([2, 6, 14]))))
        if rng.random() < 0.4:
            r = b"<think>" + r + b"</think>" + rng.choice([b"", b"\n", b"\xc2\xa0", b"\xc2"]) + b"<answer>" + \
                b"".join(rng.choice(frag) for _ in range(rng.randint(0, 5))) + b"</answer>" + rng.choice([b"", b" ", b"\xe2\x80\x83", b"\x83", b"\xe2\x80"])
        rows.append(r)
    return rows


def expected(row: bytes, mode: int):
    """(parsed | None, format flag) the reference computes for these bytes."""
    t = row.decode("utf-8", errors="replace")
    if mode == 0:
        return ref_train(t)
    return ref_eval(t, mode == 2), False


def assert_rows_match(rows, mode, s, e, v, f):
    for i, r in enumerate(rows):
        want, fmt = expected(r, mode)
        if mode == 0:
            assert bool(f[i]) == fmt, ("format", r[:200])
        assert bool(v[i]) == (want is not None), ("valid", mode, r[:200], want)
        if want is not None:
            assert bits(s[i]) == bits(want[0]) and bits(e[i]) == bits(want[1]), (r[:200], want, s[i], e[i])
