"""Deterministic synthetic corpora, numpy-vectorised.

All generators are pure integer functions of a seed built on the 64-bit LCG
    x <- x * 6364136223846793005 + 1442695040888963407 (mod 2**64),   draw = x >> 33
so that inputs are reproducible without Python's `random`. The byte-for-byte definitions are the
pure-Python ones in SURVEY.md Appendix E; tests/test_corpus.py pins the vectorised versions against the
sha256 values recorded there.
"""
from __future__ import annotations

import numpy as np

A = np.uint64(6364136223846793005)
C = np.uint64(1442695040888963407)
M64 = (1 << 64) - 1

_jump_cache: dict[int, tuple[np.ndarray, np.ndarray]] = {}


def _jump_tables(m: int):
    """A_k = a^k, C_k = c*(a^(k-1)+...+1) for k = 1..m (uint64, wrapping)."""
    if m in _jump_cache:
        return _jump_cache[m]
    with np.errstate(over="ignore"):
        Ak = np.empty(m, dtype=np.uint64)
        Ck = np.empty(m, dtype=np.uint64)
        Ak[0] = A
        Ck[0] = C
        have = 1
        while have < m:
            take = min(have, m - have)
            Am, Cm = Ak[have - 1], Ck[have - 1]
            Ak[have:have + take] = Ak[:take] * Am
            Ck[have:have + take] = Ak[:take] * Cm + Ck[:take]
            have += take
    if len(_jump_cache) > 4:
        _jump_cache.clear()
    _jump_cache[m] = (Ak, Ck)
    return Ak, Ck


class LCG:
    """Stateful vectorised LCG: draws(n) returns the next n values of `x >> 33` as uint64."""

    CHUNK = 1 << 22

    def __init__(self, seed: int):
        self.x = np.uint64(seed & M64)

    def states(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint64)
        done = 0
        with np.errstate(over="ignore"):
            while done < n:
                m = min(self.CHUNK, n - done)
                Ak, Ck = _jump_tables(self.CHUNK if n - done >= self.CHUNK else m)
                out[done:done + m] = Ak[:m] * self.x + Ck[:m]
                self.x = out[done + m - 1]
                done += m
        return out

    def draws(self, n: int) -> np.ndarray:
        return self.states(n) >> np.uint64(33)

    def next(self) -> int:
        return int(self.draws(1)[0])


def rand_bytes(n: int, seed: int) -> bytes:
    """bytes((draw >> 23) & 0xff) == top byte of the LCG state (Appendix E `rand_bytes`, config C4)."""
    if n == 0:
        return b""
    g = LCG(seed)
    return (g.states(n) >> np.uint64(56)).astype(np.uint8).tobytes()


def _vocab(g: LCG, words: int = 4096):
    vocab = []
    for _ in range(words):
        L = 2 + g.next() % 8
        vocab.append(bytes(97 + int(v) % 26 for v in g.draws(L)))
    return vocab


def _emit_words(vocab, idx: np.ndarray, k0: int, n: int) -> np.ndarray:
    """Concatenate vocab[idx[i]] + separator (newline after every 16th word counted from k0+1, else space)."""
    W = max(len(w) for w in vocab) + 1
    table = np.zeros((len(vocab), W), dtype=np.uint8)
    lens = np.zeros(len(vocab), dtype=np.int64)
    for i, w in enumerate(vocab):
        table[i, :len(w)] = np.frombuffer(w, dtype=np.uint8)
        lens[i] = len(w)
    wl = lens[idx] + 1
    ends = np.cumsum(wl)
    total = int(ends[-1])
    starts = ends - wl
    widx = np.repeat(np.arange(len(idx)), wl)
    off = np.arange(total) - starts[widx]
    out = table[idx[widx], off]
    is_sep = off == lens[idx[widx]]
    k = (k0 + 1 + widx)  # word counter after increment
    sep = np.where(k % 16 == 0, 10, 32).astype(np.uint8)
    out = np.where(is_sep, sep, out)
    return out[:n] if total >= n else out


def word_text(n: int, seed: int, markov: bool = False) -> bytes:
    """Appendix E `word_text`; markov=True adds the order-1 twist of config C1:
    one extra LCG draw per word; if its low bit is set the index becomes (13*prev + pick) & 4095."""
    g = LCG(seed)
    vocab = _vocab(g)
    if n == 0:
        return b""
    parts = []
    have = 0
    k = 0
    prev = 0
    while have < n:
        nwords = max(1024, (n - have) // 4 + 16)
        per = 3 if markov else 2
        d = g.draws(nwords * per).reshape(nwords, per)
        a = d[:, 0] % np.uint64(4096)
        b = d[:, 1] % np.uint64(4096)
        pick = ((a * b) >> np.uint64(12)).astype(np.int64)
        if markov:
            twist = (d[:, 2] & np.uint64(1)).astype(bool)
            idx = np.empty(nwords, dtype=np.int64)
            p = prev
            pk = pick.tolist()
            tw = twist.tolist()
            for i in range(nwords):  # serial dependency on prev; cheap (one int op per word)
                p = ((13 * p + pk[i]) & 4095) if tw[i] else pk[i]
                idx[i] = p
            prev = p
        else:
            idx = pick
        chunk = _emit_words(vocab, idx, k, n - have)
        k += nwords
        parts.append(chunk)
        have += len(chunk)
    return np.concatenate(parts)[:n].tobytes()


def g_runs() -> bytes:
    return b"".join(bytes([i % 251]) * (1 + (i * 7) % 300) for i in range(4000))


_KEYWORDS = [b"if", b"else", b"for", b"while", b"return", b"int", b"char", b"void", b"static", b"const", b"struct",
             b"switch", b"case", b"break", b"continue", b"sizeof", b"unsigned", b"long", b"float", b"double",
             b"typedef", b"enum", b"union", b"goto", b"do", b"default", b"extern", b"register", b"volatile",
             b"inline", b"bool", b"auto", b"class", b"public", b"private", b"template", b"typename", b"namespace",
             b"using", b"new", b"delete", b"this", b"virtual", b"override", b"nullptr", b"true", b"false", b"try",
             b"catch", b"throw", b"operator", b"friend", b"explicit", b"mutable", b"constexpr", b"noexcept",
             b"static_cast", b"uint8_t", b"uint32_t", b"size_t", b"std", b"vector", b"string", b"include"]


def code_text(n: int, seed: int) -> bytes:
    """Source-code-like lines (config C2): indent*4 spaces + keyword(64) + ' ' + identifier(512) + punctuation + newline."""
    g = LCG(seed)
    idents = []
    for _ in range(512):
        L = 3 + g.next() % 10
        idents.append(bytes(97 + int(v) % 26 for v in g.draws(L)))
    punct = b"(){};=,"
    out = bytearray()
    nl = max(16, n // 12 + 16)
    d = g.draws(nl * 4).reshape(nl, 4).astype(np.int64)
    for i in range(nl):
        if len(out) >= n:
            break
        out += b" " * (4 * int(d[i, 0] % 6)) + _KEYWORDS[int(d[i, 1] % 64)] + b" " + idents[int(d[i, 2] % 512)]
        out += bytes([punct[int(d[i, 3] % 7)]]) + b"\n"
    while len(out) < n:  # pathological tiny-line case
        out += b"\n"
    return bytes(out[:n])


def binary_records(n: int, seed: int) -> bytes:
    """16-byte records (config C2): u32 LE counter, u16 from 8 values, u16 zero, 4 LCG bytes, 4 bytes of the previous record."""
    g = LCG(seed)
    nrec = (n + 15) // 16
    if nrec == 0:
        return b""
    st = g.states(nrec)
    rec = np.zeros((nrec, 16), dtype=np.uint8)
    cnt = np.arange(nrec, dtype=np.uint32)
    rec[:, 0:4] = cnt.view(np.uint8).reshape(nrec, 4)
    vals = np.array([0, 1, 2, 3, 256, 1000, 4096, 65535], dtype=np.uint16)
    rec[:, 4:6] = vals[((st >> np.uint64(40)) & np.uint64(7)).astype(np.int64)].view(np.uint8).reshape(nrec, 2)
    rec[:, 8:12] = (st >> np.uint64(32)).astype(np.uint32).view(np.uint8).reshape(nrec, 4)
    rec[1:, 12:16] = rec[:-1, 8:12]
    return rec.reshape(-1)[:n].tobytes()


def _mixed_segment(args):
    k, m = args
    kind = k % 3
    if kind == 0:
        return word_text(m, 200 + k)
    if kind == 1:
        return code_text(m, 300 + k)
    return binary_records(m, 400 + k)


def mixed_corpus(n: int, seg: int = 16 << 20, workers: int = 1) -> bytes:
    """Config C2: segments of `seg` bytes cycling text (seed 200+k) / code (300+k) / binary (400+k).
    Segments are independent, so `workers` > 1 generates them in a process pool (same bytes)."""
    jobs = []
    have = 0
    k = 0
    while have < n:
        m = min(seg, n - have)
        jobs.append((k, m))
        have += m
        k += 1
    if workers > 1 and len(jobs) > 1:
        from concurrent.futures import ProcessPoolExecutor
        with ProcessPoolExecutor(max_workers=min(workers, len(jobs))) as ex:
            parts = list(ex.map(_mixed_segment, jobs))
    else:
        parts = [_mixed_segment(j) for j in jobs]
    return b"".join(parts)


def skewed_bytes(n: int, seed: int, alpha: int = 40, num: int = 5, den: int = 3) -> bytes:
    """i.i.d. bytes 0..alpha-1 with P(i) proportional to (den/num)^i, thresholds in exact integer arithmetic.  With the default ratio 5/3
    the symbol counts of a 900k block fall off like a Fibonacci sequence: an unrestricted Huffman code would be 21-22 bits deep, so the
    length-limited allocator (HUFFMAN_ENCODE_MAXIMUM_CODE_LENGTH = 20, kernel.cpp:2748-2805) has to relocate nodes."""
    if n == 0:
        return b""
    w = [den ** i * num ** (alpha - 1 - i) for i in range(alpha)]
    tot, acc, thr = sum(w), 0, []
    for x in w:
        acc += x
        thr.append((acc << 32) // tot)
    u = LCG(seed).states(n) >> np.uint64(32)
    return np.minimum(np.searchsorted(np.array(thr, dtype=np.uint64), u, side="right"), alpha - 1).astype(np.uint8).tobytes()


def run_fuzz(n: int, seed: int) -> bytes:
    """About n bytes of run-structured noise for the parity fuzz tests: runs whose lengths sit on the RLE1 edges (3/4/5, 254..261, 510..515:
    BlockCompressor.hpp:134-154 writes a run as 4 bytes + count, at most 255 + 4 per chunk), literal stretches over small and full alphabets,
    short periods, and copies of earlier stretches -- so that block ends fall into every phase of a run."""
    g = LCG(seed)
    edges = (1, 2, 3, 4, 5, 6, 7, 254, 255, 256, 257, 258, 259, 260, 261, 262, 300, 510, 513, 514, 515, 516, 1000, 2600)
    out = bytearray()
    while len(out) < n:
        kind, a, b, c = (int(v) for v in g.draws(4))
        if kind % 5 == 0:
            out += bytes([a & 0xff]) * edges[b % len(edges)]
        elif kind % 5 == 1:
            k = 1 + c % 400
            alpha = (2, 3, 4, 16, 256)[b % 5]
            out += bytes((int(v) % alpha + (a & 0x3f)) & 0xff for v in g.draws(k))
        elif kind % 5 == 2:
            p = 1 + b % 7
            pat = bytes((a + 31 * i) & 0xff for i in range(p))
            out += (pat * (c % 300 // p + 2))[:c % 300 + 1]
        elif kind % 5 == 3 and len(out) > 16:
            lo = a % (len(out) - 8)
            out += out[lo:lo + 1 + b % 700]
        else:
            out += bytes([a & 0xff]) * (1 + b % 9) + bytes([c & 0xff])
    return bytes(out[:n])


def repetitive_corpus(n: int) -> bytes:
    """Config C3 scaled to n bytes (n/4 each): zeros; short periods + 'a' run; long-LCP text; near-periodic 'ab'."""
    q = n // 4
    parts = [bytes(q)]
    e = q // 8
    seg = []
    for p in (2, 3, 4, 5, 7, 71, 255):
        pat = bytes((i * 37 + 11) % 251 + 1 for i in range(p))
        seg.append((pat * (e // p + 1))[:e])
    seg.append(b"a" * (q - 7 * e))
    parts.append(b"".join(seg))
    base = word_text(100_000, 500)
    parts.append((base * (q // len(base) + 1))[:q])
    ab = np.frombuffer((b"ab" * (q // 2 + 1))[:n - 3 * q], dtype=np.uint8).copy()
    g = LCG(7)
    pos = np.arange(65535, len(ab), 65536)
    if len(pos):
        ab[pos] = (g.states(len(pos)) >> np.uint64(56)).astype(np.uint8)
    parts.append(ab.tobytes())
    return b"".join(parts)


GOLDEN_INPUTS = {
    "g_empty": lambda: b"",
    "g_one": lambda: b"x",
    "g_zero4": lambda: bytes(4),
    "g_aaa": lambda: b"aaa",
    "g_zeros300k": lambda: bytes(300000),
    "g_runs": g_runs,
    "g_rand1m": lambda: rand_bytes(1000003, 1),
    "g_text3m": lambda: word_text(3000000, 2),
    "g_ababx": lambda: b"ab" * 400000 + b"x",
    "g_abab": lambda: b"ab" * 5000,
    "g_abc": lambda: b"abc" * 400000,
    "g_zeros50m": lambda: bytes(50_000_000),
}
