"""The packed-sequence predicates of csrc/wd_seq.cuh (Hamming, shifted-Hamming
filter, multi-word Myers) built for the host and checked against the oracle's
textbook dynamic programme.  Host build is test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import ref_port as R

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("seq") / "seq_harness.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(HERE, "cpu_seq_harness.cpp")])
    lib = ctypes.CDLL(so)
    lib.seq_check.argtypes = [ctypes.c_char_p, ctypes.c_char_p] + [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_int)] * 3
    lib.seq_first_reject.argtypes = [ctypes.c_char_p, ctypes.c_char_p] + [ctypes.c_int] * 4
    lib.seq_head32_rejects.argtypes = [ctypes.c_char_p, ctypes.c_char_p] + [ctypes.c_int] * 4
    lib.seq_prefix_dp.argtypes = [ctypes.c_char_p, ctypes.c_char_p] + [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_int)]
    return lib


def check(lib, a, b, e, ham, words):
    d, x, s = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.seq_check(bytes(a), bytes(b), len(a), words, e, ham, d, x, s) == 0
    return d.value, x.value, s.value


def words_for(n):
    for w in (1, 2, 4, 8, 16):
        if n <= 64 * w:
            return w
    raise ValueError(n)


def mutate(rng, a):
    b = list(a)
    for _ in range(int(rng.integers(0, 5))):
        op = int(rng.integers(0, 3))
        if op == 0 and b:
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 5))
        elif op == 1 and b:
            del b[int(rng.integers(0, len(b)))]
            b.append(int(rng.integers(0, 5)))
        else:
            b.insert(int(rng.integers(0, len(b) + 1)), int(rng.integers(0, 5)))
            b.pop()
    return b


@pytest.mark.parametrize("lengths", [(1, 20), (40, 64), (65, 130), (190, 260), (500, 520)])
def test_predicates_match_oracle(harness, lengths):
    rng = np.random.default_rng(lengths[0])
    n_cases = 400 if lengths[1] < 200 else 60
    for _ in range(n_cases):
        n = int(rng.integers(lengths[0], lengths[1] + 1))
        a = [int(v) for v in rng.integers(0, 5, n)]
        b = mutate(rng, a) if rng.random() < 0.7 else [int(v) for v in rng.integers(0, 5, n)]
        sa = "".join("ACGTN"[v] for v in a)
        sb = "".join("ACGTN"[v] for v in b)
        lev, hd = R.levenshtein(sa, sb), R.hamming(sa, sb)
        w = words_for(n)
        for e in (0, 1, 2, 3, 4, 7, n, n + 3):
            d, x, s = check(harness, a, b, e, 0, w)
            assert x == lev, (sa, sb)
            assert d == int(lev <= e), (sa, sb, e)
            assert not (s and lev <= e), "shifted-Hamming filter rejected a true duplicate"
            d, x, _ = check(harness, a, b, e, 1, w)
            assert x == hd and d == int(hd <= e)


def test_wider_word_count_gives_same_answer(harness):
    rng = np.random.default_rng(3)
    for _ in range(100):
        n = int(rng.integers(1, 64))
        a = [int(v) for v in rng.integers(0, 5, n)]
        b = mutate(rng, a)
        ref = check(harness, a, b, 2, 0, 1)
        for w in (2, 4, 16):
            assert check(harness, a, b, 2, 0, w) == ref


@pytest.mark.parametrize("lengths", [(1, 30), (40, 64), (65, 200)])
def test_prefix_rejection_is_safe_and_useful(harness, lengths):
    """The early exit of the gather never drops a true duplicate, and it does
    fire early on unrelated reads."""
    rng = np.random.default_rng(100 + lengths[0])
    fired, unrelated = 0, 0
    for _ in range(500):
        n = int(rng.integers(lengths[0], lengths[1] + 1))
        a = [int(v) for v in rng.integers(0, 5, n)]
        related = rng.random() < 0.5
        b = mutate(rng, a) if related else [int(v) for v in rng.integers(0, 4, n)]
        sa = "".join("ACGTN"[v] for v in a)
        sb = "".join("ACGTN"[v] for v in b)
        lev, hd = R.levenshtein(sa, sb), R.hamming(sa, sb)
        w = words_for(n)
        for e in (0, 1, 2, 3, 5, 8):
            for ham, dist in ((0, lev), (1, hd)):
                k = harness.seq_first_reject(bytes(a), bytes(b), n, w, e, ham)
                assert k >= 0
                if k:
                    assert dist > e, (sa, sb, e, ham, k)
        if not related and n >= 40:
            unrelated += 1
            fired += 0 < harness.seq_first_reject(bytes(a), bytes(b), n, w, 2, 0) <= 32
    if unrelated:
        assert fired > 0.9 * unrelated


def _band_min_bruteforce(a, b, k):
    """out[p] = min over |j-p| <= k of ed(b[:p], a[:j]) + |j-p|, textbook DP."""
    n = len(a)
    col = list(range(n + 1))                      # D[j][0] = j
    out = [0] * (n + 1)
    for p in range(1, n + 1):
        new = [p] + [0] * n
        for j in range(1, n + 1):
            new[j] = min(col[j] + 1, new[j - 1] + 1, col[j - 1] + (a[j - 1] != b[p - 1]))
        col = new
        out[p] = min(col[j] + abs(j - p) for j in range(max(0, p - k), min(n, p + k) + 1))
    return out


@pytest.mark.parametrize("lengths", [(1, 20), (40, 64), (65, 130), (190, 260)])
def test_incremental_prefix_dp(harness, lengths):
    """PrefixDP of the fused kernel: exact band minimum after every symbol with the
    centre known only k symbols ahead; never rejects a true duplicate; ends on
    the edit distance test itself."""
    rng = np.random.default_rng(7 + lengths[0])
    n_cases = 150 if lengths[1] < 200 else 25
    for _ in range(n_cases):
        n = int(rng.integers(lengths[0], lengths[1] + 1))
        a = [int(v) for v in rng.integers(0, 5, n)]
        b = mutate(rng, a) if rng.random() < 0.7 else [int(v) for v in rng.integers(0, 5, n)]
        sa = "".join("ACGTN"[v] for v in a)
        sb = "".join("ACGTN"[v] for v in b)
        lev = R.levenshtein(sa, sb)
        w = words_for(n)
        for e in (2, 3, 4, 7, 12):
            k = e // 2
            want = _band_min_bruteforce(a, b, k)
            out = (ctypes.c_int * (n + 1))()
            assert harness.seq_prefix_dp(bytes(a), bytes(b), n, w, k, int(rng.integers(1, 40)), out) == 0
            got = list(out)
            # values above e only have to stay above e (cells outside the band are over-estimated)
            for p in range(1, n + 1):
                assert (got[p] <= e) == (want[p] <= e), (sa, sb, e, p, got[p], want[p])
                if want[p] <= e:
                    assert got[p] == want[p]
                    assert lev >= want[p]
            assert (got[n] <= e) == (lev <= e)
            if lev <= e:
                assert all(got[p] <= e for p in range(1, n + 1)), "prefix test rejected a true duplicate"


@pytest.mark.parametrize("lengths", [(1, 31), (32, 64), (65, 200)])
def test_head32_prefilter_is_safe_and_useful(harness, lengths):
    """The 32-symbol pre-filter of exhaustive mode never rejects a pair within
    distance e, and it does reject nearly all unrelated pairs."""
    rng = np.random.default_rng(200 + lengths[0])
    fired, unrelated = 0, 0
    for _ in range(600):
        n = int(rng.integers(lengths[0], lengths[1] + 1))
        a = [int(v) for v in rng.integers(0, 5, n)]
        related = rng.random() < 0.6
        b = mutate(rng, a) if related else [int(v) for v in rng.integers(0, 4, n)]
        sa = "".join("ACGTN"[v] for v in a)
        sb = "".join("ACGTN"[v] for v in b)
        lev, hd = R.levenshtein(sa, sb), R.hamming(sa, sb)
        w = words_for(n)
        for e in (0, 1, 2, 3, 5, 8, 31):
            if e >= n:
                continue
            for ham, dist in ((0, lev), (1, hd)):
                if harness.seq_head32_rejects(bytes(a), bytes(b), n, w, e, ham):
                    assert dist > e, (sa, sb, e, ham)
        if not related and n >= 32:
            unrelated += 1
            fired += harness.seq_head32_rejects(bytes(a), bytes(b), n, w, 2, 0)
    if unrelated:
        assert fired > 0.97 * unrelated
