"""CPU checks of the cache-builder oracle (oracle/cache_oracle.py) and of the host half of the C ABI.

The reference's own builder (src/data/build_cache_v1.py) imports Polars, which is not installed and whose hash is unpinned
(SURVEY §8c), so the oracle is pinned by published known-answer vectors (FNV-1a 64, splitmix64) and by hand-worked cases of the
reference's parse / impute rules instead of by running the reference."""
import ctypes

import numpy as np
import pytest

from oracle import cache_oracle as co


def test_fnv1a64_published_vectors():
    assert co.fnv1a64(b"") == 0xCBF29CE484222325
    assert co.fnv1a64(b"a") == 0xAF63DC4C8601EC8C
    assert co.fnv1a64(b"foobar") == 0x85944171F73967E8


def test_splitmix64_first_output_of_seed_zero():
    # splitmix64: state += 0x9e3779b97f4a7c15, output = finalize(state); the first output for seed 0 is published
    assert co.splitmix64_finalize(0x9E3779B97F4A7C15) == 0xE220A8397B1DCDAF


def test_hash_value_forms():
    assert co.hash_value(None) == co.hash_bytes(b"NA")
    assert co.hash_value(-12) == co.hash_bytes(b"-12")
    assert co.hash_value(np.int32(7)) == co.hash_bytes(b"7")
    assert co.hash_value("café") == co.hash_bytes("café".encode("utf-8"))
    assert co.hash_bytes(b"x", seed=1) != co.hash_bytes(b"x", seed=2)


def test_host_hash_of_the_c_abi_matches_the_oracle():
    from b200ctr import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 7, 64, 300):
        data = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        buf = ctypes.create_string_buffer(data, max(n, 1))
        got = lib.b200ctr_cache_hash_bytes_host(ctypes.cast(buf, ctypes.c_void_p), n, 2025) & ((1 << 64) - 1)
        assert got == co.hash_bytes(data, 2025)


def test_parse_seq_rules():
    got = co.parse_seq(["1,2,3", "", None, "5,,6,", "1,2,3,4,5,6,7", "42"], 4, 0)
    want = np.array([[0, 1, 2, 3], [0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 5, 6], [4, 5, 6, 7], [0, 0, 0, 42]], dtype=np.int32)
    assert got.dtype == np.int32 and np.array_equal(got, want)
    assert np.array_equal(co.parse_seq(["9"], 3, 7), np.array([[7, 7, 9]], dtype=np.int32))
    with pytest.raises(ValueError):
        co.parse_seq(["1,x"], 4, 0)


def test_numeric_block_and_median():
    x = np.array([[1.0, np.nan, np.inf], [np.nan, np.nan, -np.inf], [3.0, np.nan, 2.0], [2.0, np.nan, np.nan]], dtype=np.float32)
    med = [co.median_ignore_missing(x[:, j]) for j in range(3)]
    assert med[0] == 2.0 and med[1] == 0.0            # all-missing column -> 0.0 (build_cache_v1.py:62-64)
    assert med[2] == 2.0                              # sorted non-missing: -inf, 2, inf
    X, M = co.numeric_block(x, med)
    assert np.array_equal(M, np.array([[0, 1, 0], [1, 1, 0], [0, 1, 0], [0, 1, 1]], dtype=np.uint8))
    assert np.array_equal(X, np.array([[1, 0, 1e6], [2, 0, -1e6], [3, 0, 2], [2, 0, 2]], dtype=np.float32))
    assert co.median_ignore_missing(np.array([4, 1, 3, 2], np.float32)) == 2.5


# ---- property tests (hypothesis): the parse rule and the two hash implementations on arbitrary inputs ------------------------
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=150, deadline=None)
@given(st.lists(st.one_of(st.none(), st.lists(st.integers(0, 2 ** 31 - 1), max_size=40)), max_size=6), st.integers(1, 12),
       st.integers(0, 5), st.integers(0, 3))
def test_parse_seq_keeps_the_last_tokens_right_aligned(rows, L, pad, extra_commas):
    strings = []
    for toks in rows:
        if toks is None:
            strings.append(None)
        else:
            s = ",".join(map(str, toks))
            strings.append("," * extra_commas + s.replace(",", "," * (1 + extra_commas % 2)) + "," * extra_commas)
    got = co.parse_seq(strings, L, pad)
    assert got.shape == (len(rows), L) and got.dtype == np.int32
    for i, toks in enumerate(rows):
        last = (toks or [])[-L:]
        assert list(got[i, L - len(last):]) == last and all(v == pad for v in got[i, :L - len(last)])


@settings(max_examples=100, deadline=None)
@given(st.binary(max_size=200), st.integers(0, 2 ** 32 - 1))
def test_host_hash_equals_oracle_on_arbitrary_bytes(data, seed):
    from b200ctr import _lib
    lib = _lib.load()
    buf = ctypes.create_string_buffer(data, max(len(data), 1))
    got = lib.b200ctr_cache_hash_bytes_host(ctypes.cast(buf, ctypes.c_void_p), len(data), seed) & ((1 << 64) - 1)
    assert got == co.hash_bytes(data, seed)
