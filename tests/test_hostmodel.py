"""CPU suite, part 2: the host compilation of the GPU kernels' integer building blocks
(moonbit_flate_b200/csrc/huff_build.cuh, common.cuh -- the very code k_build_codes runs)
and the lane-by-lane emulation of K1's 32-wide batched probe (tests/hostmodel/hostmodel.cu)
against the oracle.  No GPU needed; this is what makes the CUDA path's logic testable here."""
import os
import subprocess

import numpy as np
import pytest

from helpers import BLK_DYNAMIC, BLK_HUFF, ROOT, Corpus, HostModel, Oracle, fuzz_streams


@pytest.fixture(scope="module")
def hm():
    so = os.path.join(ROOT, "tests", "hostmodel", "libfb_hostmodel.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "hostmodel")])
    return HostModel()


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.fixture(scope="module")
def corpus():
    return Corpus()


def test_length_and_offset_codes_all_values(hm, oracle):
    """length_code_of / offset_code_of (computed) == token.mbt:30-61,:107-123 LUTs + hbw:49-78 extra-bit tables."""
    lbase = [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192,
             224, 255]
    for xlen in range(256):
        c = hm.codes(xlen, 0)
        assert c[0] == oracle.L.orc_length_code(xlen), xlen
        assert xlen - lbase[c[0]] == c[2] and c[2] < (1 << c[1]) or (c[1] == 0 and c[2] == 0), xlen
    for xoff in range(32768):
        c = hm.codes(0, xoff)
        assert c[3] == oracle.L.orc_offset_code(xoff), xoff
        base = xoff if c[3] < 4 else ((2 + (c[3] & 1)) << c[4])
        assert base + c[5] == xoff, xoff


def _freq_cases(rng):
    yield np.array([5, 0, 0, 1] + [0] * 26)                                 # two symbols: codes 0/1 (hc:326-336)
    yield np.array([0] * 29 + [9])                                          # one symbol
    yield np.ones(286, np.int64)                                            # flat
    fib = [1, 1]
    while len(fib) < 40:
        fib.append(fib[-1] + fib[-2])
    yield np.array(fib[:32] + [0] * 254)                                    # forces the 15-bit limit
    yield np.array(fib[:24][::-1] + [3] * 262)
    for _ in range(200):
        n = int(rng.choice([19, 30, 286]))
        f = (rng.pareto(0.7, n) * rng.integers(1, 50)).astype(np.int64)
        f[rng.random(n) < rng.random() * 0.8] = 0
        yield np.minimum(f, 60000)


def test_generate_matches_oracle(hm, oracle):
    """generate_dev == HuffmanEncoder::generate (huffman-code.mbt:295-343) incl. bit_counts tie-breaking."""
    rng = np.random.default_rng(1)
    for f in _freq_cases(rng):
        for mb in (15, 7):
            if mb == 7 and f.size != 19:
                continue
            gl, gc = hm.generate(f, mb)
            ol, oc = oracle.huff_generate(f, mb)
            assert np.array_equal(gl, ol) and np.array_equal(gc, oc), (f.tolist(), mb)
            assert gl.max(initial=0) <= mb


def test_generate_tie_heavy_randomized(hm, oracle):
    """The eager, level-by-level package-merge of huff_build.cuh against the reference's lazy boundary
    package-merge on tie-heavy frequency vectors (tie-breaking decides the bit counts)."""
    rng = np.random.default_rng(12345)
    for it in range(3000):
        n = int(rng.choice([19, 30, 286]))
        mode = it % 6
        if mode == 0:
            f = rng.integers(0, 3, n)
        elif mode == 1:
            f = rng.integers(0, 8, n)
        elif mode == 2:
            f = (rng.geometric(0.3, n) - 1) * rng.integers(1, 4)
        elif mode == 3:
            f = (2 ** rng.integers(0, 16, n)) * (rng.random(n) < 0.6)
        elif mode == 4:
            f = np.minimum((rng.pareto(0.5, n) * 3).astype(np.int64), 65535)
        else:
            k = int(rng.integers(3, n + 1))
            f = np.zeros(n, np.int64)
            f[rng.choice(n, k, replace=False)] = rng.integers(1, 5, k)
        f = np.asarray(f, np.int64)
        mb = 7 if n == 19 and it % 2 else 15
        gl, gc = hm.generate(f, mb)
        ol, oc = oracle.huff_generate(f, mb)
        assert np.array_equal(gl, ol) and np.array_equal(gc, oc), (f.tolist(), mb)


SIZES = [128, 129, 300, 4096, 30000, 65534, 65535, 65536, 65662, 65663, 70000, 131070, 200000]


@pytest.mark.parametrize("klass", range(6))
def test_batched_parse_model_matches_oracle(hm, oracle, corpus, klass):
    """The 32-wide batch schedule / intra-batch bucket resolution / multi-match walk of parse.cu, emulated lane by
    lane on the CPU, yields the oracle's (sequential) token arrays -- single- and multi-block streams."""
    for i, n in enumerate(SIZES):
        d = corpus.unit(n, seed=17, index=i, klass=klass)
        _, wtok, wntok, _, _ = oracle.deflate_ex(d)
        for kw in ({}, {"v2": True}, {"v3": True}, {"v3": True, "lowest_wins": False}, {"v4": True},
                   {"v4": True, "lowest_wins": False}):
            toks, ntok = hm.parse_stream(d, **kw)
            assert list(ntok) == list(wntok), (klass, n, kw)
            assert np.array_equal(toks, wtok), (klass, n, kw)


def _histogram(tokens):
    f = np.zeros(320, np.uint32)
    lit = tokens[tokens < (1 << 30)]
    np.add.at(f, lit, 1)
    m = tokens[tokens >= (1 << 30)]
    f[256] = 1
    return f, m


def test_build_block_matches_oracle(hm, oracle, corpus):
    """build_block_dev (codes, codegen, header bit string, exact block size) against the oracle's stream:
    the header words are the first bits of the stream and header + payload bits equal the oracle's block size."""
    for klass in range(6):
        for i, n in enumerate((128, 1000, 65535)):
            d = corpus.unit(n, seed=23, index=i, klass=klass)
            comp, toks, ntok, kind, bits = oracle.deflate_ex(d)
            f = np.zeros(320, np.uint32)
            if kind[0] == BLK_DYNAMIC:
                for t in toks:
                    t = int(t)
                    if t < (1 << 30):
                        f[t] += 1
                    else:
                        c = hm.codes((t - (1 << 30)) >> 22, t & ((1 << 22) - 1))
                        f[257 + c[0]] += 1
                        f[286 + c[3]] += 1
            else:
                assert kind[0] == BLK_HUFF
                np.add.at(f, np.frombuffer(d, np.uint8), 1)
            f[256] = 1
            k, codes, hdr, hdr_nbits, blk_bits = hm.build_block(f, int(kind[0]), n)
            assert k == kind[0]
            assert blk_bits == int(bits[0]), (klass, n)
            got = np.unpackbits(hdr.view(np.uint8), bitorder="little")[:hdr_nbits]
            want = np.unpackbits(np.frombuffer(comp, np.uint8), bitorder="little")[:hdr_nbits]
            assert np.array_equal(got, want), (klass, n)


def test_batched_parse_model_fuzz(hm, oracle):
    """Structured random inputs (tests/helpers.py: fuzz_streams) through the lane-by-lane emulation of parse.cu."""
    rng = np.random.default_rng(7)
    for i, d in enumerate(fuzz_streams(rng, 140)):
        _, wtok, wntok, _, _ = oracle.deflate_ex(d)
        for kw in ({"v2": True}, {"v3": True}, {"v3": True, "lowest_wins": False}, {"v4": True},
                   {"v4": True, "lowest_wins": False}):
            toks, ntok = hm.parse_stream(d, **kw)
            assert list(ntok) == list(wntok), (i, len(d), kw)
            assert np.array_equal(toks, wtok), (i, len(d), kw)


@pytest.mark.parametrize("klass,nblk", [(0, 40), (1, 40), (2, 12), (3, 40), (-1, 60)])
def test_block_parallel_parse_fixpoint(klass, nblk):
    """CPU model of the block-parallel parse of multi-block streams (tests/hostmodel/blockpar_model.c, on the
    oracle's own encode): parse every block from an empty table, then re-parse the blocks whose predecessor's
    normalised end table changed, until nothing changes; the tokens of every block then equal the sequential
    parse's, and the fixpoint is reached in a handful of rounds whatever the number of blocks."""
    import ctypes as C
    so = os.path.join(ROOT, "tests", "hostmodel", "libfb_blockpar.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "hostmodel"), "libfb_blockpar.so"])
    L = C.CDLL(so)
    L.fbm_blockpar_check.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_long)]
    n = nblk * 65535
    src = Corpus().fill((n + 65535) // 65536, 65536, seed=5, klass=klass)[:n]
    rounds, parses = C.c_int(), C.c_long()
    assert L.fbm_blockpar_check(src.ctypes.data, nblk, C.byref(rounds), C.byref(parses)) == 1
    assert rounds.value <= 10, rounds.value
    assert parses.value <= 6 * nblk, parses.value
