"""CPU suite, part 1: pins the oracle (oracle/flate_oracle.c, the C restatement of the
reference) against every known answer the reference's own tests hold for this path
(tests/golden/reference_kats.json, lifted from /root/reference by make_golden.py),
against the data-independent stream shapes SURVEY.md 8c derives from the source, and
against stdlib zlib as an independent inflater / foreign-stream producer."""
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

from helpers import (BLK_DYNAMIC, BLK_HUFF, BLK_STORED, ORC_CORRUPT, ORC_EOF_AT_REFILL, ORC_OK, ORC_UNEXPECTED_EOF,
                     Corpus, Oracle)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KATS = json.load(open(os.path.join(GOLD, "reference_kats.json")))


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.fixture(scope="module")
def corpus():
    return Corpus()


# ---------------------------------------------------------------- reference KATs

def test_kat_writer_dict_38_bytes(oracle):
    """deflate_test.mbt:12-35 -- the only compressed-size known answer in the reference."""
    k = KATS["writer_dict"]
    d, t = k["dict"].encode(), k["text"].encode()
    want = oracle.writer_roundtrip([d, t])
    assert len(want) == k["compressed_len"] == 38
    assert oracle.writer_roundtrip([t], dict_=d) == want  # new_dict(dict)+write(text) == write(dict)+write(text)
    assert zlib.decompress(want, -15) == d + t
    # derived during the survey from an independent Python restatement
    assert want.hex() == "04c0c10900300843d155b29aa0d4c2a7012faedfd705d67ac82eb0e2c47d5a0ff9030000ffff"


def test_kat_unit_values(oracle):
    u = KATS["units"]
    assert oracle.L.orc_token_offset(u["token_offset"][0]) == u["token_offset"][1]      # token.mbt:95-99
    assert oracle.L.orc_reverse16(u["reverse16"][0]) == u["reverse16"][1]               # bits.mbt:24-27
    assert oracle.L.orc_reverse_bits(u["reverse_bits"][0], u["reverse_bits"][1]) == u["reverse_bits"][2]


def test_kat_best_speed_matrix(oracle):
    """deflate-fast_test.mbt:14-100: 16 write patterns x 6 first-write sizes, round trip through Writer + Reader."""
    b = KATS["best_speed"]
    abcabc = bytes(range(b["abc_len"])) * (b["total"] // b["abc_len"])
    n = 0
    for tc in b["test_cases"]:
        for first in b["first_n"]:
            sizes = [first] + tc[1:]
            writes = [abcabc[:s] for s in sizes]
            want = b"".join(writes)
            comp = oracle.writer_roundtrip(writes)
            assert comp == oracle.deflate(want)  # write chunking never changes the bytes (deflate.mbt:222-241)
            st, out, _, cons = oracle.inflate(comp, len(want) + 1)
            assert st == ORC_OK and out == want and cons == len(comp)
            assert zlib.decompress(comp, -15) == want
            n += 1
    assert n == 96


def test_kat_dict_decoder_scenario(oracle):
    """dict-decoder_wbtest.mbt:9-291 replayed against the oracle's DictDecoder (2 KiB window)."""
    import ctypes as C

    k = KATS["dict_decoder"]
    L = oracle.L
    size = k["window"]
    dd = C.c_void_p(L.orc_dict_new(size, None, 0))
    got = bytearray()
    flush_buf = np.zeros(size, np.uint8)

    def flush():
        n = L.orc_dict_read_flush(dd, flush_buf.ctypes.data)
        got.extend(flush_buf[:n].tobytes())

    def write_copy(dist, length):
        while length > 0:
            cnt = L.orc_dict_try_write_copy(dd, dist, length)
            if cnt == 0:
                cnt = L.orc_dict_write_copy(dd, dist, length)
            length -= cnt
            if L.orc_dict_avail_write(dd) == 0:
                flush()

    def write_string(s: bytes):
        while s:
            a = np.frombuffer(s, dtype=np.uint8)
            cnt = L.orc_dict_write(dd, a.ctypes.data, len(s))
            s = s[cnt:]
            if L.orc_dict_avail_write(dd) == 0:
                flush()

    poem, abc, fox = k["poem"].encode(), k["abc"].encode(), k["fox"].encode()
    want = bytearray()
    write_string(b".")
    want += b"."
    pos = 0
    for dist, length in k["poem_refs"]:
        if dist == 0:
            write_string(poem[pos: pos + length])
        else:
            write_copy(dist, length)
        pos += length
    want += poem
    write_copy(L.orc_dict_hist_size(dd), 33)
    want += want[:33]
    write_string(abc)
    write_copy(len(abc), 59 * len(abc))
    want += abc * 60
    write_string(fox)
    write_copy(len(fox), 9 * len(fox))
    want += fox * 10
    write_string(b".")
    write_copy(1, 9)
    want += b"." * 10
    write_string(poem.upper())
    write_copy(len(poem), 7 * len(poem))
    want += poem.upper() * 8
    hs = L.orc_dict_hist_size(dd)
    write_copy(hs, 10)
    drop = len(want) - hs
    want += want[drop: drop + 10]
    flush()
    L.orc_dict_free(dd)
    assert bytes(got) == bytes(want)


# ---------------------------------------------------------------- stream shapes derived from the source

def test_data_independent_sizes(oracle):
    assert oracle.deflate(b"") == bytes.fromhex("010000ffff")                       # close only (deflate.mbt:171)
    assert oracle.deflate(b"x") == bytes.fromhex("000100feff") + b"x" + bytes.fromhex("010000ffff")
    assert len(oracle.deflate(bytes(16))) == 26                                      # stored(16) + trailer
    assert len(oracle.deflate(bytes(65535))) == 82                                   # 3 literals + 254 x (258, dist 1) + 3-lit tail
    assert len(oracle.deflate(bytes(65536))) == 88                                   # + 1-byte stored tail


def test_block_policy(oracle, corpus):
    """Compressor::enc_speed (deflate.mbt:236-277): <=16 stored, 17..127 huff-only, else parsed; 15/16 rule."""
    for n, kinds in ((1, [BLK_STORED]), (16, [BLK_STORED]), (17, [BLK_HUFF]), (127, [BLK_HUFF]),
                     (128, [BLK_DYNAMIC]), (65535, [BLK_DYNAMIC]), (65536, [BLK_DYNAMIC, BLK_STORED]),
                     (65552, [BLK_DYNAMIC, BLK_HUFF]), (65663, [BLK_DYNAMIC, BLK_DYNAMIC])):
        d = corpus.unit(n, seed=1, index=0, klass=Corpus.CONST)  # one repeated byte: always worth a dynamic block
        _, _, ntok, kind, _ = oracle.deflate_ex(d)
        assert list(kind) == kinds, n
    # incompressible data: tokens > n - n/16 -> literal-only block, never stored (quirk D2)
    d = corpus.unit(65535, seed=1, index=0, klass=Corpus.RANDOM)
    c, _, ntok, kind, bits = oracle.deflate_ex(d)
    assert list(kind) == [BLK_HUFF] and len(c) > 65535 and len(c) < 65535 + 128


def test_cross_block_matches_are_four_bytes(oracle, corpus):
    """Quirk D1 (deflate-fast.mbt:152-159, :309-313): `prev` is never populated, so a match whose candidate lies
    in the previous block is exactly 4 bytes long (Go would extend it)."""
    a = corpus.unit(65535, seed=5, index=0, klass=Corpus.TEXT)
    d = a + a[40000:44000]                       # block 2 repeats block-1 bytes at distance 25535 <= 32768
    c, toks, ntok, kind, _ = oracle.deflate_ex(d)
    assert zlib.decompress(c, -15) == d and len(ntok) == 2
    pos, cross = 0, 0
    for t in toks[int(ntok[0]):]:
        t = int(t)
        if t < (1 << 30):
            pos += 1
            continue
        length, dist = ((t - (1 << 30)) >> 22) + 3, (t & ((1 << 22) - 1)) + 1
        if pos - dist < 0:
            assert length == 4, (pos, dist, length)
            cross += 1
        pos += length
    assert pos == 4000 and cross > 10


@pytest.mark.parametrize("klass", range(6))
def test_zlib_inflates_every_oracle_stream(oracle, corpus, klass):
    for i, n in enumerate([0, 1, 15, 16, 17, 127, 128, 129, 1000, 65534, 65535, 65536, 65537, 65551, 65552, 65662,
                           65663, 131072, 200000]):
        d = corpus.unit(n, seed=13, index=i, klass=klass)
        c = oracle.deflate(d)
        assert zlib.decompress(c, -15) == d, (klass, n)
        st, out, _, cons = oracle.inflate(c, n + 3)
        assert st == ORC_OK and out == d and cons == len(c), (klass, n)


def test_oracle_digests_frozen(oracle, corpus):
    g = json.load(open(os.path.join(GOLD, "oracle_digests.json")))
    for u in g["units"]:
        c = oracle.deflate(corpus.unit(u["n"], seed=u["seed"], index=u["index"], klass=u["klass"]))
        assert (len(c), hashlib.sha256(c).hexdigest()) == (u["clen"], u["sha256"]), u


# ---------------------------------------------------------------- decoder: foreign streams and the error model

def test_oracle_inflates_foreign_streams(oracle, corpus):
    for n in (0, 1, 300, 70000):
        for klass in (0, 2, 3):
            d = corpus.unit(n, seed=3, index=n, klass=klass)
            for lvl, strat in ((0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                               (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY)):
                co = zlib.compressobj(lvl, zlib.DEFLATED, -15, 9, strat)
                c = co.compress(d) + co.flush()
                st, out, _, cons = oracle.inflate(c, n + 1)
                assert st == ORC_OK and out == d and cons == len(c)


def test_oracle_error_model(oracle, corpus):
    """Error classes and offsets of inflate.mbt: corrupt (:38-40) carries roffset; input exhausted inside huff_sym is
    unexpected EOF (:818-826) but inside more_bits it is plain eof (:789-799, quirk D5); partial output is delivered."""
    assert oracle.inflate(bytes([0x07]), 16)[0] == ORC_CORRUPT                        # reserved block type
    st, out, eo, _ = oracle.inflate(bytes([0x01, 0x05, 0x00, 0x00, 0x00]), 16)         # LEN / NLEN mismatch
    assert (st, eo) == (ORC_CORRUPT, 5)
    assert oracle.inflate(b"", 16)[0] == ORC_EOF_AT_REFILL                            # nothing to read at next_block
    d = corpus.unit(5000, seed=2, index=0, klass=0)
    c = oracle.deflate(d)
    seen = set()
    for k in range(1, len(c) - 5, 13):
        st, out, eo, cons = oracle.inflate(c[:k], len(d) + 8)
        assert st in (ORC_UNEXPECTED_EOF, ORC_EOF_AT_REFILL)
        assert d.startswith(out) and cons == k
        seen.add(st)
        try:  # zlib agrees on how much is decodable from the truncated prefix
            z = zlib.decompressobj(-15).decompress(c[:k])
            assert z.startswith(out) or out.startswith(z)
        except zlib.error:
            pass
    assert ORC_UNEXPECTED_EOF in seen


def test_reader_hands_out_window_flushes(oracle, corpus):
    """Decompressor.read returns at most one 32 KiB window flush per call (inflate.mbt:382-407)."""
    import ctypes as C

    d = corpus.unit(100000, seed=4, index=0, klass=0)
    c = oracle.deflate(d)
    a = np.frombuffer(c, dtype=np.uint8)
    r = C.c_void_p(oracle.L.orc_reader_new(a.ctypes.data, len(c)))
    buf = np.zeros(1 << 20, np.uint8)
    st, eo = C.c_int(), C.c_int64()
    got, chunks = b"", []
    while True:
        n = oracle.L.orc_reader_read(r, buf.ctypes.data, buf.size, C.byref(st), C.byref(eo))
        got += buf[:n].tobytes()
        chunks.append(n)
        if st.value >= 0:
            break
    oracle.L.orc_reader_free(r)
    assert got == d and st.value == ORC_OK
    assert chunks[:3] == [32768, 32768, 32768] and max(chunks) == 32768


def test_deflate_bound_covers_worst_cases():
    """fb200_deflate_stream_bound / fb200_deflate_bound (the capacity a caller allocates) against what the
    reference's encoder really emits for its worst inputs: incompressible bytes (literal-only Huffman blocks,
    ~1.001 x, never stored: D2), tiny streams (5..26 bytes of framing), skewed alphabets (15-bit codes)."""
    import ctypes as C
    import moonbit_flate_b200 as fb
    L = fb._lib
    o = Oracle()
    rng = np.random.default_rng(3)
    worst = 0.0
    for n in [0, 1, 2, 15, 16, 17, 100, 127, 128, 129, 1000, 65534, 65535, 65536, 65537, 65662, 65663, 131070, 200001]:
        for kind in range(3):
            if kind == 0:
                d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
            elif kind == 1:
                p = 0.93 ** np.arange(256); p /= p.sum()
                d = rng.choice(256, size=n, p=p).astype(np.uint8).tobytes()
            else:
                d = bytes(n)
            c = len(o.deflate(d))
            b = L.fb200_deflate_stream_bound(n)
            assert c <= b, (n, kind, c, b)
            worst = max(worst, c / max(b, 1))
            assert c <= L.fb200_deflate_bound(n, 65536) or n == 0, (n, kind)
    assert worst < 0.75  # the bound is generous, not tight
