"""GPU parity tests proper: the CUDA path, called through the C ABI, against the
CPU oracle on the same seeded inputs (bit-exact: integer / byte work)."""
import os
import zlib

import numpy as np
import pytest

from helpers import BLK_DYNAMIC, Corpus, Oracle, fuzz_streams

pytestmark = pytest.mark.gpu


def C_u64():
    import ctypes
    return ctypes.c_uint64()


def need_addr(v):
    import ctypes
    return ctypes.addressof(v)


@pytest.fixture(scope="module", params=["cta", "warp"])
def ctx(request):
    """Every test runs twice: calls with few streams inflate with one CTA per stream (the default), or -- with
    FB200_INFLATE_CTA_STREAMS=0, read when the context is created -- with one warp per stream like big batches."""
    import moonbit_flate_b200 as fb

    old = os.environ.get("FB200_INFLATE_CTA_STREAMS")
    if request.param == "warp":
        os.environ["FB200_INFLATE_CTA_STREAMS"] = "0"
    try:
        c = fb.Context()
    finally:
        if old is None:
            os.environ.pop("FB200_INFLATE_CTA_STREAMS", None)
        else:
            os.environ["FB200_INFLATE_CTA_STREAMS"] = old
    c.inflate_kernel = request.param
    yield c
    c.close()


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.fixture(scope="module")
def corpus():
    return Corpus()


# sizes around every threshold of Compressor::enc_speed / write (deflate.mbt:236-294)
SIZES = [0, 1, 15, 16, 17, 100, 127, 128, 129, 300, 4096, 65534, 65535, 65536, 65537, 65551, 65552,
         65662, 65663, 70000, 131070, 131072, 200000]


@pytest.mark.parametrize("klass", [0, 1, 2, 3, 4, 5])
def test_deflate_stream_bytes_and_tokens(ctx, oracle, corpus, klass):
    """Each stream's compressed bytes (and the parse's tokens) are identical to the oracle's."""
    datas = [corpus.unit(n, seed=3, index=i, klass=klass) for i, n in enumerate(SIZES)]
    src = np.frombuffer(b"".join(datas), dtype=np.uint8)
    off = np.concatenate([[0], np.cumsum([len(d) for d in datas])]).astype(np.uint64)
    comp, doff = ctx.deflate_streams(src, off)
    nblocks = int(ctx.last_stats().nblocks)
    ntok, kind, bits, toks = ctx.last_blocks(nblocks, src.size + 16)
    b = 0
    t = 0
    for i, d in enumerate(datas):
        want, wtok, wntok, wkind, wbits = oracle.deflate_ex(d)
        got = comp[int(doff[i]): int(doff[i + 1])].tobytes()
        nb = len(wntok)
        assert list(kind[b: b + nb]) == list(wkind), (klass, len(d))
        # the parse runs for every block >= 128 bytes; compare tokens where both ran
        assert list(ntok[b: b + nb]) == list(wntok), (klass, len(d))
        k = int(wntok.sum())
        assert np.array_equal(toks[t: t + k], wtok), (klass, len(d))
        for q in range(nb):
            if wkind[q] != 0:
                assert int(bits[b + q]) == int(wbits[q]), (klass, len(d), q)
        assert got == want, (klass, len(d))
        assert zlib.decompress(got, -15) == d
        b += nb
        t += k
    assert b == nblocks


def test_deflate_segments_mixed(ctx, oracle, corpus):
    """north-star unit: 65536-byte segments = block(65535) + stored(1) + trailer, each an independent stream."""
    nseg, seg = 96, 65536
    src = corpus.fill(nseg, seg, seed=1)
    comp, off = ctx.deflate_segments(src, seg)
    assert off[0] == 0 and off[-1] == comp.size
    for i in range(nseg):
        d = src[i * seg: (i + 1) * seg].tobytes()
        got = comp[int(off[i]): int(off[i + 1])].tobytes()
        assert got == oracle.deflate(d), i


def test_deflate_ragged_last_segment(ctx, oracle, corpus):
    src = corpus.fill(5, 65536, seed=9)[: 4 * 65536 + 777]
    comp, off = ctx.deflate_segments(src, 65536)
    assert len(off) == 6
    for i in range(5):
        d = src[i * 65536: (i + 1) * 65536].tobytes()
        assert comp[int(off[i]): int(off[i + 1])].tobytes() == oracle.deflate(d)


def test_deflate_empty_batch(ctx):
    comp, off = ctx.deflate_segments(np.zeros(0, np.uint8), 65536)
    assert comp.size == 0 and list(off) == [0]


def test_deflate_1mib_single_stream(ctx, oracle, corpus):
    """BASELINE config 1: one 1 MiB text stream = 16 parsed blocks (table persists, quirk D1) + 16-byte stored tail."""
    d = corpus.unit(1 << 20, seed=1, index=0, klass=0)
    got = ctx.deflate(d)
    assert got == oracle.deflate(d)


def test_inflate_reference_streams(ctx, oracle, corpus):
    datas = [corpus.unit(n, seed=5, index=i, klass=i % 6) for i, n in enumerate(SIZES * 2)]
    comps = [oracle.deflate(d) for d in datas]
    comp = np.frombuffer(b"".join(comps), dtype=np.uint8)
    coff = np.concatenate([[0], np.cumsum([len(c) for c in comps])]).astype(np.uint64)
    ooff = np.concatenate([[0], np.cumsum([len(d) + 7 for d in datas])]).astype(np.uint64)
    out, olen, st, eo, cons = ctx.inflate_batch(comp, coff, ooff)
    for i, d in enumerate(datas):
        assert st[i] == 0, (i, st[i])
        assert int(olen[i]) == len(d)
        assert out[int(ooff[i]): int(ooff[i]) + len(d)].tobytes() == d
        assert int(cons[i]) == len(comps[i])


def test_inflate_foreign_streams(ctx, oracle, corpus):
    """zlib-produced streams: fixed Huffman, dynamic, stored, multi-block."""
    datas, comps = [], []
    for i, n in enumerate([0, 1, 100, 5000, 70000, 200000]):
        for klass in (0, 2, 3):
            d = corpus.unit(n, seed=11, index=i, klass=klass)
            for lvl, strat in ((0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                               (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY)):
                co = zlib.compressobj(lvl, zlib.DEFLATED, -15, 9, strat)
                datas.append(d)
                comps.append(co.compress(d) + co.flush())
    comp = np.frombuffer(b"".join(comps), dtype=np.uint8)
    coff = np.concatenate([[0], np.cumsum([len(c) for c in comps])]).astype(np.uint64)
    ooff = np.concatenate([[0], np.cumsum([len(d) for d in datas])]).astype(np.uint64)
    out, olen, st, eo, cons = ctx.inflate_batch(comp, coff, ooff)
    for i, d in enumerate(datas):
        ost, oout, oeo, ocons = oracle.inflate(comps[i], len(d))
        assert (int(st[i]), int(olen[i]), int(cons[i])) == (ost, len(oout), ocons), i
        assert out[int(ooff[i]): int(ooff[i + 1])].tobytes() == d


def _error_cases(oracle, corpus):
    d = corpus.unit(6000, seed=2, index=0, klass=0)
    c = oracle.deflate(d)
    cases = [c[:k] for k in range(0, len(c), 7)]
    rng = np.random.default_rng(4)
    for _ in range(300):
        b = bytearray(c)
        i = int(rng.integers(0, len(b)))
        b[i] ^= 1 << int(rng.integers(0, 8))
        cases.append(bytes(b))
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_FIXED)
    z = co.compress(d) + co.flush()
    cases += [z[:k] for k in range(0, len(z), 41)]
    for _ in range(100):
        b = bytearray(z)
        i = int(rng.integers(0, len(b)))
        b[i] ^= 1 << int(rng.integers(0, 8))
        cases.append(bytes(b))
    # reserved block type, bad stored LEN/NLEN, HLIT / HDIST out of range
    cases += [bytes([0x07]), bytes([0x01, 0x05, 0x00, 0x00, 0x00]), bytes([0x05, 0xfe, 0xff]), bytes([0xfd, 0xff, 0x03])]
    return cases, len(d) * 4 + 70000


def test_inflate_error_parity(ctx, oracle, corpus):
    """Truncated and bit-flipped streams: same status class, same `corrupt input before offset N`, same partial output."""
    cases, cap = _error_cases(oracle, corpus)
    comp = np.frombuffer(b"".join(cases), dtype=np.uint8)
    coff = np.concatenate([[0], np.cumsum([len(c) for c in cases])]).astype(np.uint64)
    ooff = (np.arange(len(cases) + 1) * cap).astype(np.uint64)
    out, olen, st, eo, cons = ctx.inflate_batch(comp, coff, ooff)
    for i, c in enumerate(cases):
        ost, oout, oeo, ocons = oracle.inflate(c, cap)
        assert int(st[i]) == ost, (i, int(st[i]), ost)
        assert int(olen[i]) == len(oout), (i, ost)
        assert out[int(ooff[i]): int(ooff[i]) + len(oout)].tobytes() == oout, i
        if ost == 1:
            assert int(eo[i]) == oeo, (i, int(eo[i]), oeo)
        assert int(cons[i]) == ocons, (i, ost, int(cons[i]), ocons)


def test_roundtrip_property_large(ctx, corpus):
    """Size-independent property at a larger size: inflate(deflate(x)) == x for 1024 mixed segments, and every
    GPU stream inflates under zlib."""
    nseg, seg = 1024, 65536
    src = corpus.fill(nseg, seg, seed=21)
    comp, off = ctx.deflate_segments(src, seg)
    ooff = (np.arange(nseg + 1) * seg).astype(np.uint64)
    out, olen, st, eo, cons = ctx.inflate_batch(comp, off, ooff)
    assert (st == 0).all() and (olen == seg).all()
    assert np.array_equal(out, src)
    for i in range(0, nseg, 37):
        assert zlib.decompress(comp[int(off[i]): int(off[i + 1])].tobytes(), -15) == src[i * seg:(i + 1) * seg].tobytes()


def test_writer_reader_mirror(ctx, oracle):
    """TestBestSpeed restated (deflate-fast_test.mbt:14-100) through the Writer / Reader mirror."""
    import io

    import moonbit_flate_b200 as fb

    abc = bytes(range(128)) * (131072 // 128)
    cases = [[65536, 0], [65536, 1, 256], [65536, 16, 65536], [65536, 127], [65536, 128, 256], [65536, 65536, 65536]]
    for tc in cases:
        for first_n in (1, 65534, 65535, 65536, 65537, 131072):
            sizes = [first_n] + tc[1:]
            buf = io.BytesIO()
            w = fb.Writer.new(buf, ctx)
            want = b""
            for n in sizes:
                assert w.write(abc[:n]) == (n, None)
                want += abc[:n]
            assert w.close() is None
            assert w.close() is None
            assert w.write(b"x") == (0, fb.WRITER_CLOSED_ERROR)
            comp = buf.getvalue()
            assert comp == oracle.writer_roundtrip([abc[:n] for n in sizes])
            got, err = fb.Reader.new(comp, ctx).read_all()
            assert err is None and got == want


@pytest.mark.parametrize("klass", [0, 1, 2, 3])
def test_writer_streams_window_by_window(ctx, oracle, corpus, klass):
    """Compressor::write encodes every 65535-byte window the moment it is full (deflate.mbt:280-294): the Writer
    hands the bytes of the windows a write has completed to its sink before the write returns -- whatever the
    chunking, the concatenation is the oracle's stream (the hash table, the 32768 bytes of history for the 4-byte
    cross-block check, the table-reset count and the bit position carry over from call to call)."""
    import io

    import moonbit_flate_b200 as fb

    data = corpus.unit(5 * 65535 + 4321, seed=77, index=klass, klass=klass)
    want = oracle.deflate(data)
    rng = np.random.default_rng(klass)
    patterns = [
        [len(data)],
        [65535] * 5 + [4321],
        [1, 65534, 65535 * 2, 1, len(data) - 65535 * 3 - 1],
        [3 * 65535, 2 * 65535 + 4321],
        [5 * 65535, 4321],          # close() with a small parsed tail
        [5 * 65535 + 4321 - 10, 10],
    ]
    cuts = sorted(rng.choice(len(data), 40, replace=False).tolist())
    patterns.append([b - a for a, b in zip([0] + cuts, cuts + [len(data)])])
    for sizes in patterns:
        assert sum(sizes) == len(data)
        buf = io.BytesIO()
        w = fb.Writer.new(buf, ctx)
        pos = 0
        for n in sizes:
            assert w.write(data[pos: pos + n]) == (n, None)
            pos += n
            if pos >= 65535:  # the completed windows have left before close
                assert len(buf.getvalue()) > 0, sizes
            # nothing of a window that is not full yet has been emitted: the prefix so far is a prefix of the stream
            assert want.startswith(buf.getvalue()), sizes
        assert w.close() is None
        assert buf.getvalue() == want, (klass, sizes[:6])
    # exact multiple of the window, then close with nothing pending; and a stream of one window + stored tail
    for n in (2 * 65535, 65535, 65535 + 7, 65535 + 100, 65535 + 128):
        d = data[:n]
        buf = io.BytesIO()
        w = fb.Writer.new(buf, ctx)
        assert w.write(d[:65535]) == (65535, None)
        assert w.write(d[65535:]) == (n - 65535, None)
        assert w.close() is None
        assert buf.getvalue() == oracle.deflate(d), n


def test_writer_pull_form_and_reader_consumed(ctx, oracle, corpus):
    """The pull form of the Writer (no callback: hosts like the MoonBit shim collect the bytes with
    fb200_writer_take) yields the same stream; the Reader reports how much of its input the decoder consumed, so
    that what follows the deflate stream -- a gzip / zlib trailer, the next member -- stays with the caller
    (the reference pulls its input byte by byte and stops behind the final block, inflate.mbt:789-799)."""
    import moonbit_flate_b200 as fb

    L = fb._lib
    data = corpus.unit(200000, seed=5, index=3, klass=0)
    want = oracle.deflate(data)
    w = L.fb200_writer_new(ctx._h, fb.SINK_FN(0), None)
    assert w
    got = bytearray()
    buf = np.empty(1 << 16, np.uint8)
    for a in range(0, len(data), 50000):
        chunk = np.frombuffer(data[a: a + 50000], dtype=np.uint8)
        assert L.fb200_writer_write(w, chunk.ctypes.data, chunk.size) == chunk.size
        while L.fb200_writer_pending(w):
            k = L.fb200_writer_take(w, buf.ctypes.data, buf.size)
            got += buf[:k].tobytes()
        if a + 50000 >= 65535:
            assert len(got) > 0  # the first window has left with the write that completed it
    assert L.fb200_writer_close(w) == 0
    while L.fb200_writer_pending(w):
        k = L.fb200_writer_take(w, buf.ctypes.data, buf.size)
        got += buf[:k].tobytes()
    L.fb200_writer_free(w)
    assert bytes(got) == want
    # consumed: exactly the deflate stream, whatever follows it
    for tail in (b"", b"\x01\x02\x03\x04trailer", want):
        r = fb.Reader.new(want + tail, ctx)
        out, err = r.read_all()
        assert err is None and out == data
        assert r.consumed() == len(want)
    st, _, _, cons = oracle.inflate(want[: len(want) // 2], len(data) + 1)
    r = fb.Reader.new(want[: len(want) // 2], ctx)
    r.read_all()
    assert r.consumed() == cons


def test_writer_dict_kat(ctx):
    """deflate_test.mbt:12-35: 28 bytes -> exactly 38 bytes; new_dict(dict)+write(text) == write(dict)+write(text)."""
    import io

    import moonbit_flate_b200 as fb

    b = io.BytesIO()
    w = fb.Writer.new(b, ctx)
    assert w.write(b"hello world") == (11, None)
    assert w.write(b"hello again world") == (17, None)
    assert w.close() is None
    want = b.getvalue()
    assert len(want) == 38
    b1 = io.BytesIO()
    w = fb.Writer.new_dict(b1, b"hello world", ctx)
    assert w.write(b"hello again world") == (17, None)
    assert w.close() is None
    assert b1.getvalue() == want


def test_host_api_chunk_pipeline(oracle, corpus, monkeypatch):
    """The host-buffer entry points cut a batch into chunks of whole streams and pipeline H2D / kernels / D2H,
    and the deflate packs and copies back group by group (groups share output words at their edges);
    neither may change a byte.  Forced here with 1 MiB chunks and 1 MiB groups over ragged stream sizes."""
    import moonbit_flate_b200 as fb

    monkeypatch.setenv("FB200_CHUNK_MB", "1")
    monkeypatch.setenv("FB200_GROUP_MB", "1")
    c = fb.Context()
    monkeypatch.delenv("FB200_CHUNK_MB")
    monkeypatch.delenv("FB200_GROUP_MB")
    try:
        rng = np.random.default_rng(8)
        sizes = [int(x) for x in rng.integers(0, 300000, 40)] + [0, 1, 65536, 65536, 2_500_000]
        datas = [corpus.unit(n, seed=31, index=i, klass=i % 6) for i, n in enumerate(sizes)]
        src = np.frombuffer(b"".join(datas), dtype=np.uint8)
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
        comp, doff = c.deflate_streams(src, off)
        assert int(doff[-1]) == comp.size
        for i in range(len(datas)):
            assert comp[int(doff[i]): int(doff[i + 1])].tobytes() == oracle.deflate(datas[i]), i
        ooff = np.concatenate([[0], np.cumsum([n + 5 for n in sizes])]).astype(np.uint64)
        out, olen, st, eo, cons = c.inflate_batch(comp, doff, ooff)
        assert (st == 0).all() and list(olen) == sizes
        for i, d in enumerate(datas):
            assert out[int(ooff[i]): int(ooff[i]) + len(d)].tobytes() == d, i
            assert int(cons[i]) == int(doff[i + 1] - doff[i])
        # fixed-size segments with a ragged tail, several chunks
        seg = corpus.fill(70, 65536, seed=33)[: 69 * 65536 + 123]
        comp2, off2 = c.deflate_segments(seg, 65536)
        for i in range(70):
            assert comp2[int(off2[i]): int(off2[i + 1])].tobytes() == oracle.deflate(seg[i * 65536:(i + 1) * 65536].tobytes()), i
        # capacity error reports the size the call needs
        need = C_u64()
        small = np.empty(1000, np.uint8)
        so = np.zeros(71, np.uint64)
        rc = fb._lib.fb200_deflate_segments(c._h, seg.ctypes.data, seg.size, 65536, small.ctypes.data, small.size,
                                            so.ctypes.data, need_addr(need))
        assert rc == fb.ERR_DST_TOO_SMALL and need.value == comp2.size
    finally:
        c.close()


def test_config3_many_small_streams(ctx, oracle, corpus):
    """BASELINE configs[2] in small: a batch of independent 1..16 KiB streams (ragged offsets, every class):
    every GPU stream equals the oracle's on a sample, all of them inflate back bit-exactly, and the decoder
    consumes exactly each stream."""
    nstreams = 20000
    src, off = corpus.fill_var(nstreams, seed=5)
    comp, doff = ctx.deflate_streams(src, off)
    assert int(doff[-1]) == comp.size
    for i in range(0, nstreams, 397):
        d = src[int(off[i]): int(off[i + 1])].tobytes()
        assert comp[int(doff[i]): int(doff[i + 1])].tobytes() == oracle.deflate(d), i
    out, olen, st, eo, cons = ctx.inflate_batch(comp, doff, off)
    assert (st == 0).all()
    assert np.array_equal(olen, np.diff(off))
    assert np.array_equal(cons, np.diff(doff))
    assert np.array_equal(out, src)


@pytest.mark.parametrize("klass", [2, 4, 5, 3])
def test_config5_worst_cases(ctx, oracle, corpus, klass):
    """BASELINE configs[4] in small: incompressible random bytes (every block a literal-only Huffman block --
    the reference never stores, quirk D2) and long-run data (maximum-length matches), 512 segments each."""
    nseg, seg = 512, 65536
    src = corpus.fill(nseg, seg, seed=41, klass=klass)
    comp, off = ctx.deflate_segments(src, seg)
    for i in range(0, nseg, 61):
        d = src[i * seg:(i + 1) * seg].tobytes()
        assert comp[int(off[i]): int(off[i + 1])].tobytes() == oracle.deflate(d), (klass, i)
    if klass == 2:
        assert comp.size > src.size  # ~1.001x: dynamic literal-only blocks, no stored data blocks
    out, olen, st, eo, cons = ctx.inflate_batch(comp, off, np.arange(nseg + 1, dtype=np.uint64) * seg)
    assert (st == 0).all() and (olen == seg).all() and np.array_equal(out, src)


def test_flat_codes_stay_on_the_fast_path(ctx, corpus):
    """Incompressible streams of ragged sizes, single- and multi-block: their literal-only codes (255 symbols of 8
    bits) give the speculative decoder nothing to synchronise on, so it searches the start phase of every range (a
    wrong phase runs into EOB early).  The result must be exact AND come from the fast kernel: a wrong guess that
    survived would be caught by its consistency checks and handed to the exact kernel, which this test forbids."""
    rng = np.random.default_rng(77)
    sizes = [1000, 4096, 65535, 65536, 65537, 100000, 131070, 200000, 262144, 300001] + [int(x) for x in rng.integers(2000, 150000, 90)]
    parts = [corpus.unit(n, seed=78, index=i, klass=2) for i, n in enumerate(sizes)]
    src = np.frombuffer(b"".join(parts), dtype=np.uint8)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    comp, doff = ctx.deflate_streams(src, off)
    out, olen, st, eo, cons = ctx.inflate_batch(comp, doff, off)
    assert (st == 0).all() and np.array_equal(olen, np.diff(off)) and np.array_equal(out, src)
    assert int(ctx.last_stats().inflate_fallbacks) == 0


def test_fuzz_deflate_inflate_against_oracle(ctx, oracle):
    rng = np.random.default_rng(20260101)
    datas = fuzz_streams(rng, 420)
    src = np.frombuffer(b"".join(datas), dtype=np.uint8)
    off = np.concatenate([[0], np.cumsum([len(d) for d in datas])]).astype(np.uint64)
    comp, doff = ctx.deflate_streams(src, off)
    for i, d in enumerate(datas):
        got = comp[int(doff[i]): int(doff[i + 1])].tobytes()
        assert got == oracle.deflate(d), (i, i % 7, len(d))
    out, olen, st, eo, cons = ctx.inflate_batch(comp, doff, off)
    assert (st == 0).all() and np.array_equal(out, src)
    assert np.array_equal(cons, np.diff(doff))


def test_peer_frame_single_rank(ctx, corpus):
    """Frame assembly through the IPC frame buffer (fb200_mg_*), world size 1: header, sizes and payload of the
    frame equal the rank's streams; the IPC handle of the buffer can be exported."""
    import torch
    from moonbit_flate_b200 import multigpu as mg
    nseg, seg = 64, 65536
    src = corpus.fill(nseg, seg, seed=9)
    comp, off = ctx.deflate_segments(src, seg)
    dev = torch.device("cuda", 0)
    payload = torch.from_numpy(comp.copy()).to(dev)
    sizes = torch.from_numpy(np.diff(off).astype(np.int64)).to(dev)
    pf = mg.PeerFrame(ctx, 0, 1, mg.frame_header_bytes(nseg) + comp.size + 64, dev)
    assert pf.available
    total = pf.put(payload, sizes, seg)
    pf.wait()
    torch.cuda.synchronize()
    seg_size, n, sizes_f, hdr = mg.parse_frame(pf.view)
    assert (seg_size, n, hdr, total) == (seg, nseg, mg.frame_header_bytes(nseg), hdr + comp.size)
    assert np.array_equal(sizes_f.numpy(), np.diff(off).astype(np.int64))
    assert np.array_equal(pf.view[hdr:total].cpu().numpy(), comp)
    # the way back (fb200_mg_get): ranges of segments out of the frame, inflated, against the input
    for first, count in ((0, nseg), (5, 17), (nseg - 1, 1), (nseg, 0), (20, 0)):
        d_comp = torch.zeros(comp.size + 16, dtype=torch.uint8, device=dev)
        d_coff = torch.zeros(count + 1, dtype=torch.int64, device=dev)
        seg_size, n, nbytes = pf.get(first, count, d_comp, d_coff)
        assert (seg_size, n) == (seg, nseg) and nbytes == int(off[first + count] - off[first])
        assert np.array_equal(d_coff.cpu().numpy().astype(np.uint64), (off[first: first + count + 1] - off[first]))
        assert np.array_equal(d_comp[:nbytes].cpu().numpy(), comp[int(off[first]): int(off[first + count])])
        if count:
            d_out = torch.zeros(count * seg, dtype=torch.uint8, device=dev)
            d_ooff = torch.arange(count + 1, dtype=torch.int64, device=dev) * seg
            d_olen = torch.zeros(count, dtype=torch.int64, device=dev)
            d_st = torch.zeros(count, dtype=torch.int32, device=dev)
            d_eo = torch.zeros(count, dtype=torch.int64, device=dev)
            ctx.inflate_batch_dev(d_comp.data_ptr(), d_coff.data_ptr(), count, d_out.data_ptr(), d_ooff.data_ptr(),
                                  d_olen.data_ptr(), d_st.data_ptr(), d_eo.data_ptr())
            assert int(d_st.abs().sum()) == 0
            assert np.array_equal(d_out.cpu().numpy(), src[first * seg: (first + count) * seg])
    # the asynchronous form: offsets and sizes are final on return, the payload after mg_wait
    d_comp = torch.zeros(comp.size + 16, dtype=torch.uint8, device=dev)
    d_coff = torch.zeros(31, dtype=torch.int64, device=dev)
    seg_size, n, nbytes = pf.get_begin(7, 30, d_comp, d_coff)
    assert (seg_size, n) == (seg, nseg) and nbytes == int(off[37] - off[7])
    ctx.mg_wait()
    assert np.array_equal(d_coff.cpu().numpy().astype(np.uint64), off[7:38] - off[7])
    assert np.array_equal(d_comp[:nbytes].cpu().numpy(), comp[int(off[7]): int(off[37])])
    import moonbit_flate_b200 as fb
    with pytest.raises(fb.FlateError):
        pf.get(nseg - 1, 2, torch.zeros(16, dtype=torch.uint8, device=dev), torch.zeros(3, dtype=torch.int64, device=dev))
    pf.close()


def test_stream_beyond_2gib_table_reset(ctx, oracle, corpus):
    """One stream longer than buffer_reset (deflate-fast.mbt:55, :129-132): about 2 GiB into a stream
    DeflateFast.cur reaches buffer_reset and shift_offsets clears the hash table (prev is always empty, D1), so
    the block that follows finds no 4-byte matches into its predecessor.  The GPU stream (block-parallel parse,
    closed-form reset blocks) must equal the oracle's byte for byte (the oracle restates shift_offsets) and decode
    back to the input (zlib here; the GPU's own inflate of a single 2 GiB stream is one warp's serial work and runs
    with FB200_SLOW_TESTS=1)."""
    if ctx.inflate_kernel == "warp":
        pytest.skip("deflate-side test: once is enough")
    nblk = 32770  # the reset happens at the start of block 32766
    n = nblk * 65535 + 777
    src = corpus.fill((n + 65535) // 65536, 65536, seed=77, klass=Corpus.TEXT)[:n]
    off = np.array([0, n], dtype=np.uint64)
    comp, doff = ctx.deflate_streams(src, off)
    want = np.frombuffer(oracle.deflate(src), dtype=np.uint8)
    assert comp.size == want.size
    assert np.array_equal(comp, want)
    del want
    d = zlib.decompressobj(-15)
    pos = 0
    for a in range(0, comp.size, 1 << 26):
        piece = d.decompress(comp[a: a + (1 << 26)].tobytes())
        assert piece == src[pos: pos + len(piece)].tobytes()
        pos += len(piece)
    assert pos == n and d.eof
    if os.environ.get("FB200_SLOW_TESTS") == "1":
        out, olen, st, eo, cons = ctx.inflate_batch(comp, doff, off)
        assert int(st[0]) == 0 and int(olen[0]) == n and int(cons[0]) == comp.size
        assert np.array_equal(out, src)


def test_inflate_long_codes_next_to_short_ones(ctx, oracle):
    """Literal-heavy data whose Huffman code reaches 13-15 bits for rare bytes: a rare byte followed by two common
    ones puts three literals into one 32-bit peek of the parallel decoder only if the first two codes leave room
    for the third look-up (regression: the third literal used to be read past the peek)."""
    rng = np.random.default_rng(77)
    datas = []
    for k in range(24):
        ratio = 0.80 + 0.008 * k  # geometric byte frequencies: code lengths from 2-3 bits up to 15
        p = ratio ** np.arange(256)
        p /= p.sum()
        perm = rng.permutation(256)
        n = 60000 + 211 * k
        datas.append(perm[rng.choice(256, size=n, p=p)].astype(np.uint8).tobytes())
    src = np.frombuffer(b"".join(datas), dtype=np.uint8)
    off = np.concatenate([[0], np.cumsum([len(d) for d in datas])]).astype(np.uint64)
    comp, doff = ctx.deflate_streams(src, off)
    for i, d in enumerate(datas):
        assert comp[int(doff[i]): int(doff[i + 1])].tobytes() == oracle.deflate(d), i
    out, olen, st, eo, cons = ctx.inflate_batch(comp, doff, off)
    assert (st == 0).all() and np.array_equal(olen, np.diff(off))
    assert np.array_equal(out, src)
    assert int(ctx.last_stats().inflate_fallbacks) == 0


@pytest.mark.parametrize("mode", ["0", "2"])
def test_multi_block_streams_both_parse_paths(oracle, corpus, monkeypatch, mode):
    """Multi-block streams are parsed either by one warp per stream (FB200_PARSE_BLOCKPAR=0) or block-parallel by
    fixpoint rounds (=2; the default picks by the number of such streams): both must give the oracle's bytes,
    for streams whose last block is parsed, stored (< 17 bytes), literal-only (< 128) or missing."""
    import moonbit_flate_b200 as fb

    monkeypatch.setenv("FB200_PARSE_BLOCKPAR", mode)
    c = fb.Context()
    monkeypatch.delenv("FB200_PARSE_BLOCKPAR")
    try:
        sizes = [65535 + 128, 2 * 65535, 2 * 65535 + 5, 2 * 65535 + 100, 3 * 65535 + 4000, 700000, 65536, 1000, 0,
                 5 * 65535 + 127, 5 * 65535 + 128]
        datas = [corpus.unit(n, seed=61, index=i, klass=(0, 1, 3, 5, 0, -1, 0, 0, 0, 1, 2)[i]) for i, n in enumerate(sizes)]
        src = np.frombuffer(b"".join(datas), dtype=np.uint8)
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
        comp, doff = c.deflate_streams(src, off)
        for i, d in enumerate(datas):
            assert comp[int(doff[i]): int(doff[i + 1])].tobytes() == oracle.deflate(d), (mode, i, sizes[i])
        out, olen, st, eo, cons = c.inflate_batch(comp, doff, off)
        assert (st == 0).all() and np.array_equal(out, src)
    finally:
        c.close()


def test_reader_preset_dictionary(ctx, oracle, corpus):
    """&Reader::new_dict / Decompressor::reset (inflate.mbt:310-317, :862-883; DictDecoder::new,
    dict-decoder.mbt:42-60): streams whose back-references reach into a preset dictionary (made with zlib's
    zdict: the reference's own Writer::new_dict compresses the dictionary into the output instead, D3).
    Output, error class and offset equal the oracle's, with the dictionary, without it (the reference then
    reports corrupt input: dist > hist_size), and for a dictionary longer than the 32 KiB window."""
    import moonbit_flate_b200 as fb

    for dlen in (100, 5000, 32768, 50000):
        dict_ = corpus.unit(dlen, seed=91, index=dlen, klass=0)
        data = dict_[-3000:] + corpus.unit(70000, seed=92, index=dlen, klass=0) + dict_[:2000] + dict_[-500:]
        co = zlib.compressobj(6, zlib.DEFLATED, -15, zdict=dict_)
        comp = co.compress(data) + co.flush()
        ost, oout, oeo = oracle.inflate_dict(comp, dict_)
        assert ost == 0 and oout == data
        r = fb.Reader.new_dict(comp, dict_, ctx)
        got, err = r.read_all()
        assert err is None and got == data, dlen
        assert r.close() is None
        # the same stream without its dictionary: the first reference into it is corrupt input, same offset
        ost2, oout2, oeo2 = oracle.inflate_dict(comp, b"")
        got2, err2 = fb.Reader.new(comp, ctx).read_all()
        assert ost2 == fb.ST_CORRUPT and err2 == fb.corrupt_input_error(oeo2) and got2 == oout2, (dlen, err2, oeo2)
        # reset: same object, other input and dictionary
        r.reset(comp, dict_)
        got3, err3 = r.read_all()
        assert err3 is None and got3 == data
        r.reset(oracle.deflate(b"plain stream, no dictionary"))
        assert r.read_all() == (b"plain stream, no dictionary", None)


def _oracle_read_sequence(oracle, comp: bytes, dict_: bytes, want: int):
    """(bytes, status) of every orc_reader_read(want) call until a status arrives."""
    import ctypes as C

    a = np.frombuffer(comp, dtype=np.uint8)
    d = np.frombuffer(dict_, dtype=np.uint8)
    r = C.c_void_p(oracle.L.orc_reader_new_dict(a.ctypes.data if len(comp) else None, len(comp),
                                                d.ctypes.data if len(dict_) else None, len(dict_)))
    buf = np.empty(max(want, 1), np.uint8)
    seq = []
    st, eo = C.c_int(-1), C.c_int64(0)
    while st.value < 0:
        k = oracle.L.orc_reader_read(r, buf.ctypes.data, want, C.byref(st), C.byref(eo))
        seq.append((buf[:k].tobytes(), 0 if st.value == 5 else st.value, eo.value if st.value == 1 else 0))
    oracle.L.orc_reader_free(r)
    return seq


def _gpu_read_sequence(ctx, comp: bytes, dict_: bytes, want: int):
    import ctypes as C
    import moonbit_flate_b200 as fb

    r = fb.Reader.new_dict(comp, dict_, ctx) if dict_ else fb.Reader.new(comp, ctx)
    buf = np.empty(max(want, 1), np.uint8)
    seq = []
    st, eo = C.c_int32(-1), C.c_int64(0)
    while st.value < 0:
        k = fb._lib.fb200_reader_read(r._h, buf.ctypes.data, want, C.byref(st), C.byref(eo))
        stv = 0 if st.value == fb.ST_EOF_AT_REFILL else st.value  # the reference reports plain ioeof for both
        seq.append((buf[: int(k)].tobytes(), stv, eo.value if st.value == 1 else 0))
    return seq


@pytest.mark.parametrize("want", [1 << 20, 32768, 10000])
def test_reader_read_granularity(ctx, oracle, corpus, want):
    """Decompressor.read (inflate.mbt:382-407) hands out at most one 32 KiB window flush per call, and the final
    status rides on the read that drains the last partial flush -- or arrives alone, with 0 bytes, when the output
    ended exactly on a flush.  With a preset dictionary of D bytes the window starts at wr_pos = D
    (dict-decoder.mbt:56-62), which shifts every flush boundary.  fb200_reader_read must return the same sequence
    of (bytes, status) as the oracle's reader: valid streams, exact multiples of the window, corrupt and truncated
    streams (partial output, then the error), dictionaries of 100 / 5000 / 32768 / 50000 bytes."""
    cases = []
    for n in (0, 1, 32767, 32768, 32769, 65536, 100000, 3 * 32768):
        d = corpus.unit(n, seed=31, index=n, klass=0)
        cases.append((oracle.deflate(d), b""))
    big = corpus.unit(150000, seed=32, index=1, klass=0)
    cbig = oracle.deflate(big)
    cases.append((cbig[: len(cbig) // 2], b""))                       # truncated: unexpected EOF after partial output
    bad = bytearray(cbig)
    bad[len(bad) // 2] ^= 0x55
    cases.append((bytes(bad), b""))                                   # (most likely) corrupt somewhere in the middle
    for dlen in (100, 5000, 32768, 50000):
        dict_ = corpus.unit(dlen, seed=91, index=dlen, klass=0)
        for n in (70000, 32768 - (dlen % 32768), 65536 - (dlen % 32768), 98304):
            data = dict_[-3000:] + corpus.unit(n, seed=92, index=dlen, klass=0)
            data = data[:n] if n >= 3000 else data
            co = zlib.compressobj(6, zlib.DEFLATED, -15, zdict=dict_)
            cases.append((co.compress(data) + co.flush(), dict_))
    for comp, dict_ in cases:
        want_seq = _oracle_read_sequence(oracle, comp, dict_, want)
        got_seq = _gpu_read_sequence(ctx, comp, dict_, want)
        assert [(len(b), s, e) for b, s, e in got_seq] == [(len(b), s, e) for b, s, e in want_seq], (len(comp), len(dict_))
        assert got_seq == want_seq


def test_two_contexts_two_gpus_one_process(oracle, corpus):
    """One context per GPU inside ONE process (the C ABI's contract): device scratch and kernel attributes are
    per context / per device.  Skipped on a single-GPU box."""
    import torch
    import moonbit_flate_b200 as fb

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctxs = [fb.Context(0), fb.Context(1)]
    try:
        src = corpus.fill(96, 65536, seed=13)
        long_stream = corpus.fill(12, 65536, seed=14)  # one multi-block stream: block-parallel parse
        for rep in range(2):
            for c in ctxs:
                comp, off = c.deflate_segments(src, 65536)
                for i in range(0, 96, 17):
                    assert comp[int(off[i]): int(off[i + 1])].tobytes() == oracle.deflate(src[i * 65536:(i + 1) * 65536].tobytes())
                out, olen, st, eo, cons = c.inflate_batch(comp, off, np.arange(97, dtype=np.uint64) * 65536)
                assert (st == 0).all() and np.array_equal(out, src)
                assert c.deflate(long_stream.tobytes()) == oracle.deflate(long_stream.tobytes())
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("overlap", ["auto", "forced"])
def test_smoke_with_blocking_launches(overlap):
    """Tools that make kernel launches synchronous (ncu, compute-sanitizer, cuda-gdb, CUDA_LAUNCH_BLOCKING=1) must
    not deadlock the host-buffer calls, whose kernels wait on a watermark the H2D copy stream advances: every copy
    is queued before the kernel is launched, and with such a tool attached the calls copy first and launch after
    (fb200_ctx::overlap_h2d).  `forced` keeps the watermark overlap on under blocking launches: feed-first alone
    has to be enough there."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
    env.pop("FB200_HOST_OVERLAP", None)
    if overlap == "forced":
        env["FB200_HOST_OVERLAP"] = "1"
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=root, env=env,
                       timeout=180, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "smoke ok" in r.stdout
