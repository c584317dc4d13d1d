"""Shared test helpers: ctypes views of the oracle (checker), the corpus tool and the host model."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ORC_OK, ORC_CORRUPT, ORC_UNEXPECTED_EOF, ORC_DST_TOO_SMALL, ORC_INTERNAL, ORC_EOF_AT_REFILL = 0, 1, 2, 3, 4, 5
BLK_STORED, BLK_HUFF, BLK_DYNAMIC = 0, 1, 2


def _load(rel):
    return C.CDLL(os.path.join(ROOT, rel))


class Oracle:
    def __init__(self):
        L = self.L = _load("oracle/libflate_oracle.so")
        L.orc_deflate.restype = C.c_int64
        L.orc_deflate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_deflate_bound.restype = C.c_size_t
        L.orc_deflate_bound.argtypes = [C.c_size_t]
        L.orc_deflate_ex.restype = C.c_int64
        L.orc_deflate_ex.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                     C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_inflate.restype = C.c_int
        L.orc_inflate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                  C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_huff_generate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_writer_new.restype = C.c_void_p
        L.orc_writer_new_dict.restype = C.c_void_p
        L.orc_writer_new_dict.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_writer_write.restype = C.c_int64
        L.orc_writer_write.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_writer_close.argtypes = [C.c_void_p]
        L.orc_writer_data.restype = C.c_void_p
        L.orc_writer_data.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        L.orc_writer_free.argtypes = [C.c_void_p]
        L.orc_reader_new.restype = C.c_void_p
        L.orc_reader_new.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_reader_read.restype = C.c_size_t
        L.orc_reader_read.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int64)]
        L.orc_reader_free.argtypes = [C.c_void_p]
        L.orc_reader_new_dict.restype = C.c_void_p
        L.orc_reader_new_dict.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        for f in ("orc_token_offset", "orc_reverse16", "orc_reverse_bits", "orc_fixed_chunk"):
            getattr(L, f).restype = C.c_uint32
        L.orc_token_offset.argtypes = [C.c_uint32]
        L.orc_reverse16.argtypes = [C.c_uint32]
        L.orc_reverse_bits.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_fixed_chunk.argtypes = [C.c_int]
        L.orc_length_code.argtypes = [C.c_uint32]
        L.orc_offset_code.argtypes = [C.c_uint32]
        L.orc_dict_new.restype = C.c_void_p
        L.orc_dict_new.argtypes = [C.c_int, C.c_void_p, C.c_size_t]
        for f in ("orc_dict_free", "orc_dict_hist_size", "orc_dict_avail_read", "orc_dict_avail_write"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_dict_write.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_dict_write_copy.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_dict_try_write_copy.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_dict_read_flush.argtypes = [C.c_void_p, C.c_void_p]

    def deflate(self, data: bytes) -> bytes:
        a = np.frombuffer(data, dtype=np.uint8)
        cap = self.L.orc_deflate_bound(len(data))
        out = np.empty(cap, np.uint8)
        n = self.L.orc_deflate(a.ctypes.data if len(data) else None, len(data), out.ctypes.data, cap)
        assert n >= 0
        return out[:n].tobytes()

    def deflate_ex(self, data: bytes):
        """-> (compressed bytes, tokens uint32[], blk_ntok, blk_kind, blk_bits)"""
        n = len(data)
        a = np.frombuffer(data, dtype=np.uint8)
        cap = self.L.orc_deflate_bound(n)
        out = np.empty(cap, np.uint8)
        nb_cap = n // 65535 + 2
        toks = np.zeros(n + 16, np.uint32)
        ntok = np.zeros(nb_cap, np.uint32)
        kind = np.zeros(nb_cap, np.uint8)
        bits = np.zeros(nb_cap, np.uint64)
        ol = C.c_size_t()
        nb = self.L.orc_deflate_ex(a.ctypes.data if n else None, n, out.ctypes.data, cap, C.byref(ol),
                                   toks.ctypes.data, n + 16, ntok.ctypes.data, kind.ctypes.data, bits.ctypes.data, nb_cap)
        assert nb >= 0
        return out[: ol.value].tobytes(), toks[: int(ntok[:nb].sum())], ntok[:nb], kind[:nb], bits[:nb]

    def inflate(self, comp: bytes, cap: int):
        """-> (status, output bytes, err_off, consumed)"""
        a = np.frombuffer(comp, dtype=np.uint8)
        out = np.empty(max(cap, 1), np.uint8)
        ol = C.c_size_t()
        eo = C.c_int64()
        cons = C.c_int64()
        st = self.L.orc_inflate(a.ctypes.data if len(comp) else None, len(comp), out.ctypes.data, cap, C.byref(ol),
                                C.byref(eo), C.byref(cons))
        return st, out[: ol.value].tobytes(), eo.value, cons.value

    def inflate_dict(self, comp: bytes, dict_: bytes):
        """&Reader::new_dict(dict) + read until an error / ioeof -> (status, output bytes, err_off)"""
        a = np.frombuffer(comp, dtype=np.uint8)
        d = np.frombuffer(dict_, dtype=np.uint8)
        r = C.c_void_p(self.L.orc_reader_new_dict(a.ctypes.data if len(comp) else None, len(comp),
                                                  d.ctypes.data if len(dict_) else None, len(dict_)))
        out = bytearray()
        buf = np.empty(1 << 16, np.uint8)
        st, eo = C.c_int(-1), C.c_int64(0)
        while st.value < 0:
            k = self.L.orc_reader_read(r, buf.ctypes.data, buf.size, C.byref(st), C.byref(eo))
            out += buf[:k].tobytes()
        self.L.orc_reader_free(r)
        return st.value, bytes(out), eo.value

    def huff_generate(self, freq, max_bits):
        f = np.ascontiguousarray(np.asarray(freq, dtype=np.int32))
        lens = np.zeros(f.size, np.uint8)
        codes = np.zeros(f.size, np.uint16)
        self.L.orc_huff_generate(f.ctypes.data, f.size, max_bits, lens.ctypes.data, codes.ctypes.data)
        return lens, codes

    def writer_roundtrip(self, writes, dict_=None):
        """Writer::new(/new_dict) + write(each) + close -> compressed bytes"""
        L = self.L
        if dict_ is None:
            w = L.orc_writer_new()
        else:
            d = np.frombuffer(dict_, dtype=np.uint8)
            w = L.orc_writer_new_dict(d.ctypes.data, len(dict_))
        w = C.c_void_p(w)
        for b in writes:
            a = np.frombuffer(b, dtype=np.uint8)
            assert L.orc_writer_write(w, a.ctypes.data if len(b) else None, len(b)) == len(b)
        assert L.orc_writer_close(w) == 0
        assert L.orc_writer_close(w) == 0  # second close -> None
        assert L.orc_writer_write(w, None, 0) == -1  # writer closed
        n = C.c_size_t()
        p = L.orc_writer_data(w, C.byref(n))
        out = C.string_at(p, n.value)
        L.orc_writer_free(w)
        return out


class Corpus:
    TEXT, RECORDS, RANDOM, RUNS, CONST, PERIOD7 = 0, 1, 2, 3, 4, 5
    MIXED = -1

    def __init__(self):
        L = self.L = _load("tools/libfb_corpus.so")
        L.fb_corpus_unit.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_int]
        L.fb_corpus_fill.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int]
        L.fb_corpus_class.argtypes = [C.c_uint64, C.c_uint64]
        L.fb_corpus_var_len.restype = C.c_uint32
        L.fb_corpus_var_len.argtypes = [C.c_uint64, C.c_uint64]
        L.fb_corpus_fill_var.restype = C.c_uint64
        L.fb_corpus_fill_var.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]

    def unit(self, n, seed=1, index=0, klass=-1) -> bytes:
        b = np.zeros(max(n, 1), np.uint8)
        self.L.fb_corpus_unit(b.ctypes.data, n, seed, index, klass)
        return b[:n].tobytes()

    def fill(self, nunit, unit_size, seed=1, first=0, klass=-1) -> np.ndarray:
        b = np.zeros(nunit * unit_size, np.uint8)
        self.L.fb_corpus_fill(b.ctypes.data, first, nunit, unit_size, seed, klass)
        return b

    def fill_var(self, nunit, seed=1, first=0, klass=-1):
        off = np.zeros(nunit + 1, np.uint64)
        total = self.L.fb_corpus_fill_var(None, off.ctypes.data, first, nunit, seed, klass)
        b = np.zeros(total, np.uint8)
        self.L.fb_corpus_fill_var(b.ctypes.data, off.ctypes.data, first, nunit, seed, klass)
        return b, off


class HostModel:
    def __init__(self):
        L = self.L = _load("tests/hostmodel/libfb_hostmodel.so")
        L.fbm_generate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.fbm_build_block.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p,
                                      C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.fbm_codes.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
        L.fbm_parse_stream.restype = C.c_int64
        L.fbm_parse_stream.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        for fn in (L.fbm_parse_stream_v2, L.fbm_parse_stream_v3, L.fbm_parse_stream_v4):
            fn.restype = C.c_int64
            fn.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]

    def generate(self, freq, max_bits):
        f = np.ascontiguousarray(np.asarray(freq, dtype=np.uint32))
        lens = np.zeros(f.size, np.uint8)
        codes = np.zeros(f.size, np.uint16)
        self.L.fbm_generate(f.ctypes.data, f.size, max_bits, lens.ctypes.data, codes.ctypes.data)
        return lens, codes

    def parse_stream(self, data: bytes, v2: bool = False, v3: bool = False, v4: bool = False, lowest_wins: bool = True):
        """v2 / v3 / v4: the multi-match batch / the unified fast batch / the fixed windows of parse.cu; lowest_wins selects which lane of a
        bucket group wins the speculative insert (arbitrary on the GPU: both extremes must be exact)."""
        n = len(data)
        a = np.frombuffer(data, dtype=np.uint8)
        nb_cap = n // 65535 + 2
        toks = np.zeros(n + 16, np.uint32)
        ntok = np.zeros(nb_cap, np.uint32)
        if v2 or v3 or v4:
            stats = np.zeros(8, np.uint64)
            stats[3] = 1 if lowest_wins else 0
            fn = self.L.fbm_parse_stream_v4 if v4 else (self.L.fbm_parse_stream_v3 if v3 else self.L.fbm_parse_stream_v2)
            tot = fn(a.ctypes.data if n else None, n, toks.ctypes.data, n + 16, ntok.ctypes.data, nb_cap,
                     stats.ctypes.data)
            self.last_stats = stats
        else:
            tot = self.L.fbm_parse_stream(a.ctypes.data if n else None, n, toks.ctypes.data, n + 16,
                                          ntok.ctypes.data, nb_cap)
        assert tot >= 0
        return toks[:tot], ntok[: (n + 65534) // 65535]

    def codes(self, xlen: int, xoff: int):
        out = (C.c_int * 6)()
        self.L.fbm_codes(xlen, xoff, out)
        return list(out)

    def build_block(self, freq320, kind, n):
        f = np.ascontiguousarray(np.asarray(freq320, dtype=np.uint32)).copy()
        codes = np.zeros(320, np.uint32)
        hdr = np.zeros(160, np.uint32)
        hb = C.c_uint32()
        bb = C.c_uint32()
        k = self.L.fbm_build_block(f.ctypes.data, kind, n, codes.ctypes.data, hdr.ctypes.data, C.byref(hb), C.byref(bb))
        return k, codes, hdr, hb.value, bb.value


def fuzz_streams(rng, count):
    """Structured random inputs: small alphabets, periodic data with noise, copies at every distance class,
    long runs interrupted at random places -- the shapes that stress hash collisions, the skip heuristic, the
    32768-byte distance limit, the 258-byte match cap and the 15/16 literal rule."""
    out = []
    for k in range(count):
        n = int(rng.choice([130, 200, 1000, 5000, 20000, 65535, 66000, 70000])) + int(rng.integers(0, 200))
        kind = k % 7
        if kind == 0:
            a = rng.integers(0, int(rng.integers(2, 6)), n, dtype=np.uint8)
        elif kind == 1:
            period = int(rng.integers(1, 300))
            a = np.tile(rng.integers(0, 256, period, dtype=np.uint8), n // period + 1)[:n].copy()
            noise = rng.random(n) < rng.choice([0.0, 0.001, 0.01, 0.05])
            a[noise] = rng.integers(0, 256, int(noise.sum()), dtype=np.uint8)
        elif kind == 2:
            a = rng.integers(0, 256, n, dtype=np.uint8)
            for _ in range(int(rng.integers(1, 60))):
                ln = int(rng.integers(4, 600))
                src = int(rng.integers(0, max(1, n - ln)))
                dist = int(rng.choice([1, 2, 3, 7, 255, 256, 4095, 32767, 32768, 32769, 40000]))
                dst = src + dist
                if dst + ln <= n:
                    a[dst:dst + ln] = a[src:src + ln]
        elif kind == 3:
            a = np.repeat(rng.integers(0, 256, n // 50 + 1, dtype=np.uint8), rng.integers(1, 600, n // 50 + 1))[:n]
            a = np.resize(a, n)
        elif kind == 4:
            words = [bytes(rng.integers(97, 123, int(rng.integers(1, 12)), dtype=np.uint8)) for _ in range(int(rng.integers(2, 400)))]
            s = b" ".join(words[int(i)] for i in rng.integers(0, len(words), n // 4 + 1))
            a = np.frombuffer(s[:n].ljust(n, b"."), dtype=np.uint8)
        elif kind == 5:
            a = (np.arange(n) * int(rng.integers(1, 7)) % int(rng.integers(2, 257))).astype(np.uint8)
        else:
            base = rng.integers(0, 256, 64, dtype=np.uint8)
            a = base[rng.integers(0, 64, n)]
            a[:: int(rng.integers(2, 50))] = 0
        out.append(np.ascontiguousarray(a, dtype=np.uint8).tobytes())
    return out


