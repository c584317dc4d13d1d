"""moonbit_flate_b200 -- Python host binding of libflate_b200.so.

The product is the CUDA library behind the C ABI in ``include/flate_b200.h``;
this module is the thin ctypes layer the tests and ``bench.py`` drive it
through, plus ``Writer`` / ``Reader`` classes that mirror the reference's
public API (pkg.generated.mbti:9-45: ``Writer::new`` / ``write`` / ``close``,
``&Reader::new`` -> ``Decompressor.read`` / ``close``) on top of the
``fb200_writer_*`` / ``fb200_reader_*`` entry points.

There is no CPU fallback: importing works without a GPU (so the symbol table can
be checked), but creating a ``Context`` raises unless an sm_100 device is
usable, and a missing ``libflate_b200.so`` raises at import.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# FB200_LIB selects another build of the same library (kernel experiments); the default is the in-tree build
LIB_PATH = os.environ.get("FB200_LIB") or os.path.join(_HERE, "libflate_b200.so")

OK = 0
ERR_ARG, ERR_DST_TOO_SMALL, ERR_CUDA, ERR_CLOSED, ERR_NOMEM = -1, -2, -3, -4, -5
ST_EOF, ST_CORRUPT, ST_UNEXPECTED_EOF, ST_DST_TOO_SMALL, ST_INTERNAL, ST_EOF_AT_REFILL = 0, 1, 2, 3, 4, 5

# reference error values (inflate.mbt:19, :38-46, deflate.mbt:154)
IOEOF = "EOF"
WRITER_CLOSED_ERROR = "writer closed"
ERR_UNEXPECTED_EOF = "unexpected EOF"


def corrupt_input_error(off: int) -> str:
    return f"flate: corrupt input before offset {off}"


class FlateError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(there is no CPU fallback)"
    )
_lib = C.CDLL(LIB_PATH)

_u8p = C.c_void_p
_u64p = C.c_void_p
SINK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint8), C.c_uint64)


class Stats(C.Structure):
    _fields_ = [("nblocks", C.c_uint64), ("ntokens", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("inflate_fallbacks", C.c_uint64)]


# name -> (restype, argtypes); every symbol include/flate_b200.h declares
ABI = {
    "fb200_version": (C.c_int, []),
    "fb200_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "fb200_destroy": (None, [C.c_void_p]),
    "fb200_last_error": (C.c_char_p, [C.c_void_p]),
    "fb200_deflate_stream_bound": (C.c_uint64, [C.c_uint64]),
    "fb200_deflate_bound": (C.c_uint64, [C.c_uint64, C.c_uint64]),
    "fb200_deflate_segments": (C.c_int, [C.c_void_p, _u8p, C.c_uint64, C.c_uint64, _u8p, C.c_uint64, _u64p, _u64p]),
    "fb200_deflate_streams": (C.c_int, [C.c_void_p, _u8p, _u64p, C.c_uint64, _u8p, C.c_uint64, _u64p, _u64p]),
    "fb200_deflate_streams_dev": (
        C.c_int, [C.c_void_p, _u8p, _u64p, C.c_uint64, C.c_uint64, _u8p, C.c_uint64, _u64p, _u64p]),
    "fb200_deflate_segments_dev": (
        C.c_int, [C.c_void_p, _u8p, C.c_uint64, C.c_uint64, _u8p, C.c_uint64, _u64p, _u64p]),
    "fb200_inflate_batch": (
        C.c_int, [C.c_void_p, _u8p, _u64p, C.c_uint64, _u8p, _u64p, _u64p, C.c_void_p, C.c_void_p, _u64p]),
    "fb200_inflate_batch_dev": (
        C.c_int, [C.c_void_p, _u8p, _u64p, C.c_uint64, _u8p, _u64p, _u64p, C.c_void_p, C.c_void_p, _u64p]),
    "fb200_deflate_segments_async": (C.c_int, [C.c_void_p, _u8p, C.c_uint64, C.c_uint64, _u8p, C.c_uint64, _u64p, _u64p]),
    "fb200_deflate_streams_async": (C.c_int, [C.c_void_p, _u8p, _u64p, C.c_uint64, _u8p, C.c_uint64, _u64p, _u64p]),
    "fb200_inflate_batch_async": (
        C.c_int, [C.c_void_p, _u8p, _u64p, C.c_uint64, _u8p, _u64p, _u64p, C.c_void_p, C.c_void_p, _u64p]),
    "fb200_wait": (C.c_int, [C.c_void_p]),
    "fb200_frame_header_bytes": (C.c_uint64, [C.c_uint64]),
    "fb200_mg_frame_alloc": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.c_void_p]),
    "fb200_mg_frame_open": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "fb200_mg_frame_close": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "fb200_mg_put": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]),
    "fb200_mg_wait": (C.c_int, [C.c_void_p]),
    "fb200_mg_get": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, _u8p, C.c_uint64, _u64p,
                               C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "fb200_mg_get_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, _u8p, C.c_uint64, _u64p,
                               C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "fb200_writer_new": (C.c_void_p, [C.c_void_p, SINK_FN, C.c_void_p]),
    "fb200_writer_new_dict": (C.c_void_p, [C.c_void_p, SINK_FN, C.c_void_p, _u8p, C.c_uint64]),
    "fb200_writer_write": (C.c_int64, [C.c_void_p, _u8p, C.c_uint64]),
    "fb200_writer_close": (C.c_int, [C.c_void_p]),
    "fb200_writer_pending": (C.c_uint64, [C.c_void_p]),
    "fb200_writer_take": (C.c_uint64, [C.c_void_p, _u8p, C.c_uint64]),
    "fb200_writer_free": (None, [C.c_void_p]),
    "fb200_inflate_dict": (C.c_int, [C.c_void_p, _u8p, C.c_uint64, _u8p, C.c_uint64, _u8p, C.c_uint64, _u64p, C.c_void_p,
                                     C.c_void_p, _u64p]),
    "fb200_reader_new": (C.c_void_p, [C.c_void_p, _u8p, C.c_uint64]),
    "fb200_reader_new_dict": (C.c_void_p, [C.c_void_p, _u8p, C.c_uint64, _u8p, C.c_uint64]),
    "fb200_reader_reset": (C.c_int, [C.c_void_p, _u8p, C.c_uint64, _u8p, C.c_uint64]),
    "fb200_reader_read": (C.c_uint64, [C.c_void_p, _u8p, C.c_uint64, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "fb200_reader_consumed": (C.c_uint64, [C.c_void_p]),
    "fb200_reader_close": (C.c_int, [C.c_void_p]),
    "fb200_reader_free": (None, [C.c_void_p]),
    "fb200_last_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "fb200_cuda_stream": (C.c_void_p, [C.c_void_p]),
    "fb200_last_stage_ms": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fb200_debug_block_resets": (C.c_int, [C.c_uint64]),
    "fb200_last_blocks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]),
}
for _name, (_res, _args) in ABI.items():
    _f = getattr(_lib, _name)
    _f.restype = _res
    _f.argtypes = _args


def _as_u8(buf) -> np.ndarray:
    if isinstance(buf, np.ndarray):
        a = buf if buf.dtype == np.uint8 else buf.view(np.uint8)
        return np.ascontiguousarray(a).reshape(-1)
    return np.frombuffer(bytes(buf) if not isinstance(buf, (bytes, bytearray, memoryview)) else buf, dtype=np.uint8)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


class Context:
    """One per host thread and GPU (fb200_create)."""

    def __init__(self, device: int = -1):
        h = C.c_void_p()
        rc = _lib.fb200_create(C.byref(h), device)
        if rc != OK:
            raise FlateError(f"fb200_create failed ({rc}): no usable sm_100 CUDA device (no CPU fallback)")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _lib.fb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != OK:
            raise FlateError(f"{what} failed ({rc}): {_lib.fb200_last_error(self._h).decode()}")

    # ---------------- host-buffer batch API ----------------
    def deflate_segments(self, src, seg_size: int):
        """Each seg_size slice of src becomes an independent reference stream.
        Returns (compressed uint8 array, seg_off uint64[nseg+1])."""
        s = _as_u8(src)
        n = s.size
        nseg = (n + seg_size - 1) // seg_size
        cap = int(_lib.fb200_deflate_bound(n, seg_size))
        dst = np.empty(cap, np.uint8)
        off = np.zeros(nseg + 1, np.uint64)
        olen = C.c_uint64()
        rc = _lib.fb200_deflate_segments(self._h, _ptr(s), n, seg_size, _ptr(dst), cap, _ptr(off), C.addressof(olen))
        self._check(rc, "fb200_deflate_segments")
        return dst[: olen.value], off

    def deflate_streams(self, src, src_off: Sequence[int]):
        s = _as_u8(src)
        so = np.ascontiguousarray(np.asarray(src_off, dtype=np.uint64))
        ns = so.size - 1
        cap = int(sum(int(_lib.fb200_deflate_stream_bound(int(so[i + 1] - so[i]))) for i in range(ns))) + 16 \
            if ns < 4096 else int(2 * s.size + 656 * (ns + s.size // 65535 + 1) + 16)
        dst = np.empty(cap, np.uint8)
        off = np.zeros(ns + 1, np.uint64)
        olen = C.c_uint64()
        rc = _lib.fb200_deflate_streams(self._h, _ptr(s), _ptr(so), ns, _ptr(dst), cap, _ptr(off), C.addressof(olen))
        self._check(rc, "fb200_deflate_streams")
        return dst[: olen.value], off

    def deflate(self, data) -> bytes:
        """Writer::new + write(data) + close as one call."""
        s = _as_u8(data)
        out, _ = self.deflate_streams(s, [0, s.size])
        return out.tobytes()

    def inflate_batch(self, comp, comp_off, out_off):
        """Returns (out uint8[out_off[-1]], out_len, status, err_off, consumed)."""
        c = _as_u8(comp)
        co = np.ascontiguousarray(np.asarray(comp_off, dtype=np.uint64))
        oo = np.ascontiguousarray(np.asarray(out_off, dtype=np.uint64))
        ns = co.size - 1
        out = np.zeros(int(oo[-1]) if ns else 0, np.uint8)
        out_len = np.zeros(ns, np.uint64)
        status = np.zeros(ns, np.int32)
        err_off = np.zeros(ns, np.int64)
        consumed = np.zeros(ns, np.uint64)
        rc = _lib.fb200_inflate_batch(self._h, _ptr(c), _ptr(co), ns, _ptr(out), _ptr(oo), _ptr(out_len),
                                      _ptr(status), _ptr(err_off), _ptr(consumed))
        self._check(rc, "fb200_inflate_batch")
        return out, out_len, status, err_off, consumed

    def inflate(self, comp, cap: int):
        """One stream; returns (bytes, status, err_off, consumed)."""
        c = _as_u8(comp)
        out, ol, st, eo, cons = self.inflate_batch(c, [0, c.size], [0, cap])
        return out[: int(ol[0])].tobytes(), int(st[0]), int(eo[0]), int(cons[0])

    # ---------------- device-pointer batch API (bench; torch owns the memory) ----------------
    def deflate_segments_dev(self, d_src: int, n: int, seg_size: int, d_dst: int, dst_cap: int, d_seg_off: int) -> int:
        olen = C.c_uint64()
        rc = _lib.fb200_deflate_segments_dev(self._h, d_src, n, seg_size, d_dst, dst_cap, d_seg_off, C.addressof(olen))
        self._check(rc, "fb200_deflate_segments_dev")
        return olen.value

    def deflate_streams_dev(self, d_src: int, d_src_off: int, ns: int, n_total: int, d_dst: int, dst_cap: int,
                            d_dst_off: int) -> int:
        olen = C.c_uint64()
        rc = _lib.fb200_deflate_streams_dev(self._h, d_src, d_src_off, ns, n_total, d_dst, dst_cap, d_dst_off,
                                            C.addressof(olen))
        self._check(rc, "fb200_deflate_streams_dev")
        return olen.value

    def inflate_batch_dev(self, d_comp: int, d_comp_off: int, ns: int, d_out: int, d_out_off: int, d_out_len: int,
                          d_status: int, d_err_off: int, d_consumed: int = 0):
        rc = _lib.fb200_inflate_batch_dev(self._h, d_comp, d_comp_off, ns, d_out, d_out_off, d_out_len, d_status,
                                          d_err_off, d_consumed or None)
        self._check(rc, "fb200_inflate_batch_dev")

    # raw host-pointer variants (pinned torch tensors in bench.py's e2e leg)
    def deflate_segments_ptr(self, src: int, n: int, seg_size: int, dst: int, dst_cap: int, seg_off: int) -> int:
        olen = C.c_uint64()
        rc = _lib.fb200_deflate_segments(self._h, src, n, seg_size, dst, dst_cap, seg_off, C.addressof(olen))
        self._check(rc, "fb200_deflate_segments")
        return olen.value

    def inflate_batch_ptr(self, comp: int, comp_off: int, ns: int, out: int, out_off: int, out_len: int, status: int,
                          err_off: int):
        rc = _lib.fb200_inflate_batch(self._h, comp, comp_off, ns, out, out_off, out_len, status, err_off, None)
        self._check(rc, "fb200_inflate_batch")

    # asynchronous forms (fb200_*_async + fb200_wait): raw pointers, the caller keeps the buffers alive
    def deflate_segments_async_ptr(self, src: int, n: int, seg_size: int, dst: int, dst_cap: int, seg_off: int):
        self._async_len = C.c_uint64()
        rc = _lib.fb200_deflate_segments_async(self._h, src, n, seg_size, dst, dst_cap, seg_off,
                                               C.addressof(self._async_len))
        self._check(rc, "fb200_deflate_segments_async")

    def deflate_streams_async_ptr(self, src: int, src_off: int, ns: int, dst: int, dst_cap: int, dst_off: int):
        self._async_len = C.c_uint64()
        rc = _lib.fb200_deflate_streams_async(self._h, src, src_off, ns, dst, dst_cap, dst_off, C.addressof(self._async_len))
        self._check(rc, "fb200_deflate_streams_async")

    def inflate_batch_async_ptr(self, comp: int, comp_off: int, ns: int, out: int, out_off: int, out_len: int,
                                status: int, err_off: int):
        rc = _lib.fb200_inflate_batch_async(self._h, comp, comp_off, ns, out, out_off, out_len, status, err_off, None)
        self._check(rc, "fb200_inflate_batch_async")

    def wait(self) -> int:
        """Joins the asynchronous call in flight; returns the compressed size for a deflate call."""
        self._check(_lib.fb200_wait(self._h), "fb200_wait")
        v = getattr(self, "_async_len", None)
        return int(v.value) if v is not None else 0

    # ---------------- introspection ----------------
    def last_stats(self) -> Stats:
        s = Stats()
        self._check(_lib.fb200_last_stats(self._h, C.byref(s)), "fb200_last_stats")
        return s

    STAGES = ("setup", "parse", "histogram", "build", "layout", "pack", "inflate")

    def last_stage_ms(self) -> dict:
        ms = np.zeros(len(self.STAGES), np.float32)
        self._check(_lib.fb200_last_stage_ms(self._h, _ptr(ms)), "fb200_last_stage_ms")
        return {k: float(v) for k, v in zip(self.STAGES, ms)}

    def cuda_stream(self) -> int:
        return int(_lib.fb200_cuda_stream(self._h) or 0)

    # ---------------- multi-GPU frame (CUDA IPC + peer copies) ----------------
    IPC_HANDLE_BYTES = 64

    def mg_frame_alloc(self, nbytes: int):
        """-> (device pointer, 64-byte IPC handle) of a frame buffer on this context's GPU."""
        p = C.c_void_p()
        h = (C.c_uint8 * self.IPC_HANDLE_BYTES)()
        self._check(_lib.fb200_mg_frame_alloc(self._h, nbytes, C.byref(p), h), "fb200_mg_frame_alloc")
        return int(p.value), bytes(h)

    def mg_frame_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        h = (C.c_uint8 * self.IPC_HANDLE_BYTES).from_buffer_copy(handle)
        self._check(_lib.fb200_mg_frame_open(self._h, h, C.byref(p)), "fb200_mg_frame_open")
        return int(p.value)

    def mg_frame_close(self, d_frame: int, owner: bool):
        self._check(_lib.fb200_mg_frame_close(self._h, d_frame, 1 if owner else 0), "fb200_mg_frame_close")

    def mg_put(self, d_frame: int, offset: int, d_payload: int, n: int):
        self._check(_lib.fb200_mg_put(self._h, d_frame, offset, d_payload, n), "fb200_mg_put")

    def mg_wait(self):
        self._check(_lib.fb200_mg_wait(self._h), "fb200_mg_wait")

    def mg_get(self, d_frame: int, frame_bytes: int, first: int, count: int, d_comp: int, comp_cap: int, d_comp_off: int):
        """Fetch the streams of segments [first, first + count) from a frame -> (seg_size, nseg_total, bytes)."""
        seg, nseg, nb = C.c_uint32(), C.c_uint64(), C.c_uint64()
        self._check(_lib.fb200_mg_get(self._h, d_frame, frame_bytes, first, count, d_comp, comp_cap, d_comp_off,
                                      C.byref(seg), C.byref(nseg), C.byref(nb)), "fb200_mg_get")
        return int(seg.value), int(nseg.value), int(nb.value)

    def mg_get_async(self, d_frame: int, frame_bytes: int, first: int, count: int, d_comp: int, comp_cap: int, d_comp_off: int):
        """mg_get without the final wait: the payload is on its way when this returns, ``mg_wait`` blocks until it
        has arrived -> (seg_size, nseg_total, bytes)."""
        seg, nseg, nb = C.c_uint32(), C.c_uint64(), C.c_uint64()
        self._check(_lib.fb200_mg_get_async(self._h, d_frame, frame_bytes, first, count, d_comp, comp_cap, d_comp_off,
                                      C.byref(seg), C.byref(nseg), C.byref(nb)), "fb200_mg_get_async")
        return int(seg.value), int(nseg.value), int(nb.value)

    def last_blocks(self, nblocks: int, tok_cap: int):
        ntok = np.zeros(nblocks, np.uint32)
        kind = np.zeros(nblocks, np.uint8)
        bits = np.zeros(nblocks, np.uint32)
        toks = np.zeros(tok_cap, np.uint32)
        rc = _lib.fb200_last_blocks(self._h, _ptr(ntok), _ptr(kind), _ptr(bits), nblocks, _ptr(toks), tok_cap)
        self._check(rc, "fb200_last_blocks")
        return ntok, kind, bits, toks


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


class Writer:
    """Mirror of the reference ``Writer`` (writer.mbt:10-58).

    ``Writer(buf)`` ~ ``Writer::new(buf)`` where ``buf`` has a ``write(bytes)``
    method (the ``&@io.Writer``); ``write`` returns ``(n, err)`` and ``close``
    returns ``err`` with the reference's values (``None`` on success,
    ``WRITER_CLOSED_ERROR`` for a write after close, ``None`` for a second close).
    """

    def __init__(self, buf, ctx: Optional[Context] = None, _dict: Optional[bytes] = None):
        self._ctx = ctx or default_context()
        self._buf = buf

        def _sink(_user, data, n):
            try:
                self._buf.write(C.string_at(data, n))
                return 0
            except Exception:  # sticky sink error
                return 1

        self._cb = SINK_FN(_sink)
        if _dict is None:
            self._h = _lib.fb200_writer_new(self._ctx._h, self._cb, None)
        else:
            d = _as_u8(_dict)
            self._h = _lib.fb200_writer_new_dict(self._ctx._h, self._cb, None, _ptr(d), d.size)
        if not self._h:
            raise FlateError("fb200_writer_new failed")

    @classmethod
    def new(cls, buf, ctx: Optional[Context] = None) -> "Writer":
        return cls(buf, ctx)

    @classmethod
    def new_dict(cls, buf, dict_: bytes, ctx: Optional[Context] = None) -> "Writer":
        """Writer::new_dict (writer.mbt:25-31): the dictionary is compressed into
        the output as if it had been written first (deflate_test.mbt:12-35)."""
        return cls(buf, ctx, _dict=bytes(dict_))

    def write(self, data):
        d = _as_u8(data)
        rc = _lib.fb200_writer_write(self._h, _ptr(d), d.size)
        if rc == ERR_CLOSED:
            return 0, WRITER_CLOSED_ERROR
        if rc < 0:
            return 0, f"fb200 error {rc}"
        return int(rc), None

    def close(self):
        rc = _lib.fb200_writer_close(self._h)
        if rc != OK:
            return f"fb200 error {rc}: {_lib.fb200_last_error(self._ctx._h).decode()}"
        return None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.fb200_writer_free(self._h)
                self._h = None
        except Exception:
            pass


def _status_to_err(status: int, err_off: int):
    if status < 0:
        return None
    if status in (ST_EOF, ST_EOF_AT_REFILL):
        return IOEOF
    if status == ST_CORRUPT:
        return corrupt_input_error(err_off)
    if status == ST_UNEXPECTED_EOF:
        return ERR_UNEXPECTED_EOF
    return f"flate: internal error: status {status}"


class Reader:
    """Mirror of ``&Reader::new(buf)`` -> ``Decompressor`` (inflate.mbt:305-418).

    ``read(n)`` returns ``(bytes, err)`` like ``Decompressor.read``: at most one
    32 KiB window flush per call, ``err`` is ``None`` until the stream ends and
    then ``IOEOF`` or the reference's error text."""

    def __init__(self, comp, ctx: Optional[Context] = None, _dict: Optional[bytes] = None):
        self._ctx = ctx or default_context()
        if hasattr(comp, "getvalue"):
            comp = comp.getvalue()
        self._comp = _as_u8(comp).copy()
        if _dict is None:
            self._h = _lib.fb200_reader_new(self._ctx._h, _ptr(self._comp), self._comp.size)
        else:
            d = _as_u8(_dict)
            self._h = _lib.fb200_reader_new_dict(self._ctx._h, _ptr(self._comp), self._comp.size, _ptr(d), d.size)
        if not self._h:
            raise FlateError("fb200_reader_new failed")

    @classmethod
    def new(cls, comp, ctx: Optional[Context] = None) -> "Reader":
        return cls(comp, ctx)

    @classmethod
    def new_dict(cls, comp, dict_: bytes, ctx: Optional[Context] = None) -> "Reader":
        """&Reader::new_dict (inflate.mbt:310-317): decode as if the output started with dict_ (already read)."""
        return cls(comp, ctx, _dict=dict_)

    def reset(self, comp, dict_: bytes = b""):
        """Decompressor::reset(r, dict) (inflate.mbt:862-883): new input and dictionary, state dropped."""
        if hasattr(comp, "getvalue"):
            comp = comp.getvalue()
        self._comp = _as_u8(comp).copy()
        d = _as_u8(dict_)
        rc = _lib.fb200_reader_reset(self._h, _ptr(self._comp), self._comp.size, _ptr(d) if d.size else None, d.size)
        if rc != OK:
            raise FlateError(f"fb200_reader_reset failed ({rc})")

    def read(self, n: int):
        buf = np.empty(max(n, 1), np.uint8)
        st = C.c_int32(-1)
        eo = C.c_int64(0)
        k = _lib.fb200_reader_read(self._h, _ptr(buf), n, C.byref(st), C.byref(eo))
        return buf[: int(k)].tobytes(), _status_to_err(st.value, eo.value)

    def consumed(self) -> int:
        """Input bytes the decoder consumed: what follows them (a trailer, the next member) belongs to the caller."""
        return int(_lib.fb200_reader_consumed(self._h))

    def read_all(self):
        """@io.copy(got, r): returns (bytes, err) with err None for a clean ioeof."""
        out = bytearray()
        while True:
            b, err = self.read(1 << 16)
            out += b
            if err is not None:
                return bytes(out), (None if err == IOEOF else err)

    def close(self):
        rc = _lib.fb200_reader_close(self._h)
        return None if rc == OK else f"flate error status {rc}"

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.fb200_reader_free(self._h)
                self._h = None
        except Exception:
            pass
