"""CPU suite, part 3: the C-ABI library loads without a GPU, exports every symbol
include/flate_b200.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

from helpers import ROOT


def _declared_symbols():
    h = open(os.path.join(ROOT, "include", "flate_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    names = set(re.findall(r"\b(fb200_[a-z0-9_]+)\s*\(", h))
    names.discard("fb200_sink_fn")
    return sorted(names)


def test_library_exports_every_declared_symbol():
    import moonbit_flate_b200 as fb

    names = _declared_symbols()
    assert len(names) >= 25
    lib = C.CDLL(fb.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/flate_b200.h but not exported"
    assert set(names) == set(fb.ABI), "python binding table and header disagree"


def test_pure_functions_without_gpu():
    import moonbit_flate_b200 as fb

    L = fb._lib
    assert L.fb200_version() == 1
    assert L.fb200_frame_header_bytes(10) == 16 + 40
    # bound: every stream fits (oracle sizes are checked against it in the GPU tests)
    assert L.fb200_deflate_stream_bound(0) >= 5
    assert L.fb200_deflate_stream_bound(65536) >= 65536 + 65536 // 8
    assert L.fb200_deflate_bound(1 << 20, 65536) >= 16 * L.fb200_deflate_stream_bound(65536)
    assert L.fb200_deflate_bound(100, 0) == 0


def test_no_cpu_fallback():
    """Without a usable sm_100 device fb200_create fails with FB200_ERR_CUDA and the Python layer raises;
    nothing routes to the oracle.  (On a GPU box this test just creates and destroys a context.)"""
    import moonbit_flate_b200 as fb

    h = C.c_void_p()
    rc = fb._lib.fb200_create(C.byref(h), -1)
    if rc == fb.OK:
        fb._lib.fb200_destroy(h)
        return
    assert rc == fb.ERR_CUDA and not h.value
    try:
        fb.Context()
        raise AssertionError("Context() must raise without a GPU")
    except fb.FlateError as e:
        assert "no CPU fallback" in str(e)


def test_product_never_references_the_oracle():
    """The oracle is test infrastructure: no product source may include, link or load it."""
    pkg = os.path.join(ROOT, "moonbit_flate_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".py", ".mbt", "Makefile")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "flate_oracle" not in txt and "oracle/" not in txt, os.path.join(dp, f)
    inc = open(os.path.join(ROOT, "include", "flate_b200.h")).read()
    assert "flate_oracle" not in inc


def test_table_reset_blocks_closed_form():
    """deflate-fast.mbt:129-132 / :366-374: encode clears the table when cur >= buffer_reset.  cur is 65535 before
    block 0 and grows by 65535 per full block; after a reset it restarts at 32769.  The parse uses a closed form for
    the reset blocks (common.cuh block_resets_table): compare it with the running sum over 200000 blocks."""
    import ctypes as C
    import moonbit_flate_b200 as fb
    f = fb._lib.fb200_debug_block_resets
    f.restype = C.c_int
    f.argtypes = [C.c_uint64]
    buffer_reset = 2147483647 - 2 * 65535
    cur = 65535
    resets = []
    for b in range(200000):
        if cur >= buffer_reset:
            resets.append(b)
            cur = 32768 + 1
        cur += 65535
    assert resets[:3] == [32766, 65532, 98298]
    got = [b for b in range(200000) if f(b)]
    assert got == resets
