#!/usr/bin/env python
"""Regenerates tests/golden/reference_kats.json and tests/golden/oracle_digests.json.

reference_kats.json -- the known-answer data the reference's OWN tests hold for
the deflate / inflate path, lifted verbatim (data only) from the test files
under /root/reference so they can travel to the GPU box:
  * deflate_test.mbt:12-35          "writer dict": 28 input bytes -> 38 bytes
  * deflate-fast_test.mbt:14-100    TestBestSpeed write-size matrix (round trip)
  * dict-decoder_wbtest.mbt:9-291   DictDecoder scenario (poem + (dist,len) list)
  * token.mbt:95-99, bits.mbt:24-27, huffman-code.mbt:289-292 unit values
The reference holds no compressed-byte goldens.

oracle_digests.json -- sha256 of the ORACLE's compressed output for seeded
corpus units.  These are *derived* values (not reference-pinned): they freeze
the oracle so an accidental change to it is caught; byte parity of the oracle
itself rests on source transcription + the KATs above + zlib cross-inflation.

Run from the repo root in the build container (needs /root/reference):
    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "tests"))


def reference_kats():
    wb = open(os.path.join(REF, "dict-decoder_wbtest.mbt")).read()
    poem_lines = re.findall(r"^\s*#\|(.*)$", wb, flags=re.M)
    poem = "\n".join(poem_lines)
    refs_src = wb[wb.index("let poem_refs = ["): wb.index("let got = @buffer.new()")]
    poem_refs = [[int(a), int(b)] for a, b in re.findall(r"\((\d+),\s*(\d+)\)", refs_src)]
    assert sum(l for _, l in poem_refs) == len(poem), (sum(l for _, l in poem_refs), len(poem))
    bs = open(os.path.join(REF, "deflate-fast_test.mbt")).read()
    tc_src = bs[bs.index("let test_cases = ["): bs.index("let mut i = 0")]
    test_cases = [[int(x) for x in m.split(",") if x.strip()] for m in re.findall(r"\[([\d,\s]+)\],", tc_src)]
    first_n = [int(x) for x in re.search(r"for first_n in \[([\d,\s]+)\]", bs).group(1).split(",")]
    assert len(test_cases) == 16 and len(first_n) == 6
    return {
        "source": "data lifted from the reference's test files (see make_golden.py docstring)",
        "writer_dict": {"dict": "hello world", "text": "hello again world", "compressed_len": 38,
                        "cite": "deflate_test.mbt:12-35"},
        "best_speed": {"abc_len": 128, "total": 131072, "test_cases": test_cases, "first_n": first_n,
                       "cite": "deflate-fast_test.mbt:14-100"},
        "dict_decoder": {"window": 2048, "abc": "ABC\n", "fox": "The quick brown fox jumped over the lazy dog!\n",
                         "poem": poem, "poem_refs": poem_refs, "cite": "dict-decoder_wbtest.mbt:9-291"},
        "units": {"token_offset": [2143289471, 127], "reverse16": [32768, 1], "reverse_bits": [64, 7, 1],
                  "cite": "token.mbt:95-99, bits.mbt:24-27, huffman-code.mbt:289-292"},
    }


def oracle_digests():
    from helpers import Corpus, Oracle

    orc, corpus = Oracle(), Corpus()
    out = {"source": "derived from oracle/flate_oracle.c (not reference-pinned); corpus = tools/corpus.c",
           "units": []}
    for klass in range(6):
        for n in (0, 1, 16, 17, 127, 128, 4096, 65535, 65536, 65663, 200000):
            d = corpus.unit(n, seed=7, index=n % 97, klass=klass)
            c = orc.deflate(d)
            out["units"].append({"klass": klass, "n": n, "seed": 7, "index": n % 97, "clen": len(c),
                                 "sha256": hashlib.sha256(c).hexdigest()})
    return out


if __name__ == "__main__":
    json.dump(reference_kats(), open(os.path.join(HERE, "reference_kats.json"), "w"), indent=1)
    json.dump(oracle_digests(), open(os.path.join(HERE, "oracle_digests.json"), "w"), indent=1)
    print("wrote reference_kats.json, oracle_digests.json")
