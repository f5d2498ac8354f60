"""The host library's own DEFLATE decoder (uvc_b200/csrc/host/inflate_fast.cpp: BGZF members are inflated by it)
against zlib, byte for byte: every kind of block (stored, fixed and dynamic Huffman codes), every compression level, long and short matches,
overlapping copies, incompressible data, empty input, the members of a real BAM; and damaged streams, which must be refused or decoded like zlib
decodes them, never crash."""
import ctypes as C
import os
import random
import struct
import zlib

import pytest

from uvc_b200 import capi


def _lib():
    lib = capi.load_host()
    lib.uvc_inflate_raw.restype = C.c_int64
    lib.uvc_inflate_raw.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    lib.uvc_inflate_member.restype = C.c_int64
    lib.uvc_inflate_member.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    return lib


def _inflate(lib, comp: bytes, cap: int, fn="uvc_inflate_raw"):
    src = C.create_string_buffer(comp + b"\0" * 16, len(comp) + 16)     # the decoder may read 16 bytes past the stream
    out = C.create_string_buffer(cap + 1)
    n = getattr(lib, fn)(src, len(comp), out, cap)
    return n, out.raw[:max(n, 0)]


def _raw(data: bytes, level: int, strategy=zlib.Z_DEFAULT_STRATEGY) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return c.compress(data) + c.flush()


def _samples():
    rnd = random.Random(20261018)
    yield b""
    yield b"a"
    yield b"abc" * 20000
    yield bytes(60000)                                                   # one long run: distance 1
    yield bytes(rnd.getrandbits(8) for _ in range(50000))                # incompressible: stored blocks at every level
    yield bytes(rnd.choice(b"ACGT") for _ in range(65000))               # 2 bits of entropy per byte: literal-heavy dynamic codes
    yield b"".join(bytes([rnd.choice(b"ACGTN")]) * rnd.randint(1, 40) for _ in range(4000))[:65000]
    yield b"".join(struct.pack("<iiH", rnd.randint(0, 1 << 28), rnd.randint(0, 300), rnd.randint(0, 65535)) for _ in range(6000))   # record-like
    words = [bytes(rnd.choice(b"abcdefghijklmnopqrstuvwxyz") for _ in range(rnd.randint(2, 12))) for _ in range(300)]
    yield b" ".join(rnd.choice(words) for _ in range(9000))[:65000]      # text: long matches at all distances


def test_matches_zlib_on_every_level_and_block_type():
    lib = _lib()
    n_cases = 0
    for data in _samples():
        for level in range(0, 10):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                comp = _raw(data, level, strategy)
                n, got = _inflate(lib, comp, 65536)
                assert n == len(data) and got == data, (len(data), level, strategy)
                n_cases += 1
    assert n_cases == 9 * 10 * 4


def test_output_capacity_is_respected():
    lib = _lib()
    data = b"abcdefgh" * 1000
    comp = _raw(data, 6)
    assert _inflate(lib, comp, len(data))[0] == len(data)
    assert _inflate(lib, comp, len(data) - 1)[0] == -1
    assert _inflate(lib, comp, 0)[0] == -1


def test_damaged_streams_are_refused_or_decoded_like_zlib():
    lib = _lib()
    rnd = random.Random(7)
    data = b"".join(bytes([rnd.choice(b"ACGT")]) * rnd.randint(1, 9) for _ in range(6000))
    for level in (1, 6):
        comp = _raw(data, level)
        for trial in range(400):
            bad = bytearray(comp)
            kind = trial % 3
            if kind == 0:
                bad[rnd.randrange(len(bad))] ^= 1 << rnd.randrange(8)
            elif kind == 1:
                del bad[rnd.randrange(1, len(bad)):]
            else:
                i = rnd.randrange(len(bad)); bad[i:i + 4] = bytes(rnd.getrandbits(8) for _ in range(4))
            bad = bytes(bad)
            try:
                d = zlib.decompressobj(-15)
                want = d.decompress(bad, 65536)
                ok = d.eof and len(want) <= 65536 and not d.unconsumed_tail
            except zlib.error:
                ok = False
            n, got = _inflate(lib, bad, 65536)
            if n >= 0:      # accepted: zlib must accept it with the same bytes (trailing garbage after the final block is not an error for either)
                assert ok and got == want
            n2, got2 = _inflate(lib, bad, 65536, "uvc_inflate_member")
            assert (n2 >= 0) == ok and (not ok or got2 == want)


def test_members_of_a_bam(synth_small):
    """Every BGZF member of a BAM written by the test generator (pysam-free writer, zlib level 6) and re-compressed at level 1 (what aligners write)."""
    lib = _lib()
    raw = open(synth_small["bam"], "rb").read()
    off, n_blocks = 0, 0
    while off < len(raw):
        xlen = struct.unpack_from("<H", raw, off + 10)[0]
        bsize = struct.unpack_from("<H", raw, off + 16)[0] + 1
        comp = raw[off + 12 + xlen:off + bsize - 8]
        isize = struct.unpack_from("<I", raw, off + bsize - 4)[0]
        want = zlib.decompress(comp, -15)
        assert len(want) == isize
        for c in (comp, _raw(want, 1)):
            n, got = _inflate(lib, c, 65536)
            assert n == isize and got == want
        off += bsize
        n_blocks += 1
    assert n_blocks > 10
