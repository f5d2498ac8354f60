"""ctypes bindings of the two in-tree native libraries.

* ``libuvcgpu.so``  - the C ABI of include/uvcgpu.h (CUDA kernels; the product path).
* ``libuvchost.so`` - host substrate (BAM/BAI/FASTA decoding into the SoA record layout).

The product never falls back to a CPU implementation: ``load_gpu()`` raises if the CUDA library is missing, and
``uvcgpu_create`` fails with UVCGPU_ENODEVICE when there is no device.  ``load_gpu(emulate=True)`` loads the test-only
emulation build (tests/emu/libuvcgpu_emu.so) and is used exclusively by the ``-m "not gpu"`` unit tests.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "uvcgpu.h")
LIBDIR = os.path.join(ROOT, "uvc_b200", "lib")

_CT = {"int32_t": C.c_int32, "uint32_t": C.c_uint32, "int64_t": C.c_int64, "uint64_t": C.c_uint64, "double": C.c_double,
       "uint16_t": C.c_uint16, "uint8_t": C.c_uint8}
_NP = {"int32_t": "<i4", "uint32_t": "<u4", "int64_t": "<i8", "uint64_t": "<u8", "double": "<f8"}


def _struct_fields(name: str) -> List[Tuple[str, str, int]]:
    """Parses ``typedef struct <name> { ... } <name>;`` from the header: [(ctype name, field, array length or 0)]."""
    text = open(HEADER).read()
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, re.S)
    if not m:
        raise RuntimeError("struct %s not found in %s" % (name, HEADER))
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    out = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        parts = stmt.split(None, 1)
        ctype, names = parts[0], parts[1]
        for nm in names.split(","):
            nm = nm.strip()
            am = re.match(r"(\w+)\[(\d+)\]", nm)
            if am:
                out.append((ctype, am.group(1), int(am.group(2))))
            else:
                out.append((ctype, nm, 0))
    return out


def _make_ctypes_struct(name: str):
    fields = []
    for ctype, fname, alen in _struct_fields(name):
        t = _CT[ctype]
        fields.append((fname, t * alen if alen else t))
    return type(name, (C.Structure,), {"_fields_": fields})


def struct_dtype(name: str) -> np.dtype:
    """numpy dtype with C alignment mirroring a per-position record of the header."""
    spec = []
    for ctype, fname, alen in _struct_fields(name):
        spec.append((fname, _NP[ctype], (alen,)) if alen else (fname, _NP[ctype]))
    return np.dtype(spec, align=True)


Params = _make_ctypes_struct("uvcgpu_params")
Tile = _make_ctypes_struct("uvcgpu_tile")
BatchStats = _make_ctypes_struct("uvcgpu_batch_stats")


class ReadsSoA(C.Structure):
    _fields_ = [("n_reads", C.c_int64)] + [(n, C.c_void_p) for n in (
        "pos", "mpos", "isize", "mtid", "l_qseq", "n_cigar", "nm", "flag", "mapq",
        "seq_off", "qual_off", "cigar_off", "qname_off", "seq", "qual", "cigar", "qname")]


SECTIONS = {"meta": 0, "rtr": 1, "baq": 2, "baq2": 3, "prep": 4, "thres": 5, "seginfo": 6, "faminfo": 7, "fragdepth0": 8,
            "fragdepth1": 9, "famdepth0": 10, "famdepth1": 11, "duplex": 12, "vq": 13, "families": 14, "rtr_initial": 15, "indelmaps": 16, "haplinks": 17, "vcf": 18}

_gpu_libs: Dict[bool, C.CDLL] = {}
_host_lib: Optional[C.CDLL] = None


def load_gpu(emulate: bool = False) -> C.CDLL:
    if emulate in _gpu_libs:
        return _gpu_libs[emulate]
    path = (os.path.join(ROOT, "tests", "emu", "libuvcgpu_emu.so") if emulate else os.path.join(LIBDIR, "libuvcgpu.so"))
    if not os.path.exists(path):
        raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    lib.uvcgpu_params_default.argtypes = [C.POINTER(Params)]
    lib.uvcgpu_params_default.restype = None
    lib.uvcgpu_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Params)]
    lib.uvcgpu_destroy.argtypes = [C.c_void_p]
    lib.uvcgpu_destroy.restype = None
    lib.uvcgpu_last_error.argtypes = [C.c_void_p]
    lib.uvcgpu_last_error.restype = C.c_char_p
    lib.uvcgpu_set_contig.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_int64]
    lib.uvcgpu_set_contig_name.argtypes = [C.c_void_p, C.c_int32, C.c_char_p]
    lib.uvcgpu_score.argtypes = [C.c_void_p, C.c_int64, C.POINTER(BatchStats)]
    lib.uvcgpu_tile_vcf.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.uvcgpu_batch_vcf.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.uvcgpu_submit.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Tile), C.POINTER(ReadsSoA), C.POINTER(C.c_int64)]
    lib.uvcgpu_collect.argtypes = [C.c_void_p, C.c_int64, C.POINTER(BatchStats)]
    lib.uvcgpu_release.argtypes = [C.c_void_p, C.c_int64]
    lib.uvcgpu_dump_counters.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.uvcgpu_device_count.restype = C.c_int
    lib.uvcgpu_sizeof_params.restype = C.c_size_t
    if lib.uvcgpu_sizeof_params() != C.sizeof(Params):
        raise RuntimeError("uvcgpu_params mirror mismatch: %d vs %d" % (lib.uvcgpu_sizeof_params(), C.sizeof(Params)))
    _gpu_libs[emulate] = lib
    return lib


def load_host() -> C.CDLL:
    global _host_lib
    if _host_lib is not None:
        return _host_lib
    path = os.path.join(LIBDIR, "libuvchost.so")
    if not os.path.exists(path):
        raise RuntimeError("%s is missing: run __graft_entry__.build()" % path)
    lib = C.CDLL(path)
    lib.uvchost_bam_open.argtypes = [C.c_char_p]
    lib.uvchost_bam_open.restype = C.c_void_p
    lib.uvchost_bam_close.argtypes = [C.c_void_p]
    lib.uvchost_bam_close.restype = None
    lib.uvchost_bam_n_targets.argtypes = [C.c_void_p]
    lib.uvchost_bam_n_targets.restype = C.c_int32
    lib.uvchost_bam_target_name.argtypes = [C.c_void_p, C.c_int32]
    lib.uvchost_bam_target_name.restype = C.c_char_p
    lib.uvchost_bam_target_len.argtypes = [C.c_void_p, C.c_int32]
    lib.uvchost_bam_target_len.restype = C.c_int64
    lib.uvchost_readbuf_new.restype = C.c_void_p
    lib.uvchost_readbuf_free.argtypes = [C.c_void_p]
    lib.uvchost_readbuf_free.restype = None
    lib.uvchost_readbuf_clear.argtypes = [C.c_void_p]
    lib.uvchost_readbuf_clear.restype = None
    lib.uvchost_readbuf_size.argtypes = [C.c_void_p]
    lib.uvchost_readbuf_size.restype = C.c_int64
    lib.uvchost_readbuf_view.argtypes = [C.c_void_p, C.POINTER(ReadsSoA)]
    lib.uvchost_readbuf_view.restype = None
    lib.uvchost_bam_fetch.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p]
    lib.uvchost_bam_fetch.restype = C.c_int64
    lib.uvchost_fasta_open.argtypes = [C.c_char_p]
    lib.uvchost_fasta_open.restype = C.c_void_p
    lib.uvchost_fasta_close.argtypes = [C.c_void_p]
    lib.uvchost_fasta_close.restype = None
    lib.uvchost_fasta_fetch_contig.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]
    lib.uvchost_fasta_fetch_contig.restype = C.c_void_p
    _host_lib = lib
    return lib


class UvcGpuError(RuntimeError):
    pass


class Context:
    """One uvcgpu context (one CUDA device, one stream)."""

    def __init__(self, device: int = 0, emulate: bool = False, **param_overrides):
        self.lib = load_gpu(emulate)
        self.params = Params()
        self.lib.uvcgpu_params_default(C.byref(self.params))
        for k, v in param_overrides.items():
            setattr(self.params, k, v)
        self.handle = C.c_void_p()
        rc = self.lib.uvcgpu_create(C.byref(self.handle), device, C.byref(self.params))
        if rc != 0:
            raise UvcGpuError("uvcgpu_create failed with code %d (no CUDA device? there is no CPU fallback)" % rc)
        self._keep = []

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise UvcGpuError("%s failed (%d): %s" % (what, rc, self.lib.uvcgpu_last_error(self.handle).decode()))

    def set_contig(self, tid: int, bases: Optional[bytes]) -> None:
        self._check(self.lib.uvcgpu_set_contig(self.handle, tid, bases, len(bases) if bases is not None else 0), "uvcgpu_set_contig")

    def set_contig_name(self, tid: int, name: str) -> None:
        self._check(self.lib.uvcgpu_set_contig_name(self.handle, tid, name.encode()), "uvcgpu_set_contig_name")

    def score(self, ticket: int) -> BatchStats:
        st = BatchStats()
        self._check(self.lib.uvcgpu_score(self.handle, ticket, C.byref(st)), "uvcgpu_score")
        return st

    def tile_vcf(self, ticket: int, tile_index: int) -> bytes:
        need = C.c_size_t()
        self._check(self.lib.uvcgpu_tile_vcf(self.handle, ticket, tile_index, None, 0, C.byref(need)), "uvcgpu_tile_vcf")
        buf = C.create_string_buffer(max(1, need.value))
        self._check(self.lib.uvcgpu_tile_vcf(self.handle, ticket, tile_index, buf, need.value, C.byref(need)), "uvcgpu_tile_vcf")
        return buf.raw[:need.value]

    def batch_vcf(self, ticket: int) -> bytearray:
        """VCF bodies of all tiles of the batch in tile order (one copy into a bytearray)."""
        need = C.c_size_t()
        self._check(self.lib.uvcgpu_batch_vcf(self.handle, ticket, None, 0, C.byref(need)), "uvcgpu_batch_vcf")
        out = bytearray(max(1, need.value))
        buf = (C.c_char * len(out)).from_buffer(out)
        self._check(self.lib.uvcgpu_batch_vcf(self.handle, ticket, buf, need.value, C.byref(need)), "uvcgpu_batch_vcf")
        del buf
        if need.value < len(out):
            del out[need.value:]
        return out

    def submit(self, tiles: Sequence[Tile], reads: ReadsSoA) -> int:
        arr = (Tile * len(tiles))(*tiles)
        ticket = C.c_int64()
        self._keep.append((arr, reads))
        self._check(self.lib.uvcgpu_submit(self.handle, len(tiles), arr, C.byref(reads), C.byref(ticket)), "uvcgpu_submit")
        return ticket.value

    def collect(self, ticket: int) -> BatchStats:
        st = BatchStats()
        self._check(self.lib.uvcgpu_collect(self.handle, ticket, C.byref(st)), "uvcgpu_collect")
        return st

    def release(self, ticket: int) -> None:
        self._check(self.lib.uvcgpu_release(self.handle, ticket), "uvcgpu_release")

    def dump(self, ticket: int, tile_index: int, section: str) -> bytes:
        need = C.c_size_t()
        sec = SECTIONS[section]
        self._check(self.lib.uvcgpu_dump_counters(self.handle, ticket, tile_index, sec, None, 0, C.byref(need)), "uvcgpu_dump_counters")
        buf = C.create_string_buffer(max(1, need.value))
        self._check(self.lib.uvcgpu_dump_counters(self.handle, ticket, tile_index, sec, buf, need.value, C.byref(need)), "uvcgpu_dump_counters")
        return buf.raw[:need.value]

    def close(self) -> None:
        if self.handle:
            self.lib.uvcgpu_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BamFile:
    def __init__(self, path: str):
        self.lib = load_host()
        self.handle = self.lib.uvchost_bam_open(path.encode())
        if not self.handle:
            raise IOError("cannot open BAM (or its .bai): " + path)
        self.targets = [(self.lib.uvchost_bam_target_name(self.handle, i).decode(), self.lib.uvchost_bam_target_len(self.handle, i))
                        for i in range(self.lib.uvchost_bam_n_targets(self.handle))]

    def fetch_into(self, readbuf: "ReadBuf", tid: int, beg: int, end: int) -> int:
        n = self.lib.uvchost_bam_fetch(self.handle, tid, beg, end, readbuf.handle)
        if n < 0:
            raise IOError("BAM fetch failed")
        return n

    def close(self):
        if self.handle:
            self.lib.uvchost_bam_close(self.handle)
            self.handle = None


class ReadBuf:
    def __init__(self):
        self.lib = load_host()
        self.handle = self.lib.uvchost_readbuf_new()

    def __len__(self):
        return self.lib.uvchost_readbuf_size(self.handle)

    def view(self) -> ReadsSoA:
        v = ReadsSoA()
        self.lib.uvchost_readbuf_view(self.handle, C.byref(v))
        return v

    def clear(self):
        self.lib.uvchost_readbuf_clear(self.handle)

    def close(self):
        if self.handle:
            self.lib.uvchost_readbuf_free(self.handle)
            self.handle = None


def read_fasta_contig(path: str, name: str) -> bytes:
    lib = load_host()
    f = lib.uvchost_fasta_open(path.encode())
    if not f:
        raise IOError("cannot open FASTA (or its .fai): " + path)
    n = C.c_int64()
    p = lib.uvchost_fasta_fetch_contig(f, name.encode(), C.byref(n))
    lib.uvchost_fasta_close(f)
    if not p:
        raise KeyError(name)
    data = C.string_at(p, n.value)
    C.CDLL(None).free(C.c_void_p(p))
    return data


def make_tile(tid: int, beg: int, end: int, region_flag: int, contig_len: int, read_begin: int, read_end: int,
              prev: Tuple[int, int, int] = (-1, 0, 0)) -> Tile:
    t = Tile()
    t.tid, t.beg_pos, t.end_pos, t.region_flag = tid, beg, end, region_flag
    t.prev_tid, t.prev_beg_pos, t.prev_end_pos = prev
    t.contig_len = contig_len
    t.read_begin, t.read_end = read_begin, read_end
    return t
