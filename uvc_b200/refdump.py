"""Reader of the oracle harness dump format (oracle/harness_dump.cpp) and section dtypes shared with the C ABI dump."""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

from .capi import struct_dtype

NSYM = 14


def section_dtype(name: str) -> np.dtype:
    if name in ("prep",):
        return struct_dtype("uvcgpu_prep_set")
    if name == "thres":
        return struct_dtype("uvcgpu_thres_set")
    if name == "seginfo":
        return np.dtype((struct_dtype("uvcgpu_seginfo_set"), (NSYM,)))
    if name == "faminfo":
        return np.dtype((struct_dtype("uvcgpu_faminfo_set"), (NSYM,)))
    if name in ("rtr", "rtr_initial", "rtr_final"):
        return struct_dtype("uvcgpu_rtr")
    if name in ("baq", "baq2", "meta"):
        return np.dtype("<i8")
    if name in ("fragdepth0", "fragdepth1"):
        return np.dtype(("<i4", (NSYM, 3)))
    if name in ("famdepth0", "famdepth1"):
        return np.dtype(("<i4", (NSYM, 8)))
    if name == "duplex":
        return np.dtype(("<i4", (NSYM, 2)))
    if name == "vq":
        return np.dtype(("<i4", (NSYM, 27)))
    raise KeyError(name)


def read_dump(path: str) -> Dict[str, object]:
    """Returns {section: numpy array or str} of an oracle dump file."""
    out: Dict[str, object] = {}
    with open(path, "rb") as f:
        if f.read(8) != b"UVCDUMP1":
            raise IOError("not an oracle dump: " + path)
        while True:
            hdr = f.read(48)
            if len(hdr) < 48:
                break
            name = hdr[:32].split(b"\0", 1)[0].decode()
            esz, cnt = struct.unpack("<QQ", hdr[32:])
            raw = f.read(esz * cnt)
            if name in ("families", "indelmaps", "haplinks", "refstring"):
                out[name] = raw.decode()
            else:
                dt = section_dtype(name)
                if dt.itemsize != esz and name != "meta":
                    raise IOError("section %s: element size %d != expected %d" % (name, esz, dt.itemsize))
                out[name] = np.frombuffer(raw, dtype=dt)
    return out
