"""Shared helpers of the parity tests: run one tile through the C ABI (CUDA or test-only emulation) and through the
oracle harness (oracle/_ref/uvc_ref_dump, the unmodified reference compiled against the htslib shim), then compare
every per-position section field by field."""
from __future__ import annotations

import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from uvc_b200 import capi, refdump

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "uvc_ref_dump")
REF_UVC1 = os.path.join(ROOT, "oracle", "_ref", "uvc1")


def have_oracle() -> bool:
    return os.path.exists(REF_DUMP)


def run_oracle_dump(bam: str, fasta: str, tid: int, beg: int, end: int, flag: int, out: str,
                    prev: Tuple[int, int, int] = (-1, 0, 0), extra: Sequence[str] = ()) -> Dict[str, object]:
    cmd = [REF_DUMP, bam, fasta, str(tid), str(beg), str(end), str(flag), out, str(prev[0]), str(prev[1]), str(prev[2])] + list(extra)
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return refdump.read_dump(out)


def run_tiles(bam: str, fasta: str, tiles: Sequence[Tuple[int, int, int, int]], emulate: bool, sections: Sequence[str],
              **params) -> Tuple[List[Dict[str, object]], capi.BatchStats]:
    """tiles: (tid, beg, end, region_flag) in order; prev of tile k is tile k-1 (as main.cpp:1513-1515)."""
    bf = capi.BamFile(bam)
    rb = capi.ReadBuf()
    # counter sections are compared with the reference's arrays over the whole extended range of a tile: the library fills them there only when
    # asked to (all_positions); the VCF text is checked on the product's default path
    if any(sec != "vcf" for sec in sections):
        params.setdefault("all_positions", 1)
    ctx = capi.Context(0, emulate=emulate, **params)
    tids = sorted(set(t[0] for t in tiles))
    for tid in tids:
        name, _ = bf.targets[tid]
        ctx.set_contig(tid, capi.read_fasta_contig(fasta, name))
        ctx.set_contig_name(tid, name)
    ctiles = []
    prev = (-1, 0, 0)
    for (tid, beg, end, flag) in tiles:
        r0 = len(rb)
        bf.fetch_into(rb, tid, max(0, beg - 2000), end + 2000)
        ctiles.append(capi.make_tile(tid, beg, end, flag, bf.targets[tid][1], r0, len(rb), prev))
        prev = (tid, beg, end)
    view = rb.view()
    ticket = ctx.submit(ctiles, view)
    stats = ctx.collect(ticket)
    out = []
    for ti in range(len(tiles)):
        d: Dict[str, object] = {}
        for sec in sections:
            raw = ctx.dump(ticket, ti, sec)
            if sec in ("families", "indelmaps", "haplinks", "vcf"):
                d[sec] = raw.decode()
            else:
                d[sec] = np.frombuffer(raw, dtype=refdump.section_dtype(sec)).copy()
        out.append(d)
    ctx.release(ticket)
    ctx.close()
    rb.close()
    bf.close()
    return out, stats


def diff_section(name: str, ours: np.ndarray, ref: np.ndarray, ext_beg: int = 0, max_report: int = 8) -> List[str]:
    """Field-wise comparison; returns human-readable mismatch lines (empty = bit-exact)."""
    msgs: List[str] = []
    if ours.shape != ref.shape:
        return ["%s: shape %s vs reference %s" % (name, ours.shape, ref.shape)]
    if ours.dtype.names is None:
        bad = np.argwhere(ours != ref)
        for idx in bad[:max_report]:
            idx = tuple(idx)
            msgs.append("%s[pos %d%s]: ours %d ref %d" % (name, ext_beg + idx[0], list(idx[1:]) if len(idx) > 1 else "", ours[idx], ref[idx]))
        if len(bad) > max_report:
            msgs.append("%s: ... %d mismatching elements in total" % (name, len(bad)))
        return msgs
    for f in ours.dtype.names:
        a, b = ours[f], ref[f]
        bad = np.argwhere(a != b)
        for idx in bad[:max_report]:
            idx = tuple(idx)
            msgs.append("%s.%s[pos %d%s]: ours %d ref %d" % (name, f, ext_beg + idx[0], list(idx[1:]) if len(idx) > 1 else "", a[idx], b[idx]))
        if len(bad) > max_report:
            msgs.append("%s.%s: ... %d mismatching elements in total" % (name, f, len(bad)))
    return msgs
