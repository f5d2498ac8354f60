"""Seeded synthetic inputs for the pileup-and-score path.

Writes a reference FASTA (+.fai) and a coordinate-sorted BAM (+.bai) of simulated
2x150 paired-end reads carrying spiked-in SNVs/indels at known allele fractions,
optionally with duplex UMIs in the read name (``QNAME#AGTA+TGGT``), which is the
input format the reference ``uvc1`` consumes (reference CmdLineArgs.cpp:1015-1022:
``<bam>.bai`` and ``<ref>.fai`` must sit next to the inputs).  There is no network
and the reference ships no fixtures, so every parity test and
every benchmark in this repository starts from this generator.

The bulk of the reads (no indel, no clip) is produced with vectorised numpy code;
reads overlapping a carried indel, reads with a sequencing-error indel and reads
with a soft clip go through a per-read Python path.  Everything is a pure function
of ``SynthConfig`` (including the seed).
"""
from __future__ import annotations

import dataclasses
import os
import struct
import zlib
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
# nt16 code of A,C,G,T (BAM 4-bit encoding "=ACMGRSVTWYHKDBN")
NT16 = np.array([1, 2, 4, 8], dtype=np.uint8)
BGZF_BLOCK = 0xFF00
READ_LEN = 150


@dataclasses.dataclass
class Variant:
    contig: int
    pos: int          # 0-based position of the first affected reference base (SNV) or of the anchor base (indel)
    kind: str         # "snv" | "ins" | "del"
    ref: str
    alt: str
    vaf: float


@dataclasses.dataclass
class SynthConfig:
    name: str = "c1"
    seed: int = 1001
    contigs: Sequence[Tuple[str, int]] = (("chrS1", 1_000_000),)
    depth: float = 100.0                # raw read depth (reads * 150 / covered bases)
    n_snv: int = 200
    n_indel: int = 60
    vafs: Sequence[float] = (0.05, 0.10, 0.25, 0.50)
    max_indel_len: int = 30
    # targeted panels: list of (contig, beg, end); None = whole contigs
    targets: Optional[List[Tuple[int, int, int]]] = None
    amplicon_frac: float = 0.0          # fraction of targets whose fragments share both ends (amplicon-like)
    # UMI / duplex
    umi: bool = False
    umi_len: int = 4
    family_mean: float = 8.0
    duplex_frac: float = 0.6
    swapped_umi_frac: float = 0.5       # bottom-strand families tagged B+A instead of A+B (debarcode -D style)
    pcr_err_rate: float = 1e-3
    # error model
    sub_err: float = 5e-4
    indel_err: float = 1e-5
    clip_frac: float = 0.01
    lowmapq_frac: float = 0.02
    str_every: int = 2000
    insert_mean: float = 350.0
    insert_sd: float = 50.0
    insert_min: int = 160
    insert_max: int = 1000


def _rng(cfg: SynthConfig, salt: int) -> np.random.Generator:
    return np.random.default_rng([cfg.seed, salt])


# --------------------------------------------------------------------------- reference

def make_reference(cfg: SynthConfig) -> List[np.ndarray]:
    """Uniform-random ACGT with planted homopolymers (6-20) and di/tri-nucleotide tracks every ~str_every bp."""
    out = []
    for ci, (_, length) in enumerate(cfg.contigs):
        rng = _rng(cfg, 100 + ci)
        seq = BASES[rng.integers(0, 4, size=length)]
        pos = int(rng.integers(200, cfg.str_every))
        while pos + 80 < length:
            kind = int(rng.integers(0, 3))
            if kind == 0:
                n = int(rng.integers(6, 21))
                seq[pos:pos + n] = BASES[rng.integers(0, 4)]
            else:
                ulen = kind + 1
                unit = BASES[rng.integers(0, 4, size=ulen)]
                if np.all(unit == unit[0]):
                    unit[-1] = BASES[(int(np.where(BASES == unit[0])[0][0]) + 1) % 4]
                n = int(rng.integers(4, 13))
                seq[pos:pos + n * ulen] = np.tile(unit, n)
            pos += int(rng.integers(cfg.str_every // 2, cfg.str_every * 3 // 2))
        out.append(seq)
    return out


def write_fasta(path: str, cfg: SynthConfig, ref: List[np.ndarray], width: int = 60) -> None:
    with open(path, "wb") as fa, open(path + ".fai", "w") as fai:
        off = 0
        for (name, length), seq in zip(cfg.contigs, ref):
            hdr = (">%s\n" % name).encode()
            fa.write(hdr)
            off += len(hdr)
            fai.write("%s\t%d\t%d\t%d\t%d\n" % (name, length, off, width, width + 1))
            nfull = length // width
            body = seq[:nfull * width].reshape(nfull, width)
            lines = np.concatenate([body, np.full((nfull, 1), 10, dtype=np.uint8)], axis=1).tobytes()
            fa.write(lines)
            off += len(lines)
            rest = seq[nfull * width:]
            if len(rest):
                fa.write(rest.tobytes() + b"\n")
                off += len(rest) + 1


# --------------------------------------------------------------------------- variants

def make_variants(cfg: SynthConfig, ref: List[np.ndarray]) -> List[Variant]:
    rng = _rng(cfg, 200)
    variants: List[Variant] = []
    regions = cfg.targets if cfg.targets is not None else [(ci, 0, l) for ci, (_, l) in enumerate(cfg.contigs)]
    weights = np.array([max(1, e - b) for (_, b, e) in regions], dtype=np.float64)
    weights /= weights.sum()
    taken: Dict[int, List[int]] = {}
    n_total = cfg.n_snv + cfg.n_indel
    kinds = ["snv"] * cfg.n_snv + ["indel"] * cfg.n_indel
    tries = 0
    while len(variants) < n_total and tries < n_total * 50:
        tries += 1
        kind = kinds[len(variants)]
        ci, b, e = regions[int(rng.choice(len(regions), p=weights))]
        lo, hi = max(b + 5, 300), min(e - 5, cfg.contigs[ci][1] - 300)
        if hi <= lo:
            continue
        pos = int(rng.integers(lo, hi))
        if any(abs(pos - p) < 120 for p in taken.get(ci, [])):
            continue
        vaf = float(cfg.vafs[int(rng.integers(0, len(cfg.vafs)))])
        seq = ref[ci]
        if kind == "snv":
            r = chr(seq[pos])
            alt = "ACGT".replace(r, "")[int(rng.integers(0, 3))]
            variants.append(Variant(ci, pos, "snv", r, alt, vaf))
        else:
            ilen = int(min(cfg.max_indel_len, 1 + rng.geometric(0.25)))
            if rng.random() < 0.5:
                ins = "".join("ACGT"[i] for i in rng.integers(0, 4, size=ilen))
                variants.append(Variant(ci, pos, "ins", chr(seq[pos]), chr(seq[pos]) + ins, vaf))
            else:
                variants.append(Variant(ci, pos, "del", seq[pos:pos + 1 + ilen].tobytes().decode(), chr(seq[pos]), vaf))
        taken.setdefault(ci, []).append(pos)
    variants.sort(key=lambda v: (v.contig, v.pos))
    return variants


# --------------------------------------------------------------------------- fragments

def _sample_inserts(cfg: SynthConfig, rng: np.random.Generator, n: int) -> np.ndarray:
    ins = np.rint(rng.normal(cfg.insert_mean, cfg.insert_sd, size=n)).astype(np.int64)
    return np.clip(ins, cfg.insert_min, cfg.insert_max)


def make_fragments(cfg: SynthConfig) -> Dict[str, np.ndarray]:
    """Returns per-fragment arrays: contig, start, end (exclusive), top (1 = R1 forward), molecule id, umi codes."""
    rng = _rng(cfg, 300)
    regions = cfg.targets if cfg.targets is not None else [(ci, 0, l) for ci, (_, l) in enumerate(cfg.contigs)]
    contig_l, start_l, end_l = [], [], []
    for ri, (ci, b, e) in enumerate(regions):
        clen = cfg.contigs[ci][1]
        n_reads = cfg.depth * (e - b) / READ_LEN
        n_mol = int(round(n_reads / 2.0 / (cfg.family_mean * (1.0 + cfg.duplex_frac) if cfg.umi else 1.0)))
        n_mol = max(n_mol, 1)
        ins = _sample_inserts(cfg, rng, n_mol)
        is_amplicon = (cfg.targets is not None and rng.random() < cfg.amplicon_frac)
        if is_amplicon:
            st = np.full(n_mol, max(0, b - 20), dtype=np.int64)
            en = np.full(n_mol, min(clen, e + 20), dtype=np.int64)
            en = np.maximum(en, st + cfg.insert_min)
        else:
            if cfg.targets is None:
                st = rng.integers(0, np.maximum(1, clen - ins), size=n_mol)
            else:
                # fragments overlapping the target: start in [b - insert + 30, e - 30)
                st = rng.integers(b - ins + 30, e - 30, size=n_mol)
                st = np.clip(st, 0, clen - ins)
            en = st + ins
        en = np.minimum(en, clen)
        contig_l.append(np.full(n_mol, ci, dtype=np.int32))
        start_l.append(st.astype(np.int64))
        end_l.append(en.astype(np.int64))
    m_contig = np.concatenate(contig_l)
    m_start = np.concatenate(start_l)
    m_end = np.concatenate(end_l)
    n_mol = len(m_start)
    m_top = rng.integers(0, 2, size=n_mol).astype(np.int8)
    if not cfg.umi:
        return dict(contig=m_contig, start=m_start, end=m_end, top=m_top, mol=np.arange(n_mol, dtype=np.int64),
                    umi_a=np.zeros(n_mol, np.int64), umi_b=np.zeros(n_mol, np.int64), swapped=np.zeros(n_mol, np.int8))
    # UMI: each molecule yields a top-strand family and, with prob duplex_frac, a bottom-strand family
    umi_a = rng.integers(0, 4 ** cfg.umi_len, size=n_mol)
    umi_b = rng.integers(0, 4 ** cfg.umi_len, size=n_mol)
    has_bottom = rng.random(n_mol) < cfg.duplex_frac
    p = 1.0 / cfg.family_mean  # geometric-like NegBin(r=1) family size with the requested mean
    fam_top = rng.geometric(p, size=n_mol)
    fam_bot = np.where(has_bottom, rng.geometric(p, size=n_mol), 0)
    swapped_m = (rng.random(n_mol) < cfg.swapped_umi_frac).astype(np.int8)
    reps = fam_top + fam_bot
    idx = np.repeat(np.arange(n_mol), reps)
    # within each molecule the first fam_top copies are top-strand, the rest bottom-strand
    offs = np.arange(len(idx)) - np.repeat(np.cumsum(reps) - reps, reps)
    is_bottom = offs >= fam_top[idx]
    top = np.where(is_bottom, 1 - m_top[idx], m_top[idx]).astype(np.int8)
    return dict(contig=m_contig[idx], start=m_start[idx], end=m_end[idx], top=top, mol=idx.astype(np.int64),
                umi_a=umi_a[idx], umi_b=umi_b[idx], swapped=(swapped_m[idx] * is_bottom).astype(np.int8),
                bottom=is_bottom.astype(np.int8))


# --------------------------------------------------------------------------- BAM helpers

def reg2bin(beg: np.ndarray, end: np.ndarray) -> np.ndarray:
    end = end - 1
    out = np.zeros(len(beg), dtype=np.int64)
    done = np.zeros(len(beg), dtype=bool)
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        m = (~done) & ((beg >> shift) == (end >> shift))
        out[m] = base + (beg[m] >> shift)
        done |= m
    return out


def _pack_seq(seq_codes: np.ndarray) -> np.ndarray:
    """[n, L] nt16 codes -> [n, (L+1)//2] packed bytes."""
    n, L = seq_codes.shape
    if L % 2:
        seq_codes = np.concatenate([seq_codes, np.zeros((n, 1), np.uint8)], axis=1)
    return ((seq_codes[:, 0::2] << 4) | seq_codes[:, 1::2]).astype(np.uint8)


def _qname_matrix(ids: np.ndarray, width: int = 10) -> np.ndarray:
    """'r' + zero-padded decimal id as a uint8 matrix [n, width+1]."""
    n = len(ids)
    out = np.empty((n, width + 1), dtype=np.uint8)
    out[:, 0] = ord("r")
    v = ids.astype(np.int64).copy()
    for k in range(width, 0, -1):
        out[:, k] = 48 + (v % 10)
        v //= 10
    return out


def _umi_matrix(codes: np.ndarray, ulen: int) -> np.ndarray:
    n = len(codes)
    out = np.empty((n, ulen), dtype=np.uint8)
    v = codes.astype(np.int64).copy()
    for k in range(ulen - 1, -1, -1):
        out[:, k] = BASES[v % 4]
        v //= 4
    return out


class _BgzfWriter:
    """Writes BGZF blocks and records the compressed offset of every block."""

    def __init__(self, path: str, level: int = 1):
        self.f = open(path, "wb")
        self.level = level
        self.coff = 0
        self.block_coffs: List[int] = []   # compressed offset of each data block written through write_stream

    def _block(self, data: bytes) -> bytes:
        co = zlib.compressobj(self.level, zlib.DEFLATED, -15)
        cdata = co.compress(data) + co.flush()
        bsize = len(cdata) + 25
        hdr = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize)
        return hdr + cdata + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data))

    def write_blocks(self, data: bytes, record: bool) -> None:
        for i in range(0, len(data), BGZF_BLOCK):
            blk = self._block(data[i:i + BGZF_BLOCK])
            if record:
                self.block_coffs.append(self.coff)
            self.f.write(blk)
            self.coff += len(blk)

    def close(self) -> None:
        self.f.write(self._block(b""))
        self.f.close()


# --------------------------------------------------------------------------- read construction

def _hap_ops(ref: np.ndarray, beg: int, end: int, carried: List[Variant]):
    """Haplotype of reference interval [beg, end) as a list of (op, refpos, bases) with op in 'M','I','D'."""
    ops = []
    p = beg
    for v in carried:
        if v.kind == "snv":
            if beg <= v.pos < end:
                if v.pos > p:
                    ops.append(("M", p, ref[p:v.pos].copy()))
                ops.append(("M", v.pos, np.frombuffer(v.alt.encode(), dtype=np.uint8).copy()))
                p = v.pos + 1
        elif v.kind == "ins":
            if beg <= v.pos and v.pos + 1 < end and v.pos + 1 >= p:
                ops.append(("M", p, ref[p:v.pos + 1].copy()))
                ops.append(("I", v.pos + 1, np.frombuffer(v.alt[1:].encode(), dtype=np.uint8).copy()))
                p = v.pos + 1
        else:
            dlen = len(v.ref) - 1
            if beg <= v.pos and v.pos + 1 + dlen < end and v.pos + 1 >= p:
                ops.append(("M", p, ref[p:v.pos + 1].copy()))
                ops.append(("D", v.pos + 1, dlen))
                p = v.pos + 1 + dlen
    if p < end:
        ops.append(("M", p, ref[p:end].copy()))
    return ops


def _read_from_ops(ops, forward: bool, n: int):
    """Take the first (forward) or last (reverse) n query bases of a haplotype; returns (pos, cigar list, seq array)."""
    if not forward:
        ops = ops[::-1]
    taken = []
    need = n
    for op, rp, payload in ops:
        if need <= 0:
            break
        if op == "D":
            if taken:
                taken.append((op, rp, payload))
            continue
        L = len(payload)
        if L <= need:
            taken.append((op, rp, payload))
            need -= L
        else:
            if forward:
                taken.append((op, rp, payload[:need]))
            else:
                taken.append((op, rp + (L - need) if op == "M" else rp, payload[L - need:]))
            need = 0
    # a read may not end (or start) with a deletion / insertion
    while taken and taken[-1][0] in ("D", "I"):
        taken.pop()
    if not forward:
        taken = taken[::-1]
    while taken and taken[0][0] in ("D", "I"):
        taken.pop(0)
    if not taken:
        return None
    pos = taken[0][1]
    cigar: List[Tuple[int, int]] = []
    seq = []
    for op, rp, payload in taken:
        code = {"M": 0, "I": 1, "D": 2}[op]
        ln = payload if op == "D" else len(payload)
        if cigar and cigar[-1][0] == code:
            cigar[-1] = (code, cigar[-1][1] + ln)
        else:
            cigar.append((code, ln))
        if op != "D":
            seq.append(payload)
    return pos, cigar, np.concatenate(seq)


def _cigar_reflen(cigar) -> int:
    return sum(l for c, l in cigar if c in (0, 2, 3, 7, 8))


def _nm(ref: np.ndarray, pos: int, cigar, seq: np.ndarray) -> int:
    nm, q, r = 0, 0, pos
    for c, l in cigar:
        if c == 0:
            nm += int(np.count_nonzero(seq[q:q + l] != ref[r:r + l]))
            q += l
            r += l
        elif c == 1:
            nm += l
            q += l
        elif c == 2:
            nm += l
            r += l
        elif c == 4:
            q += l
    return nm


def _sample_quals(rng: np.random.Generator, shape) -> np.ndarray:
    u = rng.random(shape)
    q = np.full(shape, 37, dtype=np.uint8)
    mid = (u >= 0.80) & (u < 0.95)
    low = u >= 0.95
    q[mid] = rng.integers(30, 37, size=int(mid.sum()))
    q[low] = rng.integers(2, 26, size=int(low.sum()))
    return q


def generate(cfg: SynthConfig, outdir: str, bam_name: Optional[str] = None, chunk: int = 500_000) -> Dict[str, object]:
    """Generates <outdir>/<name>.fa(.fai), <name>.bam(.bai), <name>.truth.tsv and, for panels, <name>.bed."""
    os.makedirs(outdir, exist_ok=True)
    base = os.path.join(outdir, bam_name or cfg.name)
    ref = make_reference(cfg)
    write_fasta(base + ".fa", cfg, ref)
    variants = make_variants(cfg, ref)
    with open(base + ".truth.tsv", "w") as f:
        for v in variants:
            f.write("%s\t%d\t%s\t%s\t%s\t%g\n" % (cfg.contigs[v.contig][0], v.pos + 1, v.kind, v.ref, v.alt, v.vaf))
    if cfg.targets is not None:
        with open(base + ".bed", "w") as f:
            for ci, b, e in cfg.targets:
                f.write("%s\t%d\t%d\n" % (cfg.contigs[ci][0], b, e))

    fr = make_fragments(cfg)
    nfrag = len(fr["start"])
    rng = _rng(cfg, 400)

    # ---- which molecule carries which variant (decided per original molecule, so PCR copies agree)
    n_mol = int(fr["mol"].max()) + 1 if nfrag else 0
    var_by_contig: Dict[int, List[int]] = {}
    for vi, v in enumerate(variants):
        var_by_contig.setdefault(v.contig, []).append(vi)
    carried_indel: Dict[int, List[int]] = {}      # fragment -> variant indices (indels) it carries and overlaps
    snv_frag, snv_pos, snv_alt = [], [], []
    for vi, v in enumerate(variants):
        mol_carry = _rng(cfg, 1000 + vi).random(n_mol) < v.vaf
        span = (len(v.ref) if v.kind == "del" else 1)
        ov = np.nonzero((fr["contig"] == v.contig) & (fr["start"] <= v.pos + span) & (fr["end"] > v.pos - 1)
                        & mol_carry[fr["mol"]])[0]
        if v.kind == "snv":
            snv_frag.append(ov)
            snv_pos.append(np.full(len(ov), v.pos, dtype=np.int64))
            snv_alt.append(np.full(len(ov), ord(v.alt), dtype=np.uint8))
        else:
            for fi in ov.tolist():
                carried_indel.setdefault(fi, []).append(vi)
    # PCR error at "cycle 1": a molecule-level substitution carried by a random half of the copies
    if cfg.umi and cfg.pcr_err_rate > 0:
        mol_len = np.zeros(n_mol, dtype=np.int64)
        np.maximum.at(mol_len, fr["mol"], fr["end"] - fr["start"])
        has_err = rng.random(n_mol) < cfg.pcr_err_rate * mol_len
        err_off = (rng.random(n_mol) * np.maximum(mol_len, 1)).astype(np.int64)
        err_alt = rng.integers(0, 4, size=n_mol)
        sel = np.nonzero(has_err[fr["mol"]] & (rng.random(nfrag) < 0.5))[0]
        p = fr["start"][sel] + err_off[fr["mol"][sel]]
        ok = p < fr["end"][sel]
        sel, p = sel[ok], p[ok]
        snv_frag.append(sel)
        snv_pos.append(p)
        snv_alt.append(BASES[err_alt[fr["mol"][sel]]])
    snv_frag = np.concatenate(snv_frag) if snv_frag else np.zeros(0, np.int64)
    snv_pos = np.concatenate(snv_pos) if snv_pos else np.zeros(0, np.int64)
    snv_alt = np.concatenate(snv_alt) if snv_alt else np.zeros(0, np.uint8)

    # ---- reads: index 2*f = R1, 2*f+1 = R2.  "top" fragments have R1 forward at the fragment start.
    nreads = 2 * nfrag
    frag_of = np.repeat(np.arange(nfrag), 2)
    is_r2 = np.tile(np.array([0, 1], dtype=np.int8), nfrag)
    forward = (fr["top"][frag_of] ^ is_r2).astype(bool)          # top: R1 fwd, R2 rev; bottom: R1 rev, R2 fwd
    flen = (fr["end"] - fr["start"])[frag_of]
    rlen = np.minimum(READ_LEN, flen)
    pos = np.where(forward, fr["start"][frag_of], fr["end"][frag_of] - rlen).astype(np.int64)
    rend = pos + rlen
    contig = fr["contig"][frag_of]
    # slow path: carried indel overlapping the read's neighbourhood, sequencing-error indel, soft clip, short fragments
    slow = np.zeros(nreads, dtype=bool)
    for fi in carried_indel:
        slow[2 * fi] = slow[2 * fi + 1] = True
    slow |= rng.random(nreads) < cfg.indel_err * READ_LEN
    slow |= rng.random(nreads) < cfg.clip_frac
    slow |= rlen < READ_LEN
    mapq = np.full(nreads, 60, dtype=np.uint8)
    lowm = rng.random(nreads) < cfg.lowmapq_frac
    mapq[lowm] = rng.integers(0, 31, size=int(lowm.sum()))

    # slow-path reads are fully built now (they may move pos/rend); the fast ones are serialised chunk-wise later
    slow_idx = np.nonzero(slow)[0]
    slow_rec: Dict[int, Tuple[List[Tuple[int, int]], np.ndarray, np.ndarray]] = {}
    srng = _rng(cfg, 500)
    slow_edits: Dict[int, List[int]] = {}
    if len(snv_frag):
        for k in np.nonzero(np.isin(snv_frag, np.unique(slow_idx // 2)))[0].tolist():
            slow_edits.setdefault(int(snv_frag[k]), []).append(k)
    for ri in slow_idx.tolist():
        fi = ri // 2
        ci = int(contig[ri])
        rseq = ref[ci]
        carried = [variants[vi] for vi in carried_indel.get(fi, [])]
        # SNVs carried by this fragment take part in the haplotype too
        for k in slow_edits.get(fi, []):
            p = int(snv_pos[k])
            carried.append(Variant(ci, p, "snv", chr(rseq[p]), chr(snv_alt[k]), 1.0))
        carried.sort(key=lambda v: v.pos)
        fb, fe = int(fr["start"][fi]), int(fr["end"][fi])
        ops = _hap_ops(rseq, fb, fe, carried)
        got = _read_from_ops(ops, bool(forward[ri]), READ_LEN)
        if got is None:
            got = (int(pos[ri]), [(0, int(rlen[ri]))], rseq[int(pos[ri]):int(rend[ri])].copy())
        p0, cigar, seq = got
        seq = seq.copy()
        qual = _sample_quals(srng, len(seq))
        # sequencing-error indel in the middle of the read
        if srng.random() < 0.15 and len(cigar) == 1 and cigar[0][1] > 60:
            L = cigar[0][1]
            at = int(srng.integers(25, L - 25))
            if srng.random() < 0.5:
                seq = np.concatenate([seq[:at], BASES[srng.integers(0, 4, size=1)], seq[at:]])[:L]
                qual = _sample_quals(srng, L)
                cigar = [(0, at), (1, 1), (0, L - at - 1)]
            elif p0 + L + 1 <= len(rseq):
                seq = np.concatenate([rseq[p0:p0 + at], rseq[p0 + at + 1:p0 + L + 1]]).copy()
                cigar = [(0, at), (2, 1), (0, L - at)]
        # soft clip at the 3' end of the read (right end if forward, left end if reverse)
        if srng.random() < 0.6 and cigar[0][0] == 0 and cigar[-1][0] == 0 and cigar[0][1] > 45 and cigar[-1][1] > 45:
            k = int(srng.integers(5, 31))
            if forward[ri]:
                seq[len(seq) - k:] = BASES[srng.integers(0, 4, size=k)]
                cigar = cigar[:-1] + [(0, cigar[-1][1] - k), (4, k)]
            else:
                seq[:k] = BASES[srng.integers(0, 4, size=k)]
                cigar = [(4, k), (0, cigar[0][1] - k)] + cigar[1:]
                p0 += k
        # substitution errors
        perr = np.power(10.0, -qual.astype(np.float64) / 10.0) * (qual <= 25) + cfg.sub_err
        e = np.nonzero(srng.random(len(seq)) < perr)[0]
        for q in e.tolist():
            seq[q] = BASES[(int(np.where(BASES == seq[q])[0][0]) + 1 + int(srng.integers(0, 3))) % 4]
        pos[ri] = p0
        rend[ri] = p0 + _cigar_reflen(cigar)
        slow_rec[ri] = (cigar, seq, qual)

    # ---- mate fields (after slow reads settled)
    mate = np.arange(nreads) ^ 1
    mpos = pos[mate]
    left = np.minimum(pos, mpos)
    right = np.maximum(rend, rend[mate])
    tl = right - left
    isize = np.where((pos < mpos) | ((pos == mpos) & (is_r2 == 0)), tl, -tl).astype(np.int64)
    flag = (0x1 | 0x2 | np.where(is_r2 == 1, 0x80, 0x40) | np.where(forward, 0, 0x10)
            | np.where(forward[mate], 0, 0x20)).astype(np.int64)

    # ---- global coordinate order
    order = np.lexsort((np.arange(nreads), pos, contig))
    rank_of = np.empty(nreads, dtype=np.int64)
    rank_of[order] = np.arange(nreads)

    # SNV edits keyed by read: (read, query offset, alt) for fast-path reads only
    ed_read = np.concatenate([2 * snv_frag, 2 * snv_frag + 1])
    ed_pos = np.concatenate([snv_pos, snv_pos])
    ed_alt = np.concatenate([snv_alt, snv_alt])
    keep = (~slow[ed_read]) & (ed_pos >= pos[ed_read]) & (ed_pos < rend[ed_read])
    ed_read, ed_pos, ed_alt = ed_read[keep], ed_pos[keep], ed_alt[keep]
    ed_rank = rank_of[ed_read]
    eo = np.argsort(ed_rank, kind="stable")
    ed_rank, ed_read, ed_pos, ed_alt = ed_rank[eo], ed_read[eo], ed_pos[eo], ed_alt[eo]

    # ---- names
    qn = _qname_matrix(frag_of)
    if cfg.umi:
        ua = _umi_matrix(fr["umi_a"][frag_of], cfg.umi_len)
        ub = _umi_matrix(fr["umi_b"][frag_of], cfg.umi_len)
        sw = fr["swapped"][frag_of].astype(bool)
        first = np.where(sw[:, None], ub, ua)
        second = np.where(sw[:, None], ua, ub)
        qn = np.concatenate([qn, np.full((nreads, 1), ord("#"), np.uint8), first,
                             np.full((nreads, 1), ord("+"), np.uint8), second], axis=1)
    qn = np.concatenate([qn, np.zeros((nreads, 1), np.uint8)], axis=1)
    l_qname = qn.shape[1]

    # ---- BAM header
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (n, l) for n, l in cfg.contigs)
    hdr = b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(cfg.contigs))
    for n, l in cfg.contigs:
        hdr += struct.pack("<i", len(n) + 1) + n.encode() + b"\x00" + struct.pack("<i", l)
    bw = _BgzfWriter(base + ".bam")
    bw.write_blocks(hdr, record=False)

    # ---- records, chunk by chunk in coordinate order
    rec_size_fast = 4 + 32 + l_qname + 4 + (READ_LEN + 1) // 2 + READ_LEN + 4
    rec_uoff = np.zeros(nreads + 1, dtype=np.int64)      # uncompressed stream offset of each record, by rank
    pending = b""
    stream_off = 0
    qrng = _rng(cfg, 600)
    for c0 in range(0, nreads, chunk):
        ids = order[c0:c0 + chunk]
        n = len(ids)
        fast = ~slow[ids]
        fid = ids[fast]
        nf = len(fid)
        # fast records as one byte matrix
        mat = np.zeros((nf, rec_size_fast), dtype=np.uint8)
        if nf:
            hdrv = np.zeros(nf, dtype=[("bs", "<i4"), ("tid", "<i4"), ("pos", "<i4"), ("lq", "u1"), ("mq", "u1"),
                                       ("bin", "<u2"), ("nc", "<u2"), ("fl", "<u2"), ("ls", "<i4"), ("mt", "<i4"),
                                       ("mp", "<i4"), ("tl", "<i4")])
            hdrv["bs"] = rec_size_fast - 4
            hdrv["tid"] = contig[fid]
            hdrv["pos"] = pos[fid]
            hdrv["lq"] = l_qname
            hdrv["mq"] = mapq[fid]
            hdrv["bin"] = reg2bin(pos[fid], rend[fid])
            hdrv["nc"] = 1
            hdrv["fl"] = flag[fid]
            hdrv["ls"] = READ_LEN
            hdrv["mt"] = contig[fid]
            hdrv["mp"] = mpos[fid]
            hdrv["tl"] = isize[fid]
            mat[:, :36] = hdrv.view(np.uint8).reshape(nf, 36)
            o = 36
            mat[:, o:o + l_qname] = qn[fid]
            o += l_qname
            mat[:, o:o + 4] = np.frombuffer(struct.pack("<I", READ_LEN << 4), dtype=np.uint8)
            o += 4
            # sequence from the reference (+ SNV edits + errors)
            seq = np.empty((nf, READ_LEN), dtype=np.uint8)
            for ci in range(len(cfg.contigs)):
                mci = contig[fid] == ci
                if mci.any():
                    seq[mci] = ref[ci][pos[fid][mci][:, None] + np.arange(READ_LEN)[None, :]]
            refseq = seq.copy()
            lo, hi = np.searchsorted(ed_rank, [c0, c0 + n])
            if hi > lo:
                # row of each edit inside the fast matrix
                row_of_rank = np.full(n, -1, dtype=np.int64)
                row_of_rank[np.nonzero(fast)[0]] = np.arange(nf)
                rows = row_of_rank[ed_rank[lo:hi] - c0]
                cols = ed_pos[lo:hi] - pos[ed_read[lo:hi]]
                seq[rows, cols] = ed_alt[lo:hi]
            qual = _sample_quals(qrng, (nf, READ_LEN))
            perr = np.where(qual <= 25, np.power(10.0, -qual.astype(np.float64) / 10.0), 0.0) + cfg.sub_err
            err = qrng.random((nf, READ_LEN)) < perr
            if err.any():
                cur = np.searchsorted(BASES, seq[err])
                seq[err] = BASES[(cur + 1 + qrng.integers(0, 3, size=int(err.sum()))) % 4]
            nm = np.count_nonzero(seq != refseq, axis=1)
            mat[:, o:o + READ_LEN // 2] = _pack_seq(NT16[np.searchsorted(BASES, seq)])
            o += (READ_LEN + 1) // 2
            mat[:, o:o + READ_LEN] = qual
            o += READ_LEN
            mat[:, o] = ord("N"); mat[:, o + 1] = ord("M"); mat[:, o + 2] = ord("C"); mat[:, o + 3] = np.minimum(nm, 255)
        # interleave with slow records in rank order
        sizes = np.full(n, rec_size_fast, dtype=np.int64)
        slow_blobs: Dict[int, bytes] = {}
        for k in np.nonzero(~fast)[0].tolist():
            ri = int(ids[k])
            cigar, sq, ql = slow_rec[ri]
            ci = int(contig[ri])
            nmv = min(255, _nm(ref[ci], int(pos[ri]), cigar, sq))
            body = struct.pack("<iiBBHHHiiii", ci, int(pos[ri]), l_qname, int(mapq[ri]),
                               int(reg2bin(pos[ri:ri + 1], rend[ri:ri + 1])[0]), len(cigar), int(flag[ri]), len(sq),
                               ci, int(mpos[ri]), int(isize[ri]))
            body += qn[ri].tobytes()
            body += b"".join(struct.pack("<I", (l << 4) | c) for c, l in cigar)
            body += _pack_seq(NT16[np.searchsorted(BASES, sq)][None, :]).tobytes()
            body += ql.tobytes() + b"NMC" + bytes([nmv])
            blob = struct.pack("<i", len(body)) + body
            slow_blobs[k] = blob
            sizes[k] = len(blob)
        rec_uoff[c0:c0 + n] = stream_off + np.concatenate([[0], np.cumsum(sizes)[:-1]])
        stream_off += int(sizes.sum())
        if slow_blobs:
            parts = []
            frow = 0
            prev = 0
            for k in sorted(slow_blobs):
                cnt = k - prev
                if cnt:
                    parts.append(mat[frow:frow + cnt].tobytes())
                    frow += cnt
                parts.append(slow_blobs[k])
                prev = k + 1
            if frow < nf:
                parts.append(mat[frow:].tobytes())
            data = b"".join(parts)
        else:
            data = mat.tobytes()
        pending += data
        nfull = (len(pending) // BGZF_BLOCK) * BGZF_BLOCK
        bw.write_blocks(pending[:nfull], record=True)
        pending = pending[nfull:]
    bw.write_blocks(pending, record=True)
    rec_uoff[nreads] = stream_off
    total_coff = bw.coff
    bw.close()

    # ---- BAI
    coffs = np.array(bw.block_coffs + [total_coff], dtype=np.int64)
    def voff(u: np.ndarray) -> np.ndarray:
        return (coffs[u // BGZF_BLOCK] << 16) | (u % BGZF_BLOCK)
    v_beg = voff(rec_uoff[:-1])
    v_end = voff(rec_uoff[1:])
    s_contig, s_pos, s_end = contig[order], pos[order], rend[order]
    with open(base + ".bam.bai", "wb") as f:
        f.write(b"BAI\x01" + struct.pack("<i", len(cfg.contigs)))
        for ci in range(len(cfg.contigs)):
            m = np.nonzero(s_contig == ci)[0]
            if len(m) == 0:
                f.write(struct.pack("<i", 0) + struct.pack("<i", 0))
                continue
            bins = reg2bin(s_pos[m], s_end[m])
            starts = np.nonzero(np.concatenate([[True], bins[1:] != bins[:-1]]))[0]
            ends = np.concatenate([starts[1:], [len(m)]]) - 1
            run_bin = bins[starts]
            run_beg = v_beg[m][starts]
            run_end = v_end[m][ends]
            ro = np.argsort(run_bin, kind="stable")
            run_bin, run_beg, run_end = run_bin[ro], run_beg[ro], run_end[ro]
            ub, ustart = np.unique(run_bin, return_index=True)
            uend = np.concatenate([ustart[1:], [len(run_bin)]])
            f.write(struct.pack("<i", len(ub)))
            for b, s, e in zip(ub.tolist(), ustart.tolist(), uend.tolist()):
                f.write(struct.pack("<Ii", b, e - s))
                f.write(np.stack([run_beg[s:e], run_end[s:e]], axis=1).astype("<u8").tobytes())
            nwin = int((s_end[m].max() - 1) >> 14) + 1
            lin = np.full(nwin, np.iinfo(np.int64).max, dtype=np.int64)
            w0 = s_pos[m] >> 14
            w1 = (s_end[m] - 1) >> 14
            for d in range(int((w1 - w0).max()) + 1):
                sel = (w0 + d) <= w1
                np.minimum.at(lin, (w0 + d)[sel], v_beg[m][sel])
            # empty windows inherit the next filled window's offset (records are sorted)
            for w in range(nwin - 2, -1, -1):
                if lin[w] == np.iinfo(np.int64).max:
                    lin[w] = lin[w + 1]
            f.write(struct.pack("<i", nwin) + lin.astype("<u8").tobytes())
    return dict(fasta=base + ".fa", bam=base + ".bam", truth=base + ".truth.tsv",
                bed=(base + ".bed" if cfg.targets is not None else None),
                n_reads=int(nreads), n_fragments=int(nfrag), variants=variants)


# --------------------------------------------------------------------------- named configurations (BASELINE.json configs)

def named_config(name: str, scale: float = 1.0) -> SynthConfig:
    """c1..c3 of SURVEY.md section 8d; ``scale`` shrinks the region length (depth is kept)."""
    if name == "c1":
        L = max(20_000, int(1_000_000 * scale))
        return SynthConfig(name="c1", seed=1001, contigs=(("chrS1", L),), depth=100.0,
                           n_snv=max(4, int(200 * scale)), n_indel=max(2, int(60 * scale)),
                           vafs=(0.05, 0.10, 0.25, 0.50))
    if name == "c2":
        n_targets = max(4, int(10_000 * scale))
        L = n_targets * 1000 + 2000
        targets = [(0, 1000 + 1000 * i, 1200 + 1000 * i) for i in range(n_targets)]
        return SynthConfig(name="c2", seed=1002, contigs=(("chrS2", L),), depth=2000.0, targets=targets,
                           amplicon_frac=0.2, n_snv=max(4, int(400 * scale)), n_indel=max(2, int(100 * scale)),
                           vafs=(0.005, 0.01, 0.02, 0.05))
    if name == "c3":
        L = max(5_000, int(1_000_000 * scale))
        return SynthConfig(name="c3", seed=1003, contigs=(("chrS3", L),), depth=20000.0, umi=True,
                           n_snv=max(4, int(200 * scale)), n_indel=max(2, int(50 * scale)),
                           vafs=(0.001, 0.005, 0.01))
    raise ValueError("unknown config " + name)


if __name__ == "__main__":
    import argparse
    import time
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("outdir")
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    t0 = time.time()
    info = generate(named_config(args.config, args.scale), args.outdir)
    print("generated %d reads (%d fragments, %d variants) in %.1fs: %s" % (
        info["n_reads"], info["n_fragments"], len(info["variants"]), time.time() - t0, info["bam"]))
