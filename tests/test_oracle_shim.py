"""Self-check of the oracle's htslib shim (oracle/htslib_compat): the reference reaches its input only through htslib, which is not vendored,
so the oracle links the unmodified reference against a shim - and parity at that boundary is pinned by nothing in the reference (SURVEY.md
section 8c, step 5). Here everything the shim decodes from a BAM / BAI / FASTA is compared with an independent pure-Python decoder
(gzip multi-member inflate + struct), record by record, for a whole-file pass and for index queries with htslib's overlap semantics
(pos < end and bam_endpos > beg), including records that span BGZF blocks (the Python generator writes them)."""
import gzip
import os
import struct
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "oracle", "_ref", "shim_selfcheck")

pytestmark = pytest.mark.skipif(not os.path.exists(TOOL), reason="oracle/_ref/shim_selfcheck not built")


def _python_bam(path):
    """[(tid, pos, endpos, flag, mapq, mpos, isize, mtid, qname, nm, cigar, seq, qual)] and the reference sequences of a BAM file."""
    d = gzip.open(path, "rb").read()
    assert d[:4] == b"BAM\x01"
    l_text, = struct.unpack_from("<i", d, 4)
    o = 8 + l_text
    n_ref, = struct.unpack_from("<i", d, o)
    o += 4
    refs = []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", d, o)
        name = d[o + 4:o + 4 + l_name - 1].decode()
        l_ref, = struct.unpack_from("<i", d, o + 4 + l_name)
        refs.append((name, l_ref))
        o += 8 + l_name
    recs = []
    while o < len(d):
        bs, = struct.unpack_from("<i", d, o)
        tid, pos, l_qname, mapq, _bin, n_cigar, flag, l_seq, mtid, mpos, isize = struct.unpack_from("<iiBBHHHiiii", d, o + 4)
        p = o + 36
        qname = d[p:p + l_qname - 1].decode()
        p += l_qname
        cig = struct.unpack_from("<%dI" % n_cigar, d, p)
        p += 4 * n_cigar
        seqb = d[p:p + (l_seq + 1) // 2]
        p += (l_seq + 1) // 2
        qual = d[p:p + l_seq]
        p += l_seq
        aux = d[p:o + 4 + bs]
        nm = -1
        a = 0
        while a + 3 <= len(aux):
            tag, ty = aux[a:a + 2], chr(aux[a + 2])
            size = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}.get(ty)
            if size is None:
                break
            if tag == b"NM":
                nm = int.from_bytes(aux[a + 3:a + 3 + size], "little", signed=ty in "csi")
                break
            a += 3 + size
        reflen = sum(c >> 4 for c in cig if (c & 0xf) in (0, 2, 3, 7, 8)) if not (flag & 4) else 0
        endpos = pos + (reflen if reflen else 1)
        cigar = "".join("%d%s" % (c >> 4, "MIDNSHP=XB"[c & 0xf]) for c in cig)
        seq = "".join("=ACMGRSVTWYHKDBN"[(seqb[i >> 1] >> ((~i & 1) << 2)) & 0xf] for i in range(l_seq))
        recs.append((tid, pos, endpos, flag, mapq, mpos, isize, mtid, qname, nm, cigar, seq, "".join(chr(33 + q) for q in qual)))
        o += 4 + bs
    return refs, recs


def _shim(args):
    out = subprocess.run([TOOL] + [str(a) for a in args], check=True, stdout=subprocess.PIPE, text=True).stdout.split("\n")
    refs = [tuple(l.split("\t")[1:]) for l in out if l.startswith("@\t")]
    recs = []
    for l in out:
        if l and not l.startswith("@\t"):
            f = l.split("\t")
            recs.append(tuple(int(x) for x in f[:8]) + (f[8], int(f[9]), f[10], f[11], f[12]))
    return [(n, int(l)) for n, l in refs], recs


def test_shim_whole_file_and_region_queries(synth_small, synth_umi):
    for info, queries in ((synth_small, [(0, 0, 1), (0, 3000, 3001), (0, 5900, 6100), (0, 11900, 12000), (0, 0, 12000)]), (synth_umi, [(0, 900, 1100), (0, 2400, 2600)])):
        refs, truth = _python_bam(info["bam"])
        s_refs, s_recs = _shim([info["bam"], "all"])
        assert s_refs == refs
        assert s_recs == truth and len(truth) > 1000
        for tid, beg, end in queries:
            _, got = _shim([info["bam"], "region", tid, beg, end])
            want = [r for r in truth if r[0] == tid and r[1] < end and r[2] > beg]
            assert got == want, (tid, beg, end, len(got), len(want))


def test_shim_fasta_fetch(synth_small):
    fa = open(synth_small["fasta"]).read().split("\n")
    name = fa[0][1:]
    seq = "".join(fa[1:])
    for beg, end in ((0, 9), (55, 130), (11990, 11999)):      # faidx_fetch_seq takes an inclusive end
        out = subprocess.run([TOOL, synth_small["fasta"], "fetch", name, str(beg), str(end)], check=True, stdout=subprocess.PIPE, text=True).stdout.strip().split("\t")
        assert int(out[0]) == end - beg + 1 and out[1] == seq[beg:end + 1]
