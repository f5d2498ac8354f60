"""The uvc1 command line (SURVEY.md rows a-0 and b): same tile list as the reference's SamIter for the same -t / --mem-per-thread
(compared through --bed-out-fname, byte for byte), same VCF body as the reference executable (byte for byte after the ## header lines,
which carry date / version / command line), valid BGZF framing, and output independent of batching.

CPU tests drive tests/emu/uvc1_emu (the same host program linked against the test-only emulation library); the ``gpu`` tests drive the
product binary uvc_b200/bin/uvc1 on the B200. The reference executable is oracle/_ref/uvc1 (the unmodified reference)."""
import gzip
import os
import subprocess

import pytest

import parity_util as pu

ROOT = pu.ROOT
EMU = os.path.join(ROOT, "tests", "emu", "uvc1_emu")
GPU = os.path.join(ROOT, "uvc_b200", "bin", "uvc1")

needs_ref = pytest.mark.skipif(not os.path.exists(pu.REF_UVC1), reason="oracle/_ref/uvc1 not built")


def _run(exe, bam, fasta, out, extra):
    cmd = [exe, bam, "-f", fasta, "-o", out, "-s", "S"] + list(extra)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, "%s failed: %s" % (" ".join(cmd), p.stderr[-2000:])
    return p


def _body(path):
    with gzip.open(path, "rt") as f:
        return [l for l in f.read().split("\n") if l and not l.startswith("##")]


def _compare(exe, data, tmp_path, extra, bed=True):
    ref_out, our_out = str(tmp_path / "ref.vcf.gz"), str(tmp_path / "ours.vcf.gz")
    ref_bed, our_bed = str(tmp_path / "ref.bed"), str(tmp_path / "ours.bed")
    _run(pu.REF_UVC1, data["bam"], data["fasta"], ref_out, list(extra) + (["--bed-out-fname", ref_bed] if bed else []))
    _run(exe, data["bam"], data["fasta"], our_out, list(extra) + (["--bed-out-fname", our_bed] if bed else []))
    if bed:
        assert open(our_bed).read() == open(ref_bed).read()
    ref, ours = _body(ref_out), _body(our_out)
    assert len(ref) > 1
    for i, (a, b) in enumerate(zip(ours, ref)):
        assert a == b, "line %d differs:\n ours %s\n ref  %s" % (i, a[:300], b[:300])
    assert len(ours) == len(ref)
    return our_out


@pytest.fixture(scope="module")
def synth_cli(tmp_path_factory):
    """Two contigs (a contig change inside the run), 60x, spiked variants."""
    from uvc_b200 import synth
    d = tmp_path_factory.mktemp("synth_cli")
    cfg = synth.SynthConfig(name="cli", seed=99, contigs=(("chrA", 24000), ("chrB", 9000)), depth=50.0, n_snv=12, n_indel=6)
    return synth.generate(cfg, str(d))


@needs_ref
def test_cli_default_matches_reference(synth_cli, tmp_path):
    out = _compare(EMU, synth_cli, tmp_path, ["-t", "4"])
    raw = open(out, "rb").read()
    assert raw[:4] == b"\x1f\x8b\x08\x04" and raw[12:14] == b"BC"          # BGZF member with the BC extra field
    assert raw.endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))   # end-of-file block


@needs_ref
def test_cli_tiler_under_memory_pressure(synth_cli, tmp_path):
    """--mem-per-thread 8 forces sub-memory cuts every ~220 bp and several tier-1 iterations (the read that triggers a tier-1 cut is dropped)."""
    _compare(EMU, synth_cli, tmp_path, ["-t", "2", "--mem-per-thread", "8"])
    n_iter = len(set(l.split("\t")[10] for l in open(str(tmp_path / "ref.bed"))))
    assert n_iter > 1


@needs_ref
def test_cli_platform_offsets_are_added_to_user_values(synth_cli, tmp_path):
    """--syserr-minABQ-* given on the command line: the reference adds the Illumina offsets (200 / 100) to the user value (CmdLineArgs.cpp:127-134)."""
    _compare(EMU, synth_cli, tmp_path, ["-t", "4", "--syserr-minABQ-pcr-snv", "50", "--syserr-minABQ-cap-snv", "50", "--syserr-minABQ-cap-indel", "20"], bed=False)


def test_cli_fails_on_bad_index(synth_cli, tmp_path):
    """A .bai with a foreign magic or a truncated .bai is an error (the reference: 'Failed to load BAM index'), not an empty VCF."""
    import shutil
    bam = str(tmp_path / "x.bam")
    shutil.copy(synth_cli["bam"], bam)
    bai = open(synth_cli["bam"] + ".bai", "rb").read()
    for bad in (b"XXXX" + bai[4:], bai[:60]):
        open(bam + ".bai", "wb").write(bad)
        p = subprocess.run([EMU, bam, "-f", synth_cli["fasta"], "-o", str(tmp_path / "x.vcf.gz")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode != 0 and "BAM index" in p.stderr


def test_cli_rejects_germline_records(synth_cli, tmp_path):
    p = subprocess.run([EMU, synth_cli["bam"], "-f", synth_cli["fasta"], "-o", str(tmp_path / "x.vcf.gz"), "--outvar-flag", "63"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 105 and "GERMLINE" in p.stderr


@needs_ref
def test_cli_output_independent_of_batching(synth_cli, tmp_path):
    a, b = str(tmp_path / "a.vcf.gz"), str(tmp_path / "b.vcf.gz")
    _run(EMU, synth_cli["bam"], synth_cli["fasta"], a, ["-t", "3", "--mem-per-thread", "30", "--gpu-batch-positions", "3000", "--lanes-per-gpu", "3"])
    _run(EMU, synth_cli["bam"], synth_cli["fasta"], b, ["-t", "3", "--mem-per-thread", "30", "--gpu-batch-positions", "100000000", "--lanes-per-gpu", "1"])
    assert _body(a) == _body(b)


@needs_ref
def test_cli_bed_regions_umi(synth_umi, tmp_path):
    bed = tmp_path / "r.bed"
    bed.write_text("chrU\t1200\t1500\nchrU\t2000\t2300\n")
    _compare(EMU, synth_umi, tmp_path, ["-t", "2", "-R", str(bed)])


@needs_ref
def test_cli_dense_panel_bed(synth_cli, tmp_path):
    """40 small targets 500 bp apart: the tiler counts their reads in one sweep over the BAM and every lane decodes its run of targets once
    (span fetch) - both must give what the reference's per-target index queries give. Also an unsorted BED (falls back to per-target queries)."""
    bed = tmp_path / "panel.bed"
    bed.write_text("".join("chrA\t%d\t%d\n" % (1000 + 500 * i, 1150 + 500 * i) for i in range(40)) + "chrB\t2000\t2300\n")
    (tmp_path / "s").mkdir()
    _compare(EMU, synth_cli, tmp_path / "s", ["-t", "3", "-R", str(bed)])
    bed2 = tmp_path / "unsorted.bed"
    bed2.write_text("".join("chrA\t%d\t%d\n" % (1000 + 500 * i, 1150 + 500 * i) for i in reversed(range(36))))
    (tmp_path / "u").mkdir()
    _compare(EMU, synth_cli, tmp_path / "u", ["-t", "2", "-R", str(bed2)])


@needs_ref
def test_cli_targets(synth_cli, tmp_path):
    _compare(EMU, synth_cli, tmp_path, ["-t", "2", "--targets", "chrA:5000-7000,chrB:100-900"])


def test_cli_plain_text_to_stdout_and_errors(synth_cli, tmp_path):
    p = _run(EMU, synth_cli["bam"], synth_cli["fasta"], "-", ["-t", "2", "--targets", "chrA:5000-5400"])
    lines = p.stdout.split("\n")
    assert lines[0] == "##fileformat=VCFv4.2"
    assert any(l.startswith("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS") for l in lines)
    assert any("MGVCF_BLOCK" in l and not l.startswith("#") for l in lines)
    q = subprocess.run([EMU, str(tmp_path / "missing.bam"), "-f", "NA"], capture_output=True, text=True)
    assert q.returncode != 0 and "does not exist" in q.stderr
    q = subprocess.run([EMU, synth_cli["bam"], "-f", synth_cli["fasta"], "--no-such-option", "1"], capture_output=True, text=True)
    assert q.returncode != 0


def test_tiler_library_api(synth_cli):
    """The tiler through libuvchost.so's C API: tiles partition the covered genome in order and end with the end-of-file flag."""
    import ctypes as C
    from uvc_b200 import capi
    lib = capi.load_host()

    class BedLine(C.Structure):
        _fields_ = [("tid", C.c_int32), ("beg_pos", C.c_int32), ("end_pos", C.c_int32), ("region_flag", C.c_uint32), ("n_reads", C.c_int64)]
    lib.uvchost_tiler_open.restype = C.c_void_p
    lib.uvchost_tiler_open.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32]
    lib.uvchost_tiler_next.restype = C.c_int64
    lib.uvchost_tiler_next.argtypes = [C.c_void_p, C.POINTER(C.POINTER(BedLine)), C.POINTER(C.c_int64)]
    lib.uvchost_tiler_close.argtypes = [C.c_void_p]
    t = lib.uvchost_tiler_open(synth_cli["bam"].encode(), b"", b"", 2, 20, -1, 0)
    tiles = []
    while True:
        p, n = C.POINTER(BedLine)(), C.c_int64()
        nreads = lib.uvchost_tiler_next(t, C.byref(p), C.byref(n))
        assert nreads >= 0
        if nreads == 0 and n.value == 0:
            break
        tiles += [(p[i].tid, p[i].beg_pos, p[i].end_pos, p[i].region_flag, p[i].n_reads) for i in range(n.value)]
    lib.uvchost_tiler_close(t)
    assert len(tiles) > 10
    for a, b in zip(tiles, tiles[1:]):
        assert (a[0], a[2]) <= (b[0], b[1])            # ordered, non-overlapping
    assert tiles[-1][3] & 0x2                          # end of file
    assert any(t[3] & 0x10 for t in tiles)             # contig change
    assert sum(t[4] for t in tiles) <= synth_cli["n_reads"]


@pytest.mark.gpu
@needs_ref
def test_cuda_cli_matches_reference(synth_cli, synth_umi, tmp_path):
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir(); (tmp_path / "c").mkdir()
    _compare(GPU, synth_cli, tmp_path / "a", ["-t", "4"])
    _compare(GPU, synth_cli, tmp_path / "b", ["-t", "2", "--mem-per-thread", "8"])
    bed = tmp_path / "r.bed"
    bed.write_text("chrU\t1200\t1500\nchrU\t2000\t2300\n")
    _compare(GPU, synth_umi, tmp_path / "c", ["-t", "2", "-R", str(bed)])


@pytest.mark.gpu
def test_gpu_cli_two_gpus_equal_one(synth_cli, tmp_path):
    """Tiles sharded over the lanes of two GPUs, VCF text concatenated in tile order: byte-identical to the one-GPU run."""
    import ctypes
    from uvc_b200 import capi
    if capi.load_gpu().uvcgpu_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    outs = []
    for n in (1, 2):
        out = str(tmp_path / ("g%d.vcf.gz" % n))
        _run(GPU, synth_cli["bam"], synth_cli["fasta"], out, ["-t", "4", "--mem-per-thread", "8", "--gpus", str(n), "--gpu-batch-positions", "3000"])
        outs.append(_body(out))
    assert len(outs[0]) > 1 and outs[0] == outs[1]
