"""Scoring-stage parity (SURVEY.md rows a-13 ... a-19): the VCF body our C ABI returns for a tile must equal, byte for byte, what the
reference prints for the same region - every FORMAT value, INFO field, QUAL and FILTER, the MGVCF block lines and the
additional-indel-candidate lines, in the same order. (The stated tolerance for the floating-point scoring is 0.01 phred with identical
FILTER calls; identical text is stricter, and it is what we get.)

Checked against (1) the committed golden fixtures generated from the unmodified reference binary, and (2) where oracle/_ref exists,
the reference itself on freshly generated data."""
import gzip
import importlib.util
import os

import pytest

import parity_util as pu

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_vcf", os.path.join(HERE, "golden", "make_golden_vcf.py"))
mgv = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mgv)


def _our_lines(bam, fasta, tiles, emulate, **params):
    out, stats = pu.run_tiles(bam, fasta, tiles, emulate, ["vcf"], **params)
    return [l for o in out for l in o["vcf"].split("\n") if l], stats


def _assert_same(ours, ref):
    assert len(ref) > 0
    for i, (a, b) in enumerate(zip(ours, ref)):
        if a != b:
            fa, fb = a.split("\t"), b.split("\t")
            detail = [(j, x, y) for j, (x, y) in enumerate(zip(fa, fb)) if x != y and j != 9]
            if len(fa) > 9 and len(fb) > 9 and fa[9] != fb[9]:
                detail += [(t, x, y) for t, x, y in zip(fb[8].split(":"), fa[9].split(":"), fb[9].split(":")) if x != y]
            raise AssertionError("line %d (%s) differs: %s" % (i, "\t".join(fb[:5]), detail[:12]))
    assert len(ours) == len(ref)


def _golden_case(case_index, emulate, tmp_path):
    fname, tile, _, params = mgv.CASES[case_index]
    info = mgv.mg.golden_inputs(str(tmp_path))
    with gzip.open(os.path.join(HERE, "golden", fname), "rt") as f:
        ref = [l for l in f.read().split("\n") if l]
    ours, stats = _our_lines(info["bam"], info["fasta"], [tile], emulate, **params)
    _assert_same(ours, ref)
    return stats


@pytest.mark.parametrize("case_index", range(len(mgv.CASES)))
def test_emulation_matches_golden_vcf(case_index, tmp_path):
    _golden_case(case_index, True, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("case_index", range(len(mgv.CASES)))
def test_cuda_matches_golden_vcf(case_index, tmp_path):
    st = _golden_case(case_index, False, tmp_path)
    assert st.gpu_launches > 0


def _vs_reference(paths, contig, tiles, opts, params, emulate, tmp_path):
    ref = []
    for k, t in enumerate(tiles):
        d = tmp_path / ("r%d" % k)
        d.mkdir()
        ref += mgv.reference_vcf_lines(paths["bam"], paths["fasta"], contig, t, opts, str(d))
    ours, stats = _our_lines(paths["bam"], paths["fasta"], tiles, emulate, **params)
    _assert_same(ours, ref)
    return stats


@pytest.mark.skipif(not os.path.exists(pu.REF_UVC1), reason="oracle/_ref/uvc1 not built")
def test_emulation_vcf_vs_reference_small(synth_small, tmp_path):
    _vs_reference(synth_small, "chrA", [(0, 0, 6000, 0), (0, 6000, 12000, 0)], [], {}, True, tmp_path)


@pytest.mark.skipif(not os.path.exists(pu.REF_UVC1), reason="oracle/_ref/uvc1 not built")
def test_emulation_vcf_vs_reference_umi_allout(synth_umi, tmp_path):
    _vs_reference(synth_umi, "chrU", [(0, 2000, 2400, 0)], ["-A"], {"should_output_all": 1}, True, tmp_path)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(pu.REF_UVC1), reason="oracle/_ref/uvc1 not built")
def test_cuda_vcf_vs_reference(synth_small, synth_umi, tmp_path):
    st = _vs_reference(synth_small, "chrA", [(0, 0, 6000, 0), (0, 6000, 12000, 0)], ["-A"], {"should_output_all": 1}, False, tmp_path)
    assert st.gpu_launches > 0
    (tmp_path / "u").mkdir()
    _vs_reference(synth_umi, "chrU", [(0, 1000, 4000, 0)], [], {}, False, tmp_path / "u")


def _restricted_equals_full(paths, tiles, emulate, **params):
    """The product computes the output-only position counters on the positions that can reach the output (all_positions = 0); the text must be
    what the whole-extent run (the reference's arrays, all_positions = 1) gives."""
    a, st = _our_lines(paths["bam"], paths["fasta"], tiles, emulate, all_positions=0, **params)
    b, _ = _our_lines(paths["bam"], paths["fasta"], tiles, emulate, all_positions=1, **params)
    assert len(a) > 0 and a == b
    return st


def test_emulation_vcf_needed_positions_equal_all_positions(synth_small, synth_umi):
    # tiles much shorter than the extent of their reads (the halo is most of the tile's arrays), with MGVCF block lines that read 1000 positions ahead
    _restricted_equals_full(synth_small, [(0, 3000, 3200, 0), (0, 3200, 3250, 0), (0, 5000, 5400, 0)], True)
    _restricted_equals_full(synth_umi, [(0, 2000, 2300, 0)], True, should_output_all=1)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(pu.REF_UVC1), reason="oracle/_ref/uvc1 not built")
def test_cuda_vcf_c2_depth_panel_vs_reference(synth_c2_depth, tmp_path):
    """Panel tiles at 2000x (configs[1]): each target its own tile, so most of a tile's extended range is halo that the output-only kernels skip."""
    tiles = [(0, b, e, 0) for (_, b, e) in synth_c2_depth["targets"]]
    st = _vs_reference(synth_c2_depth, "chrP", tiles, [], {}, False, tmp_path)
    assert st.gpu_launches > 0
    _restricted_equals_full(synth_c2_depth, tiles, False)


def test_emulation_vcf_text_in_many_ranges(synth_small, synth_umi, monkeypatch):
    """The text of a tile is formatted in independent ranges of positions (512 records each by default): tiny ranges must give the same bytes,
    including the MGVCF block lines, the additional-indel-candidate lines (they look at the position before the range) and empty ranges."""
    tiles = [(0, 0, 6000, 0), (0, 6000, 12000, 0)]
    a, _ = _our_lines(synth_small["bam"], synth_small["fasta"], tiles, True, should_output_all=1)
    u, _ = _our_lines(synth_umi["bam"], synth_umi["fasta"], [(0, 1500, 2500, 0)], True)
    monkeypatch.setenv("UVC_TEXT_RECS_PER_RANGE", "3")
    b, _ = _our_lines(synth_small["bam"], synth_small["fasta"], tiles, True, should_output_all=1)
    w, _ = _our_lines(synth_umi["bam"], synth_umi["fasta"], [(0, 1500, 2500, 0)], True)
    assert len(a) > 1000 and a == b and len(u) > 0 and u == w


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_emulation_needed_positions_random_tilings(synth_small, seed):
    """Random cuts of a region into tiles of 1 to ~900 bases (tiny tiles, tiles that start where the previous one ends, gaps between tiles): the
    text with the output-only counters restricted to the needed positions must equal the whole-extent text."""
    import random
    rnd = random.Random(seed)
    tiles, pos = [], rnd.randint(0, 300)
    while pos < 5200 and len(tiles) < 9:
        length = rnd.choice([1, 2, 3, 17, 64, 200, 333, 900])
        tiles.append((0, pos, pos + length, 0))
        pos += length + rnd.choice([0, 0, 1, 50, 700])
    _restricted_equals_full(synth_small, tiles, True)


def _two_allele_site(tmp_path_factory_dir):
    """configs[0] as the bench generates it has a site (chrS1:78101) with two different insertions of the same length class: the reference formats
    the alleles of one indel symbol on one object, and the second allele's record repeats the first one's nNFA / nAFA / nBCFA in front of its own."""
    import bench
    ds = bench.dataset(str(tmp_path_factory_dir), "c1", 1.0, 0, 4)
    return ds, [(0, 77000, 79500, 0)]


@pytest.mark.skipif(not os.path.exists(pu.REF_UVC1), reason="oracle/_ref/uvc1 not built")
def test_emulation_two_alleles_of_one_indel_symbol(tmp_path_factory, tmp_path):
    ds, tiles = _two_allele_site(tmp_path_factory.mktemp("c1_full"))
    _vs_reference({"bam": ds["bam"], "fasta": ds["fasta"]}, "chrS1", tiles, [], {}, True, tmp_path)
    ours, _ = _our_lines(ds["bam"], ds["fasta"], tiles, True)
    site = [l.split("\t") for l in ours if l.split("\t")[1] == "78101" and len(l.split("\t")[4]) > 1]
    assert len(site) == 2
    n_nfa = [len(dict(zip(f[8].split(":"), f[9].split(":")))["nNFA"].split(",")) for f in site]
    assert sorted(n_nfa) == [6, 12]


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(pu.REF_UVC1), reason="oracle/_ref/uvc1 not built")
def test_cuda_two_alleles_of_one_indel_symbol(tmp_path_factory, tmp_path):
    ds, tiles = _two_allele_site(tmp_path_factory.mktemp("c1_full_gpu"))
    st = _vs_reference({"bam": ds["bam"], "fasta": ds["fasta"]}, "chrS1", tiles, [], {}, False, tmp_path)
    assert st.gpu_launches > 0
