"""Parity of the pileup stages against the oracle (the unmodified reference run through oracle/_ref/uvc_ref_dump).

Bar: bit-exact (all sections are integer counters). The ``gpu`` tests run the CUDA path through the C ABI; the others
run the same kernel bodies in the test-only emulation build so that the logic is covered in a container without a GPU.
"""
import os

import numpy as np
import pytest

import parity_util as pu

SECTIONS = ["meta", "families", "rtr_initial", "baq", "baq2", "prep", "thres", "rtr", "seginfo", "vq", "fragdepth0", "fragdepth1",
            "famdepth0", "famdepth1", "faminfo", "duplex", "indelmaps", "haplinks"]
IMPLEMENTED_VQ_TAGS = 14   # VQ_a1BQf .. VQ_cIDQr: everything updateByRegion3Aln fills (main_conversion.hpp:743-762)


def _check(info, tiles, emulate, tmp_path, extra=(), **params):
    if not pu.have_oracle():
        pytest.skip("oracle/_ref not built")
    ours, stats = pu.run_tiles(info["bam"], info["fasta"], tiles, emulate, SECTIONS, **params)
    assert emulate or stats.gpu_launches > 0
    prev = (-1, 0, 0)
    for ti, (tid, beg, end, flag) in enumerate(tiles):
        ref = pu.run_oracle_dump(info["bam"], info["fasta"], tid, beg, end, flag, str(tmp_path / ("t%d.dump" % ti)), prev, extra)
        prev = (tid, beg, end)
        o = ours[ti]
        assert list(o["meta"][:9]) == list(ref["meta"][:9])
        assert o["families"] == ref["families"]
        assert sorted(o["indelmaps"].split("\n")) == sorted(ref["indelmaps"].split("\n"))
        assert sorted(o["haplinks"].split("\n")) == sorted(ref["haplinks"].split("\n"))
        ext_beg = int(ref["meta"][6])
        msgs = []
        for ours_name, ref_name in [("rtr_initial", "rtr_initial"), ("baq", "baq"), ("baq2", "baq2"), ("prep", "prep"),
                                    ("thres", "thres"), ("rtr", "rtr_final"), ("seginfo", "seginfo"), ("fragdepth0", "fragdepth0"),
                                    ("fragdepth1", "fragdepth1"), ("famdepth0", "famdepth0"), ("famdepth1", "famdepth1"),
                                    ("faminfo", "faminfo"), ("duplex", "duplex")]:
            msgs += pu.diff_section(ours_name, o[ours_name], ref[ref_name], ext_beg)
        msgs += pu.diff_section("vq", o["vq"][:, :, :IMPLEMENTED_VQ_TAGS], ref["vq"][:, :, :IMPLEMENTED_VQ_TAGS], ext_beg)
        assert not msgs, "\n".join(msgs[:40])
    return stats


def test_emulated_pileup_matches_oracle_small(synth_small, tmp_path):
    _check(synth_small, [(0, 0, 6000, 4), (0, 6000, 11995, 2)], True, tmp_path)


def test_emulated_pileup_matches_oracle_umi(synth_umi, tmp_path):
    _check(synth_umi, [(0, 1000, 2500, 4), (0, 2500, 4000, 2)], True, tmp_path)


# Non-default parameters that steer the kernels onto their general paths: a mutation neighbourhood wider than the 32-bit chunk masks
# (K3a scans bit by bit), and fam_thres_dup1add = 1, with which every single-fragment family enters a quality bucket (K4 runs family loop 2
# for every family: its need-list overflows and the window is walked a second time, log and all).
GENERAL_PATH_ARGS = ["--syserr-mut-region-n-bases", "40", "--fam-thres-dup1add", "1"]
GENERAL_PATH_PARAMS = dict(syserr_mut_region_n_bases=40, fam_thres_dup1add=1)


def test_emulated_pileup_general_paths(synth_small, tmp_path):
    _check(synth_small, [(0, 0, 6000, 4)], True, tmp_path, GENERAL_PATH_ARGS, **GENERAL_PATH_PARAMS)


@pytest.mark.gpu
def test_cuda_pileup_general_paths(synth_small, synth_umi, tmp_path):
    _check(synth_small, [(0, 0, 6000, 4), (0, 6000, 11995, 2)], False, tmp_path, GENERAL_PATH_ARGS, **GENERAL_PATH_PARAMS)
    _check(synth_umi, [(0, 1000, 2500, 4)], False, tmp_path, GENERAL_PATH_ARGS, **GENERAL_PATH_PARAMS)


@pytest.mark.gpu
@pytest.mark.parametrize("block", ["64", "128"])
def test_cuda_pileup_block_shapes(synth_small, tmp_path, block, monkeypatch):
    """The position kernels pick 32, 64 or 128 threads per block from the batch size (small test tiles get 32): force the other shapes."""
    monkeypatch.setenv("UVC_POS_BLOCK", block)
    _check(synth_small, [(0, 0, 6000, 4), (0, 6000, 11995, 2)], False, tmp_path)


@pytest.mark.gpu
def test_cuda_pileup_matches_oracle_small(synth_small, tmp_path):
    _check(synth_small, [(0, 0, 6000, 4), (0, 6000, 11995, 2)], False, tmp_path)


@pytest.mark.gpu
def test_cuda_pileup_matches_oracle_umi(synth_umi, tmp_path):
    _check(synth_umi, [(0, 1000, 2500, 4), (0, 2500, 4000, 2)], False, tmp_path)


@pytest.mark.gpu
def test_cuda_pileup_c2_depth(synth_c2_depth, tmp_path):
    """Panel tiles at 2000x (configs[1]): one tile per target, amplicon-like and capture-like targets, neighbouring tiles share their read halos."""
    tiles = [(0, b, e, 8) for (_, b, e) in synth_c2_depth["targets"]]
    st = _check(synth_c2_depth, tiles, False, tmp_path)
    assert st.n_reads_kept > 10000


@pytest.mark.gpu
def test_cuda_pileup_c3_depth(synth_c3_depth, tmp_path):
    """One tile at 20 000x raw depth with duplex UMIs (configs[2]): deep windows, multi-fragment families on both strands (the wide K4 shape)."""
    st = _check(synth_c3_depth, [(0, 1000, 2000, 4)], False, tmp_path)
    assert st.n_reads_kept > 100000 and st.n_families > 1000
