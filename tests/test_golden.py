"""Parity against the committed golden fixture (generated from the oracle by tests/golden/make_golden.py).

The fixture covers every per-position counter array after updateByRegion3Aln, the family grouping, the indel identity maps
and the haplotype links of one seeded duplex-UMI tile. Bar: bit-exact."""
import importlib.util
import os

import numpy as np
import pytest

import parity_util as pu
from uvc_b200 import refdump

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)

PAIRS = [("rtr_initial", "rtr_initial"), ("rtr", "rtr_final"), ("baq", "baq"), ("baq2", "baq2"), ("prep", "prep"), ("thres", "thres"),
         ("seginfo", "seginfo"), ("faminfo", "faminfo"), ("fragdepth0", "fragdepth0"), ("fragdepth1", "fragdepth1"),
         ("famdepth0", "famdepth0"), ("famdepth1", "famdepth1"), ("duplex", "duplex")]


def _golden():
    z = np.load(os.path.join(HERE, "golden", "tile_umi.npz"))
    out = {}
    for k in z.files:
        if k in mg.TEXT_SECTIONS:
            out[k] = z[k].tobytes().decode()
        else:
            out[k] = np.frombuffer(z[k].tobytes(), dtype=refdump.section_dtype(k))
    return out


def _run(emulate, tmp_path):
    info = mg.golden_inputs(str(tmp_path))
    gold = _golden()
    secs = ["meta", "families", "indelmaps", "haplinks", "vq"] + [a for a, _ in PAIRS]
    ours, stats = pu.run_tiles(info["bam"], info["fasta"], [mg.GOLDEN_TILE], emulate, secs)
    o = ours[0]
    assert list(o["meta"][:9]) == list(gold["meta"][:9])
    assert o["families"] == gold["families"]
    assert sorted(o["indelmaps"].split("\n")) == sorted(gold["indelmaps"].split("\n"))
    assert sorted(o["haplinks"].split("\n")) == sorted(gold["haplinks"].split("\n"))
    ext_beg = int(gold["meta"][6])
    msgs = []
    for a, b in PAIRS:
        msgs += pu.diff_section(a, o[a], gold[b], ext_beg)
    msgs += pu.diff_section("vq", o["vq"][:, :, :14], gold["vq"][:, :, :14], ext_beg)
    assert not msgs, "\n".join(msgs[:40])
    # the fixture must exercise the family paths
    assert np.count_nonzero(gold["duplex"]) > 0 and np.count_nonzero(gold["famdepth0"][:, :, 2]) > 0
    return stats


def test_emulation_matches_golden(tmp_path):
    _run(True, tmp_path)


@pytest.mark.gpu
def test_cuda_matches_golden(tmp_path):
    st = _run(False, tmp_path)
    assert st.gpu_launches > 0
