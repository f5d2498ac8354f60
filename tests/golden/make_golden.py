"""Generates the committed golden fixture tests/golden/tile_umi.npz from the ORACLE (the unmodified reference run through
oracle/_ref/uvc_ref_dump). Run in the build container where /root/reference exists:

    python tests/golden/make_golden.py

The fixture pins every per-position counter array of Symbol2CountCoverageSet after updateByRegion3Aln plus the family
grouping, indel maps and haplotype links for one seeded synthetic duplex-UMI tile, so that parity can be checked where
the oracle binary is absent."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from uvc_b200 import synth  # noqa: E402
import parity_util as pu  # noqa: E402

GOLDEN_CFG = dict(name="golden_umi", seed=90210, contigs=(("chrG", 3200),), depth=500.0, umi=True, family_mean=3.0,
                  n_snv=5, n_indel=4, vafs=(0.02, 0.1, 0.4), targets=[(0, 900, 2300)])
GOLDEN_TILE = (0, 1000, 2200, 4)
ARRAY_SECTIONS = ["meta", "rtr_initial", "rtr_final", "baq", "baq2", "prep", "thres", "seginfo", "faminfo", "fragdepth0", "fragdepth1",
                  "famdepth0", "famdepth1", "duplex", "vq"]
TEXT_SECTIONS = ["families", "indelmaps", "haplinks"]


def golden_inputs(outdir):
    return synth.generate(synth.SynthConfig(**GOLDEN_CFG), outdir)


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as tmp:
        info = golden_inputs(tmp)
        tid, beg, end, flag = GOLDEN_TILE
        ref = pu.run_oracle_dump(info["bam"], info["fasta"], tid, beg, end, flag, os.path.join(tmp, "g.dump"))
        out = {}
        for s in ARRAY_SECTIONS:
            out[s] = np.frombuffer(ref[s].tobytes(), dtype=np.uint8)
        for s in TEXT_SECTIONS:
            out[s] = np.frombuffer(ref[s].encode(), dtype=np.uint8)
        path = os.path.join(ROOT, "tests", "golden", "tile_umi.npz")
        np.savez_compressed(path, **out)
        print("wrote %s (%d bytes), %d reads" % (path, os.path.getsize(path), info["n_reads"]))
