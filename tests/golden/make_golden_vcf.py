"""Generates the committed golden VCF fixtures from the ORACLE (the unmodified reference binary oracle/_ref/uvc1). Run in the build
container where /root/reference exists:

    python tests/golden/make_golden_vcf.py

Two fixtures on the seeded duplex-UMI data set of make_golden.py: the default output of a 1200 bp region, and the base-pair resolution
output (--all-out) of a 300 bp region. They pin the scoring stage (FORMAT/INFO/QUAL/FILTER of every record, MGVCF block lines,
additional-indel-candidate lines) byte for byte where the oracle binary is absent."""
import gzip
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import importlib.util  # noqa: E402

import parity_util as pu  # noqa: E402

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)

# (fixture file, tile (tid, beg, end, region_flag), reference command-line options, uvcgpu_params overrides)
CASES = [
    ("umi_default.vcf.gz", (0, 1000, 2200, 0), [], {}),
    ("umi_allout.vcf.gz", (0, 1500, 1800, 0), ["-A"], {"should_output_all": 1}),
]


def reference_vcf_lines(bam, fasta, contig_name, tile, opts, workdir):
    """Body lines of the reference's VCF for one BED region."""
    bed = os.path.join(workdir, "region.bed")
    with open(bed, "w") as f:
        f.write("%s\t%d\t%d\n" % (contig_name, tile[1], tile[2]))
    out = os.path.join(workdir, "ref.vcf.gz")
    subprocess.run([pu.REF_UVC1, "-f", fasta, "-o", out, "-s", "S", "-t", "1", "-R", bed, bam] + list(opts), check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    with gzip.open(out, "rt") as f:
        return [l for l in f.read().split("\n") if l and not l.startswith("#")]


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as tmp:
        info = mg.golden_inputs(tmp)
        for fname, tile, opts, _ in CASES:
            lines = reference_vcf_lines(info["bam"], info["fasta"], mg.GOLDEN_CFG["contigs"][0][0], tile, opts, tmp)
            path = os.path.join(ROOT, "tests", "golden", fname)
            with gzip.GzipFile(path, "wb", mtime=0) as f:
                f.write(("\n".join(lines) + "\n").encode())
            print("wrote %s: %d lines, %d bytes" % (path, len(lines), os.path.getsize(path)))
