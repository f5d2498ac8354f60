"""CPU-side checks: the C-ABI library loads and exports every symbol include/uvcgpu.h declares (no compute without a GPU), the
struct mirrors agree, the product refuses to run without a CUDA device, the BAM substrate round-trips the generator's records,
and the reference's own compile-time known-answer checks (main_conversion.hpp:205-209, 251-254) hold for our restatement."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

from uvc_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_libuvcgpu_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "uvcgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(uvcgpu_[a-z_]+)\s*\(", hdr))
    assert {"uvcgpu_create", "uvcgpu_submit", "uvcgpu_collect", "uvcgpu_dump_counters", "uvcgpu_destroy", "uvcgpu_set_contig"} <= names
    lib = C.CDLL(os.path.join(ROOT, "uvc_b200", "lib", "libuvcgpu.so"))
    for n in sorted(names):
        assert hasattr(lib, n), "libuvcgpu.so does not export " + n


def test_struct_mirrors_and_record_sizes():
    lib = capi.load_gpu()          # raises if the mirror of uvcgpu_params disagrees with the compiled library
    assert C.sizeof(capi.Params) == lib.uvcgpu_sizeof_params()
    # record sizes of the reference layout (SURVEY.md probe: 208, 72, 152, 72, 28 bytes)
    assert capi.struct_dtype("uvcgpu_prep_set").itemsize == 208
    assert capi.struct_dtype("uvcgpu_thres_set").itemsize == 72
    assert capi.struct_dtype("uvcgpu_seginfo_set").itemsize == 152
    assert capi.struct_dtype("uvcgpu_faminfo_set").itemsize == 72
    assert capi.struct_dtype("uvcgpu_rtr").itemsize == 28


def test_product_library_has_no_cpu_fallback():
    lib = capi.load_gpu()
    if lib.uvcgpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.UvcGpuError):
        capi.Context(0)


def test_bam_substrate_roundtrip(tmp_path):
    cfg = synth.SynthConfig(name="rt", seed=5, contigs=(("c0", 4000), ("c1", 3000)), depth=30.0, n_snv=2, n_indel=2)
    info = synth.generate(cfg, str(tmp_path))
    bf = capi.BamFile(info["bam"])
    assert bf.targets == [("c0", 4000), ("c1", 3000)]
    rb = capi.ReadBuf()
    n0 = bf.fetch_into(rb, 0, 0, 4000)
    n1 = bf.fetch_into(rb, 1, 0, 3000)
    assert n0 + n1 == info["n_reads"]
    v = rb.view()
    pos = np.ctypeslib.as_array(C.cast(v.pos, C.POINTER(C.c_int32)), (len(rb),)).copy()
    assert (np.diff(pos[:n0]) >= 0).all() and (np.diff(pos[n0:]) >= 0).all()
    # windowed query = records with pos < end and endpos > beg
    rb2 = capi.ReadBuf()
    k = bf.fetch_into(rb2, 0, 1000, 1200)
    lq = np.ctypeslib.as_array(C.cast(v.l_qseq, C.POINTER(C.c_int32)), (len(rb),))[:n0]
    approx = int(((pos[:n0] < 1200) & (pos[:n0] + 200 > 1000)).sum())
    assert 0 < k <= approx
    assert capi.read_fasta_contig(info["fasta"], "c1")[:10] == open(info["fasta"]).read().split(">c1\n")[1].replace("\n", "")[:10].encode()
    rb.close(); rb2.close(); bf.close()


def _binom_10log10_likeratio(prob, a, b):
    """calc_binom_10log10_likeratio<false,false> restated (main_conversion.hpp:222-237)."""
    eps = np.finfo(np.float64).eps
    prob = (prob + eps) / (1.0 + 2.0 * eps)
    a += eps
    b += eps
    A = prob * (a + b)
    B = (1.0 - prob) * (a + b)
    if a > A:
        return 10.0 / math.log(10.0) * (a * math.log(a / A) + b * math.log(b / B))
    return 0.0


def test_reference_static_asserts():
    assert abs(_binom_10log10_likeratio(0.1, 10, 90)) < 1e-4
    assert 763 < _binom_10log10_likeratio(0.1, 90, 10) < 764
    assert abs(_binom_10log10_likeratio(0.1, 1, 99)) < 1e-4
    prob2odds = lambda p: p / (1.0 - p)       # noqa: E731
    odds2prob = lambda o: o / (o + 1.0)       # noqa: E731
    assert 0.99 < prob2odds(odds2prob(1)) < 1.01
    assert 0.65 < odds2prob(prob2odds(0.66)) < 0.67


def _device_math(emulate):
    """The reference's compile-time known-answer checks (main_conversion.hpp:205-209, 251-254) evaluated by the library's own scoring code."""
    ctx = capi.Context(0, emulate=emulate)
    ctx.lib.uvcgpu_selftest_math.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.c_int32, C.POINTER(C.c_double)]

    def ev(which, pts):
        flat = (C.c_double * (3 * len(pts)))(*[x for p in pts for x in p])
        out = (C.c_double * len(pts))()
        assert ctx.lib.uvcgpu_selftest_math(ctx.handle, which, flat, len(pts), out) == 0
        return list(out)
    lr = ev(0, [(0.1, 10, 90), (0.1, 90, 10), (0.1, 1, 99)])
    assert abs(lr[0]) < 1e-4 and 763 < lr[1] < 764 and abs(lr[2]) < 1e-4
    assert abs(lr[1] - _binom_10log10_likeratio(0.1, 90, 10)) < 1e-9
    assert 0.99 < ev(1, [(1.0, 0, 0)])[0] < 1.01
    assert 0.65 < ev(2, [(0.66, 0, 0)])[0] < 0.67
    assert abs(ev(3, [(0.0, 30.0, 10.0)])[0] - math.log(3.0)) < 1e-9
    ctx.close()


def test_emulated_scoring_math_known_answers():
    _device_math(True)


@pytest.mark.gpu
def test_cuda_scoring_math_known_answers():
    _device_math(False)
