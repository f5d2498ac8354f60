import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Native components are built in-tree once per session if anything is missing."""
    need = [os.path.join(ROOT, "uvc_b200", "lib", "libuvcgpu.so"), os.path.join(ROOT, "uvc_b200", "lib", "libuvchost.so"),
            os.path.join(ROOT, "tests", "emu", "libuvcgpu_emu.so")]
    if os.path.isdir("/root/reference"):
        need += [os.path.join(ROOT, "oracle", "_ref", "uvc1"), os.path.join(ROOT, "oracle", "_ref", "uvc_ref_dump")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()
    yield


@pytest.fixture(scope="session")
def synth_small(tmp_path_factory):
    """c1-like data: 12 kbp at 80x with spiked SNVs/indels."""
    from uvc_b200 import synth
    d = tmp_path_factory.mktemp("synth_small")
    cfg = synth.SynthConfig(name="small", seed=4242, contigs=(("chrA", 12000),), depth=80.0, n_snv=8, n_indel=6)
    return synth.generate(cfg, str(d))


@pytest.fixture(scope="session")
def synth_umi(tmp_path_factory):
    """c3-like data: 3 kbp at 3000x raw depth with duplex UMIs."""
    from uvc_b200 import synth
    d = tmp_path_factory.mktemp("synth_umi")
    cfg = synth.SynthConfig(name="umi", seed=4343, contigs=(("chrU", 5000),), depth=3000.0, umi=True, n_snv=4, n_indel=3,
                            vafs=(0.01, 0.05), targets=[(0, 1000, 4000)])
    return synth.generate(cfg, str(d))


@pytest.fixture(scope="session")
def synth_c2_depth(tmp_path_factory):
    """BASELINE configs[1] shape at its real depth: 200 bp targets every 1000 bp at 2000x, half of them amplicon-like (shared fragment ends)."""
    from uvc_b200 import synth
    d = tmp_path_factory.mktemp("synth_c2_depth")
    targets = [(0, 1000 + 1000 * i, 1200 + 1000 * i) for i in range(6)]
    cfg = synth.SynthConfig(name="c2d", seed=1202, contigs=(("chrP", 8000),), depth=2000.0, targets=targets, amplicon_frac=0.5, n_snv=5, n_indel=3,
                            vafs=(0.005, 0.01, 0.02, 0.05))
    info = synth.generate(cfg, str(d))
    info["targets"] = targets
    return info


@pytest.fixture(scope="session")
def synth_c3_depth(tmp_path_factory):
    """BASELINE configs[2] shape at its real depth: 20 000x raw depth, duplex UMIs, both the A+B and the B+A flavour of the bottom strand."""
    from uvc_b200 import synth
    d = tmp_path_factory.mktemp("synth_c3_depth")
    cfg = synth.SynthConfig(name="c3d", seed=1303, contigs=(("chrD", 3000),), depth=20000.0, umi=True, n_snv=3, n_indel=2, vafs=(0.001, 0.005, 0.01),
                            targets=[(0, 1000, 2000)], swapped_umi_frac=0.5)
    return synth.generate(cfg, str(d))
