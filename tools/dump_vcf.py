"""Debug helper: writes the VCF body our C ABI returns for the golden duplex-UMI data set (all-out) so that the CUDA
build's output (GPU box) can be diffed with the emulation build's (container). Usage: dump_vcf.py <out> [emu] [beg end]"""
import importlib.util, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as pu
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
out = sys.argv[1]; emu = (len(sys.argv) > 2 and sys.argv[2] == "emu")
beg, end = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1000, 2200)
with tempfile.TemporaryDirectory() as tmp:
    info = mg.golden_inputs(tmp)
    res, st = pu.run_tiles(info["bam"], info["fasta"], [(0, beg, end, 0)], emu, ["vcf"], should_output_all=1)
    open(out, "w").write(res[0]["vcf"])
    print("wrote", out, len(res[0]["vcf"]), "bytes; launches", st.gpu_launches)
