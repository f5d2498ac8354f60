#!/usr/bin/env python
"""Times the host staging stage (P0/P1, uvc_build_host_batch) alone on the bench workload, with the emulation library and its kernels skipped.
usage: UVC_PREP_PROFILE=1 python tools/prep_profile.py [config] [scale] [threads] [repeats]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["UVC_EMU_PREP_ONLY"] = "1"
import ctypes as C
import bench
from uvc_b200 import capi

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
threads = int(sys.argv[3]) if len(sys.argv) > 3 else os.cpu_count()
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ds = bench.dataset("/tmp/uvc_bench", name, scale)
tiles = bench.tile_list(ds)
bf = capi.BamFile(ds["bam"])
rb = capi.ReadBuf()
ctiles, prev = [], (-1, 0, 0)
for (tid, beg, end, flag) in tiles:
    r0 = len(rb)
    bf.fetch_into(rb, tid, max(0, beg - 2000), end + 2000)
    ctiles.append(capi.make_tile(tid, beg, end, flag, ds["contigs"][tid][1], r0, len(rb), prev))
    prev = (tid, beg, end)
view = rb.view()
ctx = capi.Context(0, emulate=True)
ctx.lib.uvcgpu_set_host_threads.argtypes = [C.c_void_p, C.c_int32]
ctx.lib.uvcgpu_set_host_threads(ctx.handle, threads)
for tid, (cname, _) in enumerate(ds["contigs"]):
    ctx.set_contig(tid, capi.read_fasta_contig(ds["fasta"], cname))
    ctx.set_contig_name(tid, cname)
for _ in range(reps):
    t0 = time.time()
    ticket = ctx.submit(ctiles, view)
    st = ctx.collect(ticket)
    t1 = time.time()
    ctx.release(ticket)
    print("tiles %d reads %d: submit %.1f ms (host_prep %.1f ms, upload/alloc %.1f ms), %d threads" % (len(ctiles), st.n_reads_kept, (t1 - t0) * 1e3, st.host_prep_ms, st.h2d_ms, threads))
