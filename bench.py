#!/usr/bin/env python
"""bench.py - throughput of the pileup-and-score hot path on synthetic input (contract: DESIGN.md section "Measurement").

A "step" is one pass of the hot path over one batch of synthetic input: all tier-3 tiles of the named workload (tile list from the
reference-identical tiler, -t 16). Metric (BASELINE.json): aligned reads/sec (`value`) and genomic positions/sec (`positions_per_s`).

* `value`    - device throughput with the inputs resident in HBM: sum of the CUDA-event times of every kernel of the step (events on the
               library's own stream), one context, whole batch per launch.
* `e2e`      - the same metric through the C ABI from HOST buffers (decoded BAM records in SoA form): host staging, H2D, kernels, D2H and
               VCF text inside the timed wall clock. The batch is cut into sub-batches that run on a few contexts (one CUDA stream each),
               so that staging, copies and kernels of different sub-batches overlap - the same schedule the uvc1 host uses.
* `roofline` - the kernel with the largest share of the step, against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
* `cpu_baseline` / `--impl reference` - the unmodified reference uvc1 (oracle/_ref) with all host threads on the same BAM.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STAGE_NAMES = ["K0 per-read", "K1 prep+thres", "K2 bias pileup", "K2e indel events", "KF fragment columns", "K3a fragment stats", "K3b fragment consensus",
               "KM family columns", "K4a family ends", "K4 family+duplex consensus", "K4c family haplotypes", "K6 block-line inputs", "K5 candidate scoring"]
STAGES_IMPLEMENTED = ("P0 read filter+family grouping (host), P1 repeat context (host), K0..K4c = updateByRegion3Aln, K6 block-line inputs, "
                      "K5a/K5 candidate scoring, VCF text (host) = process_batch")
WORKLOADS = {"c1": "uvc1 tumor-only, synthetic 1 Mbp @100x, non-UMI (BASELINE.json configs[0])",
             "c2": "targeted panel 2 Mbp @2000x non-UMI, low-VAF spikes (BASELINE.json configs[1])",
             "c3": "UMI duplex panel 1 Mbp @20000x (BASELINE.json configs[2])"}
# c2 and c3 are run on a fraction of their region (same depth, same tile shapes): the pure-Python generator needs minutes per million reads,
# and the default run has to finish within minutes on a fresh box. The fraction is part of `config.workload`.
DEFAULT_SCALE = {"c1": 1.0, "c2": 0.05, "c3": 0.002}
TILER_THREADS = 16


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2")
    ap.add_argument("--scale", type=float, default=None, help="fraction of the named config's region (default: per-config)")
    ap.add_argument("--workdir", default=os.environ.get("UVC_BENCH_DIR", "/tmp/uvc_bench"))
    ap.add_argument("--contexts", type=int, default=4, help="contexts (CUDA streams) the e2e step pipelines its sub-batches over")
    ap.add_argument("--sub-batches", type=int, default=8)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    return ap.parse_args()


def dataset(workdir: str, name: str, scale: float):
    """Generates (once) the seeded synthetic BAM/FASTA of a named config; returns the generator's info dict."""
    from uvc_b200 import synth
    cfg = synth.named_config(name, scale)
    key = hashlib.sha1(repr(cfg).encode()).hexdigest()[:12]
    d = os.path.join(workdir, "%s_%s" % (name, key))
    meta = os.path.join(d, "meta.json")
    if os.path.exists(meta):
        return json.load(open(meta))
    os.makedirs(d, exist_ok=True)
    t0 = time.time()
    info = synth.generate(cfg, d)
    out = dict(bam=info["bam"], fasta=info["fasta"], n_reads=info["n_reads"], contigs=[list(c) for c in cfg.contigs],
               targets=cfg.targets, gen_s=time.time() - t0, bed=info.get("bed"))
    json.dump(out, open(meta, "w"))
    return out


class _BedLine(C.Structure):
    _fields_ = [("tid", C.c_int32), ("beg_pos", C.c_int32), ("end_pos", C.c_int32), ("region_flag", C.c_uint32), ("n_reads", C.c_int64)]


def tile_list(ds, nthreads: int = TILER_THREADS):
    """Tier-3 tiles exactly as the uvc1 host cuts them (uvc_b200/csrc/host/tiler.cpp = the reference's SamIter for -t nthreads)."""
    from uvc_b200 import capi
    lib = capi.load_host()
    lib.uvchost_tiler_open.restype = C.c_void_p
    lib.uvchost_tiler_open.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32]
    lib.uvchost_tiler_next.restype = C.c_int64
    lib.uvchost_tiler_next.argtypes = [C.c_void_p, C.POINTER(C.POINTER(_BedLine)), C.POINTER(C.c_int64)]
    lib.uvchost_tiler_close.argtypes = [C.c_void_p]
    t = lib.uvchost_tiler_open(ds["bam"].encode(), (ds.get("bed") or "").encode(), b"", nthreads, 1536, -1, 0)
    tiles = []
    while True:
        p, n = C.POINTER(_BedLine)(), C.c_int64()
        nreads = lib.uvchost_tiler_next(t, C.byref(p), C.byref(n))
        if nreads < 0:
            raise RuntimeError("tiler failed")
        if nreads == 0 and n.value == 0:
            break
        tiles += [(p[i].tid, p[i].beg_pos, p[i].end_pos, p[i].region_flag) for i in range(n.value)]
    lib.uvchost_tiler_close(t)
    return tiles


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.sm_max = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.sm_max = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(s)}


def run_reference(args, ds, name):
    """Times the unmodified reference uvc1 (oracle/_ref) with all host threads on the bench workload's own BAM (bounded: the c2/c3 workloads
    are already fractions of the named configs; c1 takes a few seconds)."""
    uvc1 = os.path.join(ROOT, "oracle", "_ref", "uvc1")
    cores = os.cpu_count() or 1
    best = None
    runs = max(1, min(args.steps, 2))
    for _ in range(runs):
        out_vcf = os.path.join(args.workdir, "ref_%s.vcf.gz" % name)
        cmd = [uvc1, ds["bam"], "-f", ds["fasta"], "-o", out_vcf, "-s", "S", "-t", str(cores)]
        if ds.get("bed"):
            cmd += ["-R", ds["bed"]]
        t0 = time.time()
        p = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        wall = time.time() - t0
        if p.returncode != 0:
            raise RuntimeError("reference uvc1 failed: " + p.stderr[-500:])
        m = re.search(r"Wall clock time passed: ([0-9.]+) seconds", p.stderr)
        wall_ref = float(m.group(1)) if m else wall
        best = wall_ref if best is None else min(best, wall_ref)
    npos = sum(l for _, l in ds["contigs"]) if not ds.get("targets") else sum(e - b for _, b, e in ds["targets"])
    return dict(value=ds["n_reads"] / best, unit="reads/s", cores=cores, kind="reference",
                sample="the bench workload itself: %d reads, %d positions; uvc1 -t %d (BAM decode, tiling and BGZF output included), best of %d, %.2f s" % (
                    ds["n_reads"], npos, cores, runs, best),
                positions_per_s=npos / best, seconds=best)


def kernel_algorithmic_bytes(stage: int, st, n_reads: int) -> float:
    """Compulsory bytes of one launch of a kernel: its inputs read once plus its outputs written once."""
    P, R = float(st.n_ext_positions), float(n_reads)
    read_rec = 1.5 * 150 + 64                      # packed bases + qualities + cigar + scalars of one read
    table = {
        1: R * read_rec + P * (208 + 72),                                   # K1: reads -> prep + thres
        2: R * read_rec + P * (72 + 14 * 152 + 14 * 4 + 14 * 4 * 4),        # K2: reads + thres -> seginfo, bqsum, 4 VQ tags
        4: R * read_rec + R * 150 * 8 / 2,                                  # KF: reads -> 8 B column entry per fragment base (2 reads per fragment)
        5: R * 150 * 8 / 2,                                                 # K3a: fragment columns
        6: R * 150 * 8 / 2 + P * (2 * 14 * 3 * 4 + 14 * 4 * 4),             # K3b: fragment columns -> fragdepth + 4 VQ tags
        7: R * 150 * (8 + 32) / 2,                                          # KM: fragment columns -> family columns
        9: R * 150 * 32 / 2 + P * (72 + 2 * 14 * 8 * 4 + 14 * 72 + 14 * 2 * 4 + 14 * 6 * 4),   # K4: family columns + thres -> famdepth, faminfo, duplex, 6 VQ tags
        10: R * 150 * 32 / 2,                                               # K4c
        11: P * 6272,                                                       # K6 reads the depth arrays of every position
        12: P * 6272,                                                       # K5a/K5 read the position state once
    }
    return table.get(stage, P * 6272)


def main():
    args = parse_args()
    # stdout carries exactly one JSON line: libraries that print to file descriptor 1 (NCCL's version banner, ...) are sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    json_out = os.fdopen(json_fd, "w")
    name = args.config
    scale = args.scale if args.scale is not None else DEFAULT_SCALE.get(name, 1.0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = "%s; region fraction %g" % (WORKLOADS.get(name, name), scale)
    metric = "aligned reads/sec (and positions/sec) called"

    if args.impl == "reference":
        if rank != 0:
            return
        ds = dataset(args.workdir, name, scale)
        res = run_reference(args, ds, name)
        line = {"impl": "reference", "metric": metric, "value": res["value"], "unit": "reads/s",
                "positions_per_s": res["positions_per_s"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": res["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/int64 counters, f64 scoring",
                "data": "synthetic", "config": {"workload": workload, "reference_sample": res["sample"]},
                "cpu_baseline": {"value": res["value"], "unit": "reads/s", "cores": res["cores"], "kind": "reference", "sample": res["sample"]},
                "e2e": {"value": res["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
        return

    import torch
    import torch.distributed as dist
    from uvc_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl")

    ds = dataset(args.workdir, name, scale) if rank == 0 else None
    if world > 1:
        dist.barrier()
        if rank != 0:
            ds = dataset(args.workdir, name, scale)
    # Regions are independent: no data-path collective. Weak scaling: every rank (GPU) processes its own copy of the per-GPU workload.
    tiles = tile_list(ds)
    host_threads = max(1, (os.cpu_count() or 1) // world)

    # host-side decode (untimed): BAM -> SoA records of every tile's fetch window (what sam_itr_queryi(tid, beg - 2000, end + 2000) yields)
    bf = capi.BamFile(ds["bam"])
    rb = capi.ReadBuf()
    ctiles = []
    prev = (-1, 0, 0)
    t_dec0 = time.time()
    for (tid, beg, end, flag) in tiles:
        r0 = len(rb)
        bf.fetch_into(rb, tid, max(0, beg - 2000), end + 2000)
        ctiles.append(capi.make_tile(tid, beg, end, flag, ds["contigs"][tid][1], r0, len(rb), prev))
        prev = (tid, beg, end)
    decode_s = time.time() - t_dec0
    view = rb.view()
    contig_bases = {tid: capi.read_fasta_contig(ds["fasta"], cname) for tid, (cname, _) in enumerate(ds["contigs"])}

    def make_ctx(threads):
        ctx = capi.Context(local_rank)
        ctx.lib.uvcgpu_set_host_threads.argtypes = [C.c_void_p, C.c_int32]
        ctx.lib.uvcgpu_set_host_threads(ctx.handle, threads)
        for tid, (cname, _) in enumerate(ds["contigs"]):
            ctx.set_contig(tid, contig_bases[tid])
            ctx.set_contig_name(tid, cname)
        return ctx

    def submit_tiles(ctx, sub):
        t0 = time.time()
        ticket = ctx.submit(sub, view)       # host staging (P0/P1), H2D copies and every pileup kernel enqueued on the context's stream
        return (ticket, sub, t0, time.time())

    def finish_tiles(ctx, pending):
        ticket, sub, t0, t1 = pending
        tw0 = time.time()
        ctx.collect(ticket)
        t2 = time.time()
        st = ctx.score(ticket)               # candidate scoring on the device + D2H of the kept records and block-line inputs
        t3 = time.time()
        nbytes = 0
        for ti in range(len(sub)):           # the step's result: every tile's VCF body text
            nbytes += len(ctx.tile_vcf(ticket, ti))
        t4 = time.time()
        ctx.release(ticket)
        st.vcf_bytes = nbytes
        st.phase_s = (t1 - t0, t2 - tw0, t3 - t2, t4 - t3, time.time() - t4)   # submit (staging + H2D enqueue), wait, score, text, release
        return st

    def run_tiles(ctx, sub):
        return finish_tiles(ctx, submit_tiles(ctx, sub))

    # ---- phase 1: device-resident throughput (`value`): one context, the whole batch per launch, CUDA-event time of every kernel
    ctx0 = make_ctx(host_threads)
    n_warm = max(args.warmup, 3)
    for _ in range(n_warm):
        run_tiles(ctx0, ctiles)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    kernel_ms = 0.0
    stage_ms = [0.0] * 16
    launches = 0
    last = None
    for _ in range(args.steps):
        last = run_tiles(ctx0, ctiles)
        kernel_ms += last.kernel_ms
        launches += int(last.gpu_launches)
        for i in range(16):
            stage_ms[i] += last.kernel_ms_by_stage[i]
    torch.cuda.synchronize()

    # ---- phase 2: end to end from host buffers: sub-batches pipelined over a few contexts (streams)
    n_ctx = max(1, min(args.contexts, len(ctiles)))
    n_sub = max(n_ctx, min(args.sub_batches, len(ctiles)))
    subs = [ctiles[len(ctiles) * k // n_sub: len(ctiles) * (k + 1) // n_sub] for k in range(n_sub)]
    subs = [s for s in subs if s]
    ctxs = [ctx0] + [make_ctx(max(1, host_threads // n_ctx + 1)) for _ in range(n_ctx - 1)]
    ctx0.lib.uvcgpu_set_host_threads(ctx0.handle, max(1, host_threads // n_ctx + 1))
    totals = {"h2d": 0, "d2h": 0, "vcf": 0, "rec": 0, "launch": 0, "prep_ms": 0.0, "submit_s": 0.0, "wait_s": 0.0, "score_s": 0.0, "text_s": 0.0, "release_s": 0.0}

    def e2e_steps(n_steps):
        # the n_steps passes over the workload are one continuous stream of sub-batches (no drain between steps), as in a long run of the uvc1 host
        work_items = [sub for _ in range(n_steps) for sub in subs]
        nxt = [0]
        lock = threading.Lock()
        acc = {k: 0 for k in totals}
        errs = []

        def work(ctx):
            try:
                # two sub-batches in flight per context: the next one is staged and enqueued before the previous one is collected, so the
                # stream always has work queued while the host stages
                pending = None
                while True:
                    with lock:
                        k = nxt[0]
                        nxt[0] += 1
                    nxt_pending = submit_tiles(ctx, work_items[k]) if k < len(work_items) else None
                    if pending is None and nxt_pending is None:
                        return
                    if pending is None:
                        pending = nxt_pending
                        continue
                    st = finish_tiles(ctx, pending)
                    pending = nxt_pending
                    with lock:
                        acc["h2d"] += int(st.h2d_bytes)
                        acc["d2h"] += int(st.d2h_bytes)
                        acc["vcf"] += int(st.vcf_bytes)
                        acc["rec"] += int(st.n_vcf_records)
                        acc["launch"] += int(st.gpu_launches)
                        acc["prep_ms"] += st.host_prep_ms
                        for key, val in zip(("submit_s", "wait_s", "score_s", "text_s", "release_s"), st.phase_s):
                            acc[key] += val
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        ths = [threading.Thread(target=work, args=(c,)) for c in ctxs]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if errs:
            raise errs[0]
        return acc

    e2e_steps(n_warm)
    # steady state: the library page-locks its staging blocks in the background; warm up until none is outstanding (bounded)
    # (page-locking while batches are in flight also slows every CUDA call down, so the timed region must not trigger any: warm up until the
    # cache has stopped growing for two rounds in a row)
    ctx0.lib.uvcgpu_staging_backlog.restype = C.c_int
    ctx0.lib.uvcgpu_staging_pinned_bytes.restype = C.c_int64
    stable = 0
    for _ in range(16):
        t_wait = time.time()
        while ctx0.lib.uvcgpu_staging_backlog() > 0 and time.time() - t_wait < 10.0:
            time.sleep(0.05)
        before = int(ctx0.lib.uvcgpu_staging_pinned_bytes())
        e2e_steps(2)
        grown = (ctx0.lib.uvcgpu_staging_backlog() > 0 or int(ctx0.lib.uvcgpu_staging_pinned_bytes()) != before)
        stable = 0 if grown else stable + 1
        if stable >= 2:
            break
    backlog0 = int(ctx0.lib.uvcgpu_staging_backlog())
    pinned0 = int(ctx0.lib.uvcgpu_staging_pinned_bytes())
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    acc = e2e_steps(args.steps)
    for k in totals:
        totals[k] += acc[k]
    torch.cuda.synchronize()
    wall_s = time.time() - t0
    backlog1 = int(ctx0.lib.uvcgpu_staging_backlog())
    sampler.stop_flag = True
    if world > 1:
        tt = torch.tensor([kernel_ms, wall_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        kernel_ms, wall_s = float(tt[0]), float(tt[1])
        dist.barrier()
    for c in ctxs:                           # explicit teardown (contexts own CUDA streams and pool memory)
        c.close()
    rb.close()
    bf.close()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    n_reads = int(ds["n_reads"])             # primary mapped records of the BAM, each counted once, not once per overlapping tile
    n_positions = int(last.n_positions)
    units = world
    value = units * n_reads * args.steps / (kernel_ms / 1e3)
    e2e = units * n_reads * args.steps / wall_s
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dom = max(range(16), key=lambda i: stage_ms[i])
    dom_name = STAGE_NAMES[dom] if dom < len(STAGE_NAMES) else "stage%d" % dom
    dom_ms = stage_ms[dom] / args.steps
    dom_bytes = kernel_algorithmic_bytes(dom, last, int(last.n_reads_kept))
    achieved = dom_bytes / (dom_ms / 1e3) / 1e9
    step_ms = kernel_ms / args.steps
    path_bytes = last.n_reads_kept * (1.5 * 150 + 64) + last.n_ext_positions * 2 * 6272
    traffic = None
    issue = None
    try:   # dram bytes of that kernel from the committed ncu --set full capture of the same workload (profiles/), per launch
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        ent = tj.get("%s@%g" % (name, scale), {}).get(dom_name)
        if ent:
            traffic = ent["dram_bytes"]
            if ent.get("ipc_per_sm"):
                # the kernel is bound by instruction issue, not by bytes: the same capture's issue rate against the 4 warp
                # instructions per cycle an SM can issue
                issue = {"bound": "issue", "achieved": ent["ipc_per_sm"], "peak": 4.0, "unit": "warp instructions/cycle/SM", "frac": ent["ipc_per_sm"] / 4.0,
                         "warp_instructions": ent.get("warp_instructions"), "resident_warps_pct": ent.get("resident_warps_pct"), "source": "profiles/r01_traffic.json (ncu --set full)"}
    except Exception:
        pass
    line = {"metric": metric, "value": value, "unit": "reads/s",
            "positions_per_s": units * n_positions * args.steps / (kernel_ms / 1e3),
            "n_gpus": args.gpus, "steps": args.steps, "warmup": n_warm, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/int64 counters, f64 scoring",
            "data": "synthetic",
            "config": {"workload": workload, "tiles": len(ctiles), "tiler": "reference SamIter semantics, -t %d" % TILER_THREADS,
                       "reads_per_step": n_reads, "read_records_per_step_incl_tile_halos": int(last.n_reads_kept),
                       "positions_per_step": n_positions, "ext_positions_per_step": int(last.n_ext_positions), "stages": STAGES_IMPLEMENTED,
                       "l2": "per-position state of a step (%.0f MB) is larger than L2, no flush needed" % (last.n_ext_positions * 6272 / 1e6),
                       "host_decode_s_untimed": decode_s, "dataset_generation_s_untimed": ds.get("gen_s"),
                       "e2e_schedule": "%d sub-batches per step over %d contexts (streams), two in flight per context, steps back to back, %d host threads" % (len(subs), len(ctxs), host_threads)},
            "e2e": {"value": e2e, "unit": "reads/s", "positions_per_s": units * n_positions * args.steps / wall_s,
                    "h2d_bytes_per_step": totals["h2d"] // args.steps, "d2h_bytes_per_step": totals["d2h"] // args.steps,
                    "vcf_bytes_per_step": totals["vcf"] // args.steps, "vcf_records_per_step": totals["rec"] // args.steps,
                    "host_prep_ms_per_step_summed_over_contexts": totals["prep_ms"] / args.steps,
                    "call_ms_per_step_summed_over_contexts": {k[:-2]: totals[k] * 1e3 / args.steps for k in ("submit_s", "wait_s", "score_s", "text_s", "release_s")},
                    "wall_ms_per_step": wall_s * 1e3 / args.steps,
                    "staging_blocks_not_yet_page_locked": {"at_start": backlog0, "at_end": backlog1},
                    "staging_page_locked_bytes": {"at_start": pinned0, "at_end": int(ctx0.lib.uvcgpu_staging_pinned_bytes())}},
            "gpu_launches": launches + totals["launch"],
            "stage_ms_per_step": {n: stage_ms[i] / args.steps for i, n in enumerate(STAGE_NAMES)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": dom_name, "kernel_ms": dom_ms, "kernel_share_of_step": dom_ms / step_ms, "algorithmic_bytes": dom_bytes,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "whole_path": {"algorithmic_bytes": path_bytes, "achieved": path_bytes / (step_ms / 1e3) / 1e9,
                                        "frac": path_bytes / (step_ms / 1e3) / 1e9 / peak},
                         "issue": issue},
            "clocks": sampler.summary()}
    if world == 1 and not args.skip_cpu_baseline and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "uvc1")):   # (reported at N = 1 only)
        try:
            res = run_reference(args, ds, name)
            line["cpu_baseline"] = {"value": res["value"], "unit": "reads/s", "cores": res["cores"], "kind": "reference", "sample": res["sample"],
                                    "positions_per_s": res["positions_per_s"]}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % e}
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()


if __name__ == "__main__":
    main()
