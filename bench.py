#!/usr/bin/env python
"""bench.py - throughput of the pileup-and-score hot path on synthetic input (contract: see DESIGN.md section "Measurement").

A "step" is one pass of the hot path over one batch of synthetic tiles. Metric (BASELINE.json): aligned reads/sec
(`value`) and genomic positions/sec (`positions_per_s`). `value` is kernel throughput with inputs resident in HBM (CUDA
events on the library's stream); `e2e` is the same metric through the C ABI from HOST buffers (host staging + H2D + kernels
+ D2H of the per-batch result inside the timed region). `--impl reference` times the unmodified reference `uvc1` (built
under oracle/_ref) on the box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
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

STAGES_IMPLEMENTED = "P0 grouping(host)+P1 refctx(host)+K0 read consts+K1 prep/thres+K2 bias pileup+K2e indel events+KF fragment columns+K3 fragment consensus+KM family columns+K4 family/duplex consensus (= updateByRegion3Aln)+K6 block-line inputs+K5 candidate scoring+VCF text(host) = process_batch"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2")
    ap.add_argument("--scale", type=float, default=None, help="fraction of the named config's region (default: per-config)")
    ap.add_argument("--tile", type=int, default=20000, help="tier-3 tile length used until the reference tiler is ported")
    ap.add_argument("--workdir", default=os.environ.get("UVC_BENCH_DIR", "/tmp/uvc_bench"))
    ap.add_argument("--ref-scale", type=float, default=None, help="sample of the workload for the CPU reference")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    return ap.parse_args()


DEFAULT_SCALE = {"c1": 1.0, "c2": 0.05, "c3": 0.02}
DEFAULT_REF_SCALE = {"c1": 0.1, "c2": 0.003, "c3": 0.001}


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


def make_tiles(ds, tile_len: int):
    tiles = []
    for tid, (name, length) in enumerate(ds["contigs"]):
        if ds.get("targets"):
            # panel: contiguous runs of targets, cut every tile_len
            beg = None
            last = None
            for (ci, b, e) in ds["targets"]:
                if ci != tid:
                    continue
                if beg is None:
                    beg, last = max(0, b - 200), e + 200
                elif b - 200 - last > 200 or (e + 200 - beg) > tile_len:
                    tiles.append((tid, beg, min(last, length), 8))
                    beg, last = b - 200, e + 200
                else:
                    last = e + 200
            if beg is not None:
                tiles.append((tid, beg, min(last, length), 2))
        else:
            for b in range(0, length, tile_len):
                tiles.append((tid, b, min(length, b + tile_len), 4))
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
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(s)}


def run_reference(args, name, ref_scale):
    """Times the unmodified reference uvc1 (oracle/_ref) with all host threads on a bounded sample of the workload."""
    uvc1 = os.path.join(ROOT, "oracle", "_ref", "uvc1")
    ds = dataset(args.workdir, name, ref_scale)
    cores = os.cpu_count() or 1
    best = None
    runs = max(1, min(args.steps, 3))
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
                sample="%s at scale %g: %d reads, %d positions, uvc1 -t %d, best of %d, %.2f s" % (name, ref_scale, ds["n_reads"], npos, cores, runs, best),
                positions_per_s=npos / best, seconds=best)


def main():
    args = parse_args()
    name = args.config
    scale = args.scale if args.scale is not None else DEFAULT_SCALE.get(name, 1.0)
    ref_scale = args.ref_scale if args.ref_scale is not None else DEFAULT_REF_SCALE.get(name, 0.01)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = "%s scale %g (BASELINE.json configs: %s)" % (name, scale, {"c1": "1 Mbp @100x non-UMI", "c2": "targeted panel 2 Mbp @2000x", "c3": "UMI duplex 1 Mbp @20000x"}.get(name, name))

    if args.impl == "reference":
        if rank != 0:
            return
        res = run_reference(args, name, ref_scale)
        line = {"impl": "reference", "metric": "aligned reads/sec (and positions/sec) called", "value": res["value"], "unit": "reads/s",
                "positions_per_s": res["positions_per_s"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": res["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/int64 counters, f64 scoring",
                "data": "synthetic", "config": {"workload": workload, "reference_sample": res["sample"]},
                "cpu_baseline": {"value": res["value"], "unit": "reads/s", "cores": res["cores"], "kind": "reference", "sample": res["sample"]},
                "e2e": {"value": res["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
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
    tiles = make_tiles(ds, args.tile)
    my_tiles = tiles[rank::world] if world > 1 else tiles   # regions are independent: weak scaling = every rank gets its own copy below
    if world > 1:
        my_tiles = tiles  # weak scaling: each rank processes the whole per-GPU workload (independent tiles, no collective on the path)

    # host-side decode (untimed): BAM -> SoA records of every tile's fetch window
    bf = capi.BamFile(ds["bam"])
    rb = capi.ReadBuf()
    ctx = capi.Context(local_rank)
    for tid, (cname, _) in enumerate(ds["contigs"]):
        ctx.set_contig(tid, capi.read_fasta_contig(ds["fasta"], cname))
        ctx.set_contig_name(tid, cname)
    ctiles = []
    prev = (-1, 0, 0)
    t_dec0 = time.time()
    for (tid, beg, end, flag) in my_tiles:
        r0 = len(rb)
        bf.fetch_into(rb, tid, max(0, beg - 2000), end + 2000)
        ctiles.append(capi.make_tile(tid, beg, end, flag, ds["contigs"][tid][1], r0, len(rb), prev))
        prev = (tid, beg, end)
    decode_s = time.time() - t_dec0
    view = rb.view()

    def step():
        ticket = ctx.submit(ctiles, view)
        ctx.collect(ticket)
        st = ctx.score(ticket)               # candidate scoring on the device (K5/K6) + D2H of the kept records
        nbytes = 0
        for ti in range(len(ctiles)):        # the step's result: every tile's VCF body text
            nbytes += len(ctx.tile_vcf(ticket, ti))
        ctx.release(ticket)
        st.vcf_bytes = nbytes
        return st

    sampler = ClockSampler(local_rank)
    for _ in range(max(args.warmup, 3)):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    kernel_ms = 0.0
    stage_ms = [0.0] * 16
    t0 = time.time()
    last = None
    for _ in range(args.steps):
        last = step()
        kernel_ms += last.kernel_ms
        for i in range(16):
            stage_ms[i] += last.kernel_ms_by_stage[i]
    torch.cuda.synchronize()
    wall_s = time.time() - t0
    sampler.stop_flag = True
    if world > 1:
        tt = torch.tensor([kernel_ms, wall_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        kernel_ms, wall_s = float(tt[0]), float(tt[1])
        dist.barrier()
    if rank != 0:
        return
    n_reads = last.n_reads_kept
    n_positions = last.n_positions
    units = world
    value = units * n_reads * args.steps / (kernel_ms / 1e3)
    e2e = units * n_reads * args.steps / wall_s
    # roofline of the dominant kernel: algorithmic bytes of the stages it implements / its event time
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dom = max(range(16), key=lambda i: stage_ms[i])
    STAGE_NAMES = ["K0 per-read", "K1 prep+thres", "K2 bias pileup", "K2e indel events", "KF fragment columns", "K3a fragment stats", "K3b fragment consensus", "KM family columns",
                   "K4a family ends", "K4 family+duplex consensus", "K4c family haplotypes", "K6 block-line inputs", "K5 candidate scoring"]
    dom_name = STAGE_NAMES[dom] if dom < len(STAGE_NAMES) else "stage%d" % dom
    bytes_alg = n_reads * (1.5 * 150 + 64) + last.n_ext_positions * 2 * 6272
    achieved = bytes_alg / (stage_ms[dom] / args.steps / 1e3) / 1e9
    line = {"metric": "aligned reads/sec (and positions/sec) called", "value": value, "unit": "reads/s",
            "positions_per_s": units * n_positions * args.steps / (kernel_ms / 1e3),
            "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": kernel_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/int64 counters",
            "data": "synthetic",
            "config": {"workload": workload, "tiles": len(ctiles), "tile_len": args.tile, "reads_per_step": int(n_reads), "positions_per_step": int(n_positions),
                       "ext_positions_per_step": int(last.n_ext_positions), "stages": STAGES_IMPLEMENTED, "l2": "inputs (%.0f MB counters) larger than L2" % (last.n_ext_positions * 6272 / 1e6),
                       "host_decode_s_untimed": decode_s},
            "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": int(last.h2d_bytes), "d2h_bytes_per_step": int(last.d2h_bytes), "vcf_bytes_per_step": int(last.vcf_bytes), "vcf_records_per_step": int(last.n_vcf_records),
                    "host_score_ms_per_step": last.host_score_ms,
                    "host_prep_ms_per_step": last.host_prep_ms, "wall_ms_per_step": wall_s * 1e3 / args.steps},
            "gpu_launches": int(last.gpu_launches) * args.steps,
            "stage_ms_per_step": {n: stage_ms[i] / args.steps for i, n in enumerate(STAGE_NAMES)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "kernel": dom_name, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                         "algorithmic_bytes": bytes_alg},
            "clocks": sampler.summary()}
    if not args.skip_cpu_baseline and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "uvc1")):
        try:
            res = run_reference(args, name, ref_scale)
            line["cpu_baseline"] = {"value": res["value"], "unit": "reads/s", "cores": res["cores"], "kind": "reference", "sample": res["sample"],
                                    "positions_per_s": res["positions_per_s"]}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % e}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
