#!/usr/bin/env python
"""bench.py - throughput of the pileup-and-score hot path on synthetic input (contract: DESIGN.md section "Measurement").

A "step" is one pass of the hot path over one batch of synthetic input: all tier-3 tiles of the named workload (tile list from the
reference-identical tiler, -t 16), cut into sub-batches of ~1.3 M reads that are submitted one after the other. Default workload =
BASELINE.json configs[1] at its FULL size (targeted panel, 2 Mbp as 10 000 targets @2000x, 26.7 M reads), produced on the box by the
threaded generator tools/synthgen.cpp. Metric (BASELINE.json): aligned reads/sec (`value`) and genomic positions/sec (`positions_per_s`).

* `value`    - device throughput with the inputs resident in HBM: sum of the CUDA-event times of every kernel of the step (events on the
               library's own stream), one context, one sub-batch per launch.
* `e2e`      - the same metric through the C ABI from HOST buffers (decoded BAM records in SoA form): host staging, H2D, kernels, D2H and
               VCF text inside the timed wall clock; sub-batches are pipelined over a few contexts (one CUDA stream pair each), the same
               schedule the uvc1 host uses. BAM decode (BGZF inflate + record parsing) is measured separately (`decode`) and the whole
               program (uvc1: tiling, decode, GPU, VCF text, BGZF output) in `pipeline`, the like-for-like of the reference arm.
* `roofline` - the kernel with the largest share of the step, against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
* `cpu_baseline` / `--impl reference` - the unmodified reference uvc1 (oracle/_ref) with all host threads on a bounded sample
               (a contiguous subset of the targets / of the contig) of the same BAM.
* N > 1      - one process per GPU, no data-path collective (regions are independent). Weak scaling: rank r processes shard r of an N-shard
               job (its own panel of the same shape, generated with seed + r); rank 0 concatenates the ranks' VCF bodies in shard order.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import math
import os
import re
import resource
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STAGE_NAMES = ["K0 per-read", "K1 prep+thres", "K2 bias pileup", "K2e indel events", "KF fragment columns", "K3a fragment stats", "K3b fragment consensus",
               "KM family columns", "K4a family ends", "K4 family+duplex consensus", "K4c family haplotypes", "K6 block-line inputs", "K5 candidate scoring",
               "P0/P1 device staging"]
STAGES_IMPLEMENTED = ("P0 read filter+family grouping, P1 repeat context, K0..K4c = updateByRegion3Aln, K6 block-line inputs, "
                      "K5a/K5 candidate scoring, VCF text (host) = process_batch")
WORKLOADS = {"c1": "uvc1 tumor-only, synthetic 1 Mbp @100x, non-UMI (BASELINE.json configs[0])",
             "c2": "targeted panel 2 Mbp (10000 targets x 200 bp) @2000x non-UMI, low-VAF spikes (BASELINE.json configs[1])",
             "c3": "UMI duplex panel 1 Mbp @20000x (BASELINE.json configs[2])"}
# c3 at full size is 133 M reads (20 GB of BAM): benched on a stated fraction of its region, same depth and tile shapes.
DEFAULT_SCALE = {"c1": 1.0, "c2": 1.0, "c3": 0.05}
TILER_THREADS = 16
# a sub-batch is full when it has enough positions to fill the GPU or when its reads reach the memory bound (the uvc1 host packs its batches the same way)
SUB_BATCH_POSITIONS = 512_000
SUB_BATCH_READS = 8_000_000
# rough reference throughput on 16 cores (reads/s), only used to size the bounded CPU samples
REF_RATE_GUESS = {"c1": 170e3, "c2": 95e3, "c3": 36e3}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2")
    ap.add_argument("--scale", type=float, default=None, help="fraction of the named config's region (default: per-config)")
    ap.add_argument("--workdir", default=os.environ.get("UVC_BENCH_DIR", "/tmp/uvc_bench"))
    ap.add_argument("--contexts", type=int, default=6, help="contexts (CUDA streams) the e2e step pipelines its sub-batches over: a context spends most of a sub-batch waiting (staging kernels, pileup, downloads), so several are needed to keep the GPU fed")
    ap.add_argument("--sub-batches", type=int, default=0, help="sub-batches per step (0: packed to ~512 k positions or 8 M reads each)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-pipeline", action="store_true", help="do not time the whole uvc1 program")
    ap.add_argument("--no-register", action="store_true", help="leave the decoded host buffers pageable (the library then stages them through its own page-locked copy)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ data

def synth_spec(name: str, scale: float, shard: int = 0):
    """Generator arguments of the named configs; `scale` shrinks the region (depth is kept), `shard` > 0 is another
    panel of the same shape (different seed) for the weak-scaling ranks."""
    sfx = ("" if shard == 0 else "_%d" % shard)
    if name == "c1":
        L = max(20_000, int(1_000_000 * scale))
        return dict(name="c1" + sfx, seed=1001 + 7919 * shard, contigs=[("chrS1" + sfx, L)], depth=100.0, n_snv=max(4, int(200 * scale)), n_indel=max(2, int(60 * scale)),
                    vafs=(0.05, 0.10, 0.25, 0.50), targets=None, amplicon_frac=0.0, umi=0)
    if name == "c2":
        n_targets = max(4, int(10_000 * scale))
        return dict(name="c2" + sfx, seed=1002 + 7919 * shard, contigs=[("chrS2" + sfx, n_targets * 1000 + 2000)], depth=2000.0, n_snv=max(4, int(400 * scale)),
                    n_indel=max(2, int(100 * scale)), vafs=(0.005, 0.01, 0.02, 0.05), targets=(n_targets, 1000, 1000, 200), amplicon_frac=0.2, umi=0)
    if name == "c3":
        L = max(5_000, int(1_000_000 * scale))
        return dict(name="c3" + sfx, seed=1003 + 7919 * shard, contigs=[("chrS3" + sfx, L)], depth=20000.0, n_snv=max(4, int(200 * scale)), n_indel=max(2, int(50 * scale)),
                    vafs=(0.001, 0.005, 0.01), targets=None, amplicon_frac=0.0, umi=1)
    raise ValueError("unknown config " + name)


def dataset(workdir: str, name: str, scale: float, shard: int = 0, threads: int = 0):
    """Generates (once per work directory) the seeded synthetic BAM/FASTA of a named config with uvc_b200/bin/uvc_synthgen."""
    spec = synth_spec(name, scale, shard)
    key = hashlib.sha1(repr(sorted(spec.items())).encode()).hexdigest()[:12]
    d = os.path.join(workdir, "%s_%s" % (spec["name"], key))
    meta = os.path.join(d, spec["name"] + ".meta.json")
    lock = d + ".lock"
    os.makedirs(workdir, exist_ok=True)
    while not os.path.exists(meta):
        try:
            fd = os.open(lock, os.O_CREAT | os.O_EXCL | os.O_WRONLY)
        except FileExistsError:
            try:                   # another process (rank, or the other bench arm) is generating it; a lock left behind by a killed run goes stale
                if time.time() - os.path.getmtime(lock) > 900:
                    os.unlink(lock)
            except OSError:
                pass
            time.sleep(0.2)
            continue
        try:
            os.close(fd)
            if os.path.exists(meta):
                break
            os.makedirs(d, exist_ok=True)
            gen = os.path.join(ROOT, "uvc_b200", "bin", "uvc_synthgen")
            if not os.path.exists(gen):
                raise RuntimeError(gen + " is missing: run __graft_entry__.build()")
            cmd = [gen, "--name", spec["name"], "--out", d, "--seed", str(spec["seed"]), "--depth", str(spec["depth"]), "--n-snv", str(spec["n_snv"]),
                   "--n-indel", str(spec["n_indel"]), "--vafs", ",".join("%g" % v for v in spec["vafs"]), "--amplicon-frac", str(spec["amplicon_frac"]),
                   "--umi", str(spec["umi"]), "--threads", str(threads)]
            for cn, cl in spec["contigs"]:
                cmd += ["--contig", "%s:%d" % (cn, cl)]
            if spec["targets"]:
                cmd += ["--targets", ",".join(str(x) for x in spec["targets"])]
            t0 = time.time()
            subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            m = json.load(open(meta))
            m["gen_s"] = time.time() - t0
            json.dump(m, open(meta + ".tmp", "w"))
            os.replace(meta + ".tmp", meta)
        finally:
            try:
                os.unlink(lock)
            except OSError:
                pass
    ds = json.load(open(meta))
    ds["targets"] = spec["targets"]
    ds["dir"] = d
    return ds


class _BedLine(C.Structure):
    _fields_ = [("tid", C.c_int32), ("beg_pos", C.c_int32), ("end_pos", C.c_int32), ("region_flag", C.c_uint32), ("n_reads", C.c_int64)]


def tile_list(ds, nthreads: int = TILER_THREADS, bed=None):
    """Tier-3 tiles exactly as the uvc1 host cuts them (uvc_b200/csrc/host/tiler.cpp = the reference's SamIter for -t nthreads)."""
    from uvc_b200 import capi
    lib = capi.load_host()
    lib.uvchost_tiler_open.restype = C.c_void_p
    lib.uvchost_tiler_open.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32]
    lib.uvchost_tiler_next.restype = C.c_int64
    lib.uvchost_tiler_next.argtypes = [C.c_void_p, C.POINTER(C.POINTER(_BedLine)), C.POINTER(C.c_int64)]
    lib.uvchost_tiler_close.argtypes = [C.c_void_p]
    lib.uvchost_tiler_set_scan_threads.argtypes = [C.c_void_p, C.c_int32]
    t = lib.uvchost_tiler_open(ds["bam"].encode(), (bed if bed is not None else (ds.get("bed") or "")).encode(), b"", nthreads, 1536, -1, 0)
    lib.uvchost_tiler_set_scan_threads(t, max(1, min(8, (os.cpu_count() or 1))))
    tiles = []
    while True:
        p, n = C.POINTER(_BedLine)(), C.c_int64()
        nreads = lib.uvchost_tiler_next(t, C.byref(p), C.byref(n))
        if nreads < 0:
            raise RuntimeError("tiler failed")
        if nreads == 0 and n.value == 0:
            break
        tiles += [(p[i].tid, p[i].beg_pos, p[i].end_pos, p[i].region_flag, p[i].n_reads) for i in range(n.value)]
    lib.uvchost_tiler_close(t)
    return tiles


def decode_sub_batches(ds, tiles, n_sub, threads, only=None):
    """Host-side decode: BAM -> SoA records of every tile's fetch window (what sam_itr_queryi(tid, beg - 2000, end + 2000) yields), one read
    buffer per sub-batch, sub-batches decoded on `threads` threads (each with its own BAM handle). Every record is stored once: the slices of
    neighbouring tiles overlap (uvchost_bam_fetch_span)."""
    from uvc_b200 import capi
    lib = capi.load_host()
    lib.uvchost_bam_fetch_span.restype = C.c_int64
    lib.uvchost_bam_fetch_span.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    if n_sub > 0:
        bounds = [len(tiles) * k // n_sub for k in range(n_sub + 1)]
    else:   # greedy packing by estimated positions (a read reaches about a fragment length beyond its tile) and reads, contig by contig
        bounds = [0]
        pos_sum = read_sum = 0
        tile_bases = float(sum(t[2] - t[1] for t in tiles))
        for k, t in enumerate(tiles):
            # (BED lines carry no read count: estimated from the tile's share of the called positions)
            lp, lr = (t[2] - t[1]) + 1000, (t[4] if t[4] > 0 else int(ds["n_reads"] * (t[2] - t[1]) / tile_bases))
            if k > bounds[-1] and (pos_sum + lp > SUB_BATCH_POSITIONS or read_sum + lr > SUB_BATCH_READS or tiles[k - 1][0] != t[0]):
                bounds.append(k)
                pos_sum = read_sum = 0
            pos_sum += lp
            read_sum += lr
        bounds.append(len(tiles))
        n_sub = len(bounds) - 1
    subs = [None] * n_sub
    nxt = [0]
    lock = threading.Lock()
    errs = []

    def work():
        bf = capi.BamFile(ds["bam"])
        try:
            while True:
                with lock:
                    k = nxt[0]
                    nxt[0] += 1
                if k >= n_sub:
                    return
                if only is not None and not only(k):     # (a rank of the strong-scaling run decodes only its own share)
                    continue
                sl = tiles[bounds[k]:bounds[k + 1]]
                rb = capi.ReadBuf()
                ctiles = []
                prev = (-1, 0, 0) if bounds[k] == 0 else tiles[bounds[k] - 1][:3]
                i = 0
                while i < len(sl):     # runs of tiles on the same contig
                    j = i
                    while j < len(sl) and sl[j][0] == sl[i][0]:
                        j += 1
                    n = j - i
                    begs = (C.c_int64 * n)(*[max(0, t[1] - 2000) for t in sl[i:j]])
                    ends = (C.c_int64 * n)(*[t[2] + 2000 for t in sl[i:j]])
                    rb0, rb1 = (C.c_int64 * n)(), (C.c_int64 * n)()
                    if lib.uvchost_bam_fetch_span(bf.handle, sl[i][0], n, begs, ends, rb.handle, rb0, rb1) < 0:
                        raise IOError("BAM fetch failed")
                    for q in range(n):
                        tid, beg, end, flag = sl[i + q][:4]
                        ctiles.append(capi.make_tile(tid, beg, end, flag, ds["contigs"][tid][1], rb0[q], rb1[q], prev))
                        prev = (tid, beg, end)
                    i = j
                subs[k] = (ctiles, rb, rb.view())
        except Exception as e:  # noqa: BLE001
            errs.append(e)
        finally:
            bf.close()
    ths = [threading.Thread(target=work) for _ in range(max(1, min(threads, n_sub)))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if errs:
        raise errs[0]
    return subs


def host_call_stats(lib):
    """uvcgpu_host_call_stats: total ms, calls and longest call per kind of driver call, process-wide."""
    buf = (C.c_double * 21)()
    lib.uvcgpu_host_call_stats.restype = C.c_int
    lib.uvcgpu_host_call_stats(buf, 21)
    return list(buf)


def host_cpu_times():
    """Aggregate jiffies of /proc/stat: (user + nice, system + irq + softirq, idle + iowait, steal)."""
    try:
        f = open("/proc/stat").readline().split()[1:]
        v = [int(x) for x in f] + [0] * 8
        return (v[0] + v[1], v[2] + v[5] + v[6], v[3] + v[4], v[7])
    except Exception:  # noqa: BLE001
        return None


def host_cpu_load(a, b):
    """Share of the box's CPU time between two host_cpu_times() samples: the e2e number depends on how much of the host this process really
    got (the GPU boxes are virtual machines: `steal` is time the hypervisor gave to somebody else)."""
    if not a or not b:
        return None
    d = [y - x for x, y in zip(a, b)]
    tot = max(1, sum(d))
    try:
        load = open("/proc/loadavg").read().split()[0]
    except Exception:  # noqa: BLE001
        load = None
    return {"user": d[0] / tot, "system": d[1] / tot, "idle": d[2] / tot, "steal": d[3] / tot, "cores": os.cpu_count(), "loadavg_1min": load}


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU during the timed regions. In-process NVML queries (a few microseconds each): spawning
    nvidia-smi ten times a second per rank takes driver-wide locks and measurably stalls the CUDA calls of every process on the box."""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.sm_max = None
        self.how = "nvml"

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu_index
            if visible:      # NVML enumerates all GPUs of the box; map the CUDA ordinal through CUDA_VISIBLE_DEVICES when it is a list of indices
                try:
                    idx = int(visible.split(",")[self.gpu_index])
                except Exception:
                    idx = self.gpu_index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = [("hw_slowdown", pynvml.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", pynvml.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", pynvml.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", pynvml.nvmlClocksEventReasonSwPowerCap)]
            while not self.stop_flag:
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                r = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                for n, bit in names:
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.25)
            return
        except Exception:
            self.how = "nvidia-smi"
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
            time.sleep(1.0)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(s), "how": self.how}


# ------------------------------------------------------------------------------------------------ reference arm / CPU baseline

def n_positions_of(ds):
    return int(ds["n_positions"])


def reference_samples(ds, name, seconds_per_step, n_samples, workdir):
    """Bounded samples of the workload for the reference: `n_samples` disjoint contiguous subsets (targets of the panel / stretches of the contig),
    each sized for about `seconds_per_step` of the reference on this box's cores. Returns [(extra uvc1 arguments, n_reads, n_positions)]."""
    cores = os.cpu_count() or 1
    want_reads = REF_RATE_GUESS.get(name, 100e3) * min(1.0, cores / 16.0) * seconds_per_step
    total_reads, total_pos = float(ds["n_reads"]), float(n_positions_of(ds))
    frac = min(1.0, want_reads / total_reads)
    out = []
    if ds.get("targets"):
        n_t, first, step, tlen = ds["targets"]
        per = max(1, min(n_t, int(round(n_t * frac))))
        cname = ds["contigs"][0][0]
        for s in range(n_samples):
            k0 = (s * per) % max(1, n_t - per + 1)
            bed = os.path.join(workdir, "ref_sample_%s_%d.bed" % (name, s))
            with open(bed, "w") as f:
                for k in range(k0, k0 + per):
                    f.write("%s\t%d\t%d\n" % (cname, first + step * k, first + step * k + tlen))
            out.append((["-R", bed], total_reads * per / n_t, per * tlen, "%d of the %d targets" % (per, n_t)))
    else:
        # whole-contig configs: BED lines of about the size of the tiles the reference would cut itself (21-42 kbp at 100x, 2.6-5.2 kbp at 20000x)
        cname, clen = ds["contigs"][0]
        line_len = min(clen, 2500 if name == "c3" else 20000)
        n_lines = max(1, clen // line_len)
        per = max(1, min(n_lines, int(round(n_lines * frac))))
        for s in range(n_samples):
            k0 = (s * per) % max(1, n_lines - per + 1)
            bed = os.path.join(workdir, "ref_sample_%s_%d.bed" % (name, s))
            with open(bed, "w") as f:
                for k in range(k0, k0 + per):
                    f.write("%s\t%d\t%d\n" % (cname, k * line_len, min(clen, (k + 1) * line_len)))
            out.append((["-R", bed], total_reads * per * line_len / clen, per * line_len, "%d of the %d bases of the contig (BED lines of %d bases)" % (per * line_len, clen, line_len)))
    return out


def run_uvc1(exe, ds, extra, out_vcf, threads, more=()):
    cmd = [exe, ds["bam"], "-f", ds["fasta"], "-o", out_vcf, "-s", "S", "-t", str(threads)] + list(extra) + list(more)
    t0 = time.time()
    p = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    wall = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError("%s failed: %s" % (exe, p.stderr[-500:]))
    m = re.search(r"Wall clock time passed: ([0-9.]+) seconds", p.stderr)
    return (float(m.group(1)) if m else wall), wall, p.stderr


def run_reference(ds, name, workdir, n_steps, n_warm, seconds_per_step):
    """Times the unmodified reference uvc1 (oracle/_ref) with all host threads: every step is a different bounded sample of the workload."""
    uvc1 = os.path.join(ROOT, "oracle", "_ref", "uvc1")
    cores = os.cpu_count() or 1
    samples = reference_samples(ds, name, seconds_per_step, n_steps + n_warm, workdir)
    reads = pos = secs = 0.0
    per_step = []
    for i, (extra, n_r, n_p, what) in enumerate(samples):
        t, _, _ = run_uvc1(uvc1, ds, extra, os.path.join(workdir, "ref_%s.vcf.gz" % name), cores)
        if i >= n_warm:
            reads += n_r
            pos += n_p
            secs += t
            per_step.append(t)
    what = samples[0][3]
    return dict(value=reads / secs, unit="reads/s", cores=cores, kind="reference",
                sample="%s per step (%.0f reads), a different subset every step; oracle/_ref/uvc1 -t %d on the bench BAM (index seek, BAM decode, tiling and BGZF output "
                       "included); %d timed steps after %d warm-up, %.2f s per step" % (what, reads / max(1, len(per_step)), cores, len(per_step), n_warm, secs / max(1, len(per_step))),
                positions_per_s=pos / secs, seconds_per_step=secs / max(1, len(per_step)), steps_run=len(per_step))


def kernel_algorithmic_bytes(stage: int, n_ext_positions: float, n_reads: float, n_pos_pileup: float = None, n_pos_consensus: float = None) -> float:
    """Compulsory bytes of one launch of a kernel: its inputs read once plus its outputs written once. The output-only
    position kernels (K2; K3b, K4) write the positions they run on (the needed ones, uvcgpu.h: all_positions)."""
    P, R = float(n_ext_positions), float(n_reads)
    if stage == 2 and n_pos_pileup is not None:
        P = float(n_pos_pileup)
    if stage in (6, 9) and n_pos_consensus is not None:
        P = float(n_pos_consensus)
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
    workload = WORKLOADS.get(name, name) + ("; full size" if scale == 1.0 else "; region fraction %g" % scale)
    metric = "aligned reads/sec (and positions/sec) called"
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        ds = dataset(args.workdir, name, scale, 0, cores)
        # every step is a bounded sample: ~4 s of the reference each, so that warmup + steps end within a few minutes
        res = run_reference(ds, name, args.workdir, args.steps, args.warmup, 4.0)
        line = {"impl": "reference", "metric": metric, "value": res["value"], "unit": "reads/s",
                "positions_per_s": res["positions_per_s"], "n_gpus": args.gpus, "steps": res["steps_run"], "warmup": args.warmup,
                "ms_per_step": res["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/int64 counters, f64 scoring",
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
    host_threads = max(1, cores // world)

    # Regions are independent: no data-path collective. Weak scaling: rank r processes shard r of an N-shard job.
    t_gen0 = time.time()
    ds = dataset(args.workdir, name, scale, rank, host_threads)
    gen_s = time.time() - t_gen0
    t_tile0 = time.time()
    tiles = tile_list(ds)
    tile_s = time.time() - t_tile0
    n_reads = int(ds["n_reads"])             # primary mapped records of the BAM, each counted once, not once per overlapping tile
    n_sub = max(0, min(args.sub_batches, len(tiles)))       # 0: packed by positions and reads
    t_dec0 = time.time()
    subs = decode_sub_batches(ds, tiles, n_sub, host_threads)
    decode_s = time.time() - t_dec0
    n_sub = len(subs)
    n_records = sum(len(s[1]) for s in subs)
    bam_bytes = os.path.getsize(ds["bam"])
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
        ticket = ctx.submit(sub[0], sub[2])       # host staging, H2D copies and every pileup kernel enqueued on the context's stream
        return (ticket, sub, t0, time.time())

    def finish_tiles(ctx, pending, keep_text=None):
        ticket, sub, t0, t1 = pending
        tw0 = time.time()
        ctx.collect(ticket)
        t2 = time.time()
        st = ctx.score(ticket)               # candidate scoring on the device + D2H of the kept records and block-line inputs
        t3 = time.time()
        txt = ctx.batch_vcf(ticket)          # the step's result: the VCF body text of the sub-batch's tiles, in tile order
        nbytes = len(txt)
        if keep_text is not None:
            keep_text.append(txt)
        t4 = time.time()
        ctx.release(ticket)
        st.vcf_bytes = nbytes
        st.phase_s = (t1 - t0, t2 - tw0, t3 - t2, t4 - t3, time.time() - t4)   # submit (staging + H2D enqueue), wait, score, text, release
        return st

    # ---- phase 1: device-resident throughput (`value`): one context, one sub-batch per launch, CUDA-event time of every kernel
    ctx0 = make_ctx(host_threads)
    # the decoded inputs are page-locked once (the e2e contract reads them "from pinned host memory"): submits then upload straight from them
    t_reg0 = time.time()
    ctx0.lib.uvcgpu_host_register_reads.argtypes = [C.c_void_p]
    n_registered = 0
    if not args.no_register:
        for sub in subs:
            if sub is not None and ctx0.lib.uvcgpu_host_register_reads(C.byref(sub[2])) == 0:
                n_registered += 1
    register_s = time.time() - t_reg0
    n_warm = max(args.warmup, 3)
    agg = {"n_reads_kept": 0, "n_ext_positions": 0, "n_positions": 0, "n_reads_in": 0, "n_positions_pileup": 0, "n_positions_consensus": 0}

    def device_pass(acc_stage=None):
        ms, launches = 0.0, 0
        for sub in subs:
            st = finish_tiles(ctx0, submit_tiles(ctx0, sub))
            ms += st.kernel_ms
            launches += int(st.gpu_launches)
            if acc_stage is not None:
                for i in range(16):
                    acc_stage[i] += st.kernel_ms_by_stage[i]
            for k in agg:
                agg[k] += int(getattr(st, k))
        return ms, launches

    for _ in range(n_warm):
        device_pass()
    for k in agg:
        agg[k] = 0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    kernel_ms = 0.0
    stage_ms = [0.0] * 16
    launches = 0
    for _ in range(args.steps):
        ms, ln = device_pass(stage_ms)
        kernel_ms += ms
        launches += ln
    torch.cuda.synchronize()
    for k in agg:
        agg[k] //= args.steps

    # ---- phase 2: end to end from host buffers: sub-batches pipelined over a few contexts (streams)
    # (two sub-batches are in flight per context: deep panels have few, large sub-batches - the contexts are limited so that about 28 M reads
    #  are in flight at most, which keeps the column caches of all of them inside the 180 GB of HBM)
    reads_per_sub = max(1, n_records // max(1, len(subs)))
    n_ctx = max(1, min(args.contexts, len(subs), max(1, 14_000_000 // reads_per_sub)))
    # ... and by what one context was seen to hold on the device during the phase above (its slabs: one sub-batch at a time there, two in
    # flight per context here, with their size classes): four fifths of the HBM at most
    mem_free1, mem_total1 = torch.cuda.mem_get_info()
    one_ctx_bytes = max(1, int(mem_total1 - mem_free1))
    n_ctx = max(1, min(n_ctx, int(0.8 * mem_total1 / (2.0 * one_ctx_bytes))))
    ctxs = [ctx0] + [make_ctx(max(1, host_threads // n_ctx + 1)) for _ in range(n_ctx - 1)]
    ctx0.lib.uvcgpu_set_host_threads(ctx0.handle, max(1, host_threads // n_ctx + 1))
    totals = {"h2d": 0, "d2h": 0, "vcf": 0, "rec": 0, "launch": 0, "prep_ms": 0.0, "sync_ms": 0.0, "stage_call_ms": 0.0, "sc2": 0.0, "sc3": 0.0, "sc4": 0.0, "sc5": 0.0, "submit_s": 0.0, "wait_s": 0.0, "score_s": 0.0, "text_s": 0.0, "release_s": 0.0}
    last_text = {}

    def e2e_steps(n_steps, keep_last=False):
        # the n_steps passes over the workload are one continuous stream of sub-batches (no drain between steps), as in a long run of the uvc1 host
        work_items = [(si, sub) for _ in range(n_steps) for si, sub in enumerate(subs)]
        first_of_last = len(work_items) - len(subs)
        nxt = [0]
        lock = threading.Lock()
        acc = {k: 0 for k in totals}
        errs = []

        def work(ctx):
            try:
                # two sub-batches in flight per context: the next one is staged and enqueued before the previous one is collected, so the
                # stream always has work queued while the host stages
                pending = None
                pending_k = -1
                while True:
                    with lock:
                        k = nxt[0]
                        nxt[0] += 1
                    nxt_pending = submit_tiles(ctx, work_items[k][1]) if k < len(work_items) else None
                    if pending is None and nxt_pending is None:
                        return
                    if pending is None:
                        pending, pending_k = nxt_pending, k
                        continue
                    keep = [] if (keep_last and pending_k >= first_of_last) else None
                    st = finish_tiles(ctx, pending, keep)
                    if keep is not None:
                        last_text[work_items[pending_k][0]] = b"".join(keep)
                    pending, pending_k = nxt_pending, k
                    with lock:
                        acc["h2d"] += int(st.h2d_bytes)
                        acc["d2h"] += int(st.d2h_bytes)
                        acc["vcf"] += int(st.vcf_bytes)
                        acc["rec"] += int(st.n_vcf_records)
                        acc["launch"] += int(st.gpu_launches)
                        acc["prep_ms"] += st.host_prep_ms
                        acc["sync_ms"] += st.reserved[0]
                        acc["stage_call_ms"] += st.reserved[1]
                        for j in (2, 3, 4, 5):
                            acc["sc%d" % j] += st.reserved[j]
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

    e2e_steps(min(n_warm, 3))
    # steady state: the library page-locks its staging blocks in the background; warm up until none is outstanding (bounded)
    # (page-locking while batches are in flight also slows every CUDA call down, so the timed region must not trigger any: warm up until the
    # cache has stopped growing for two rounds in a row)
    ctx0.lib.uvcgpu_staging_backlog.restype = C.c_int
    ctx0.lib.uvcgpu_staging_pinned_bytes.restype = C.c_int64
    stable = 0
    for _ in range(12):
        t_wait = time.time()
        while ctx0.lib.uvcgpu_staging_backlog() > 0 and time.time() - t_wait < 10.0:
            time.sleep(0.05)
        before = int(ctx0.lib.uvcgpu_staging_pinned_bytes())
        e2e_steps(2)      # (two steps back to back: the overlap across a step boundary needs its staging blocks too)
        grown = (ctx0.lib.uvcgpu_staging_backlog() > 0 or int(ctx0.lib.uvcgpu_staging_pinned_bytes()) != before)
        stable = 0 if grown else stable + 1
        if stable >= 3:
            break
    backlog0 = int(ctx0.lib.uvcgpu_staging_backlog())
    pinned0 = int(ctx0.lib.uvcgpu_staging_pinned_bytes())
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    cpu0 = host_cpu_times()
    ru0 = resource.getrusage(resource.RUSAGE_SELF)
    calls0 = host_call_stats(ctx0.lib)
    t0 = time.time()
    acc = e2e_steps(args.steps, keep_last=True)
    for k in totals:
        totals[k] += acc[k]
    torch.cuda.synchronize()
    wall_s = time.time() - t0
    host_load = host_cpu_load(cpu0, host_cpu_times())
    ru1 = resource.getrusage(resource.RUSAGE_SELF)
    if host_load is not None:
        host_load["this_process_cpu_seconds_per_step"] = ((ru1.ru_utime - ru0.ru_utime) + (ru1.ru_stime - ru0.ru_stime)) / args.steps
    calls1 = host_call_stats(ctx0.lib)
    mem_free, mem_total = torch.cuda.mem_get_info()
    call_stats = {name: {"ms_per_step": (calls1[3 * i] - calls0[3 * i]) / args.steps, "calls_per_step": (calls1[3 * i + 1] - calls0[3 * i + 1]) / args.steps,
                         "longest_ms_since_start": calls1[3 * i + 2]}
                  for i, name in enumerate(("device_alloc", "device_free", "memset", "download_enqueue", "kernel_launches", "event_waits", "upload_enqueue"))}
    backlog1 = int(ctx0.lib.uvcgpu_staging_backlog())
    sampler.stop_flag = True
    body = b"".join(last_text[k] for k in sorted(last_text))
    shard_sha = hashlib.sha1(body).hexdigest()
    reads_all, pos_all = n_reads, agg["n_positions"]
    concat = None
    if world > 1:
        tt = torch.tensor([kernel_ms, wall_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        kernel_ms, wall_s = float(tt[0]), float(tt[1])
        cnt = torch.tensor([n_reads, agg["n_positions"]], device="cuda", dtype=torch.int64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        reads_all, pos_all = int(cnt[0]), int(cnt[1])
        # host-side concatenation of the shards' VCF bodies in shard order (what uvc1 does for its lanes, main.cpp:1541-1551 in the reference)
        with open(os.path.join(args.workdir, "vcf_shard_%d_of_%d.txt" % (rank, world)), "wb") as f:
            f.write(body)
        dist.barrier()
        if rank == 0:
            h = hashlib.sha1()
            total = 0
            for r in range(world):
                b = open(os.path.join(args.workdir, "vcf_shard_%d_of_%d.txt" % (r, world)), "rb").read()
                h.update(b)
                total += len(b)
            concat = {"shards": world, "bytes": total, "sha1": h.hexdigest()}
        dist.barrier()
    strong = None
    if world > 1:
        # ---- strong scaling: ONE tile list (shard 0, the N = 1 workload) interleaved over the ranks' GPUs; rank 0 concatenates the VCF text in
        # tile order (the reference's ordered flush, main.cpp:1541-1551) and compares it with its own single-GPU output of the same shard
        ds0 = dataset(args.workdir, name, scale, 0, host_threads)
        tiles0 = tile_list(ds0) if rank != 0 else tiles
        subs0 = decode_sub_batches(ds0, tiles0, 0, host_threads, only=(lambda k: k % world == rank)) if rank != 0 else subs
        if rank != 0 and not args.no_register:
            for sub in subs0:
                if sub is not None:
                    ctx0.lib.uvcgpu_host_register_reads(C.byref(sub[2]))
        mine0 = [(k, s0) for k, s0 in enumerate(subs0) if s0 is not None and k % world == rank]
        if rank != 0:
            bases0 = {tid: capi.read_fasta_contig(ds0["fasta"], cname) for tid, (cname, _) in enumerate(ds0["contigs"])}
            for c in ctxs:
                for tid, (cname, _) in enumerate(ds0["contigs"]):
                    c.set_contig(tid, bases0[tid])
                    c.set_contig_name(tid, cname)
        texts0 = {}

        def strong_pass(keep):
            nxt0 = [0]
            lock0 = threading.Lock()
            errs0 = []

            def work0(ctx):
                try:
                    pending = None
                    while True:
                        with lock0:
                            i = nxt0[0]
                            nxt0[0] += 1
                        nxt_p = (mine0[i][0], submit_tiles(ctx, mine0[i][1])) if i < len(mine0) else None
                        if pending is None and nxt_p is None:
                            return
                        if pending is not None:
                            kk = [] if keep else None
                            finish_tiles(ctx, pending[1], kk)
                            if keep:
                                texts0[pending[0]] = b"".join(kk)
                        pending = nxt_p
                except Exception as e:  # noqa: BLE001
                    errs0.append(e)
            ths0 = [threading.Thread(target=work0, args=(c,)) for c in ctxs]
            for t in ths0:
                t.start()
            for t in ths0:
                t.join()
            if errs0:
                raise errs0[0]
        strong_pass(False)
        n_strong = max(1, min(args.steps, 3))
        dist.barrier()
        torch.cuda.synchronize()
        ts0 = time.time()
        for it in range(n_strong):
            strong_pass(it == n_strong - 1)
        torch.cuda.synchronize()
        tt = torch.tensor([time.time() - ts0], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        strong_s = float(tt[0])
        with open(os.path.join(args.workdir, "vcf_strong_%d_of_%d.bin" % (rank, world)), "wb") as f:
            import pickle
            pickle.dump(texts0, f)
        dist.barrier()
        if rank == 0:
            import pickle
            allt = {}
            for r in range(world):
                allt.update(pickle.load(open(os.path.join(args.workdir, "vcf_strong_%d_of_%d.bin" % (r, world)), "rb")))
            body0 = b"".join(allt[k] for k in sorted(allt))
            strong = {"value": n_reads * n_strong / strong_s, "unit": "reads/s", "steps": n_strong, "wall_ms_per_step": strong_s * 1e3 / n_strong,
                      "what": "the N = 1 workload (shard 0) with its sub-batches interleaved over the %d GPUs, one process per GPU, rank 0 concatenates the VCF text in tile order" % world,
                      "vcf_body_sha1": hashlib.sha1(body0).hexdigest(), "identical_to_single_gpu_output": hashlib.sha1(body0).hexdigest() == shard_sha}
        dist.barrier()
    for c in ctxs:                           # explicit teardown (contexts own CUDA streams and pool memory)
        c.close()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    value = reads_all * args.steps / (kernel_ms / 1e3)
    e2e = reads_all * args.steps / wall_s
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dom = max(range(13), key=lambda i: stage_ms[i])
    dom_name = STAGE_NAMES[dom] if dom < len(STAGE_NAMES) else "stage%d" % dom
    dom_ms = stage_ms[dom] / args.steps                      # per step = per pass over the sub-batches (n_sub launches)
    dom_bytes = kernel_algorithmic_bytes(dom, agg["n_ext_positions"], agg["n_reads_kept"], agg["n_positions_pileup"], agg["n_positions_consensus"])
    achieved = dom_bytes / (dom_ms / 1e3) / 1e9
    step_ms = kernel_ms / args.steps
    path_bytes = agg["n_reads_kept"] * (1.5 * 150 + 64) + agg["n_ext_positions"] * 2 * 6272
    traffic = None
    issue = None
    try:   # dram bytes of that kernel from the committed ncu --set full capture of the same workload shape (profiles/), per launch
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        ent = tj.get(name, {}).get(dom_name)
        if ent:
            traffic = ent["dram_bytes"]
            if ent.get("ipc_per_sm"):
                issue = {"bound": "issue", "achieved": ent["ipc_per_sm"], "peak": 4.0, "unit": "warp instructions/cycle/SM", "frac": ent["ipc_per_sm"] / 4.0,
                         "warp_instructions": ent.get("warp_instructions"), "resident_warps_pct": ent.get("resident_warps_pct"),
                         "source": "profiles/r02_traffic.json (ncu --set full, one sub-batch launch)"}
    except Exception:
        pass
    line = {"metric": metric, "value": value, "unit": "reads/s",
            "positions_per_s": pos_all * args.steps / (kernel_ms / 1e3),
            "n_gpus": args.gpus, "steps": args.steps, "warmup": n_warm, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/int64 counters, f64 scoring",
            "data": "synthetic (tools/synthgen.cpp, seeded)",
            "config": {"workload": workload, "tiles": len(tiles), "tiler": "reference SamIter semantics, -t %d" % TILER_THREADS,
                       "reads_per_step": n_reads, "read_records_per_step_incl_tile_halos": agg["n_reads_kept"],
                       "positions_per_step": agg["n_positions"], "ext_positions_per_step": agg["n_ext_positions"], "pileup_positions_per_step": agg["n_positions_pileup"], "consensus_positions_per_step": agg["n_positions_consensus"], "stages": STAGES_IMPLEMENTED,
                       "sub_batches_per_step": len(subs),
                       "l2": "per-position state of a sub-batch (%.0f MB) is larger than L2, no flush needed" % (agg["n_ext_positions"] * 6272 / 1e6 / len(subs)),
                       "dataset_generation_s_untimed": ds.get("gen_s"), "tiler_s_untimed": tile_s, "host_threads": host_threads,
                       "device_bytes_held_by_one_context_after_the_kernel_phase": one_ctx_bytes,
                       "host_inputs": ("%d of %d decoded sub-batch buffers page-locked with uvcgpu_host_register_reads (%.1f s, untimed): uploaded straight from them" % (n_registered, len(subs), register_s)
                                       if n_registered else "pageable: staged through the library's page-locked copy"),
                       "multi_gpu": ("rank r processes shard r (same shape, seed + r) of an %d-shard job; no data-path collective; rank 0 concatenates the VCF bodies" % world
                                     if world > 1 else "single GPU"),
                       "e2e_schedule": "%d sub-batches per step over %d contexts (streams), two in flight per context, steps back to back, %d host threads" % (len(subs), len(ctxs), host_threads)},
            "e2e": {"value": e2e, "unit": "reads/s", "positions_per_s": pos_all * args.steps / wall_s,
                    "h2d_bytes_per_step": totals["h2d"] // args.steps, "d2h_bytes_per_step": totals["d2h"] // args.steps,
                    "vcf_bytes_per_step": totals["vcf"] // args.steps, "vcf_records_per_step": totals["rec"] // args.steps,
                    "vcf_body_sha1_last_step": shard_sha, "vcf_concatenation": concat, "strong_scaling": strong,
                    "scope": "C ABI from decoded host SoA buffers: host staging + H2D + kernels + D2H + VCF text; excludes BAM decode (see `decode`) and BGZF output (see `pipeline`)",
                    "host_prep_ms_per_step_summed_over_contexts": totals["prep_ms"] / args.steps,
                    "submit_wait_for_batch_sizes_ms_per_step_summed_over_contexts": totals["sync_ms"] / args.steps,
                    "submit_staging_part_ms_per_step_summed_over_contexts": totals["stage_call_ms"] / args.steps,
                    "score_parts_ms_per_step_summed_over_contexts": {"sparse_download": totals["sc2"] / args.steps, "sparse_maps_host": totals["sc3"] / args.steps,
                                                                     "indel_table_host": totals["sc4"] / args.steps, "scoring_kernels_and_downloads": totals["sc5"] / args.steps},
                    "call_ms_per_step_summed_over_contexts": {k[:-2]: totals[k] * 1e3 / args.steps for k in ("submit_s", "wait_s", "score_s", "text_s", "release_s")},
                    "wall_ms_per_step": wall_s * 1e3 / args.steps,
                    "host_cpu_during_timed_region": host_load,
                    "driver_calls_summed_over_threads": call_stats,
                    "device_memory_in_use_bytes": int(mem_total - mem_free),
                    "staging_blocks_not_yet_page_locked": {"at_start": backlog0, "at_end": backlog1},
                    "staging_page_locked_bytes": {"at_start": pinned0, "at_end": int(ctx0.lib.uvcgpu_staging_pinned_bytes())}},
            "decode": {"seconds": decode_s, "threads": min(host_threads, max(1, n_sub)), "records": n_records, "records_per_s": n_records / decode_s,
                       "bam_bytes": bam_bytes, "bam_MB_per_s": bam_bytes / decode_s / 1e6,
                       "what": "BGZF inflate + BAM record parsing into the SoA buffers of every tile's fetch window (each record stored once), untimed in `e2e`"},
            "gpu_launches": launches + totals["launch"],
            "stage_ms_per_step": {n: stage_ms[i] / args.steps for i, n in enumerate(STAGE_NAMES)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": dom_name, "kernel_ms": dom_ms, "kernel_ms_per_launch": dom_ms / len(subs), "launches_per_step": len(subs),
                         "kernel_share_of_step": dom_ms / step_ms, "algorithmic_bytes": dom_bytes,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "whole_path": {"algorithmic_bytes": path_bytes, "achieved": path_bytes / (step_ms / 1e3) / 1e9,
                                        "frac": path_bytes / (step_ms / 1e3) / 1e9 / peak},
                         "issue": issue},
            "clocks": sampler.summary()}
    if n_registered:
        capi.load_gpu().uvcgpu_host_unregister_reads.argtypes = [C.c_void_p]
    for s in subs:
        if n_registered and s is not None:
            capi.load_gpu().uvcgpu_host_unregister_reads(C.byref(s[2]))
        s[1].close()
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "uvc1"))
    if world == 1 and not args.skip_pipeline:
        # the whole program, like for like with the reference arm: our uvc1 on the full BAM (tiling, BAM decode, GPU, VCF text, BGZF output)
        try:
            exe = os.path.join(ROOT, "uvc_b200", "bin", "uvc1")
            extra = (["-R", ds["bed"]] if ds.get("bed") else [])
            walls = []
            for _ in range(3):     # a 3 s process on a shared box: run to run 2.7 - 3.8 s (tools/gpu_pipeline_check.sh); the median of three is reported
                t_ref, wall, _ = run_uvc1(exe, ds, extra, os.path.join(args.workdir, "ours_%s.vcf.gz" % name), TILER_THREADS, ["--gpus", "1"])
                walls.append(wall)
            wall = sorted(walls)[1]
            line["pipeline"] = {"seconds": wall, "seconds_of_every_run": walls, "reads_per_s": n_reads / wall, "positions_per_s": n_positions_of(ds) / wall,
                                "what": "uvc_b200/bin/uvc1 -t %d --gpus 1 on the whole bench BAM, process start to exit (CUDA start-up, tiling, decode, kernels, VCF text, BGZF output); median of three runs" % TILER_THREADS}
        except Exception as e:  # noqa: BLE001
            line["pipeline"] = {"seconds": None, "what": "failed: %s" % e}
    if world == 1 and not args.skip_cpu_baseline and have_ref:   # (reported at N = 1 only)
        try:
            res = run_reference(ds, name, args.workdir, 2, 0, 8.0)      # bounded: two samples of ~8 s each
            line["cpu_baseline"] = {"value": res["value"], "unit": "reads/s", "cores": res["cores"], "kind": "reference", "sample": res["sample"],
                                    "positions_per_s": res["positions_per_s"]}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": cores, "kind": "reference", "sample": "failed: %s" % e}
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()


if __name__ == "__main__":
    main()
