#!/usr/bin/env python
"""Summarises ncu outputs brought back from the GPU box into the small tracked files under profiles/:
  launches CSV (--metrics gpu__time_duration.sum)  -> per-kernel launch count, mean ms, share
  full report (--set full)                          -> per-kernel duration, DRAM bytes, hit rates, occupancy, issue utilisation; traffic JSON
usage: ncu_summarize.py launches <csv> <out.md> | full <ncu-rep> <out.md> <traffic.json> <workload-key>"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

STAGE_OF = {"uvc_k0_read_consts": "K0 per-read", "uvc_k1_prep_thres": "K1 prep+thres", "uvc_k2_bias_pileup": "K2 bias pileup", "uvc_k2e_indel_events": "K2e indel events",
            "uvc_kf_fragment_columns": "KF fragment columns", "uvc_k3a_fragment_stats": "K3a fragment stats", "uvc_k3b_fragment_consensus": "K3b fragment consensus",
            "uvc_km_family_columns": "KM family columns", "uvc_k4a_family_ends": "K4a family ends", "uvc_k4_family_consensus": "K4 family+duplex consensus", "uvc_k4_family_consensus_umi": "K4 family+duplex consensus",
            "uvc_k4c_family_haplotypes": "K4c family haplotypes", "uvc_k6_gvcf_inputs": "K6 block-line inputs", "uvc_k5_score_candidates": "K5 candidate scoring",
            "uvc_k5a_flag_candidates": "K5 candidate scoring", "uvc_k5c_candidate_depths": "K5 candidate scoring"}


def launches(path, out):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        name = r[ik].split("(")[0]
        if name.startswith("void "):        # templated kernels: "void uvc_for_each<uvc::P0eKept>"
            name = name[5:]
        v = float(r[iv].replace(",", ""))
        v = v / 1e6 if r[iu] in ("ns", "nsecond") else (v / 1e3 if r[iu] in ("us", "usecond") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    ours = {k: v for k, v in agg.items() if k.startswith("uvc_")}
    tot = sum(v[1] for v in ours.values())
    with open(out, "w") as f:
        f.write("| kernel | launches | mean ms | share of our kernels |\n|---|---|---|---|\n")
        for k, (n, ms) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.3f | %.1f%% |\n" % (k, n, ms / n, 100 * ms / tot))
        others = {k: v for k, v in agg.items() if not k.startswith("uvc_")}
        f.write("\nOther launches in the process (torch/driver): %d kernels, %.3f ms in total.\n" % (sum(v[0] for v in others.values()), sum(v[1] for v in others.values())))


def full(rep, out, traffic_json, key):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    want = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "GB read"), ("dram__bytes_write.sum", "GB written"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst"),
            ("launch__registers_per_thread", "regs"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak")]
    tj = {}
    try:
        tj = json.load(open(traffic_json))
    except Exception:
        pass
    ent = tj.setdefault(key, {})
    with open(out, "w") as f:
        f.write("| kernel | " + " | ".join(w[1] for w in want) + " |\n|---|" + "---|" * len(want) + "\n")
        for r in rows[2:]:
            name = r[ix["Kernel Name"]].split("(")[0]
            vals = []
            for m, _ in want:
                vals.append(r[ix[m]] if m in ix else "")
            f.write("| %s | %s |\n" % (name, " | ".join(("%.3f" % float(v.replace(",", ""))) if v else "-" for v in vals)))

            def gb(m):
                v = float(r[ix[m]].replace(",", ""))
                u = units[ix[m]]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
            if name in STAGE_OF:
                def num(m):
                    return float(r[ix[m]].replace(",", "")) if m in ix and r[ix[m]] else None
                ent[STAGE_OF[name]] = {"dram_bytes": gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"), "ms_under_profiler": float(r[ix["gpu__time_duration.sum"]].replace(",", "")),
                                       "warp_instructions": num("smsp__inst_executed.sum"), "ipc_per_sm": num("sm__inst_executed.avg.per_cycle_elapsed"),
                                       "resident_warps_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"), "registers": num("launch__registers_per_thread"),
                                       "source": rep.split("/")[-1]}
    json.dump(tj, open(traffic_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5])
