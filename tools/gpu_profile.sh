#!/bin/bash
# Runs on the GPU box: the two ncu passes of /opt/skills/guides/B200_PROFILING.md for the default bench workload.
#   1. launch list (per-launch durations, cold-cache and serialised: compare SHARES with bench.py's live CUDA-event times)
#   2. one --set full capture of the kernels that dominate the step (dram bytes -> roofline.traffic; source-level stall samples)
CFG=${1:-c2}; TAG=${2:-r01}
mkdir -p gpurun_out
python bench.py --config $CFG --steps 1 --warmup 3 --skip-cpu-baseline > /dev/null 2>&1   # generates the dataset outside the profiler
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_${CFG}.csv python bench.py --config $CFG --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/${TAG}_launches_${CFG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"uvc_k4_family_consensus|uvc_k2_bias_pileup|uvc_k3b_fragment|uvc_km_family|uvc_kf_fragment|uvc_k1_prep|uvc_k0_read|uvc_k5_score" -c 8 -f -o gpurun_out/${TAG}_full_${CFG} python bench.py --config $CFG --steps 1 --warmup 3 --skip-cpu-baseline > gpurun_out/${TAG}_full_${CFG}.log 2>&1
cp uvc_b200/lib/engine.o gpurun_out/${TAG}_engine.o   # tools/ncu_lines.py joins the report with the line table of exactly this build
ls -la gpurun_out
