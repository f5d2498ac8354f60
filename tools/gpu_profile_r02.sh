#!/bin/bash
# Runs on the GPU box: the two ncu passes of /opt/skills/guides/B200_PROFILING.md on ONE sub-batch of configs[1] (a sub-batch of the full-size run
# has the same shape: ~1.3 M reads, ~540 k extended positions at 2000x).
#   1. launch list (per-launch durations, cold-cache and serialised: compare SHARES with bench.py's live CUDA-event times)
#   2. one --set full capture of the kernels that dominate the step (dram bytes -> roofline.traffic; source-level stall samples)
TAG=${1:-r02f}
mkdir -p gpurun_out
ARGS="--config c2 --scale 0.05 --steps 1 --warmup 3 --sub-batches 1 --contexts 1 --skip-cpu-baseline --skip-pipeline"
python bench.py $ARGS > gpurun_out/${TAG}_plain.json 2> /dev/null    # generates the dataset outside the profiler; the un-profiled numbers of the same command
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py $ARGS > gpurun_out/${TAG}_launches.log 2>&1
K="uvc_k0_read_consts|uvc_k1_prep_thres|uvc_k2_bias_pileup|uvc_kf_fragment_columns|uvc_k3b_fragment_consensus|uvc_k4_family_consensus|uvc_k5c_candidate_depths|uvc_k5e_candidate_quals"
ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip 24 --launch-count 8 -f -o gpurun_out/${TAG}_full python bench.py $ARGS > gpurun_out/${TAG}_full.log 2>&1
cp uvc_b200/lib/engine.o gpurun_out/${TAG}_engine.o 2>/dev/null   # tools/ncu_lines.py joins the report with the line table of exactly this build
ls -la gpurun_out | grep ${TAG}
