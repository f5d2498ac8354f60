#!/bin/bash
# Runs on the GPU box: the whole uvc1 program on the bench BAM of configs[1] (process start to exit, with --stats), twice per lane count
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --skip-pipeline > /dev/null 2>&1
D=$(ls -d /tmp/uvc_bench/c2_* | head -1)
for L in ${@:-3}; do
  for i in 1 2; do
    uvc_b200/bin/uvc1 $D/c2.bam -f $D/c2.fa -o /tmp/o.vcf.gz -s S -t 16 -R $D/c2.bed --gpus 1 --lanes-per-gpu $L --stats 2> /tmp/e.txt > /dev/null
    echo "lanes $L run $i: $(grep 'Wall clock' /tmp/e.txt) | $(grep 'stage seconds' /tmp/e.txt | cut -c1-140) | $(grep 'timeline' /tmp/e.txt | cut -c30-200)"
  done
done
