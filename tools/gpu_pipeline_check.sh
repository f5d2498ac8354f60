#!/bin/bash
# Runs on the GPU box: the whole uvc1 program on the bench BAM of configs[1], a few times in a row (process start to exit, with --stats)
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --skip-pipeline > /dev/null 2>&1
D=$(ls -d /tmp/uvc_bench/c2_* | head -1)
for i in 1 2 3 4; do
  t0=$(date +%s.%N)
  uvc_b200/bin/uvc1 $D/c2.bam -f $D/c2.fa -o /tmp/o$i.vcf.gz -s S -t 16 -R $D/c2.bed --gpus 1 --stats 2> /tmp/e$i.txt > /dev/null
  t1=$(date +%s.%N)
  echo "run $i wall $(echo "$t1 - $t0" | bc) s"; grep -v "^$" /tmp/e$i.txt | tail -14
  echo ---
done
