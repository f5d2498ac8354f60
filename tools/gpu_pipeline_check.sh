mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --skip-pipeline > /dev/null 2>&1
D=$(ls -d /tmp/uvc_bench/c2_* | head -1)
ls $D | head
for i in 1 2 3 4; do
  /usr/bin/time -f "wall %e s" uvc_b200/bin/uvc1 $D/c2.bam -f $D/c2.fa -o /tmp/o$i.vcf.gz -s S -t 16 -R $D/c2.bed --gpus 1 --stats 2>&1 | grep -i "wall\|lane\|fetch\|decode\|prep\|gpu\|score\|text\|compress\|batches\|start" | head -12
  echo ---
done
