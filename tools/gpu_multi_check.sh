#!/bin/bash
# Runs on a multi-GPU box: the product uvc1 on ONE BAM with --gpus 1, 2, 4, 8 (as many as the box has): the tile list is sharded over the
# GPUs' lanes and the writer concatenates the VCF text in tile order (the reference's ordered flush, main.cpp:1541-1551), so the body must be
# byte-identical for every GPU count. usage: tools/gpu_multi_check.sh <config> <scale> <threads>; writes gpurun_out/multi_<config>.txt
set -u
CFG=${1:-c2}; SCALE=${2:-1}; T=${3:-16}
OUT=gpurun_out/multi_${CFG}.txt
mkdir -p gpurun_out /tmp/uvc_cli
python - "$CFG" "$SCALE" > /tmp/uvc_cli/ds.json <<'PY'
import json, sys
sys.path.insert(0, ".")
import bench
print(json.dumps(bench.dataset("/tmp/uvc_bench", sys.argv[1], float(sys.argv[2]))))
PY
BAM=$(python -c "import json;print(json.load(open('/tmp/uvc_cli/ds.json'))['bam'])")
FA=$(python -c "import json;print(json.load(open('/tmp/uvc_cli/ds.json'))['fasta'])")
BED=$(python -c "import json;print(json.load(open('/tmp/uvc_cli/ds.json')).get('bed') or '')")
NREADS=$(python -c "import json;print(json.load(open('/tmp/uvc_cli/ds.json'))['n_reads'])")
RARG=""; if [ -n "$BED" ]; then RARG="-R $BED"; fi
NGPU=$(nvidia-smi -L | wc -l)
{
echo "config=$CFG scale=$SCALE threads=$T reads=$NREADS nproc=$(nproc) gpus=$NGPU bam=$(stat -c %s $BAM) bytes"
for N in 1 2 4 8; do
  if [ $N -gt $NGPU ]; then continue; fi
  for rep in 1 2; do
    S=$(date +%s.%N)
    uvc_b200/bin/uvc1 $BAM -f $FA -o /tmp/uvc_cli/multi_$N.vcf.gz -s S -t $T $RARG --gpus $N --stats 2> /tmp/uvc_cli/multi_$N.err; RC=$?
    E=$(date +%s.%N)
    echo "gpus=$N rep=$rep rc=$RC wall=$(python -c "print('%.2f' % ($E - $S))") s reads/s=$(python -c "print('%.3g' % ($NREADS / ($E - $S)))")"
    grep -E "uvc1-b200" /tmp/uvc_cli/multi_$N.err | tail -3
  done
  zcat /tmp/uvc_cli/multi_$N.vcf.gz | grep -v '^##' | sha1sum | sed "s/^/gpus=$N body sha1 /"
done
} > $OUT 2>&1
cat $OUT
