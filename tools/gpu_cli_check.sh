#!/bin/bash
# Runs on the GPU box: full-size end-to-end comparison of the product uvc1 (B200) with the reference uvc1 (host cores) on a named synthetic config.
# usage: tools/gpu_cli_check.sh <config> <scale> <threads> [extra uvc1 options...]; writes gpurun_out/cli_<config>.txt
set -u
CFG=${1:-c1}; SCALE=${2:-1}; T=${3:-16}; shift 3 || true
OUT=gpurun_out/cli_${CFG}.txt
mkdir -p gpurun_out /tmp/uvc_cli
python - "$CFG" "$SCALE" > /tmp/uvc_cli/ds.json <<'PY'
import json, sys
sys.path.insert(0, ".")
import bench
ds = bench.dataset("/tmp/uvc_bench", sys.argv[1], float(sys.argv[2]))
print(json.dumps(ds))
PY
BAM=$(python -c "import json;print(json.load(open('/tmp/uvc_cli/ds.json'))['bam'])")
FA=$(python -c "import json;print(json.load(open('/tmp/uvc_cli/ds.json'))['fasta'])")
BED=$(python -c "import json;print(json.load(open('/tmp/uvc_cli/ds.json')).get('bed') or '')")
NREADS=$(python -c "import json;print(json.load(open('/tmp/uvc_cli/ds.json'))['n_reads'])")
RARG=""; if [ -n "$BED" ]; then RARG="-R $BED"; fi
{
echo "config=$CFG scale=$SCALE threads=$T reads=$NREADS nproc=$(nproc) bam=$(stat -c %s $BAM) bytes"
for rep in 1 2; do
  TIMEFORMAT="ours wall=%R s user=%U sys=%S"; time uvc_b200/bin/uvc1 $BAM -f $FA -o /tmp/uvc_cli/ours.vcf.gz -s S -t $T $RARG --bed-out-fname /tmp/uvc_cli/ours.bed --stats "$@" 2> /tmp/uvc_cli/ours.err; echo "ours rc=$?"; grep -E "uvc1-b200|Wall clock|ours wall" /tmp/uvc_cli/ours.err
done
TIMEFORMAT="ref wall=%R s user=%U sys=%S"; time oracle/_ref/uvc1 $BAM -f $FA -o /tmp/uvc_cli/ref.vcf.gz -s S -t $T $RARG --bed-out-fname /tmp/uvc_cli/ref.bed 2> /tmp/uvc_cli/ref.err; echo "ref rc=$?"; grep -E "Wall clock|CPU time|ref wall" /tmp/uvc_cli/ref.err
cmp /tmp/uvc_cli/ours.bed /tmp/uvc_cli/ref.bed && echo "BED identical ($(wc -l < /tmp/uvc_cli/ref.bed) tiles)"
zcat /tmp/uvc_cli/ours.vcf.gz | grep -v '^##' > /tmp/uvc_cli/ours.txt; zcat /tmp/uvc_cli/ref.vcf.gz | grep -v '^##' > /tmp/uvc_cli/ref.txt
echo "lines ours=$(wc -l < /tmp/uvc_cli/ours.txt) ref=$(wc -l < /tmp/uvc_cli/ref.txt) bytes ours=$(stat -c %s /tmp/uvc_cli/ours.txt) ref=$(stat -c %s /tmp/uvc_cli/ref.txt)"
if cmp -s /tmp/uvc_cli/ours.txt /tmp/uvc_cli/ref.txt; then echo "VCF body byte-identical"; else echo "VCF body differs: $(diff /tmp/uvc_cli/ours.txt /tmp/uvc_cli/ref.txt | grep -c '^<') lines"; diff /tmp/uvc_cli/ours.txt /tmp/uvc_cli/ref.txt | head -4 | cut -c1-600; fi
} > $OUT 2>&1
cat $OUT
