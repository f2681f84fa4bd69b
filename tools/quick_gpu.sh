#!/bin/bash
# quick GPU pass: selected test files + bench ->  gpurun_out/<tag>/
#   gpurun --timeout 900 -- 'bash tools/quick_gpu.sh tag tests/test_gpu_nms.py ...'
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
python tools/gpu_check.py "$@" > $OUT/gpu_check.log 2>&1
cp gpurun_out/gpu_check_summary.txt $OUT/
for f in "$@"; do b=$(basename $f .py); cp gpurun_out/$b.log $OUT/ 2>/dev/null; done
if [ "${NO_BENCH:-0}" != "1" ]; then
  DMAY_LAYER_TABLE=$OUT/layers.json python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err
fi
cat $OUT/gpu_check_summary.txt; cat $OUT/bench.json 2>/dev/null | cut -c1-400
