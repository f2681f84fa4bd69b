#!/bin/bash
# Round-2 GPU pass: all GPU tests (one process per file) + bench.py for every BASELINE config -> gpurun_out/<tag>/
#   gpurun --timeout 1500 -- 'bash tools/round2_gpu.sh r5a'
TAG=${1:-r5a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  python tools/gpu_check.py ${TEST_FILES:-} > $OUT/gpu_check.log 2>&1
  cp gpurun_out/gpu_check_summary.txt $OUT/
  for f in gpurun_out/test_gpu_*.log; do cp $f $OUT/ 2>/dev/null; done
fi
DMAY_LAYER_TABLE=$OUT/layers.json python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
for c in ${CONFIGS:-cfg5 cfg3 cfg4a cfg4b}; do
  python bench.py --config $c --steps 5 --warmup 3 > $OUT/bench_$c.json 2> $OUT/bench_$c.err
done
cat $OUT/gpu_check_summary.txt 2>/dev/null
for f in $OUT/bench*.json; do echo $f; cut -c1-600 $f; done
tail -5 $OUT/bench*.err
