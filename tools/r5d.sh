#!/bin/bash
OUT=gpurun_out/r5d; mkdir -p $OUT
python tools/gpu_check.py tests/test_gpu_model.py tests/test_gpu_nms.py > $OUT/gpu_check.log 2>&1; cp gpurun_out/gpu_check_summary.txt $OUT/
for v in auto reserve1 old; do
  case $v in auto) E="";; reserve1) E="DMAY_DENSE_RESERVE=1";; old) E="DMAY_DENSE_ROWS=0";; esac
  env $E python bench.py --config cfg5 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/cfg5_$v.json 2> $OUT/cfg5_$v.err
done
python tools/prof_one.py model > $OUT/plain_model.log 2>&1 &&
ncu --clock-control none --profile-from-start off --set full --import-source on -k regex:'filter_fused_rows|tile_gather|tile_scan' -c 3 \
    -o $OUT/prof_filter -f python tools/prof_one.py model > $OUT/ncu_filter.log 2>&1
ncu --clock-control none --profile-from-start off --set full --import-source on -k regex:'ca_pool_hidden_mma|ca_gate_apply_mma' -c 2 \
    -o $OUT/prof_ca -f python tools/prof_one.py model > $OUT/ncu_ca.log 2>&1
cat $OUT/gpu_check_summary.txt; for f in $OUT/cfg5_*.json; do echo $f; python -c "
import json,sys; d=json.load(open('$f')); print(d['ms_per_step'], d['detect_style']['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items()})"; done
ls -la $OUT
