#!/bin/bash
# One GPU-box pass: tests, bench, per-kernel bench, ncu launch list (+ DRAM bytes) of one warm step, and
# `ncu --set full` captures of the dominant conv shapes and of every memory-bound kernel.
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh r1g'
# Outputs land in gpurun_out/<tag>/ ; tools/summarize_profiles.py turns them into profiles/<tag>_*.md here.
set -u
TAG=${1:-r1x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1

if [ "${SKIP_TESTS:-0}" != "1" ]; then
  python tools/gpu_check.py > $OUT/gpu_check.log 2>&1
  cp gpurun_out/gpu_check_summary.txt $OUT/ 2>/dev/null
fi

DMAY_LAYER_TABLE=$OUT/layers.json python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
for c in cfg3 cfg4a cfg4b cfg5; do
  python bench.py --config $c --steps 5 --warmup 3 > $OUT/bench_$c.json 2> $OUT/bench_$c.err
done
python tools/bench_kernels.py --out $OUT/bench_kernels.json > $OUT/bench_kernels.log 2>&1

NCU="ncu --clock-control none"
# launch list of ONE warm step (profiler range), with DRAM bytes per launch
python tools/prof_one.py model > $OUT/plain_model.log 2>&1 &&
$NCU --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --csv --log-file $OUT/launches.csv python tools/prof_one.py model > $OUT/ncu_launches.log 2>&1

# the same for the other BASELINE configs
launch_list() {  # suffix, prof_one args...
  local sfx=$1; shift
  python tools/prof_one.py "$@" > $OUT/plain_$sfx.log 2>&1 &&
  $NCU --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --csv --log-file $OUT/launches_$sfx.csv python tools/prof_one.py "$@" > $OUT/ncu_launches_$sfx.log 2>&1
}
launch_list cfg3 model 8 yolov5l-ca-sppfcspc-bifpn-scconv.yaml 1536
launch_list cfg4a model 32 spdconv.yaml 1280
launch_list cfg4b model 32 C3CASPD.yaml 1280
launch_list cfg5 nms5 256

# full captures: memory-bound kernels of the warm step
$NCU --profile-from-start off --set full -c 24 \
    -k regex:'filter_|tile_|ca_|coordatt|pool_|adconcat|avgpool|upsample|nms_greedy|topk_|prep_kernel|img_' \
    -o $OUT/prof_membound -f python tools/prof_one.py model > $OUT/ncu_membound.log 2>&1

# full captures: representative conv shapes (isolated, third launch)
prof_conv() {  # name cin cout k s ho
  python tools/prof_one.py conv $2 $3 $4 $5 $6 > $OUT/plain_$1.log 2>&1 &&
  $NCU --set full --import-source on -k regex:conv_gemm -s 2 -c 1 -o $OUT/prof_$1 -f \
      python tools/prof_one.py conv $2 $3 $4 $5 $6 > $OUT/ncu_$1.log 2>&1
}
prof_conv c256k3_40 256 256 3 1 40
prof_conv c64k3_160 64 64 3 1 160
prof_conv c128k3_80 128 128 3 1 80
prof_conv c256k1_40 256 256 1 1 40
prof_conv c1024k3_20 1024 1024 3 1 20
prof_conv stem16k3_320 16 64 3 1 320
ls -la $OUT > $OUT/ls.txt
echo done
