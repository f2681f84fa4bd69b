#!/bin/bash
# Same-box A/B of the fused decode + filter kernels (rows kernel, tile scan, tile gather) over builds of libdmayolo.so:
# ncu durations inside one warm cfg-2 step.
#   gpurun -- 'bash tools/ab_filter.sh gpurun_out/r6b dma_yolo_b200/libdmayolo.so dma_yolo_b200/libdmayolo_v1.so ...'
OUT=$1; shift
mkdir -p $OUT
python tools/prof_one.py model > /dev/null 2>&1
for round in 1 2; do
  for so in "$@"; do
    tag=$(basename $so .so)
    DMAY_SO=$so ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum \
      -k regex:"filter_fused_rows|tile_scan|tile_gather" -c 3 \
      --csv --log-file $OUT/f_${tag}_$round.csv python tools/prof_one.py model > /dev/null 2>&1
    echo "$tag round $round: $(grep gpu__time_duration $OUT/f_${tag}_$round.csv | python -c "
import sys,csv
for r in csv.reader(sys.stdin):
    print(r[4].split('(')[0].replace('void ','')[:34], r[-1], end=' | ')
")"
  done
done
