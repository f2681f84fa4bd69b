#!/bin/bash
OUT=gpurun_out/r5i; mkdir -p $OUT
NCU="ncu --clock-control none"
python tools/prof_one.py convres 128 80 > $OUT/plain.log 2>&1 &&
$NCU --set full --import-source on -k regex:conv_gemm -s 2 -c 1 -o $OUT/prof_res128_80 -f python tools/prof_one.py convres 128 80 > $OUT/ncu1.log 2>&1
$NCU --set full --import-source on -k regex:conv_gemm -s 2 -c 1 -o $OUT/prof_plain128_80 -f python tools/prof_one.py conv 128 128 3 1 80 > $OUT/ncu2.log 2>&1
ls $OUT
