#!/bin/bash
OUT=gpurun_out/r5g; mkdir -p $OUT
NCU="ncu --clock-control none"
prof_conv() {  # name cin cout k s ho
  python tools/prof_one.py conv $2 $3 $4 $5 $6 > $OUT/plain_$1.log 2>&1 &&
  $NCU --set full --import-source on -k regex:conv_gemm -s 2 -c 1 -o $OUT/prof_$1 -f \
      python tools/prof_one.py conv $2 $3 $4 $5 $6 > $OUT/ncu_$1.log 2>&1
}
prof_conv stem 16 64 3 1 320
prof_conv c64k3_320 64 64 3 1 320
prof_conv c512k1_20 512 512 1 1 20
prof_conv c256k1_40 256 256 1 1 40
ls -la $OUT
