OUT=gpurun_out/r1p; mkdir -p $OUT
NCU="ncu --clock-control none"
prof_conv() {
  python tools/prof_one.py conv $2 $3 $4 $5 $6 > $OUT/plain_$1.log 2>&1 &&
  $NCU --set full --import-source on -k regex:conv_gemm -s 2 -c 1 -o $OUT/prof_$1 -f python tools/prof_one.py conv $2 $3 $4 $5 $6 > $OUT/ncu_$1.log 2>&1
}
prof_conv c64k3_160 64 64 3 1 160
prof_conv c256k1_40 256 256 1 1 40
prof_conv c256k3_40 256 256 3 1 40
ls -la $OUT
