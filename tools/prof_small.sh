OUT=gpurun_out/${1:-r2v}; mkdir -p $OUT
NCU="ncu --clock-control none"
prof_conv() {
  python tools/prof_one.py conv $2 $3 $4 $5 $6 > $OUT/plain_$1.log 2>&1 &&
  $NCU --set full --import-source on -k regex:conv_gemm -s 2 -c 1 -o $OUT/prof_$1 -f python tools/prof_one.py conv $2 $3 $4 $5 $6 > $OUT/ncu_$1.log 2>&1
}
prof_conv c256k1_80 256 256 1 1 80
prof_conv c128k1_80 128 128 1 1 80
ls -la $OUT
