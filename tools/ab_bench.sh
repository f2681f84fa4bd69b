#!/bin/bash
# A/B on ONE box: alternates two builds of libdmayolo.so (DMAY_SO) through bench.py, 2 rounds each.
#   gpurun -- 'bash tools/ab_bench.sh /path/a.so /path/b.so [steps]'
A=$1; B=$2; STEPS=${3:-20}
for r in 1 2; do
  for so in $A $B; do
    DMAY_SO=$so timeout 300 python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline --no-latency 2>/dev/null | \
      python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$so'.split('/')[-1], j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'], 'conv TF/s', j['roofline']['achieved'], 'mhz', j['clocks']['sm_mhz'])"
  done
done
