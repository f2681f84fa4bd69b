#!/bin/bash
# membound-only pass: tests for the touched kernels + launch list + full captures of the small kernels
TAG=${1:-r1x}; OUT=gpurun_out/$TAG; mkdir -p $OUT
NO_BENCH=1 bash tools/quick_gpu.sh $TAG tests/test_gpu_pool_ca.py tests/test_gpu_nms.py tests/test_gpu_model.py
NCU="ncu --clock-control none"
python tools/prof_one.py model > $OUT/plain_model.log 2>&1 &&
$NCU --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --csv --log-file $OUT/launches.csv python tools/prof_one.py model > $OUT/ncu_launches.log 2>&1
python - <<PY
import csv,io,collections
t=open('$OUT/launches.csv').read().splitlines()
st=next(i for i,l in enumerate(t) if l.startswith('"ID"'))
agg=collections.defaultdict(lambda:[0,0.0])
for r in csv.DictReader(io.StringIO('\n'.join(t[st:]))):
    if r['Metric Name']=='gpu__time_duration.sum':
        n=r['Kernel Name'].replace('void ','').replace('dmay::','').split('(')[0][:40]
        agg[n][0]+=1; agg[n][1]+=float(r['Metric Value'].replace(',',''))/1e3
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print('%-42s n=%3d us=%.1f'%(k,v[0],v[1]))
PY
