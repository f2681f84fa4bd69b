#!/bin/bash
OUT=gpurun_out/${1:-r2j}; mkdir -p $OUT
CMD="python tools/prof_one.py model ${2:-8} ${3:-yolov5l-ca-sppfcspc-bifpn-scconv.yaml} ${4:-1536}"
$CMD > $OUT/plain_cfg3.log 2>&1 &&
ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --csv --log-file $OUT/launches_cfg3.csv $CMD > $OUT/ncu_cfg3.log 2>&1
python - <<PY
import csv,io,collections
t=open('$OUT/launches_cfg3.csv').read().splitlines()
st=next(i for i,l in enumerate(t) if l.startswith('"ID"'))
agg=collections.defaultdict(lambda:[0,0.0])
for r in csv.DictReader(io.StringIO('\n'.join(t[st:]))):
    if r['Metric Name']=='gpu__time_duration.sum':
        n=r['Kernel Name'].replace('void ','').replace('dmay::','').split('(')[0][:44]
        agg[n][0]+=1; agg[n][1]+=float(r['Metric Value'].replace(',',''))/1e3
tot=sum(v[1] for v in agg.values())
print('total us', tot)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:18]: print('%-46s n=%3d us=%9.1f %5.1f%%'%(k,v[0],v[1],100*v[1]/tot))
PY
