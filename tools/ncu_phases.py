"""Bucket the SASS-level samples / executed instructions of one ncu capture by barrier-delimited phase.
   python tools/ncu_phases.py <report.ncu-rep>"""
import csv, re, subprocess, sys, io
rep = sys.argv[1]
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
tote = sum(int(r[ix['Instructions Executed']]) for r in data)
print('samples', tot, 'warp-instr', tote)
start = 0; cs = 0; ce = 0
for i, r in enumerate(data):
    src = r[ix['Source']]
    cs += int(r[ix['# Samples']]); ce += int(r[ix['Instructions Executed']])
    if any(k in src for k in ['BAR.SYNC', 'UCGABAR_ARV', 'UCGABAR_WAIT', 'SYNCS.PHASECHK', 'EXIT']) or i == len(data) - 1:
        ops = {}
        for q in data[start:i + 1]:
            m = re.match(r'\s*(@!?U?P\d\s+)?([A-Z0-9_.]+)', q[ix['Source']])
            if m:
                o = m.group(2).split('.')[0]; ops[o] = ops.get(o, 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
        if cs or ce:
            print('%4d-%4d samples %5d (%4.1f%%) instr %9d (%4.1f%%) ends %-28s %s' % (start, i, cs, 100 * cs / tot, ce, 100 * ce / tote, src.strip()[:28], top))
        start = i + 1; cs = 0; ce = 0
if len(sys.argv) > 2:
    a, b = int(sys.argv[2]), int(sys.argv[3])
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    for r in sorted(data[a:b + 1], key=lambda r: -int(r[ix['# Samples']]))[:20]:
        st = sorted(((int(r[ix[h]]), h) for h in stalls), reverse=True)[:2]
        print('   %5s %-60s exec %8s %s' % (r[ix['# Samples']], r[ix['Source']].strip()[:60], r[ix['Instructions Executed']], st))
