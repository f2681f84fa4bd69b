"""Turn gpurun_out/<tag>/ (tools/profile_round.sh) into tracked summaries under profiles/.

    python tools/summarize_profiles.py r1g

  profiles/<tag>_launches.md   per-kernel totals of ONE warm step (ncu gpu__time_duration + DRAM bytes per launch)
  profiles/<tag>_launches.csv  the raw launch list (small)
  profiles/<tag>_ncu_<name>.md key `ncu --set full` metrics per captured launch
"""
import collections
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1]
SRC = ROOT / 'gpurun_out' / tag
DST = ROOT / 'profiles'
DST.mkdir(exist_ok=True)


def read_ncu_csv(text):
    lines = text.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(io.StringIO('\n'.join(lines[start:]))))


def short(name):
    n = name.replace('void ', '').replace('dmay::', '')
    return n.split('(')[0][:70]


def launches(suffix='', what='cfg-2, batch 64, forward + fused decode/filter + NMS'):
    f = SRC / f'launches{suffix}.csv'
    if not f.exists():
        return
    rows = read_ncu_csv(f.read_text())
    per = collections.OrderedDict()
    for r in rows:
        k = r['ID']
        d = per.setdefault(k, dict(name=short(r['Kernel Name']), grid=r['Grid Size'], block=r['Block Size']))
        v = float(r['Metric Value'].replace(',', ''))
        u = r['Metric Unit']
        if r['Metric Name'] == 'gpu__time_duration.sum':
            d['us'] = v / 1e3 if u in ('ns', 'nsecond') else v if u in ('us', 'usecond') else v * 1e3
        else:
            mul = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
            d[r['Metric Name']] = v * mul
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d['name'], dict(n=0, us=0.0, rd=0.0, wr=0.0))
        a['n'] += 1
        a['us'] += d.get('us', 0)
        a['rd'] += d.get('dram__bytes_read.sum', 0)
        a['wr'] += d.get('dram__bytes_write.sum', 0)
    tot = sum(a['us'] for a in agg.values())
    out = [f'# {tag}{suffix}: launch list of ONE warm step ({what})', '',
           'Source: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none`',
           'around the second step of `tools/prof_one.py model` (profiler range). Times are serialised and cold-cache: read the',
           'SHARES, not the absolutes.', '',
           f'Total kernel time {tot / 1e3:.3f} ms over {sum(a["n"] for a in agg.values())} launches.', '',
           '| kernel | launches | total us | share | DRAM read MB | DRAM write MB | DRAM GB/s |', '|---|---|---|---|---|---|---|']
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
        gbs = (a['rd'] + a['wr']) / a['us'] / 1e3 if a['us'] else 0
        out.append(f"| `{k}` | {a['n']} | {a['us']:.1f} | {100 * a['us'] / tot:.1f}% | {a['rd'] / 1e6:.1f} | {a['wr'] / 1e6:.1f} | {gbs:.0f} |")
    (DST / f'{tag}{suffix}_launches.md').write_text('\n'.join(out) + '\n')
    with open(DST / f'{tag}{suffix}_launches.csv', 'w') as fo:
        fo.write('id,kernel,grid,block,us,dram_read_bytes,dram_write_bytes\n')
        for k, d in per.items():
            fo.write(f"{k},{d['name']},\"{d['grid']}\",\"{d['block']}\",{d.get('us', 0):.2f},{d.get('dram__bytes_read.sum', 0):.0f},{d.get('dram__bytes_write.sum', 0):.0f}\n")
    json.dump({k: a for k, a in agg.items()}, open(DST / f'{tag}{suffix}_launches.json', 'w'), indent=1)
    print('\n'.join(out[:14]))


KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size', 'sm__cycles_active.avg',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum', 'launch__occupancy_limit_shared_mem',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum',
        'smsp__cycles_active.avg', 'sm__inst_executed_pipe_uniform.sum', 'lts__t_sectors_srcunit_tex_op_read.sum']


def full(rep):
    r = subprocess.run(['ncu', '-i', str(rep), '--page', 'raw', '--csv'], capture_output=True, text=True)
    if r.returncode != 0:
        print('ncu failed for', rep, r.stderr[:300])
        return
    rows = list(csv.reader(io.StringIO(r.stdout)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    tensor_cols = [h for h in hdr if 'tensor' in h and ('pct' in h or 'cycles_active' in h)]
    cols = [k for k in KEYS if k in idx] + [h for h in tensor_cols if h not in KEYS][:8]
    out = [f'# {tag}: `ncu --set full --clock-control none` — {rep.name}', '']
    for d in data:
        out.append(f"## {short(d[idx['Kernel Name']])}  grid {d[idx['Grid Size']]} block {d[idx['Block Size']]}")
        out.append('')
        out.append('| metric | value | unit |')
        out.append('|---|---|---|')
        for c in cols:
            out.append(f'| {c} | {d[idx[c]]} | {units[idx[c]]} |')
        try:
            t = float(d[idx['gpu__time_duration.sum']].replace(',', ''))
            tu = units[idx['gpu__time_duration.sum']]
            t_s = t * {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 'nsecond': 1e-9, 'usecond': 1e-6, 'msecond': 1e-3}.get(tu, 1e-9)
            def b(k):
                v = float(d[idx[k]].replace(',', ''))
                return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(units[idx[k]], 1)
            tr = b('dram__bytes_read.sum') + b('dram__bytes_write.sum')
            out.append(f'| **derived: DRAM traffic** | {tr / 1e6:.2f} | MB |')
            out.append(f'| **derived: DRAM GB/s** | {tr / t_s / 1e9:.0f} | GB/s |')
        except Exception as e:  # noqa
            pass
        out.append('')
    name = rep.stem.replace('prof_', '')
    (DST / f'{tag}_ncu_{name}.md').write_text('\n'.join(out) + '\n')
    print(f'wrote profiles/{tag}_ncu_{name}.md ({len(data)} launches)')


launches()
launches('_cfg3', 'cfg-3 yolov5l-ca-sppfcspc-bifpn-scconv, 1536x1536, batch 8')
launches('_cfg4a', 'cfg-4a spdconv, 1280x1280, batch 32')
launches('_cfg4b', 'cfg-4b C3CASPD, 1280x1280, batch 32')
launches('_cfg5', 'cfg-5 NMS stress: non_max_suppression(rand(256,25200,15)), val-style')
for rep in sorted(SRC.glob('*.ncu-rep')):
    full(rep)
for f in ('bench.json', 'bench_cfg3.json', 'bench_cfg4a.json', 'bench_cfg4b.json', 'bench_cfg5.json', 'bench_kernels.json', 'layers.json',
          'gpu_check_summary.txt', 'scale.json'):
    if (SRC / f).exists():
        (DST / f'{tag}_{f}').write_text((SRC / f).read_text())
