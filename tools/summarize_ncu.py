"""Summarise ncu outputs into profiles/ (tracked).  Usage:
   python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches_summary.md
   python tools/summarize_ncu.py full gpurun_out/prof_lift_r1.ncu-rep profiles/r1_lift_full_summary.md"""
import collections
import csv
import re
import subprocess
import sys

mode, src, dst = sys.argv[1:4]
out = []
if mode == 'launches':
    lines = [l for l in open(src) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in rows:
        v = float(r['Metric Value'].replace(',', ''))
        u = r['Metric Unit']
        ms = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v
        name = re.sub(r'\(.*', '', r['Kernel Name'])
        name = re.sub(r'^void ', '', name)[:96]
        agg[name][0] += 1
        agg[name][1] += ms
        tot += ms
    out.append(f'# ncu launch list summary ({src})\n')
    out.append('`ncu --metrics gpu__time_duration.sum --clock-control none` over a window of consecutive launches of '
               '`bench.py --no-graph` (cold-cache, serialised: compare SHARES, not absolutes).\n')
    out.append(f'window: {len(rows)} launches, {tot:.3f} ms total\n')
    own = sum(ms for k, (n, ms) in agg.items() if k.startswith(('sgc::', 'tc::')) or 'sgc::' in k)
    out.append(f'own kernels (sgc::*, sgc::tc::*): {own:.3f} ms = {100 * own / tot:.1f} % of the window\n')
    out.append('| share | total ms | launches | kernel |\n|---|---|---|---|')
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        out.append(f'| {100 * ms / tot:5.1f} % | {ms:.3f} | {n} | `{k}` |')
else:
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
            'smsp__inst_executed.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
    out.append(f'# ncu --set full summary ({src})\n')
    out.append('One row per captured launch (`--clock-control none`, ~40 replays per launch: durations are not bench values).\n')
    out.append('| # | kernel | ' + ' | '.join(w.split('.')[0].replace('__', ' ') for w in want if w in idx) + ' |')
    out.append('|---|---|' + '---|' * len([w for w in want if w in idx]))
    for i, d in enumerate(data):
        name = re.sub(r'\(.*', '', d[idx['Kernel Name']]).replace('void ', '')
        vals = [f'{d[idx[w]]} {units[idx[w]]}'.strip() for w in want if w in idx]
        out.append(f'| {i} | `{name}` | ' + ' | '.join(vals) + ' |')
if mode == 'full' and len(sys.argv) > 4:
    # per-launch DRAM traffic of the finest-level (= longest) instance of each hot kernel, keyed like bench.py's
    # `roofline.kernel` for the headline shape (SGCDet_ScanNet, V=40): read back by bench.py as `roofline.traffic`
    import json
    keys = {'lift_bwd_kernel': 'sgc_lift_bwd[Q=6400]', 'lift_fwd_kernel': 'sgc_lift_fwd[Q=6400]',
            'project_tc_kernel<0>': 'sgc_project_tc_fwd[V=40,C=256,S=4720,N=384]'}
    best = {}

    def _bytes(v, u):
        v = float(v.replace(',', ''))
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    for d in data:
        name = d[idx['Kernel Name']]
        for frag, key in keys.items():
            if frag in name.replace('false', '0').replace('true', '1'):
                dur = float(d[idx['gpu__time_duration.sum']].replace(',', ''))
                tr = _bytes(d[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']]) + \
                    _bytes(d[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
                if key not in best or dur > best[key][0]:
                    best[key] = (dur, int(tr))
    json.dump({k: v[1] for k, v in best.items()}, open(sys.argv[4], 'w'), indent=1)
open(dst, 'w').write('\n'.join(out) + '\n')
print(open(dst).read()[:3000])
