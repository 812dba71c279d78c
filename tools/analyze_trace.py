"""Attribute every GPU kernel of a torch-profiler chrome trace to its launching ATen op + input shapes."""
import bisect
import collections
import json
import sys

d = json.load(open(sys.argv[1]))
ev = d['traceEvents']
cpu_ops = sorted([e for e in ev if e.get('cat') in ('cpu_op', 'user_annotation')], key=lambda e: e['ts'])
launch = {e['args']['correlation']: e for e in ev if e.get('cat') in ('cuda_runtime', 'cuda_driver') and 'correlation' in e.get('args', {})}


def enclosing(e):
    t, tid, best = e['ts'], e['tid'], None
    for o in cpu_ops:
        if o['ts'] > t:
            break
        if o['tid'] == tid and o['ts'] + o['dur'] >= t:
            best = o
    return best


agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for e in ev:
    if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset'):
        l = launch.get(e['args'].get('correlation'))
        op = enclosing(l) if l else None
        key = (e['name'][:46], op['name'] if op else '?', str(op['args'].get('Input Dims', ''))[:70] if op else '')
        agg[key][0] += 1
        agg[key][1] += e['dur']
        tot += e['dur']
print(f'total GPU time {tot / 1e3:.3f} ms')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 45]:
    print(f'{t:8.1f} us {n:3d}x {k[0]:46s} | {k[1]:28s} {k[2]}')
