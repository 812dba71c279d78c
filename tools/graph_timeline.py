"""Timeline of one CUDA-graph replay from a torch-profiler chrome trace (tools/profile_step.py, SGC_GRAPH_TRACE):
per-stream busy time, the union busy time, idle gaps, and the kernel sequence with start offsets.

    python tools/graph_timeline.py trace.json [min_us_to_list] [all_events_out.txt]
"""
import collections
import json
import sys

d = json.load(open(sys.argv[1]))
lim = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
ev = sorted([e for e in d['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset')], key=lambda e: e['ts'])
t0 = ev[0]['ts']
t1 = max(e['ts'] + e['dur'] for e in ev)
print(f'{len(ev)} device activities, span {(t1 - t0) / 1e3:.3f} ms, sum of durations {sum(e["dur"] for e in ev) / 1e3:.3f} ms')
per = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    s = e['args'].get('stream')
    per[s][0] += 1
    per[s][1] += e['dur']
for s, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
    print(f'  stream {s}: {n} launches, busy {t / 1e3:.3f} ms')
# union of busy intervals
iv = sorted((e['ts'], e['ts'] + e['dur']) for e in ev)
busy, gaps, cur_s, cur_e = 0.0, [], iv[0][0], iv[0][1]
for a, b in iv[1:]:
    if a > cur_e:
        busy += cur_e - cur_s
        gaps.append((cur_e - t0, a - cur_e))
        cur_s, cur_e = a, b
    else:
        cur_e = max(cur_e, b)
busy += cur_e - cur_s
print(f'union busy {busy / 1e3:.3f} ms, idle {sum(g for _, g in gaps) / 1e3:.3f} ms in {len(gaps)} gaps '
      f'(mean {sum(g for _, g in gaps) / max(1, len(gaps)):.2f} us)')
# time with exactly one small kernel running (latency-bound segments): bucket the span in 50 us bins
print('kernels >= %.0f us:' % lim)
for e in ev:
    if e['dur'] >= lim:
        print(f'  {e["ts"] - t0:9.1f} +{e["dur"]:7.1f}  s{e["args"].get("stream")}  {e["name"][:70]}')
main = max(per.items(), key=lambda kv: kv[1][0])[0]
print(f'main stream {main}: sequence (start, dur, gap-before)')
prev = None
for e in ev:
    if e['args'].get('stream') != main:
        continue
    gap = e['ts'] - prev if prev is not None else 0.0
    print(f'  {e["ts"] - t0:9.1f} {e["dur"]:7.1f} {gap:6.1f}  {e["name"][:64]}')
    prev = e['ts'] + e['dur']

if len(sys.argv) > 3:
    # every device activity: start, duration, stream, name (for dependency / overlap analysis offline)
    with open(sys.argv[3], 'w') as f:
        for e in ev:
            f.write(f'{e["ts"] - t0:9.1f} {e["dur"]:7.1f} s{e["args"].get("stream")} {e["name"][:90]}\n')
