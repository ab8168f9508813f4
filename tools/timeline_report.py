"""Wall-time attribution of a timeline written by tools/step_timeline.py: kernels sorted by END time, each one is
charged the time by which it extends the busy front (what the step would save if that kernel vanished, to first order)."""
import collections
import json
import re
import sys


def short(n):
    n = n.replace('void ', '').replace('(anonymous namespace)::', '')
    return re.split(r'[(]', n)[0][:56]


def main(path, top=40):
    d = json.load(open(path))
    rows = d['kernels']
    evs = sorted(rows, key=lambda r: r['t_us'] + r['dur_us'])
    last, agg = 0.0, collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in evs:
        e = r['t_us'] + r['dur_us']
        a = max(0.0, e - max(last, r['t_us']))
        last = max(last, e)
        k = short(r['name'])
        agg[k][0] += 1
        agg[k][1] += a
        agg[k][2] += r['dur_us']
    tot = sum(v[1] for v in agg.values())
    # the profiler window may hold more than one replay of the step: report PER STEP (optimizer launches = replays)
    steps = max(1, sum(1 for r in rows if 'adamw_kernel' in r['name']))
    print('# %s: clean %.3f ms/step, %d kernels in %d replay(s), attributed %.1f us per step' % (
        path, d['summary']['clean_ms_per_step'], len(rows), steps, tot / steps))
    print('%-58s %5s %10s %6s %10s   (per step)' % ('kernel', 'n', 'attrib_us', 'share', 'sum_dur_us'))
    for k, (n, a, dsum) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print('%-58s %5.0f %10.1f %5.1f%% %10.1f' % (k, n / steps, a / steps, 100 * a / tot, dsum / steps))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
