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
    print('# %s: clean %.3f ms/step, %d kernels, attributed %.1f us' % (path, d['summary']['clean_ms_per_step'], len(rows), tot))
    print('%-58s %5s %10s %6s %10s' % ('kernel', 'n', 'attrib_us', 'share', 'sum_dur_us'))
    for k, (n, a, dsum) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print('%-58s %5d %10.1f %5.1f%% %10.1f' % (k, n, a, 100 * a / tot, dsum))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
