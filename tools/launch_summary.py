"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv).  usage: launch_summary.py file.csv [skip_first_n_steps]"""
import collections
import csv
import re
import sys


def main(path, last_step_only=True):
    rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 5]
    hdr = rows[0]
    ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
    items = []
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        n = re.sub(r'^void ', '', r[ki]).replace('t3d::', '')
        n = re.sub(r'\(.*', '', n)
        items.append((n, v))
    if last_step_only:
        # the bench runs warm-up step(s) then the timed one: keep everything after the second-to-last adam_kernel
        idx = [i for i, (n, _) in enumerate(items) if n.startswith('adam_kernel')]
        if len(idx) >= 2:
            items = items[idx[-2] + 1: idx[-1] + 1]
    tot, cnt = collections.Counter(), collections.Counter()
    for n, v in items:
        tot[n] += v
        cnt[n] += 1
    s = sum(tot.values())
    print('launches %d  total %.1f us' % (len(items), s / 1e3))
    for n, v in tot.most_common(30):
        print('%9.1f us %5.1f%% x%-3d %s' % (v / 1e3, 100 * v / s, cnt[n], n[:100]))


if __name__ == '__main__':
    main(sys.argv[1])
