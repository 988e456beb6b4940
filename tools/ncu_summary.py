#!/usr/bin/env python
"""Summarise an ncu launch-list CSV (``--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv``)
per kernel: launches, grid, average duration, share of the summed kernel time, DRAM bytes per launch.

    python tools/ncu_summary.py gpurun_out/launches_<tag>.csv [--json profiles/<name>.json --commit <hash> --words <plane words>]

``--json`` writes the per-kernel DRAM traffic in the form ``bench.py`` reads for ``roofline.traffic`` (bytes per launch and,
with ``--words``, bytes per plane word) together with the commit the capture was taken at, so the figure cannot silently go
stale: bench.py reports the commit next to it."""
import argparse
import collections
import csv
import json
import re
import sys


def short(name):
    name = re.sub(r'^void\s+', '', name)
    name = re.sub(r'^kt::', '', name)
    m = re.match(r'([\w:]+(?:<[^(]*>)?)', name)
    return (m.group(1) if m else name).replace('kt::', '')


def load(path):
    rows = collections.OrderedDict()
    with open(path, newline='') as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        key = int(r['ID'])
        d = rows.setdefault(key, {'name': short(r['Kernel Name']), 'grid': r['Grid Size'], 'block': r['Block Size']})
        try:
            d[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
        except ValueError:
            pass
        d['unit_' + r['Metric Name']] = r['Metric Unit']
    return list(rows.values())


def to_ns(d):
    v, u = d.get('gpu__time_duration.sum', 0.0), d.get('unit_gpu__time_duration.sum', 'ns')
    return v * {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(u, 1.0)


def to_bytes(d, key):
    v, u = d.get(key, 0.0), d.get('unit_' + key, 'byte')
    return v * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('csv')
    ap.add_argument('--json')
    ap.add_argument('--commit', default='')
    ap.add_argument('--words', type=float, default=0.0, help='words per plane (rows x pitch) of the profiled workload')
    ap.add_argument('--command', default='')
    a = ap.parse_args()
    launches = load(a.csv)
    # launches behind the solver's `done` flag return at once (the host enqueues whole chunks of 16 iterations): they move no
    # bytes and would dilute the per-launch averages -- dropped when they last < 1/10 of the kernel's median launch
    by_kernel = collections.defaultdict(list)
    for d in launches:
        by_kernel[(d['name'], d['grid'])].append(to_ns(d))
    med = {k: sorted(v)[len(v) // 2] for k, v in by_kernel.items()}
    kept = [d for d in launches if to_ns(d) >= 0.1 * med[(d['name'], d['grid'])]]
    dropped = len(launches) - len(kept)
    if dropped:
        print(f'({dropped} of {len(launches)} launches were no-ops behind the done flag and are left out)\n')
    launches = kept
    agg = collections.OrderedDict()
    for d in launches:
        g = agg.setdefault((d['name'], d['grid']), {'n': 0, 'ns': 0.0, 'rd': 0.0, 'wr': 0.0})
        g['n'] += 1
        g['ns'] += to_ns(d)
        g['rd'] += to_bytes(d, 'dram__bytes_read.sum')
        g['wr'] += to_bytes(d, 'dram__bytes_write.sum')
    total = sum(g['ns'] for g in agg.values()) or 1.0
    print('| kernel | launches | grid | avg ms | share | DRAM rd GB/launch | DRAM wr GB/launch |' + (' traffic B/word |' if a.words else ''))
    print('|---|---|---|---|---|---|---|' + ('---|' if a.words else ''))
    out = {}
    for (name, grid), g in sorted(agg.items(), key=lambda kv: -kv[1]['ns']):
        n = g['n']
        line = f"| `{name}` | {n} | {grid} | {g['ns'] / n / 1e6:.4f} | {g['ns'] / total:.3f} | {g['rd'] / n / 1e9:.3f} | {g['wr'] / n / 1e9:.3f} |"
        if a.words:
            line += f" {(g['rd'] + g['wr']) / n / a.words:.1f} |"
        print(line)
        e = out.setdefault(name, {'launches': 0, 'ns': 0.0, 'bytes': 0.0})
        e['launches'] += n; e['ns'] += g['ns']; e['bytes'] += g['rd'] + g['wr']
    if a.json:
        doc = {'commit': a.commit, 'command': a.command, 'source_csv': a.csv, 'plane_words': a.words or None, 'kernels': {}}
        for name, e in out.items():
            doc['kernels'][name] = {'launches': e['launches'], 'avg_ms': e['ns'] / e['launches'] / 1e6, 'share': e['ns'] / total,
                                    'dram_bytes_per_launch': e['bytes'] / e['launches'],
                                    'dram_bytes_per_word': (e['bytes'] / e['launches'] / a.words) if a.words else None}
        with open(a.json, 'w') as f:
            json.dump(doc, f, indent=1)
        print(f'wrote {a.json}', file=sys.stderr)


if __name__ == '__main__':
    main()
