"""Summarises an ncu launch list (CSV written with
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file X.csv \
        python bench.py --steps 1 --warmup 3 --no-extra --no-e2e --no-cpu-baseline [--workload cX]
) into (a) a per-kernel table of the LAST evaluation step (launches, time, DRAM bytes) and (b) profiles/r2_traffic.json,
the record bench.py's roofline.traffic is read from (DRAM bytes per launch of the pair-level contraction kernels).

    python tools/ncu_traffic.py X.csv <workload> [--write]
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(path):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ik, im, iv, iid, iu = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID', 'Metric Unit'))
    ig = hdr.index('Grid Size')
    launches = collections.OrderedDict()
    for row in r:
        d = launches.setdefault(int(row[iid]), {'name': row[ik], 'grid': row[ig]})
        v = float(row[iv].replace(',', ''))
        unit = row[iu].lower()
        scale = {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3, 'second': 1e6,
                 'byte': 1.0, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(unit, 1.0)
        d[row[im]] = v * scale
    return list(launches.values())


def main():
    path, workload = sys.argv[1], sys.argv[2]
    L = load(path)
    starts = [i for i, d in enumerate(L) if 'k_sys_ptr' in d['name']]
    step = L[starts[-1]:]
    tot = collections.OrderedDict()
    for d in step:
        name = d['name'].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
        if 'k_gemm128' in d['name'] or 'k_message_tc' in d['name'] or 'k_node_aggregate' in d['name'] or 'k_pair_bwd' in d['name']:
            name = d['name'][:d['name'].index('(')].replace('void ', '').replace('<unnamed>::', '')
        t = tot.setdefault(name, [0, 0.0, 0.0, 0.0])
        t[0] += 1
        t[1] += d.get('gpu__time_duration.sum', 0.0)
        t[2] += d.get('dram__bytes_read.sum', 0.0)
        t[3] += d.get('dram__bytes_write.sum', 0.0)
    total_us = sum(t[1] for t in tot.values())
    total_b = sum(t[2] + t[3] for t in tot.values())
    print(f'# last evaluation step of {os.path.basename(path)}: {len(step)} launches, {total_us / 1e3:.2f} ms summed kernel time '
          f'(ncu: serialised, cold caches), {total_b / 1e9:.2f} GB DRAM traffic')
    print(f'{"kernel":58s} {"n":>4s} {"us":>10s} {"share":>6s} {"GB read":>8s} {"GB write":>8s} {"GB/s":>7s}')
    for name, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f'{name[:58]:58s} {t[0]:4d} {t[1]:10.1f} {100 * t[1] / total_us:5.1f}% {t[2] / 1e9:8.3f} {t[3] / 1e9:8.3f} '
              f'{(t[2] + t[3]) / max(t[1], 1e-9) * 1e-3:7.0f}')
    pair = [d for d in step if ('k_gemm128_chain' in d['name'] or 'k_gemm128_ts' in d['name']) and int(d['grid'].strip('()').split(',')[0]) >= 140]
    # pair-level launches run on the full grid (one CTA per SM); node-level ones on 3N rows do too - a pair-level launch moves
    # at least 1 KB per pair, i.e. it is within a factor ~3 of the largest launch, node-level ones are >= 9x smaller
    byt = lambda d: d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
    top = max((byt(d) for d in pair), default=0.0)
    pair = [d for d in pair if byt(d) > 0.3 * top]
    if pair and '--write' in sys.argv:
        commit = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], cwd=ROOT, capture_output=True, text=True).stdout.strip()
        out = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
        rec = json.load(open(out)) if os.path.exists(out) else {}
        b = sum(d['dram__bytes_read.sum'] + d['dram__bytes_write.sum'] for d in pair)
        rec[workload] = {'kernel_class': 'pair_gemm', 'launches': len(pair), 'dram_bytes_per_launch': b / len(pair),
                         'dram_bytes_per_step': b, 'us_per_step_ncu': sum(d['gpu__time_duration.sum'] for d in pair),
                         'source': 'profiles/' + os.path.basename(path), 'commit': commit}
        json.dump(rec, open(out, 'w'), indent=1)
        print(f'# wrote {out}: {len(pair)} pair-level contraction launches, {b / 1e9:.2f} GB per step')


if __name__ == '__main__':
    main()
