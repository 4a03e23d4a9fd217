"""Turns the ncu outputs brought back in gpurun_out/ into the committed text summaries under profiles/.
Usage: python tools/summarize_profiles.py r01 [suffix]"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else 'r01'
SUF = sys.argv[2] if len(sys.argv) > 2 else ''
G = os.path.join(ROOT, 'gpurun_out')
P = os.path.join(ROOT, 'profiles')
os.makedirs(P, exist_ok=True)


def launches():
    src = os.path.join(G, 'launches_%s.csv' % R)
    if not os.path.exists(src):
        return
    lines = [l for l in open(src) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    tot = collections.OrderedDict()
    for row in rows:
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e6 if u == 'ns' else v / 1e3 if u == 'us' else v
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('<unnamed>::', '')
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += v
    T = sum(v[1] for v in tot.values())
    with open(os.path.join(P, '%s_launches%s.txt' % (R, SUF)), 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 1 --warmup 3 '
                '--no-cpu-baseline\n# %d launches over the whole bench run (warm-up + timed + end-to-end steps); per-launch '
                'times are cold-cache and serialised: compare SHARES\n' % len(rows))
        f.write('%-44s %7s %11s %7s %10s\n' % ('kernel', 'count', 'total ms', 'share', 'avg ms'))
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write('%-44s %7d %11.2f %6.1f%% %10.4f\n' % (k[:44], v[0], v[1], 100 * v[1] / T, v[1] / v[0]))
        f.write('%-44s %7d %11.2f\n' % ('TOTAL', len(rows), T))


def full(name):
    rep = os.path.join(G, '%s_%s.ncu-rep' % (name, R))
    if not os.path.exists(rep):
        return
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = re.compile(r'Kernel Name|gpu__time_duration.sum$|dram__bytes_(read|write).sum$|gpu__dram_throughput.avg.pct|'
                      r'sm__pipe_tensor_cycles_active.avg.pct|sm__warps_active.avg.pct|launch__registers_per_thread$|'
                      r'launch__grid_size|launch__block_size|sm__throughput.avg.pct|lts__t_bytes.sum$|'
                      r'sm__cycles_elapsed.max$|launch__shared_mem_per_block_dynamic|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|'
                      r'smsp__warp_issue_stalled.*_per_warp_active.pct')
    idx = [i for i, h in enumerate(hdr) if keep.search(h)]
    with open(os.path.join(P, '%s_%s%s.txt' % (R, name, SUF)), 'w') as f:
        f.write('# ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel (bench.py, B=8 DSEC)\n')
        for r in rows[2:]:
            f.write('---\n')
            for i in idx:
                f.write('%s [%s] = %s\n' % (hdr[i], units[i], r[i]))


launches()
full('prof_lstm')
print(os.listdir(P))
