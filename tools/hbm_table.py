"""Achieved HBM bandwidth of the memory-bound kernels from ncu --set full captures (gpurun_out/*.ncu-rep):
per launch  (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration.sum  against the measured copy peak
(MEASURED_PEAKS.json hbm_gbs).  Usage: python tools/hbm_table.py out.md rep1.ncu-rep [rep2 ...]"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_num(s):
    try:
        return float(s.replace(',', ''))
    except ValueError:
        return None


def rows_of(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, want_unit=None):
        i = col.get(name)
        if i is None:
            return None
        v = to_num(r[i])
        if v is None:
            return None
        u = units[i]
        scale = {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1.0)
        return v * scale
    out = []
    for r in rows[2:]:
        name = re.sub(r'\(.*', '', r[col['Kernel Name']]).replace('void ', '').replace('(anonymous namespace)::', '')
        t = get(r, 'gpu__time_duration.sum')
        rd, wr = get(r, 'dram__bytes_read.sum'), get(r, 'dram__bytes_write.sum')
        if t and rd is not None and wr is not None:
            out.append(dict(kernel=name, t=t, rd=rd, wr=wr, grid=r[col['launch__grid_size']] if 'launch__grid_size' in col else '',
                            tensor=get(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
                            dram_pct=get(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')))
    return out


def main():
    out_path, reps = sys.argv[1], sys.argv[2:]
    peak = 6449.1
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peak = json.load(open(pk)).get('hbm_gbs', peak)
    lines = ['# Achieved HBM bandwidth of the memory-bound kernels (ncu --set full, --clock-control none; B=8 DSEC 440x640 bench step)',
             '',
             'Per launch: `(dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration.sum`; peak = %.1f GB/s '
             '(MEASURED_PEAKS.json `hbm_gbs`, torch copy).  ncu launches are cold-cache and serialised.' % peak, '',
             '| capture | kernel | grid | time us | DRAM read MB | DRAM write MB | GB/s | frac of peak |', '|---|---|---|---|---|---|---|---|']
    for rep in reps:
        for r in rows_of(rep):
            if 'conv_tc' in r['kernel'] or 'wgrad_tc_kernel' in r['kernel']:
                continue
            gbs = (r['rd'] + r['wr']) / r['t'] / 1e9
            lines.append('| %s | %s | %s | %.1f | %.1f | %.1f | %.0f | %.2f |' % (os.path.basename(rep).replace('.ncu-rep', ''), r['kernel'][:46],
                         r['grid'], r['t'] * 1e6, r['rd'] / 1e6, r['wr'] / 1e6, gbs, gbs / peak))
    open(out_path, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


if __name__ == '__main__':
    main()
