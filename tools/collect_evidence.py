"""Turns the files one `tools/profile_all.sh <tag>` run left under gpurun_out/ into the tracked evidence under profiles/:

    python tools/collect_evidence.py r2n r2

writes profiles/<prefix>_bench.json (the bench line, not taken under a profiler), <prefix>_bench_detail.txt (per-layer CUDA-event
times), <prefix>_launches.csv + _launches_summary.txt (every launch of one steady-state frame, cold-cache and serialised under
ncu: compare SHARES), <prefix>_ncu_frame_summary.txt (ncu --set full of every kernel of that frame: duration, DRAM bytes, DRAM /
L2 / tensor-pipe utilisation, occupancy, registers, instructions, issue utilisation), <prefix>_ncu_conv_prog.json (what bench.py's
roofline.traffic reads) + the raw page of that capture, and <prefix>_ncu_conv_prog_top_stalls.txt (the 40 instructions of the
program kernel with the most stall samples, from the source page)."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')

WANT = [('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'rdMB'), ('dram__bytes_write.sum', 'wrMB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'), ('launch__registers_per_thread', 'regs'),
        ('smsp__inst_executed.sum', 'Minst'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%')]


def to_num(v, unit):
    v = float(v.replace(',', ''))
    return v * {'ns': 1e-3, 'ms': 1e3, 'us': 1.0, 's': 1e6, 'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3, 'inst': 1e-6}.get(unit, 1.0)


def frame_summary(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    agg = {}
    for d in data:
        name = d[hdr.index('Kernel Name')].split('(')[0].replace('void ', '').replace('mftb::', '')
        key = (name, d[hdr.index('Grid Size')], d[hdr.index('Block Size')])
        rec = agg.setdefault(key, {'n': 0})
        rec['n'] += 1
        for k, short in WANT:
            if k in hdr:
                i = hdr.index(k)
                rec.setdefault(short, []).append(to_num(d[i], units[i]))
    with open(dst, 'w') as f:
        f.write('# ncu --set full --clock-control none of ONE steady-state frame (512x512, 7 chains, 12 iterations); one line per (kernel, grid):\n')
        f.write('# launches, mean duration, DRAM read / written per launch, DRAM / L2 / tensor-pipe utilisation, achieved occupancy, registers,\n')
        f.write('# warp instructions (millions), issue-slot utilisation.  Durations under ncu are cold-cache and serialised.\n')
        f.write(f'{"kernel":34s} {"grid":16s} {"block":12s} {"n":>3s} ' + ' '.join(f'{s:>8s}' for _, s in WANT) + '\n')
        tot = 0.0
        for key, rec in sorted(agg.items(), key=lambda kv: -sum(kv[1].get('us', [0]))):
            m = lambda s: (sum(rec[s]) / len(rec[s])) if s in rec else float('nan')
            tot += sum(rec.get('us', [0]))
            f.write(f'{key[0][:34]:34s} {key[1]:16s} {key[2]:12s} {rec["n"]:3d} ' + ' '.join(f'{m(s):8.2f}' for _, s in WANT) + '\n')
        f.write(f'# sum of durations: {tot:.1f} us\n')


def prog_json(src, dst, tag):
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    g = lambda d, k: to_num(d[hdr.index(k)], units[hdr.index(k)])
    out = {'kernel': 'conv_prog_kernel<false> (steady state, 7 pairs at 512x512)',
           'dram_bytes_per_launch': sum((g(d, 'dram__bytes_read.sum') + g(d, 'dram__bytes_write.sum')) * 1e6 for d in data) / len(data),
           'duration_us': [g(d, 'gpu__time_duration.sum') for d in data],
           'tensor_pipe_active_pct': [g(d, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active') for d in data],
           'l2_to_sm_read_bytes': [g(d, 'lts__t_sectors_srcunit_tex_op_read.sum') * 32 for d in data if 'lts__t_sectors_srcunit_tex_op_read.sum' in hdr],
           'registers': [g(d, 'launch__registers_per_thread') for d in data],
           'source': f'ncu --set full --clock-control none --import-source on -k regex:conv_prog_kernel -s 3 -c 1 python tools/profile_step.py ({tag})'}
    json.dump(out, open(dst, 'w'), indent=1)


def top_stalls(src, dst, n=40):
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr, data = rows[h], rows[h + 1:]
    si, ii = hdr.index('# Samples'), hdr.index('Source')
    stall_cols = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
    data = [r for r in data if len(r) > si and r[si].isdigit()]
    total = sum(int(r[si]) for r in data) or 1
    with open(dst, 'w') as f:
        f.write(f'# conv_prog_kernel: the {n} instructions with the most warp-stall samples (ncu source page; {total} samples in total)\n')
        for r in sorted(data, key=lambda r: -int(r[si]))[:n]:
            top = sorted(((int(r[i]), hdr[i]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:2]
            f.write(f'{100.0 * int(r[si]) / total:5.1f}%  {r[ii].strip()[:70]:70s}  ' + ', '.join(f'{k} {v}' for v, k in top) + '\n')


def main(tag, prefix):
    cp = lambda a, b: shutil.copyfile(os.path.join(G, a), os.path.join(P, b))
    cp(f'{tag}_bench.json', f'{prefix}_bench.json')
    cp(f'{tag}_bench_detail.txt', f'{prefix}_bench_detail.txt')
    cp(f'{tag}_launches.csv', f'{prefix}_launches.csv')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'launch_summary.py'), os.path.join(P, f'{prefix}_launches.csv'), '40'],
                         capture_output=True, text=True).stdout
    open(os.path.join(P, f'{prefix}_launches_summary.txt'), 'w').write(out)
    frame_summary(os.path.join(G, f'{tag}_frame_raw.csv'), os.path.join(P, f'{prefix}_ncu_frame_summary.txt'))
    prog_json(os.path.join(G, f'{tag}_prog_raw.csv'), os.path.join(P, f'{prefix}_ncu_conv_prog.json'), tag)
    cp(f'{tag}_prog_raw.csv', f'{prefix}_ncu_conv_prog_raw.csv')
    top_stalls(os.path.join(G, f'{tag}_prog_source.csv'), os.path.join(P, f'{prefix}_ncu_conv_prog_top_stalls.txt'))
    print(open(os.path.join(P, f'{prefix}_ncu_frame_summary.txt')).read())


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else 'r2')
