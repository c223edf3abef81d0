"""Key metrics of every kernel in an .ncu-rep (needs ncu on PATH): python tools/ncu_summary.py file.ncu-rep"""
import csv
import io
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'dur'), ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('launch__registers_per_thread', 'regs'), ('sm__inst_executed.sum', 'inst')]


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = []
    for d in data:
        name = d[hdr.index('Kernel Name')].split('(')[0].replace('void ', '')
        rec = {'kernel': name, 'grid': d[hdr.index('Grid Size')]}
        for key, short in WANT:
            if key in hdr:
                i = hdr.index(key)
                rec[short] = (d[i], units[i])
        res.append(rec)
    return res


if __name__ == '__main__':
    for r in main(sys.argv[1]):
        dur = r.get('dur', ('0', 'us'))
        rd = to_bytes(*r['dram_rd']) if 'dram_rd' in r else 0
        wr = to_bytes(*r['dram_wr']) if 'dram_wr' in r else 0
        print(f"{r['kernel'][:34]:34s} {r['grid']:14s} {dur[0]:>8s} {dur[1]:3s} dram {rd / 1e6:8.1f}+{wr / 1e6:7.1f} MB "
              f"dram% {r.get('dram%', ('?',))[0]:>6s} tensor% {r.get('tensor%', ('?',))[0]:>6s} sm% {r.get('sm%', ('?',))[0]:>6s} "
              f"l2% {r.get('l2%', ('?',))[0]:>6s} occ% {r.get('occ%', ('?',))[0]:>6s} regs {r.get('regs', ('?',))[0]}")
