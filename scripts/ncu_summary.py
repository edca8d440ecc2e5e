#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export of one kernel launch into the JSON record that
bench.py reads (profiles/k_apply_traffic.json) and into a table row for profiles/README.md.

    python scripts/ncu_summary.py RAW.csv --n-qubits 28 --commit $(git rev-parse --short HEAD) \
        --command '...' [--out profiles/k_apply_traffic.json] [--algorithmic-bytes B]

Run in the build container on the CSV brought back in gpurun_out/ (the GPU box has no .git)."""
import argparse
import csv
import json

UNIT_SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12,
              'ns': 1.0, 'us': 1e3, 'usecond': 1e3, 'ms': 1e6, 'msecond': 1e6, 's': 1e9, 'second': 1e9}


def load(path, row=-1):
    rows = [r for r in csv.reader(open(path)) if r]
    header = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    names, units = rows[header], rows[header + 1]
    values = rows[row] if row < 0 else rows[header + 2 + row]
    return {n: (u, v) for n, u, v in zip(names, units, values)}


def number(rec, key, scale=False):
    unit, value = rec[key]
    x = float(value.replace(',', ''))
    return x * UNIT_SCALE.get(unit, 1.0) if scale else x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('raw')
    ap.add_argument('--n-qubits', type=int, required=True)
    ap.add_argument('--commit', default=None)
    ap.add_argument('--command', default=None)
    ap.add_argument('--algorithmic-bytes', type=float, default=None)
    ap.add_argument('--row', type=int, default=-1, help='which captured launch (default: last)')
    ap.add_argument('--out', default=None)
    args = ap.parse_args()
    rec = load(args.raw, args.row)
    pct = lambda key: number(rec, key) / 100.0
    stalls = {k.split('issue_stalled_')[1].split('_per_issue')[0]: number(rec, k)
              for k in rec if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio')}
    top = sorted(((v, k) for k, v in stalls.items() if k != 'selected'), reverse=True)[:4]
    read = number(rec, 'dram__bytes_read.sum', True)
    write = number(rec, 'dram__bytes_write.sum', True)
    out = {
        'kernel': rec['Kernel Name'][1],
        'n_qubits': args.n_qubits,
        'commit': args.commit,
        'source': '%s (ncu --set full --clock-control none, one launch of `%s`)' % (args.raw, args.command),
        'gpu_time_ns': number(rec, 'gpu__time_duration.sum', True),
        'dram_bytes_per_launch': read + write,
        'dram_bytes_read': read,
        'dram_bytes_write': write,
        'algorithmic_bytes_per_launch': args.algorithmic_bytes,
        'l1_hit_pct': number(rec, 'l1tex__t_sector_hit_rate.pct'),
        'l2_hit_pct': number(rec, 'lts__t_sector_hit_rate.pct'),
        'l1_frac': pct('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),
        'issue_frac': pct('sm__issue_active.avg.pct_of_peak_sustained_elapsed'),
        'fp64_frac': pct('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed'),
        'alu_frac': pct('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
        'lsu_frac': pct('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed'),
        'dram_throughput_frac': pct('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
        'warps_active_frac': pct('sm__warps_active.avg.pct_of_peak_sustained_active'),
        'registers_per_thread': number(rec, 'launch__registers_per_thread'),
        'warp_instructions': number(rec, 'smsp__inst_executed.sum'),
        'top_stalls_warps_per_issue': {k: round(v, 3) for v, k in top},
    }
    text = json.dumps(out, indent=1)
    print(text)
    if args.out:
        open(args.out, 'w').write(text + '\n')


if __name__ == '__main__':
    main()
