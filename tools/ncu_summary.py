"""Summarise ncu outputs into profiles/: launch list (per-kernel totals for the last step) and selected metrics of
full captures.  Usage: python tools/ncu_summary.py launches <csv> | full <ncu-rep>"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict


def launches(path, last_step_from=None):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(io.StringIO(''.join(lines)))
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        us = v / 1e3 if unit in ('nsecond', 'ns') else (v if unit in ('usecond', 'us') else v * 1e3)
        rows.append((int(r['ID']), r['Kernel Name'], us))
    return rows


def short(name):
    name = name.split('(')[0]
    for junk in ('void ', 'rn::', '(anonymous namespace)::', '<unnamed>::'):
        name = name.replace(junk, '')
    return name[:70]


def main():
    kind, path = sys.argv[1], sys.argv[2]
    if kind == 'launches':
        rows = launches(path)
        # the bench runs warm-up step(s) then the timed step(s) then an e2e warm-up + e2e step: print totals per kernel
        tot = OrderedDict()
        for _, k, us in rows:
            k = short(k)
            a = tot.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += us
        total = sum(v[1] for v in tot.values())
        print(f'# {path}: {len(rows)} launches, {total / 1e3:.2f} ms total (cold-cache, serialised under ncu)')
        print(f'{"kernel":70s} {"launches":>8s} {"total ms":>10s} {"avg us":>10s} {"share":>7s}')
        for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            print(f'{k:70s} {n:8d} {us / 1e3:10.3f} {us / n:10.1f} {100 * us / total:6.1f}%')
    else:
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rd = list(csv.reader(io.StringIO(out)))
        hdr, units = rd[0], rd[1]
        want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max',
                'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
                'sm__inst_executed_pipe_tc.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
                'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
                'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
                'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
                'l1tex__data_pipe_tc_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
                'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
                'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum',
                'smsp__cycles_active.avg', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
                'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active']
        for r in rd[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            print('-' * 100)
            for k in want:
                if k in d:
                    print(f'{k:85s} {d[k]:>18s} {u[k]}')


if __name__ == '__main__':
    main()
