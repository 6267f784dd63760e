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


def step_table(path, step_index=1):
    """Per-kernel totals of ONE training step of a launch list captured with time + DRAM-byte metrics (tools/gpu_prof.sh).
    Steps are delimited by their first resample launch (two per step: level 0, level 1)."""
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rows = OrderedDict()
    for r in csv.DictReader(io.StringIO(''.join(lines))):
        rows.setdefault(int(r['ID']), {'name': r['Kernel Name']})[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
    ids = sorted(rows)
    res = [i for i in ids if 'resample' in rows[i]['name']]
    starts = res[0::2]
    s0, s1 = starts[step_index], (starts[step_index + 1] if step_index + 1 < len(starts) else ids[-1] + 1)
    tot = OrderedDict()
    for i in ids:
        if s0 <= i < s1:
            a = tot.setdefault(short(rows[i]['name']), [0, 0.0, 0.0, 0.0, 0.0])
            a[0] += 1
            a[1] += rows[i].get('gpu__time_duration.sum', 0.0) / 1e3
            a[2] += rows[i].get('dram__bytes_read.sum', 0.0)
            a[3] += rows[i].get('dram__bytes_write.sum', 0.0)
            a[4] = max(a[4], rows[i].get('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 0.0))
    total = sum(v[1] for v in tot.values())
    dram = sum(v[2] + v[3] for v in tot.values())
    print(f'# {path}: training step #{step_index} of the command = launches {s0}..{s1 - 1} ({s1 - s0} launches), '
          f'{total / 1e3:.2f} ms of kernel time (cold-cache, serialised under ncu: only the SHARES are meaningful), '
          f'{dram / 1e9:.1f} GB of DRAM traffic')
    print(f'{"kernel":62s} {"n":>4s} {"total ms":>9s} {"share":>6s} {"DRAM rd GB":>11s} {"DRAM wr GB":>11s}')
    for k, (n, us, rb, wb, _) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f'{k[:62]:62s} {n:4d} {us / 1e3:9.3f} {100 * us / total:5.1f}% {rb / 1e9:11.2f} {wb / 1e9:11.2f}')
    return {'step_dram_bytes': dram, 'step_kernel_ms_under_ncu': total / 1e3, 'launches': s1 - s0}


def main():
    kind, path = sys.argv[1], sys.argv[2]
    if kind == 'step':
        step_table(path, int(sys.argv[3]) if len(sys.argv) > 3 else 1)
        return
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
        want = ['Kernel Name', 'launch__grid_size', 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
                'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg', 'sm__inst_executed_pipe_uniform.sum', 'launch__block_size', 'launch__cluster_size', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max',
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
                hit = [h for h in hdr if h == k or h.endswith('.' + k)]
                if hit:
                    print(f'{k:85s} {d[hit[0]]:>18s} {u[hit[0]]}')


if __name__ == '__main__':
    main()
