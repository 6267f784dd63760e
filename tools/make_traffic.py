"""profiles/<TAG>_traffic.json: per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the profiled
kernel classes, averaged over the launches of one `ncu --set full` capture.  Usage:
  python tools/make_traffic.py TAG chain_tc=gpurun_out/chain_x3_TAG.ncu-rep,gpurun_out/chain_pair_TAG.ncu-rep \
      wgrad_tc=gpurun_out/wgrad2_TAG.ncu-rep [step=gpurun_out/launches_TAG.csv]
(several captures per class are pooled; step=... adds the DRAM bytes of one whole training step from the launch list)"""
import csv
import io
import json
import subprocess
import sys

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def dram_bytes(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    vals = []
    for r in rd[2:]:
        vals.append(float(r[ir].replace(',', '')) * UNIT[units[ir]] + float(r[iw].replace(',', '')) * UNIT[units[iw]])
    return vals


def main():
    tag = sys.argv[1]
    res = {}
    for arg in sys.argv[2:]:
        k, paths = arg.split('=')
        if k == 'step':
            sys.path.insert(0, 'tools')
            import ncu_summary
            import contextlib
            with contextlib.redirect_stdout(io.StringIO()):
                st = ncu_summary.step_table(paths, 1)
            res.update(st)
            continue
        v = []
        for path in paths.split(','):
            v += dram_bytes(path)
        res[k] = sum(v) / len(v)
        res[k + '_launches_profiled'] = len(v)
    res['source'] = ('ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, mean per launch; '
                     'step_dram_bytes: the same two counters summed over every launch of one training step (launch list)')
    json.dump(res, open(f'profiles/{tag}_traffic.json', 'w'), indent=1)
    print(res)


if __name__ == '__main__':
    main()
