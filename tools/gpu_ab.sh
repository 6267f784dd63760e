#!/bin/bash
# Development A/B on the GPU box: parity tests of the GEMM / model paths and short bench legs (extra commands as arguments).
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
B="python bench.py --steps 10 --warmup 4 --no-render --no-cpu --no-extra --no-hbm"
pick() { python - "$1" <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    if ln.startswith('{'):
        d = json.loads(ln)
        r, w = d['roofline'], d['roofline_wgrad']
        print('  value %.0f rays/s  %.2f ms  e2e %.0f (%.2f ms; single steps %s) | chains %.1f TF/s algo, exec frac %.3f, share %.3f | wgrad %.0f GB/s frac %.3f share %.3f | sm %s MHz' % (
            d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('host_wall_ms_of_single_steps'), r['achieved'], r['executed_mma']['frac'], r['kernel_share_of_step'],
            w['achieved'], w['frac'], w['kernel_share_of_step'], d['clocks']['sm_mhz']))
        kc = d['kernel_classes']
        print('  ms/step by class: ' + ', '.join('%s %.2f' % (k, v['ms'] / d['steps']) for k, v in kc.items() if v['launches']) + ' | profiled step %.2f' % r['profiled_pass_ms_per_step'])
PY
}
if [ -z "$NOTEST" ]; then
run gemm  python -m pytest tests/test_gpu_gemm.py -q -m gpu
run model python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -q -m gpu -x
fi
TAILN=1 CUT=100 run bench_a $B; pick gpurun_out/bench_a.log
TAILN=1 CUT=100 run bench_b $B; pick gpurun_out/bench_b.log
for f in "$@"; do bash -c "$f"; done
