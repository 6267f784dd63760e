#!/bin/bash
# Round-end GPU evidence: full test-suite, smoke, bench (both arms), memcheck of every fused path.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-400; }
run gpu_tests python -m pytest tests -q -m gpu
run smoke python __graft_entry__.py --smoke
run bench_ref python bench.py --impl reference --steps 3 --warmup 1
run bench_main python bench.py --steps 20 --warmup 5
TMO=1200 run memcheck compute-sanitizer --tool memcheck --print-limit 5 python tools/memcheck_run.py 300 16384
