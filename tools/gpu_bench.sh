#!/bin/bash
# Bench + profiles on the GPU box.  Usage: bash tools/gpu_bench.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/$name.log; }
run bench_bf16   python bench.py --steps 5 --warmup 3 --precision bf16
run bench_bf16x3 python bench.py --steps 3 --warmup 3 --precision bf16x3 --no-render --no-cpu --no-extra
run bench_ref    python bench.py --impl reference --steps 2 --warmup 1
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --precision bf16 --no-render --no-cpu --no-extra > gpurun_out/ncu_list.log 2>&1
echo "rc=$?"; tail -n 3 gpurun_out/ncu_list.log
echo "=== ncu full (gemm_tc)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 3 -f -o gpurun_out/gemm_tc_$TAG \
    python bench.py --steps 1 --warmup 1 --precision bf16 --no-render --no-cpu --no-extra > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -n 3 gpurun_out/ncu_full.log
ls -la gpurun_out
