#!/bin/bash
# GPU-box run for the fused split-bf16 chains: debug comparison, fixtures, trace, short bench.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
TAILN=4 run x3_debug python tools/x3_debug.py 700
TAILN=6 run ops python -m pytest tests/test_gpu_ops.py -q -m gpu -x
TAILN=6 run model python -m pytest tests/test_gpu_model.py -q -m gpu
TAILN=6 run fullsize python -m pytest tests/test_gpu_fullsize.py -q -m gpu
PREC=bf16x3 RN_CHAIN_TRACE=3 TAILN=2 run trace_x3_train python tools/chain_trace.py train
TAILN=3 run x3_bench python bench.py --precision bf16x3 --no-cpu --no-render --no-hbm --no-extra --steps 5 --warmup 3
TAILN=3 run fp16_bench python bench.py --precision fp16 --no-cpu --no-render --no-hbm --no-extra --steps 5 --warmup 3
