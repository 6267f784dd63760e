#!/bin/bash
# Development run on the GPU box: each test group in its own process (a trapped kernel poisons the CUDA
# context of the process that launched it), logs under gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run ops      python -m pytest tests/test_gpu_ops.py -q -m gpu
run gemm     python -m pytest tests/test_gpu_gemm.py -q -m gpu
run model    python -m pytest tests/test_gpu_model.py -q -m gpu
run smoke    python __graft_entry__.py --smoke
