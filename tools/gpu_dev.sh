#!/bin/bash
# Development run on the GPU box: each test group in its own process (a trapped kernel poisons the CUDA
# context of the process that launched it), logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n 25 gpurun_out/$name.log; }
run ops      python -m pytest tests/test_gpu_ops.py -q -m gpu -x
run gemm_simt python -m pytest tests/test_gpu_gemm.py -q -m gpu -k simt
run gemm_tc  python -m pytest tests/test_gpu_gemm.py -q -m gpu -k tc
run model_fp32 python -m pytest tests/test_gpu_model.py -q -m gpu -k "fp32 or checkpoint or leading"
run model_x3 python -m pytest tests/test_gpu_model.py -q -m gpu -k "bf16x3"
run model_bf16 python -m pytest tests/test_gpu_model.py -q -m gpu -k "bf16 and not bf16x3"
run smoke    python __graft_entry__.py --smoke
