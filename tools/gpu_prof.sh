#!/bin/bash
# ncu launch list of one training step + full captures of the chain kernel instances.  Usage: bash tools/gpu_prof.sh TAG [precision]
TAG=${1:-r01}
PREC=${2:-fp16}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --precision $PREC --no-render --no-cpu --no-parity --no-hbm > gpurun_out/ncu_list_$TAG.log 2>&1
echo "list rc=$?"
# chain kernel: launches 11.. of the timed step = spatial fwd (saves), normals dgrad, view fwd, view bwd, spatial bwd
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_pair -s 20 -c 10 -f -o gpurun_out/chain_pair_$TAG \
    python bench.py --steps 1 --warmup 1 --precision $PREC --no-render --no-cpu --no-parity --no-hbm > gpurun_out/ncu_chain_$TAG.log 2>&1
echo "chain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad2_tc -s 80 -c 4 -f -o gpurun_out/wgrad2_$TAG \
    python bench.py --steps 1 --warmup 1 --precision $PREC --no-render --no-cpu --no-parity --no-hbm > gpurun_out/ncu_wgrad_$TAG.log 2>&1
echo "wgrad rc=$?"
ls -la gpurun_out | tail -5
