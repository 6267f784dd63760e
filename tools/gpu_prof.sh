#!/bin/bash
# ncu launch list of one training step + full captures of the chain / wgrad kernels.  Usage: bash tools/gpu_prof.sh TAG [precision]
TAG=${1:-r02}
PREC=${2:-bf16x3}
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 1 --precision $PREC --no-render --no-cpu --no-extra --no-hbm"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_uniform.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py $ARGS > gpurun_out/ncu_list_$TAG.log 2>&1
echo "list rc=$?"
# fused chains of the profiled-pass step: x3 forward (spatial / normals / view per level), loss-backward chains
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_x3 -s 18 -c 6 -f -o gpurun_out/chain_x3_$TAG \
    python bench.py $ARGS > gpurun_out/ncu_chainx3_$TAG.log 2>&1
echo "chain_x3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_pair -s 12 -c 4 -f -o gpurun_out/chain_pair_$TAG \
    python bench.py $ARGS > gpurun_out/ncu_chainpair_$TAG.log 2>&1
echo "chain_pair rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad2_tc -s 80 -c 4 -f -o gpurun_out/wgrad2_$TAG \
    python bench.py $ARGS > gpurun_out/ncu_wgrad_$TAG.log 2>&1
echo "wgrad rc=$?"
ls -la gpurun_out | tail -5
