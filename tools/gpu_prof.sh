#!/bin/bash
# ncu launch list of one training step + full captures of the chain / pointwise / wgrad kernels.  Usage: bash tools/gpu_prof.sh TAG [precision]
# The .ncu-rep files of the chain kernels are summarised ON the box (tools/ncu_summary.py) and deleted there: gpurun only
# merges gpurun_out/ back while it stays under 64 MiB.
TAG=${1:-r02}
PREC=${2:-bf16x3}
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 1 --precision $PREC --no-render --no-cpu --no-extra --no-hbm"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_uniform.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py $ARGS > gpurun_out/ncu_list_$TAG.log 2>&1
echo "list rc=$?"
python tools/ncu_summary.py step gpurun_out/launches_$TAG.csv 1 > gpurun_out/launches_step_$TAG.txt 2>&1
# fused chains + pointwise kernels of the second step of the command (22 matching launches per step)
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"chain_x3|chain_pair|encode_kernel|heads_prologue|ipe_grad_normals|color_fwd|color_bwd" -s 22 -c 22 -f -o gpurun_out/step_$TAG \
    python bench.py $ARGS > gpurun_out/ncu_step_$TAG.log 2>&1
echo "chains + pointwise rc=$?"
python tools/ncu_summary.py full gpurun_out/step_$TAG.ncu-rep > gpurun_out/step_ncu_full_$TAG.txt 2>&1
ls -la gpurun_out/step_$TAG.ncu-rep
[ $(stat -c %s gpurun_out/step_$TAG.ncu-rep) -gt 40000000 ] && rm -f gpurun_out/step_$TAG.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad2_tc -s 80 -c 3 -f -o gpurun_out/wgrad2_$TAG \
    python bench.py $ARGS > gpurun_out/ncu_wgrad_$TAG.log 2>&1
echo "wgrad rc=$?"
python tools/ncu_summary.py full gpurun_out/wgrad2_$TAG.ncu-rep > gpurun_out/wgrad2_ncu_full_$TAG.txt 2>&1
du -sh gpurun_out
