#!/bin/bash
# ncu capture of the fused chain kernel in the eval (render) path + wgrad kernel in training
mkdir -p gpurun_out
cat > /tmp/render_only.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from bench import build_everything
from refnerf_pl_b200 import synthetic, utils
model, cfg = build_everything('bf16', 'cuda')
model.eval()
r = synthetic.blender_rays(65536, seed=3)
rays = utils.Rays(**{k: torch.from_numpy(v).cuda() for k, v in r.items()})
with torch.no_grad():
    for _ in range(3):
        model(rays, 1.0, True)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 4 -c 2 -f -o gpurun_out/chain_eval_r01 python /tmp/render_only.py > gpurun_out/ncu_chain_eval.log 2>&1
tail -n 3 gpurun_out/ncu_chain_eval.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad2_tc -s 10 -c 2 -f -o gpurun_out/wgrad2_r01 python bench.py --steps 1 --warmup 1 --precision bf16 --no-render --no-cpu --no-extra > gpurun_out/ncu_wgrad2.log 2>&1
tail -n 3 gpurun_out/ncu_wgrad2.log
