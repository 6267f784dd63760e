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
timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none -c 600 --csv --log-file gpurun_out/launches_eval.csv python /tmp/render_only.py > gpurun_out/ncu_list_eval.log 2>&1
echo rc=$?
