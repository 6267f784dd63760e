"""GPU box: achieved GB/s of the warp-per-ray kernels (same measurement as bench.py's hbm_kernels).
   python tools/hbm_bench.py [iters]   (iters = 1 under ncu: one warm-up + one timed launch per kernel)"""
import json, sys
import torch
sys.path.insert(0, '.')
import bench
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
print(json.dumps(bench.time_hbm_kernels(torch.device('cuda', 0), bench.load_peaks(), iters), indent=1))
