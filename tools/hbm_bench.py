"""GPU box: achieved GB/s of the warp-per-ray kernels (same measurement as bench.py's hbm_kernels)."""
import json, sys
import torch
sys.path.insert(0, '.')
import bench
print(json.dumps(bench.time_hbm_kernels(torch.device('cuda', 0), bench.load_peaks()), indent=1))
