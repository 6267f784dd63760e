"""Golden vectors for ray generation: the UNMODIFIED reference camera_utils.pixels_to_rays (numpy path) on Blender-
and LLFF-shaped cameras.  Run in the build container (needs /root/reference): python -m oracle.make_raygen_golden"""
import os
import sys

import numpy as np

from oracle import ref_import


def cameras(rng, n_cam, width, height, focal, radius):
    pixtocam = np.linalg.inv(np.array([[focal, 0, width / 2], [0, focal, height / 2], [0, 0, 1.]])).astype(np.float32)
    c2w = []
    for _ in range(n_cam):
        th, ph = rng.uniform(0, 2 * np.pi), rng.uniform(0.2, 1.3)
        pos = radius * np.array([np.cos(th) * np.cos(ph), np.sin(th) * np.cos(ph), np.sin(ph)])
        fwd = -pos / np.linalg.norm(pos)
        right = np.cross(fwd, [0, 0, 1.]); right /= np.linalg.norm(right)
        up = np.cross(right, fwd)
        c2w.append(np.concatenate([np.stack([right, up, -fwd], 1), pos[:, None]], 1))
    return np.broadcast_to(pixtocam, (n_cam, 3, 3)).copy(), np.stack(c2w).astype(np.float32)


def main():
    ns, _ = ref_import.load('blender_refnerf.gin')
    sys.path.insert(0, ref_import.REFERENCE_ROOT)
    from internal import camera_utils
    rng = np.random.default_rng(7)
    out = {}
    for name, (w, h, f, rad, ndc) in {'blender': (800, 800, 1111.1, 4.03, False), 'llff': (1008, 756, 815.0, 0.3, True)}.items():
        p2c, c2w = cameras(rng, 3, w, h, f, rad)
        if ndc:   # forward-facing: cameras near the origin looking down -z with small rotations
            c2w[:, :3, :3] = np.eye(3, dtype=np.float32) + 0.02 * rng.standard_normal((3, 3, 3)).astype(np.float32)
            c2w[:, :3, 3] = 0.2 * rng.standard_normal((3, 3)).astype(np.float32)
        n = 257
        px = rng.integers(0, w, n).astype(np.int32)
        py = rng.integers(0, h, n).astype(np.int32)
        ci = rng.integers(0, 3, n).astype(np.int32)
        ndc_mat = p2c[0] if ndc else None
        o, d, v, r, ip = camera_utils.pixels_to_rays(px, py, p2c[ci], c2w[ci], pixtocam_ndc=ndc_mat)
        out.update({f'{name}_px': px, f'{name}_py': py, f'{name}_cam': ci, f'{name}_pixtocams': p2c, f'{name}_camtoworlds': c2w,
                    f'{name}_origins': o, f'{name}_directions': d, f'{name}_viewdirs': v, f'{name}_radii': r,
                    f'{name}_imageplane': ip})
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'raygen.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: (v.shape, v.dtype) for k, v in out.items() if 'origins' in k or 'radii' in k})


if __name__ == '__main__':
    main()
