"""Synthetic ray generators shaped like the reference's Blender and LLFF loaders.

There is no dataset access in this environment, so benchmarks and parity tests feed rays with the
same geometry the reference's data path produces:
  * Blender: `datasets.py:519-581` + `camera_utils.pixels_to_rays` (`camera_utils.py:502-614`):
    800x800, focal 1111.1 px, camera on a radius-4.03 sphere looking at the origin (OpenGL
    convention), non-unit directions, radii = mean neighbour-ray distance * 2/sqrt(12), near 2 far 6.
  * LLFF: 1008x756 forward-facing, NDC via `camera_utils.convert_to_ndc` (`camera_utils.py:31-97`),
    viewdirs are the pre-NDC world directions, radii from NDC origin offsets, near 0 far 1.
All arrays are float32 numpy with shape [N, C], the layout `utils.Rays` carries (`utils.py:51-93`).
"""
import numpy as np


def _look_at(cam_pos):
    """camera-to-world rotation (OpenGL: camera looks down -z, +y up) for a camera at cam_pos
    looking at the origin."""
    fwd = -cam_pos / np.linalg.norm(cam_pos)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    true_up = np.cross(right, fwd)
    return np.stack([right, true_up, -fwd], axis=1)  # columns: x, y, z axes of the camera


def _pixel_dirs(px, py, focal, cx, cy):
    """OpenGL camera-space direction through the centre of pixel (px, py)."""
    return np.stack([(px + 0.5 - cx) / focal, -(py + 0.5 - cy) / focal, -np.ones_like(px)], axis=-1)


def _pack(origins, directions, viewdirs, radii, imageplane, near, far):
    n = origins.shape[0]
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(origins=f32(origins), directions=f32(directions), viewdirs=f32(viewdirs), radii=f32(radii),
                imageplane=f32(imageplane), lossmult=np.ones((n, 1), np.float32),
                near=np.full((n, 1), near, np.float32), far=np.full((n, 1), far, np.float32),
                cam_idx=np.zeros((n, 1), np.int32))


def blender_rays(n_rays=None, seed=0, width=800, height=800, focal=1111.1, radius=4.03, near=2.0, far=6.0,
                 pixels=None):
    """Random training pixels (n_rays given) or a full frame in raster order (n_rays None)."""
    rng = np.random.default_rng(seed)
    theta = rng.uniform(0, 2 * np.pi)
    phi = rng.uniform(np.deg2rad(15), np.deg2rad(75))
    cam = radius * np.array([np.cos(theta) * np.cos(phi), np.sin(theta) * np.cos(phi), np.sin(phi)])
    rot = _look_at(cam)
    if pixels is not None:
        px, py = pixels
    elif n_rays is None:
        py, px = np.meshgrid(np.arange(height), np.arange(width), indexing='ij')
        px, py = px.reshape(-1), py.reshape(-1)
    else:
        px = rng.integers(0, width, n_rays)
        py = rng.integers(0, height, n_rays)
    px = px.astype(np.float64)
    py = py.astype(np.float64)
    cx, cy = width / 2, height / 2
    cdir = _pixel_dirs(px, py, focal, cx, cy)
    d = cdir @ rot.T
    dx = _pixel_dirs(px + 1, py, focal, cx, cy) @ rot.T
    dy = _pixel_dirs(px, py + 1, focal, cx, cy) @ rot.T
    radii = (0.5 * (np.linalg.norm(dx - d, axis=-1) + np.linalg.norm(dy - d, axis=-1)))[:, None] * 2 / np.sqrt(12)
    o = np.broadcast_to(cam, d.shape)
    v = d / np.linalg.norm(d, axis=-1, keepdims=True)
    return _pack(o, d, v, radii, cdir[:, :2], near, far)


def _to_ndc(o, d, focal, width, height, ndc_near=1.0):
    t = -(ndc_near + o[:, 2]) / d[:, 2]
    o = o + t[:, None] * d
    xm = -2.0 * focal / width
    ym = -2.0 * focal / height
    o_ndc = np.stack([xm * o[:, 0] / o[:, 2], ym * o[:, 1] / o[:, 2], -np.ones_like(o[:, 2])], -1)
    inf_ndc = np.stack([xm * d[:, 0] / d[:, 2], ym * d[:, 1] / d[:, 2], np.ones_like(o[:, 2])], -1)
    return o_ndc, inf_ndc - o_ndc


def llff_rays(n_rays=None, seed=0, width=1008, height=756, near=0.0, far=1.0, pixels=None):
    rng = np.random.default_rng(seed)
    focal = 0.81 * width
    ang = np.deg2rad(rng.uniform(-5, 5, 3))
    cx_, sx_ = np.cos(ang[0]), np.sin(ang[0])
    cy_, sy_ = np.cos(ang[1]), np.sin(ang[1])
    cz_, sz_ = np.cos(ang[2]), np.sin(ang[2])
    rot = (np.array([[cz_, -sz_, 0], [sz_, cz_, 0], [0, 0, 1]]) @ np.array([[cy_, 0, sy_], [0, 1, 0], [-sy_, 0, cy_]])
           @ np.array([[1, 0, 0], [0, cx_, -sx_], [0, sx_, cx_]]))
    cam = rng.uniform(-0.2, 0.2, 3)
    if pixels is not None:
        px, py = pixels
    elif n_rays is None:
        py, px = np.meshgrid(np.arange(height), np.arange(width), indexing='ij')
        px, py = px.reshape(-1), py.reshape(-1)
    else:
        px = rng.integers(0, width, n_rays)
        py = rng.integers(0, height, n_rays)
    px = px.astype(np.float64)
    py = py.astype(np.float64)
    cxp, cyp = width / 2, height / 2
    cdir = _pixel_dirs(px, py, focal, cxp, cyp)
    d = cdir @ rot.T
    dx = _pixel_dirs(px + 1, py, focal, cxp, cyp) @ rot.T
    dy = _pixel_dirs(px, py + 1, focal, cxp, cyp) @ rot.T
    o = np.broadcast_to(cam, d.shape)
    v = d / np.linalg.norm(d, axis=-1, keepdims=True)
    o_ndc, d_ndc = _to_ndc(o, d, focal, width, height)
    ox, _ = _to_ndc(o, dx, focal, width, height)
    oy, _ = _to_ndc(o, dy, focal, width, height)
    radii = (0.5 * (np.linalg.norm(ox - o_ndc, axis=-1) + np.linalg.norm(oy - o_ndc, axis=-1)))[:, None] * 2 / np.sqrt(12)
    return _pack(o_ndc, d_ndc, v, radii, cdir[:, :2], near, far)


def gt_rgb(n_rays, seed=0):
    return np.random.default_rng(seed + 12345).random((n_rays, 3), dtype=np.float32)
