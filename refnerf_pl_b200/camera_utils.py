"""On-device ray generation with the reference's `camera_utils.pixels_to_rays` surface (camera_utils.py:502-614),
SURVEY 8(f) rank 1: pixel indices + cameras in, the `utils.Rays` geometry fields out, computed by one sm_100a kernel
instead of the per-batch numpy path of `datasets.py:433-435` and its nine host-to-device copies.

Supported: perspective cameras without lens distortion, with or without the NDC conversion -- what the Blender and
LLFF loaders of the Ref-NeRF configs produce.  Fisheye cameras and distortion parameters raise."""
import torch

from . import ops, utils


def pixels_to_rays(pix_x_int, pix_y_int, pixtocams, camtoworlds, distortion_params=None, pixtocam_ndc=None, camtype=None,
                   cam_idx=None):
    """pix_*_int: integer CUDA tensors of any shape SH.  pixtocams [..,3,3] / camtoworlds [..,3,4]: either per-camera
    tables indexed by `cam_idx` (shape SH), or -- as in the reference -- broadcastable to SH + [3,3] / SH + [3,4].
    Returns origins, directions, viewdirs SH+[3], radii SH+[1], imageplane SH+[2] (fp32)."""
    if distortion_params is not None:
        raise NotImplementedError('lens distortion is not part of the Ref-NeRF configurations')
    if camtype is not None and getattr(camtype, 'value', camtype) not in ('perspective', 0, None):
        raise NotImplementedError('only perspective cameras are supported')
    sh = tuple(pix_x_int.shape)
    dev = pix_x_int.device
    px, py = pix_x_int.reshape(-1), pix_y_int.reshape(-1)
    n = px.numel()
    p2c = torch.as_tensor(pixtocams, dtype=torch.float32, device=dev)
    c2w = torch.as_tensor(camtoworlds, dtype=torch.float32, device=dev)[..., :3, :4]
    if cam_idx is None:
        if p2c.dim() == 2 and c2w.dim() == 2:            # one camera for every pixel
            p2c, c2w = p2c[None], c2w[None]
            ci = torch.zeros(n, dtype=torch.int32, device=dev)
        else:                                            # per-pixel matrices, the reference's calling convention
            p2c = p2c.expand(sh + (3, 3)).reshape(n, 3, 3)
            c2w = c2w.expand(sh + (3, 4)).reshape(n, 3, 4)
            ci = torch.arange(n, dtype=torch.int32, device=dev)
    else:
        ci = cam_idx.reshape(-1).to(torch.int32)
        p2c, c2w = p2c.reshape(-1, 3, 3), c2w.reshape(-1, 3, 4)
    ndc = torch.as_tensor(pixtocam_ndc, dtype=torch.float32, device=dev) if pixtocam_ndc is not None else \
        torch.empty(0, dtype=torch.float32, device=dev)
    o, d, v, r, ip = ops.pixels_to_rays(px, py, ci, p2c.contiguous(), c2w.contiguous(), ndc)
    return o.reshape(sh + (3,)), d.reshape(sh + (3,)), v.reshape(sh + (3,)), r.reshape(sh + (1,)), ip.reshape(sh + (2,))


def cast_ray_batch(cameras, pix_x_int, pix_y_int, cam_idx, near, far, lossmult=None):
    """camera_utils.cast_ray_batch (camera_utils.py:617-670) for the supported camera model:
    cameras = (pixtocams [C,3,3], camtoworlds [C,3,4], distortion_params (must be None), pixtocam_ndc or None)."""
    pixtocams, camtoworlds, distortion_params, pixtocam_ndc = cameras
    o, d, v, r, ip = pixels_to_rays(pix_x_int, pix_y_int, pixtocams, camtoworlds, distortion_params, pixtocam_ndc,
                                    cam_idx=cam_idx)
    ones = torch.ones_like(r)
    return utils.Rays(origins=o, directions=d, viewdirs=v, radii=r, imageplane=ip,
                      lossmult=ones if lossmult is None else lossmult, near=ones * near, far=ones * far,
                      cam_idx=cam_idx.reshape(tuple(pix_x_int.shape) + (1,)).to(torch.int32))
