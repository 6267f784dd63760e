"""ctypes binding of librefnerf_b200.so (the C ABI declared in include/refnerf_b200.h).

There is no CPU or eager-PyTorch fallback: if the library has not been built
(`python -m refnerf_pl_b200.build`, or `__graft_entry__.build()`), importing an op raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# (REFNERF_B200_LIB: another build of the same library, for A/B measurements of kernel variants on one box)
LIB_PATH = os.environ.get('REFNERF_B200_LIB') or os.path.join(HERE, 'librefnerf_b200.so')

PREC_FP32, PREC_BF16, PREC_BF16X3, PREC_FP16 = 0, 1, 2, 3
PREC_BY_NAME = {'fp32': PREC_FP32, 'bf16': PREC_BF16, 'bf16x3': PREC_BF16X3, 'fp16': PREC_FP16}
NUM_PARAMS = 46


class RnMlpConfig(Structure):
    _fields_ = [('prec', c_int), ('srgb_mapping', c_int), ('srgb_normalization', c_int), ('density_bias', c_float),
                ('roughness_bias', c_float), ('rgb_premultiplier', c_float), ('rgb_bias', c_float),
                ('rgb_padding', c_float), ('chunk_rows', c_int), ('gemm_impl', c_int), ('deterministic_wgrad', c_int)]


class RnMlpOutputs(Structure):
    _fields_ = [(k, c_void_p) for k in ('density', 'rgb', 'normals', 'normals_pred', 'grad_pred', 'tint', 'diffuse',
                                        'specular', 'roughness')]


_P = c_void_p
_SIGNATURES = {
    'rn_last_error': (c_char_p, []),
    'rn_abi_version': (c_int, []),
    'rn_resample': (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_float, c_float, c_float, c_float, _P, _P, _P, _P, _P]),
    'rn_max_dilate_weights': (c_int, [_P, _P, c_int64, c_int, c_float, c_float, c_float, c_int, c_int, _P, _P, _P]),
    'rn_composite_fwd': (c_int, [_P] * 11 + [c_int64, c_int, c_float, _P, _P, _P, _P, _P]),
    'rn_composite_bwd': (c_int, [_P] * 15 + [c_int64, c_int, c_float] + [_P] * 7 + [_P]),
    'rn_lossfun_outer_fwd': (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, _P, _P]),
    'rn_lossfun_outer_bwd': (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, _P, _P]),
    'rn_distortion_fwd': (c_int, [_P, _P, c_int64, c_int, _P, _P]),
    'rn_distortion_bwd': (c_int, [_P, _P, _P, c_int64, c_int, _P, _P]),
    'rn_normal_losses_fwd': (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, _P, _P, _P]),
    'rn_normal_losses_bwd': (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, _P, _P, _P]),
    'rn_data_loss_fwd': (c_int, [_P, _P, _P, c_int64, c_int, c_float, _P, _P]),
    'rn_data_loss_bwd': (c_int, [_P, _P, _P, _P, c_int64, c_int, c_float, _P, _P]),
    'rn_encode': (c_int, [_P, _P, _P, _P, c_int64, c_int, _P, _P]),
    'rn_ide': (c_int, [_P, _P, c_int64, _P, _P]),
    'rn_pixels_to_rays': (c_int, [_P, _P, _P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P]),
    'rn_mlp_param_name': (c_char_p, [c_int]),
    'rn_mlp_param_numel': (c_int64, [c_int]),
    'rn_mlp_packed_bytes': (c_size_t, [c_int]),
    'rn_mlp_workspace_bytes': (c_size_t, [POINTER(RnMlpConfig), c_int]),
    'rn_mlp_pack': (c_int, [POINTER(c_void_p), _P, c_int, _P]),
    'rn_mlp_saved_bytes': (c_size_t, [POINTER(RnMlpConfig), c_int64]),
    'rn_mlp_forward': (c_int, [POINTER(RnMlpConfig), _P, _P, _P, _P, _P, _P, c_int64, c_int, POINTER(RnMlpOutputs), _P,
                               c_size_t, _P, c_size_t, _P]),
    'rn_mlp_backward': (c_int, [POINTER(RnMlpConfig), _P, _P, _P, _P, _P, _P, c_int64, c_int, POINTER(RnMlpOutputs),
                                POINTER(c_void_p), _P, c_size_t, _P, c_size_t, _P]),
    'rn_gemm_scratch_bytes': (c_size_t, [c_int64, c_int, c_int]),
    'rn_gemm_test': (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    'rn_wgrad_test': (c_int, [_P, _P, c_int64, c_int, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    'rn_gemm_bench': (c_int, [c_int64, c_int, c_int, c_int, POINTER(c_float), _P, c_size_t, _P]),
    'rn_launch_count': (c_int64, []),
    'rn_prof_enable': (c_int, [c_int]),
    'rn_prof_summary': (c_int, [c_int, POINTER(c_int64), POINTER(c_double), POINTER(c_double)]),
    'rn_fp16_saturation_count': (c_int64, [c_int]),
    'rn_prof_summary2': (c_int, [c_int, POINTER(c_int64), POINTER(c_double), POINTER(c_double), POINTER(c_double)]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib = None


def load():
    """Load (once) and return the ctypes handle; raises if the library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build the sm_100a CUDA library first (`python -m refnerf_pl_b200.build`). '
                'refnerf_pl_b200 has no CPU / eager fallback.')
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().rn_last_error()
        raise RuntimeError(f'refnerf_b200 error {rc}: {msg.decode() if msg else "?"}')


def param_names():
    lib = load()
    return [lib.rn_mlp_param_name(i).decode() for i in range(NUM_PARAMS)]
