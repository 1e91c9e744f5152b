"""ctypes binding of libhamer_b200.so (the C ABI declared in include/hamer_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('HAMER_B200_LIB') or os.path.join(_HERE, 'libhamer_b200.so')   # (the override is for A/B builds)

c_float_p = C.POINTER(C.c_float)


class HmConfig(C.Structure):
    _fields_ = [('max_batch', C.c_int32), ('max_frame_bytes', C.c_int32),
                ('gemm_cta_group', C.c_int32), ('attention_legacy', C.c_int32), ('mano_legacy', C.c_int32),
                ('deterministic', C.c_int32), ('split_weights', C.c_int32), ('reserved', C.c_int32 * 1)]


class HmOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        'pred_cam', 'global_orient', 'hand_pose', 'betas', 'pred_cam_t', 'focal_length',
        'pred_keypoints_3d', 'pred_vertices', 'pred_keypoints_2d', 'conditioning_feats')]


class HmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f'hamer_b200 error {code}: {msg}')
        self.code = code


# every symbol include/hamer_b200.h declares (tests check the library exports all of them)
EXPORTS = ('hm_create', 'hm_destroy', 'hm_last_error', 'hm_set_weight', 'hm_finalize_weights',
           'hm_set_mano', 'hm_preprocess', 'hm_preprocess_centers', 'hm_preprocess_blur', 'hm_forward', 'hm_backbone', 'hm_mano_head',
           'hm_mano_lbs', 'hm_camera_postprocess', 'hm_rotmat_to_axis_angle', 'hm_letterbox_shape', 'hm_letterbox', 'hm_scale_boxes', 'hm_rootnet_finalize', 'hm_rootnet_depth', 'hm_comm_unique_id', 'hm_comm_init', 'hm_comm_destroy', 'hm_allgather', 'hm_forward_launch_count', 'hm_version', 'hm_profile_enable',
           'hm_profile_collect', 'hm_test_rootnet_patches', 'hm_test_gemm', 'hm_test_layernorm', 'hm_test_attention', 'hm_test_mano_clocks')

PROFILE_CATEGORIES = ('im2col', 'gemm_patch', 'layernorm', 'gemm_qkv', 'attention', 'gemm_proj',
                      'gemm_fc1', 'gemm_fc2', 'head', 'mano_lbs', 'crop')

_lib = None


def load():
    """Load the shared library (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f'{LIB_PATH} not found - build it with `python -c "import __graft_entry__ as g; '
                          f'g.build()"` (nvcc -gencode arch=compute_100a,code=sm_100a); there is no CPU fallback')
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.hm_create.argtypes = [C.POINTER(vp), i32, C.POINTER(HmConfig)]
    lib.hm_destroy.argtypes = [vp]
    lib.hm_destroy.restype = None
    lib.hm_last_error.argtypes = [vp]
    lib.hm_last_error.restype = C.c_char_p
    lib.hm_set_weight.argtypes = [vp, C.c_char_p, vp, C.POINTER(i64), i32]
    lib.hm_finalize_weights.argtypes = [vp]
    lib.hm_set_mano.argtypes = [vp] + [vp] * 8
    lib.hm_preprocess.argtypes = [vp, vp, i32, i32, i64, i32, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.hm_preprocess_centers.argtypes = [vp, vp, i32, i32, i64, i32, vp, vp, i32, vp, vp, vp]
    lib.hm_preprocess_blur.argtypes = [vp, vp, i32, i32, i64, i32, vp, vp, vp, vp, i32, vp, vp, vp]
    lib.hm_forward.argtypes = [vp, vp, i32, C.POINTER(HmOutputs), vp]
    lib.hm_backbone.argtypes = [vp, vp, i32, vp, vp]
    lib.hm_mano_head.argtypes = [vp, vp, i32, C.POINTER(HmOutputs), vp]
    lib.hm_mano_lbs.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp]
    lib.hm_camera_postprocess.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp]
    lib.hm_rotmat_to_axis_angle.argtypes = [vp, vp, i32, vp, vp]
    lib.hm_letterbox_shape.argtypes = [i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    lib.hm_letterbox.argtypes = [vp, vp, i32, i32, C.c_int64, i32, i32, i32, i32, vp, C.POINTER(vp), vp]
    lib.hm_scale_boxes.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.hm_rootnet_finalize.argtypes = [vp]
    lib.hm_rootnet_depth.argtypes = [vp, vp, i32, i32, i64, i32, vp, vp, i32, vp, vp]
    lib.hm_comm_unique_id.argtypes = [vp]
    lib.hm_comm_init.argtypes = [vp, vp, i32, i32]
    lib.hm_comm_destroy.argtypes = [vp]
    lib.hm_allgather.argtypes = [vp, vp, vp, vp, i32, vp]
    lib.hm_forward_launch_count.argtypes = [vp]
    lib.hm_version.argtypes = []
    lib.hm_version.restype = C.c_char_p
    lib.hm_profile_enable.argtypes = [vp, i32]
    lib.hm_profile_collect.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int32), i32]
    lib.hm_test_rootnet_patches.argtypes = [vp, i32, vp, vp]
    lib.hm_test_gemm.argtypes =[vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, i32, i32, vp]
    lib.hm_test_mano_clocks.argtypes = [vp, vp, vp, i32, vp, vp, vp, i32, vp]
    lib.hm_test_layernorm.argtypes = [vp, vp, vp, vp, vp, i32, C.c_float, i32, vp]
    lib.hm_test_attention.argtypes = [vp, i32, vp, vp, i32, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ('hm_destroy', 'hm_last_error', 'hm_version'):
            fn.restype = i32
    _lib = lib
    return lib


def check(ctx, rc):
    if rc != 0:
        msg = load().hm_last_error(ctx)
        raise HmError(rc, msg.decode() if msg else '')
    return rc
