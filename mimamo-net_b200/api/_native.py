"""ctypes binding of libmimamo_b200.so (include/mimamo_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails the
error is raised to the caller.  Build the library with `python __graft_entry__.py` (or
`__graft_entry__.build()`), which runs nvcc for sm_100a.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libmimamo_b200.so")

E_VALUE, E_RUNTIME, E_CUDA = -1, -2, -3
MAX_LEVELS = 8

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
vp = ctypes.c_void_p


class PyrLevelDesc(ctypes.Structure):
    _fields_ = [("c", ctypes.c_int32), ("h", ctypes.c_int32), ("hp", ctypes.c_int32), ("cp", ctypes.c_int32),
                ("trig_host", c_float_p), ("masks_host", c_float_p), ("inner_sel_host", c_int32_p)]


class ScfUnitDesc(ctypes.Structure):
    _fields_ = [("s", ctypes.c_int32), ("planes", ctypes.c_int32), ("is_real", ctypes.c_int32),
                ("twist_build", ctypes.c_int32), ("twist_recon", ctypes.c_int32),
                ("src_index_host", c_int32_p), ("build_mask_host", c_float_p), ("recon_mask_host", c_float_p)]


class TensorDesc(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("data_host", c_float_p), ("ndim", ctypes.c_int32),
                ("shape", ctypes.c_int64 * 4)]


_SIGNATURES = {
    "mimamo_abi_version": (ctypes.c_int, []),
    "mimamo_last_error": (ctypes.c_char_p, []),
    "mimamo_launch_count": (ctypes.c_uint64, []),
    "mimamo_pyr_plan_create": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                              c_float_p, ctypes.c_int32, ctypes.POINTER(PyrLevelDesc),
                                              ctypes.POINTER(vp)]),
    "mimamo_pyr_plan_destroy": (None, [vp]),
    "mimamo_pyr_build_workspace_bytes": (ctypes.c_int, [vp, ctypes.c_int64, ctypes.c_int32,
                                                        ctypes.POINTER(ctypes.c_size_t)]),
    "mimamo_pyr_build": (ctypes.c_int, [vp, vp, ctypes.c_int64, ctypes.c_int32, ctypes.POINTER(vp), vp,
                                        ctypes.c_size_t, vp]),
    "mimamo_scf_plan_create": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ScfUnitDesc), ctypes.POINTER(vp)]),
    "mimamo_scf_plan_destroy": (None, [vp]),
    "mimamo_scf_workspace_bytes": (ctypes.c_int, [vp, ctypes.c_int64, ctypes.POINTER(ctypes.c_size_t)]),
    "mimamo_scf_build": (ctypes.c_int, [vp, vp, ctypes.c_int64, ctypes.POINTER(vp), vp, ctypes.c_size_t, vp]),
    "mimamo_scf_reconstruct": (ctypes.c_int, [vp, ctypes.POINTER(vp), ctypes.c_int64, vp, vp, ctypes.c_size_t, vp]),
    "mimamo_phase_extract_workspace_bytes": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                                            ctypes.c_int32, ctypes.POINTER(ctypes.c_size_t)]),
    "mimamo_phase_extract": (ctypes.c_int, [vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                            vp, vp, ctypes.c_size_t, vp]),
    "mimamo_phase_extract_ex": (ctypes.c_int, [vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                               vp, vp, ctypes.c_size_t, vp]),
    "mimamo_pyr_phase_workspace_bytes": (ctypes.c_int, [vp, ctypes.c_int64, ctypes.c_int32,
                                                        ctypes.POINTER(ctypes.c_size_t)]),
    "mimamo_pyr_phase": (ctypes.c_int, [vp, vp, ctypes.c_int64, ctypes.c_int32, ctypes.POINTER(vp), vp,
                                        ctypes.c_size_t, vp]),
    "mimamo_pyr_phase_indexed_workspace_bytes": (ctypes.c_int, [vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32,
                                                                ctypes.POINTER(ctypes.c_size_t)]),
    "mimamo_pyr_phase_indexed": (ctypes.c_int, [vp, vp, ctypes.c_int64, vp, ctypes.c_int64, ctypes.c_int32,
                                                ctypes.POINTER(vp), vp, ctypes.c_size_t, vp]),
    "mimamo_pyr_phase_indexed_nhwc16": (ctypes.c_int, [vp, vp, ctypes.c_int64, vp, ctypes.c_int64, ctypes.c_int32,
                                                       ctypes.POINTER(vp), c_int32_p, c_int32_p, vp, ctypes.c_size_t, vp]),
    "mimamo_preproc_create": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_int32_p, c_int32_p,
                                             ctypes.c_int32, ctypes.c_int32, c_int32_p, c_int32_p, ctypes.c_int32,
                                             ctypes.c_int32, c_float_p, ctypes.POINTER(vp)]),
    "mimamo_preproc_destroy": (None, [vp]),
    "mimamo_preproc_geometry": (ctypes.c_int, [vp, c_int32_p, c_int32_p, c_int32_p]),
    "mimamo_crops_to_gray": (ctypes.c_int, [vp, vp, ctypes.c_int64, vp, vp]),
    "mimamo_crops_to_rgb": (ctypes.c_int, [vp, vp, ctypes.c_int64, vp, vp]),
    "mimamo_resnet50_pool5_crops": (ctypes.c_int, [vp, vp, vp, ctypes.c_int32, vp, vp, ctypes.c_size_t, vp]),
    "mimamo_resnet50_create": (ctypes.c_int, [ctypes.POINTER(TensorDesc), ctypes.c_int32, ctypes.POINTER(vp)]),
    "mimamo_resnet50_destroy": (None, [vp]),
    "mimamo_resnet50_workspace_bytes": (ctypes.c_int, [vp, ctypes.c_int32, ctypes.POINTER(ctypes.c_size_t)]),
    "mimamo_resnet50_pool5": (ctypes.c_int, [vp, vp, ctypes.c_int32, vp, vp, ctypes.c_size_t, vp]),
    "mimamo_resnet50_calibrate": (ctypes.c_int, [vp, vp, ctypes.c_int32, vp, ctypes.c_size_t, vp]),
    "mimamo_head_create": (ctypes.c_int, [ctypes.POINTER(TensorDesc), ctypes.c_int32, ctypes.c_int32,
                                          ctypes.POINTER(vp)]),
    "mimamo_head_destroy": (None, [vp]),
    "mimamo_head_workspace_bytes": (ctypes.c_int, [vp, ctypes.c_int32, ctypes.c_int32,
                                                   ctypes.POINTER(ctypes.c_size_t)]),
    "mimamo_head_forward": (ctypes.c_int, [vp, vp, vp, vp, ctypes.c_int32, ctypes.c_int32, vp, vp,
                                           ctypes.c_size_t, vp]),
    "mimamo_head_forward_nhwc16": (ctypes.c_int, [vp, vp, ctypes.c_int32, vp, vp, ctypes.c_int32, ctypes.c_int32, vp, vp,
                                                  ctypes.c_size_t, vp]),
    "mimamo_mlp_create": (ctypes.c_int, [ctypes.POINTER(TensorDesc), ctypes.c_int32, ctypes.POINTER(vp)]),
    "mimamo_mlp_destroy": (None, [vp]),
    "mimamo_mlp_in_features": (ctypes.c_int, [vp]),
    "mimamo_mlp_workspace_bytes": (ctypes.c_int, [vp, ctypes.c_int32, ctypes.POINTER(ctypes.c_size_t)]),
    "mimamo_mlp_forward": (ctypes.c_int, [vp, vp, ctypes.c_int32, vp, vp, ctypes.c_size_t, vp]),
    "mimamo_phasenet_create": (ctypes.c_int, [ctypes.POINTER(TensorDesc), ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                              ctypes.POINTER(vp)]),
    "mimamo_phasenet_destroy": (None, [vp]),
    "mimamo_phasenet_workspace_bytes": (ctypes.c_int, [vp, ctypes.c_int32, ctypes.POINTER(ctypes.c_size_t)]),
    "mimamo_phasenet_forward": (ctypes.c_int, [vp, vp, vp, ctypes.c_int32, ctypes.c_int32, vp, vp, ctypes.c_size_t, vp]),
    "mimamo_profile_gemm": (ctypes.c_int, [ctypes.c_int32]),
    "mimamo_profile_gemm_read": (ctypes.c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64),
                                                ctypes.POINTER(ctypes.c_double)]),
    "mimamo_profile_gemm_launches": (ctypes.c_int, [c_float_p, ctypes.c_int32]),
    "mimamo_conv_bf16": (ctypes.c_int, [vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                        c_float_p, c_float_p, c_float_p, ctypes.c_int32, ctypes.c_int32,
                                        ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, vp, vp, vp]),
    "mimamo_conv_chain_bf16": (ctypes.c_int, [vp, ctypes.c_int32, ctypes.c_int32, c_float_p, c_float_p, c_float_p, ctypes.c_int32,
                                              vp, c_float_p, c_float_p, c_float_p, ctypes.c_int32, vp, vp, vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib = None


def lib():
    """The loaded library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libmimamo_b200.so not found at %s -- run `python __graft_entry__.py` to build "
                               "the sm_100a CUDA library; there is no CPU fallback" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    """Translate the C ABI's error channel into the exception type the reference raises."""
    if rc == 0:
        return
    msg = lib().mimamo_last_error().decode("utf-8", "replace")
    if rc == E_VALUE:
        raise ValueError(msg)
    raise RuntimeError(msg)


def require_cuda(what):
    if not torch.cuda.is_available():
        raise RuntimeError("%s needs a CUDA device (sm_100a); this build has no CPU fallback" % what)


def stream_ptr(device=None):
    return vp(torch.cuda.current_stream(device).cuda_stream)


def dptr(t):
    return vp(t.data_ptr())


def f32_host_ptr(arr):
    return arr.ctypes.data_as(c_float_p)


def ptr_array(tensors):
    arr = (vp * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def launch_count():
    return int(lib().mimamo_launch_count())
