"""Face-crop preprocessing on the GPU, bit-exact with the reference's PIL / torchvision transforms.

    gray : Image.open(bmp).convert('L') -> Resize(phase_size, LANCZOS) -> float / 255
           (api/sampler/snippet_sampler.py:156-185, api/utils/data_utils.py:71-120)
    RGB  : Resize(256) -> CenterCrop(224) -> ToTensor -> x * 255 -> Normalize(mean, [1,1,1])
           (api/utils/model_utils.py:26-40)

Input everywhere: uint8 CUDA tensor (n, S, S, 3), RGB, HWC -- the decoded bytes of OpenFace's aligned
`frame_det_00_%06d.bmp` crops.  No CPU fallback: the kernels live in libmimamo_b200.so.
"""
import ctypes

import numpy as np
import torch

import _native
from utils.pil_tables import resample_table

MEAN = (131.0912, 103.8827, 91.4953)


class Crop_Preprocessor(object):
    def __init__(self, save_size=112, phase_size=48, resize=256, crop=224, mean=MEAN):
        _native.require_cuda('Crop_Preprocessor')
        self.save_size, self.phase_size, self.resize, self.crop = save_size, phase_size, resize, crop
        g_taps, g_bounds, g_kk = resample_table(save_size, phase_size, 'lanczos')
        r_taps, r_bounds, r_kk = resample_table(save_size, resize, 'bilinear')
        crop_off = int(round((resize - crop) / 2.0))                # torchvision center_crop
        mean = np.asarray(mean, dtype=np.float32)
        i32p = ctypes.POINTER(ctypes.c_int32)
        self.handle = _native.vp()
        _native.check(_native.lib().mimamo_preproc_create(
            save_size, phase_size, g_taps, g_bounds.ctypes.data_as(i32p), g_kk.ctypes.data_as(i32p),
            resize, r_taps, r_bounds.ctypes.data_as(i32p), r_kk.ctypes.data_as(i32p),
            crop, crop_off, _native.f32_host_ptr(mean), ctypes.byref(self.handle)))

    def __del__(self):
        try:
            if self.handle:
                _native.lib().mimamo_preproc_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _check(self, crops):
        assert crops.is_cuda and crops.dtype == torch.uint8, 'crops must be a uint8 CUDA tensor'
        if crops.dim() != 4 or tuple(crops.shape[1:]) != (self.save_size, self.save_size, 3):
            raise ValueError('crops must have shape (n, %d, %d, 3)' % (self.save_size, self.save_size))
        return crops.contiguous()

    def gray(self, crops):
        """(n,S,S,3) uint8 -> (n, phase_size, phase_size) float32 in [0,1]."""
        crops = self._check(crops)
        out = torch.empty((crops.shape[0], self.phase_size, self.phase_size), dtype=torch.float32, device=crops.device)
        _native.check(_native.lib().mimamo_crops_to_gray(self.handle, _native.dptr(crops), crops.shape[0],
                                                         _native.dptr(out), _native.stream_ptr(crops.device)))
        return out

    def rgb(self, crops):
        """(n,S,S,3) uint8 -> (n, 3, crop, crop) float32, what Image_Sampler's transform yields."""
        crops = self._check(crops)
        out = torch.empty((crops.shape[0], 3, self.crop, self.crop), dtype=torch.float32, device=crops.device)
        _native.check(_native.lib().mimamo_crops_to_rgb(self.handle, _native.dptr(crops), crops.shape[0],
                                                        _native.dptr(out), _native.stream_ptr(crops.device)))
        return out
