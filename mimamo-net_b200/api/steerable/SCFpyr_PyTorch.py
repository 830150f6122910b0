"""Drop-in for api/steerable/SCFpyr_PyTorch.py: owner of the pyramid configuration and of the
device-resident plans (host-built masks folded into DCT-domain tables, uploaded once).

The reference rebuilds and re-uploads every mask on every `build` call
(SCFpyr_PyTorch.py:94-107,145-146,156-158,193-196); here a plan is created the first time a
(frame size, levels) pair is seen and reused for the lifetime of the object.
"""
import ctypes

import numpy as np
import torch

import _native
from steerable import plan_tables


class _Plan(object):
    """RAII wrapper of a mimamo_pyr_plan."""

    def __init__(self, tables):
        self.tables = tables
        self.crops = [lv.c for lv in tables.levels]
        lib = _native.lib()
        descs = (_native.PyrLevelDesc * len(tables.levels))()
        self._keep = []
        for d, lv in zip(descs, tables.levels):
            trig = np.ascontiguousarray(lv.trig, np.float32)
            masks = np.ascontiguousarray(lv.masks, np.float32)
            sel = np.ascontiguousarray(lv.inner_sel, np.int32)
            self._keep += [trig, masks, sel]
            d.c, d.h, d.hp, d.cp = lv.c, lv.h, lv.hp, lv.cp
            d.trig_host = _native.f32_host_ptr(trig)
            d.masks_host = _native.f32_host_ptr(masks)
            d.inner_sel_host = sel.ctypes.data_as(_native.c_int32_p)
        dct = np.ascontiguousarray(tables.dct_t, np.float32)
        handle = _native.vp()
        _native.check(lib.mimamo_pyr_plan_create(tables.H, tables.Hp, tables.Kp, tables.nbands,
                                                 _native.f32_host_ptr(dct), len(tables.levels), descs,
                                                 ctypes.byref(handle)))
        self.handle = handle
        self._keep = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _native.lib().mimamo_pyr_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class SCFpyr_PyTorch(object):
    '''Complex steerable pyramid (Portilla & Simoncelli) configuration object.

    Same constructor as the reference (api/steerable/SCFpyr_PyTorch.py:51-65), including the
    `torch.set_default_dtype` side effect (:59).  Only float32 has CUDA kernels.
    '''

    def __init__(self, height=5, nbands=4, scale_factor=2, device=None, precision=32):
        self.height = height
        self.nbands = nbands
        self.scale_factor = scale_factor
        self.device = torch.device('cpu') if device is None else device
        self.precision = precision
        assert self.precision in [32, 64]
        self.dtype = torch.float32 if precision == 32 else torch.float64
        torch.set_default_dtype(self.dtype)
        if scale_factor != 2:
            raise RuntimeError('only scale_factor=2 is supported (the reference crops the spectrum by '
                               'two regardless of scale_factor, SCFpyr_PyTorch.py:179-190)')
        self._plans = {}

    # -- argument checks, worded like the reference (SCFpyr_PyTorch.py:81-91) --------------
    def _check_frames(self, im_batch):
        assert im_batch.device == self.device, 'Devices invalid (pyr = {}, batch = {})'.format(self.device, im_batch.device)
        assert im_batch.dtype == self.dtype, 'Image batch must be torch.float{}'.format(self.precision)
        assert im_batch.dim() == 4, 'Image batch must be of shape [N,C,H,W]'
        assert im_batch.shape[1] == 1, 'Second dimension must be 1 encoding grayscale image'

    def plan_for(self, frame_size, levels):
        """Device plan for mirror-extended frame_size x frame_size frames and the given coeff levels."""
        key = (int(frame_size), tuple(int(l) for l in levels))
        plan = self._plans.get(key)
        if plan is None:
            tables = plan_tables.build_tables(key[0], self.height, self.nbands, key[1])
            _native.require_cuda('SCFpyr_PyTorch')
            if self.precision != 32:
                raise RuntimeError('the CUDA pyramid kernels compute in float32 only')
            plan = self._plans[key] = _Plan(tables)
        return plan

    def build(self, im_batch):
        '''Full pyramid [hi0, [bands...], ..., lo] of arbitrary (non-mirrored) images.

        Not on the inference hot path: `Phase_Difference_Extractor` only consumes the oriented
        bands of mirror-extended frames (api/phase_difference_extractor.py:44-47,76-86), which is
        what the CUDA kernels implement.  The high/low residuals are never read downstream
         and are not built.'''
        self._check_frames(im_batch)
        raise NotImplementedError('SCFpyr_PyTorch.build of un-mirrored images is outside the B200 hot path; '
                                  'use Phase_Difference_Extractor.build_pyramid')
