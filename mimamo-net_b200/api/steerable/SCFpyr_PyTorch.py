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


class _FullPlan(object):
    """RAII wrapper of a mimamo_scf_plan (full pyramid of un-mirrored S x S images)."""

    def __init__(self, size, height, nbands):
        self.units = plan_tables.full_pyramid_tables(size, height, nbands)
        self.size = size
        lib = _native.lib()
        descs = (_native.ScfUnitDesc * len(self.units))()
        keep = []
        for d, u in zip(descs, self.units):
            src = np.ascontiguousarray(u.src_index, np.int32)
            bm = np.ascontiguousarray(u.build_mask, np.float32)
            rm = np.ascontiguousarray(u.recon_mask, np.float32)
            keep += [src, bm, rm]
            d.s, d.planes, d.is_real = u.s, u.planes, int(u.is_real)
            d.twist_build, d.twist_recon = u.twist_build, u.twist_recon
            d.src_index_host = src.ctypes.data_as(_native.c_int32_p)
            d.build_mask_host = _native.f32_host_ptr(bm)
            d.recon_mask_host = _native.f32_host_ptr(rm)
        handle = _native.vp()
        _native.check(lib.mimamo_scf_plan_create(size, len(self.units), descs, ctypes.byref(handle)))
        self.handle = handle

    def workspace(self, n, device):
        need = ctypes.c_size_t(0)
        _native.check(_native.lib().mimamo_scf_workspace_bytes(self.handle, n, ctypes.byref(need)))
        return torch.empty((max(need.value, 8),), dtype=torch.uint8, device=device)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _native.lib().mimamo_scf_plan_destroy(self.handle)
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
        self._full_plans = {}

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

    def _full_plan(self, size):
        plan = self._full_plans.get(int(size))
        if plan is None:
            _native.require_cuda('SCFpyr_PyTorch')
            if self.precision != 32:
                raise RuntimeError('the CUDA pyramid kernels compute in float32 only')
            plan = self._full_plans[int(size)] = _FullPlan(int(size), self.height, self.nbands)
        return plan

    def build(self, im_batch):
        '''Decomposes a batch of images into a complex steerable pyramid (reference :70-125).

        im_batch [N,1,H,W] (square) -> [hi0 (N,H,W), [nbands x (N,H,W,2)], [nbands x (N,H/2,W/2,2)], ..., lo (N,h,w)].
        Computed by mimamo_scf_build (csrc/scf_generic.cu).  The inference hot path does not come through here:
        Phase_Difference_Extractor.build_pyramid uses the mirror-symmetric fast path (mimamo_pyr_build).'''
        self._check_frames(im_batch)
        n, _, rows, cols = im_batch.shape
        if self.height > int(np.floor(np.log2(min(rows, cols))) - 2):
            raise RuntimeError('Cannot build {} levels, image too small.'.format(self.height))
        if rows != cols:
            raise RuntimeError('images must be square: the reference builds its masks with swapped axes '
                               '(SCFpyr_PyTorch.py:87,94) and cannot broadcast them otherwise')
        plan = self._full_plan(rows)
        x = im_batch.reshape(n, rows, cols).contiguous()
        outs = []
        for u in plan.units:
            shape = (n, u.s, u.s) if u.is_real else (u.planes, n, u.s, u.s, 2)
            outs.append(torch.empty(shape, dtype=torch.float32, device=x.device))
        if n > 0:
            ws = plan.workspace(n, x.device)
            _native.check(_native.lib().mimamo_scf_build(plan.handle, _native.dptr(x), n, _native.ptr_array(outs),
                                                         _native.dptr(ws), ws.numel(), _native.stream_ptr(x.device)))
        return [o if u.is_real else [o[b] for b in range(u.planes)] for o, u in zip(outs, plan.units)]

    def reconstruct(self, coeff):
        '''Inverse of build (reference :214-245): coeff list -> images (N,H,W).'''
        if self.nbands != len(coeff[1]):
            raise Exception("Unmatched number of orientations")
        hi0 = coeff[0]
        n, rows, cols = hi0.shape
        plan = self._full_plan(rows)
        if len(coeff) != len(plan.units):
            raise Exception("Unmatched number of pyramid levels")
        ins = []
        for c, u in zip(coeff, plan.units):
            t = c if u.is_real else torch.stack(list(c), 0)
            expect = (n, u.s, u.s) if u.is_real else (u.planes, n, u.s, u.s, 2)
            assert tuple(t.shape) == expect and t.dtype == torch.float32 and t.device == hi0.device, \
                'coefficient shaped {} where {} is expected'.format(tuple(t.shape), expect)
            ins.append(t.contiguous())
        out = torch.empty((n, rows, cols), dtype=torch.float32, device=hi0.device)
        if n > 0:
            ws = plan.workspace(n, hi0.device)
            _native.check(_native.lib().mimamo_scf_reconstruct(plan.handle, _native.ptr_array(ins), n, _native.dptr(out),
                                                               _native.dptr(ws), ws.numel(), _native.stream_ptr(hi0.device)))
        return out
