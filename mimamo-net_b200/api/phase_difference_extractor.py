"""Drop-in for api/phase_difference_extractor.py backed by the sm_100a kernels.

Same class name, constructor, attributes, method names, tensor layouts and error behaviour as
the reference; the compute goes through libmimamo_b200.so (include/mimamo_b200.h).
"""
import ctypes

import torch

import _native
from steerable.SCFpyr_PyTorch import SCFpyr_PyTorch
from steerable.utils import get_device


class Phase_Difference_Extractor(object):
    def __init__(self, height=5, nbands=4, scale_factor=2,
                 extract_level=1, visualize=False):
        '''Steerable-pyramid phase-difference extractor (reference :7-37).

        height: pyramid levels including the high- and low-pass residuals
        nbands: orientations; scale_factor: 2; extract_level: int or list of ints (indices into
        the reference's coeff list, 1 = finest oriented level); visualize: accepted for
        signature compatibility, plotting is out of scope.
        '''
        self.pyramid = SCFpyr_PyTorch(height=height, nbands=nbands, scale_factor=scale_factor,
                                      device=get_device())
        self.height = height
        self.nbands = nbands
        self.scale_factor = scale_factor
        self.extract_level = extract_level
        self.visualize = visualize

    # ------------------------------------------------------------------------------------
    def _levels(self):
        if isinstance(self.extract_level, int):
            return [self.extract_level]
        if isinstance(self.extract_level, list):
            return list(self.extract_level)
        raise UnboundLocalError("extract_level must be an int or a list of ints")   # reference :87

    def _prepare(self, im_batch, symmetry):
        bs, num_phase_frames, W, H = im_batch.size()                   # ValueError unless 4-D (:42)
        self.pyramid._check_frames(im_batch.view(bs * num_phase_frames, 1, W, H))
        if not symmetry:
            raise NotImplementedError('symmetry=False (no mirror extension) has no CUDA kernel; the '
                                      'inference path always mirrors (reference :38,44-45)')
        if W != H:
            raise RuntimeError('frames must be square: the reference builds its masks with swapped '
                               'axes (SCFpyr_PyTorch.py:87,94) and cannot broadcast otherwise')
        plan = self.pyramid.plan_for(W, self._levels())
        return plan, bs, num_phase_frames, im_batch.contiguous()

    def build_pyramid(self, im_batch, symmetry=True):
        """im_batch (bs, T, W, H) float32 on get_device() -> coefficients (bs, nbands, T, c, c, 2)
        per requested level (a tensor for an int extract_level, a list otherwise), c = level size
        after the reference's quadrant crop (:72-75,84-85)."""
        plan, bs, T, frames = self._prepare(im_batch, symmetry)
        outs = [torch.empty((bs, self.nbands, T, c, c, 2), dtype=torch.float32, device=frames.device)
                for c in plan.crops]
        if bs * T > 0:
            lib = _native.lib()
            need = ctypes.c_size_t(0)
            _native.check(lib.mimamo_pyr_build_workspace_bytes(plan.handle, bs, T, ctypes.byref(need)))
            ws = torch.empty((max(need.value, 8),), dtype=torch.uint8, device=frames.device)
            _native.check(lib.mimamo_pyr_build(plan.handle, _native.dptr(frames), bs, T, _native.ptr_array(outs),
                                               _native.dptr(ws), ws.numel(), _native.stream_ptr(frames.device)))
        return outs[0] if isinstance(self.extract_level, int) else outs

    def extract_coeff_level(self, level, coeff_batch):
        extr_level_coeff_batch = coeff_batch[level]
        assert isinstance(extr_level_coeff_batch, list)
        return torch.stack(extr_level_coeff_batch, 0)

    def extract(self, coeff_batch):
        """coeff (bs, nbands, T, W, H, 2) -> phase differences (bs, nbands, T-1, W, H) (reference :93-134)."""
        bs, n_bands, n_phase_frames, W, H, _ = coeff_batch.size()
        _native.require_cuda('Phase_Difference_Extractor.extract')
        assert coeff_batch.is_cuda and coeff_batch.dtype == torch.float32, 'coefficients must be float32 on the GPU'
        coeff = coeff_batch.contiguous()
        out = torch.empty((bs, n_bands, n_phase_frames - 1, W, H), dtype=torch.float32, device=coeff.device)
        if out.numel() == 0:
            return out
        lib = _native.lib()
        need = ctypes.c_size_t(0)
        _native.check(lib.mimamo_phase_extract_workspace_bytes(bs * n_bands, n_phase_frames, W, H, ctypes.byref(need)))
        ws = torch.empty((max(need.value, 8),), dtype=torch.uint8, device=coeff.device)
        _native.check(lib.mimamo_phase_extract(_native.dptr(coeff), bs * n_bands, n_phase_frames, W, H,
                                               _native.dptr(out), _native.dptr(ws), ws.numel(),
                                               _native.stream_ptr(coeff.device)))
        return out

    def phase_difference(self, im_batch):
        """build_pyramid + extract in one call: (bs, T, W, H) -> per level (bs, nbands, T-1, c, c).
        This is what Tester.phase_diff_output uses; coefficients stay in a device workspace."""
        plan, bs, T, frames = self._prepare(im_batch, True)
        outs = [torch.empty((bs, self.nbands, T - 1, c, c), dtype=torch.float32, device=frames.device)
                for c in plan.crops]
        if bs == 0 or T < 2:
            return outs[0] if isinstance(self.extract_level, int) else outs
        lib = _native.lib()
        need = ctypes.c_size_t(0)
        _native.check(lib.mimamo_pyr_phase_workspace_bytes(plan.handle, bs, T, ctypes.byref(need)))
        ws = torch.empty((max(need.value, 8),), dtype=torch.uint8, device=frames.device)
        _native.check(lib.mimamo_pyr_phase(plan.handle, _native.dptr(frames), bs, T, _native.ptr_array(outs),
                                           _native.dptr(ws), ws.numel(), _native.stream_ptr(frames.device)))
        return outs[0] if isinstance(self.extract_level, int) else outs

    def phase_difference_indexed(self, frames, window_index, out=None):
        """Clip variant of phase_difference: frames (n, W, H) are transformed
        once each; window_index (n_windows, T) int32 names the frames of every window (the clamp rule
        of api/sampler/snippet_sampler.py:144-152).  Returns per level (n_windows, nbands, T-1, c, c),
        bit-identical to phase_difference(frames[window_index]).  `out`: optional list of preallocated contiguous
        float32 tensors (one per level, n_windows*nbands*(T-1)*c*c elements each) to write into."""
        n, W, H = frames.size()
        n_windows, T = window_index.size()
        assert window_index.dtype == torch.int32 and window_index.device == frames.device
        plan, _, _, frames = self._prepare(frames.view(n, 1, W, H), True)
        if out is None:
            outs = [torch.empty((n_windows, self.nbands, T - 1, c, c), dtype=torch.float32, device=frames.device)
                    for c in plan.crops]
        else:
            outs = list(out)
            for o, c in zip(outs, plan.crops):
                assert o.is_contiguous() and o.dtype == torch.float32 and o.device == frames.device \
                    and o.numel() == n_windows * self.nbands * (T - 1) * c * c, 'bad preallocated output'
        if n_windows == 0 or T < 2:
            return outs[0] if isinstance(self.extract_level, int) else outs
        window_index = window_index.contiguous()
        lib = _native.lib()
        need = ctypes.c_size_t(0)
        _native.check(lib.mimamo_pyr_phase_indexed_workspace_bytes(plan.handle, n, n_windows, T, ctypes.byref(need)))
        ws = torch.empty((max(need.value, 8),), dtype=torch.uint8, device=frames.device)
        _native.check(lib.mimamo_pyr_phase_indexed(plan.handle, _native.dptr(frames), n, _native.dptr(window_index),
                                                   n_windows, T, _native.ptr_array(outs), _native.dptr(ws), ws.numel(),
                                                   _native.stream_ptr(frames.device)))
        return outs[0] if isinstance(self.extract_level, int) else outs

    def show_3D_subplots(self, data, title, first_k_frames=None):
        raise NotImplementedError('visualisation is out of scope (the reference method uses undefined '
                                  'plt/cm names, api/phase_difference_extractor.py:136-152)')
