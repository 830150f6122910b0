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
        if W != H:
            raise RuntimeError('frames must be square: the reference builds its masks with swapped '
                               'axes (SCFpyr_PyTorch.py:87,94) and cannot broadcast otherwise')
        plan = self.pyramid.plan_for(W, self._levels())
        return plan, bs, num_phase_frames, im_batch.contiguous()

    def build_pyramid(self, im_batch, symmetry=True):
        """im_batch (bs, T, W, H) float32 on get_device() -> coefficients (bs, nbands, T, c, c, 2)
        per requested level (a tensor for an int extract_level, a list otherwise), c = level size
        after the reference's quadrant crop (:72-75,84-85)."""
        if not symmetry:
            return self._build_unmirrored(im_batch)
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

    def _build_unmirrored(self, im_batch):
        """symmetry=False (reference :38-45,76-86 without the mirror extension and the quadrant crop): the frames go
        through SCFpyr_PyTorch.build as they are and the requested levels are stacked (nb, bs*T, s, s, 2) ->
        (bs, nb, T, s, s, 2)."""
        bs, T, W, H = im_batch.size()
        coeff = self.pyramid.build(im_batch.reshape(bs * T, 1, W, H))
        if not isinstance(coeff, list):
            raise ValueError('Batch of coefficients must be a list')

        def one(level):
            stacked = self.extract_coeff_level(level, coeff)
            nb, _, s0, s1, _ = stacked.size()
            return stacked.view(nb, bs, T, s0, s1, 2).permute(1, 0, 2, 3, 4, 5).contiguous()

        return one(self.extract_level) if isinstance(self.extract_level, int) else [one(l) for l in self._levels()]

    def extract_coeff_level(self, level, coeff_batch):
        extr_level_coeff_batch = coeff_batch[level]
        assert isinstance(extr_level_coeff_batch, list)
        return torch.stack(extr_level_coeff_batch, 0)

    def extract(self, coeff_batch):
        """coeff (bs, nbands, T, W, H, 2) -> phase differences (bs, nbands, T-1, W, H) (reference :93-134)."""
        return self._tail(coeff_batch, 0)

    def _tail(self, coeff_batch, mode):
        """mimamo_phase_extract_ex: mode 0 = differences (T-1 maps), 1 = denoised phases (T), 2 = return_both layout."""
        bs, n_bands, n_phase_frames, W, H, _ = coeff_batch.size()
        _native.require_cuda('Phase_Difference_Extractor.extract')
        assert coeff_batch.is_cuda and coeff_batch.dtype == torch.float32, 'coefficients must be float32 on the GPU'
        coeff = coeff_batch.contiguous()
        slots = (n_phase_frames - 1, n_phase_frames, 2 * (n_phase_frames - 1))[mode]
        out = torch.empty((bs, n_bands, slots, W, H), dtype=torch.float32, device=coeff.device)
        if out.numel() == 0:
            return out
        if n_phase_frames < 2:
            raise ValueError('the phase tail needs at least two frames')
        lib = _native.lib()
        need = ctypes.c_size_t(0)
        _native.check(lib.mimamo_phase_extract_workspace_bytes(bs * n_bands, n_phase_frames, W, H, ctypes.byref(need)))
        ws = torch.empty((max(need.value, 8),), dtype=torch.uint8, device=coeff.device)
        _native.check(lib.mimamo_phase_extract_ex(_native.dptr(coeff), bs * n_bands, n_phase_frames, W, H, mode,
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

    def phasenet_operands(self, frames, window_index, out=None):
        """phase_difference_indexed for the two-level Tester configuration, delivered as the fp16 NHWC operands PhaseNet
        reads (mimamo_pyr_phase_indexed_nhwc16): (phase0 (n_windows, c0, c0, C) , cat (n_windows, c1, c1, 128)) with
        C = nbands * (T-1) channels, level 1 sitting at channels [64, 64 + C) of `cat` and zeros above.  Returns None when
        the configuration has no such path (callers then use phase_difference_indexed + the fp32 head entry).  `out`:
        optional preallocated (phase0, cat) pair -- `cat` must come zero-initialised (torch.zeros) the first time."""
        n, W, H = frames.size()
        n_windows, T = window_index.size()
        C = self.nbands * (T - 1)
        levels = self._levels()
        if len(levels) != 2 or (T - 1) % 4 != 0 or C > 64 or C % 8 != 0 or W != H or n_windows == 0:
            return None
        plan, _, _, frames = self._prepare(frames.view(n, 1, W, H), True)
        if plan.crops[0] > 56 or plan.crops[1] * 2 != plan.crops[0]:
            return None
        assert window_index.dtype == torch.int32 and window_index.device == frames.device
        c0, c1 = plan.crops
        if out is None:
            phase0 = torch.empty((n_windows, c0, c0, C), dtype=torch.float16, device=frames.device)
            cat = torch.zeros((n_windows, c1, c1, 128), dtype=torch.float16, device=frames.device)
        else:
            phase0, cat = out
            assert phase0.is_contiguous() and cat.is_contiguous() and phase0.dtype == cat.dtype == torch.float16
            assert tuple(phase0.shape) == (n_windows, c0, c0, C) and tuple(cat.shape) == (n_windows, c1, c1, 128)
        window_index = window_index.contiguous()
        lib = _native.lib()
        need = ctypes.c_size_t(0)
        _native.check(lib.mimamo_pyr_phase_indexed_workspace_bytes(plan.handle, n, n_windows, T, ctypes.byref(need)))
        ws = torch.empty((max(need.value, 8),), dtype=torch.uint8, device=frames.device)
        pitch = (ctypes.c_int32 * 2)(C, 128)
        offs = (ctypes.c_int32 * 2)(0, 64)
        _native.check(lib.mimamo_pyr_phase_indexed_nhwc16(plan.handle, _native.dptr(frames), n, _native.dptr(window_index),
                                                          n_windows, T, _native.ptr_array([phase0, cat]), pitch, offs,
                                                          _native.dptr(ws), ws.numel(), _native.stream_ptr(frames.device)))
        return phase0, cat

    def show_3D_subplots(self, data, title, first_k_frames=None):
        raise NotImplementedError('visualisation is out of scope (the reference method uses undefined '
                                  'plt/cm names, api/phase_difference_extractor.py:136-152)')


class Steerable_Pyramid_Phase(Phase_Difference_Extractor):
    """Drop-in for the training-side copy of the extractor, `Steerable_Pyramid_Phase` (Aff-wild-exps/utils.py:298-451,
    duplicated in OMG-exps/utils.py): same pyramid and tail, plus `extract_phase(coeff, return_phase, return_both)`
    (:367-418) -- SURVEY.md section 8(f).4.  Constructor order as in the reference (device before extract_level)."""

    def __init__(self, height=5, nbands=4, scale_factor=2, device=None, extract_level=1, visualize=False):
        super().__init__(height=height, nbands=nbands, scale_factor=scale_factor, extract_level=extract_level,
                         visualize=visualize)
        if device is not None and torch.device(device) != self.pyramid.device:
            raise RuntimeError('Steerable_Pyramid_Phase runs on {} (no CPU kernels); got device={}'.format(
                self.pyramid.device, device))
        self.device = self.pyramid.device

    def extract_phase(self, coeff_batch, return_phase=False, return_both=False):
        """coeff (bs, nbands, T, W, H, 2) -> phase differences (bs, nbands, T-1, W, H); return_phase: the denoised
        phases minus their spatial mean (bs, nbands, T, W, H); return_both: insert_tensors(differences, phases[:, :, 1:])
        (bs, nbands, 2(T-1), W, H) -- the reference's insert_tensors (:419-432) only fills the first T-1 slots."""
        if return_both:
            return self._tail(coeff_batch, 2)
        return self._tail(coeff_batch, 1 if return_phase else 0)

    def insert_tensors(self, t_a, t_b, dim):
        """Reference :419-432, kept for API completeness (torch indexing, no kernel): slots i < t_a.size(dim) of a
        tensor twice as long along `dim` alternate t_a[i // 2], t_b[i // 2]; the other half stays zero."""
        size = list(t_a.size())
        length = size[dim]
        size[dim] = 2 * length
        result = torch.zeros(size, dtype=t_a.dtype, device=t_a.device)
        for i in range(length):
            src = t_a if i % 2 == 0 else t_b
            result.narrow(dim, i, 1).copy_(src.narrow(dim, i // 2, 1))
        return result
