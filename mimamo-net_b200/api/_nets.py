"""Native net handles: a reference-keyed state_dict -> mimamo_resnet50 / mimamo_head (C ABI)."""
import ctypes

import numpy as np
import torch

import _native


def _tensor_table(state_dict):
    """(ctypes array of mimamo_tensor_desc, keep-alive list) for every floating tensor."""
    items = [(k, v) for k, v in state_dict.items() if torch.is_tensor(v) and v.is_floating_point()]
    table = (_native.TensorDesc * len(items))()
    keep = []
    for d, (k, v) in zip(table, items):
        arr = np.ascontiguousarray(v.detach().to('cpu', torch.float32).numpy())
        name = k.encode()
        keep += [arr, name]
        d.name = name
        d.data_host = _native.f32_host_ptr(arr)
        d.ndim = arr.ndim
        if arr.ndim > 4:
            raise ValueError('tensor %s has more than 4 dims' % k)
        for i, s in enumerate(arr.shape):
            d.shape[i] = s
    return table, len(items), keep


class _Handle(object):
    _destroy = None

    def __init__(self):
        self.handle = _native.vp()
        self._ws = None

    def workspace(self, nbytes, device):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
            self._ws = torch.empty((nbytes,), dtype=torch.uint8, device=device)
        return self._ws

    def __del__(self):
        try:
            if self.handle:
                getattr(_native.lib(), self._destroy)(self.handle)
                self.handle = None
        except Exception:
            pass


class NativeResNet50(_Handle):
    _destroy = 'mimamo_resnet50_destroy'

    def __init__(self, state_dict):
        super().__init__()
        _native.require_cuda('Resnet50_Extractor')
        table, n, keep = _tensor_table(state_dict)
        _native.check(_native.lib().mimamo_resnet50_create(table, n, ctypes.byref(self.handle)))

    def pool5(self, image):
        """image (bs,3,224,224) float32 CUDA, 0-255 scale minus mean -> (bs,2048) float32 CUDA."""
        assert image.is_cuda and image.dtype == torch.float32 and tuple(image.shape[1:]) == (3, 224, 224), \
            'expected a float32 CUDA batch of shape (bs,3,224,224)'
        image = image.contiguous()
        bs = image.shape[0]
        out = torch.empty((bs, 2048), dtype=torch.float32, device=image.device)
        lib = _native.lib()
        need = ctypes.c_size_t(0)
        _native.check(lib.mimamo_resnet50_workspace_bytes(self.handle, bs, ctypes.byref(need)))
        ws = self.workspace(need.value, image.device)
        _native.check(lib.mimamo_resnet50_pool5(self.handle, _native.dptr(image), bs, _native.dptr(out), _native.dptr(ws),
                                                ws.numel(), _native.stream_ptr(image.device)))
        return out


    def pool5_crops(self, preproc, crops):
        """crops (bs,S,S,3) uint8 CUDA (OpenFace face crops) -> (bs,2048) float32 CUDA; the
        Resize/CenterCrop/mean transform runs on the device and feeds conv1 directly."""
        assert crops.is_cuda and crops.dtype == torch.uint8 and crops.dim() == 4 and crops.shape[-1] == 3, \
            'expected a uint8 CUDA batch of shape (bs,S,S,3)'
        crops = crops.contiguous()
        bs = crops.shape[0]
        out = torch.empty((bs, 2048), dtype=torch.float32, device=crops.device)
        lib = _native.lib()
        need = ctypes.c_size_t(0)
        _native.check(lib.mimamo_resnet50_workspace_bytes(self.handle, bs, ctypes.byref(need)))
        ws = self.workspace(need.value, crops.device)
        _native.check(lib.mimamo_resnet50_pool5_crops(self.handle, preproc.handle, _native.dptr(crops), bs, _native.dptr(out),
                                                      _native.dptr(ws), ws.numel(), _native.stream_ptr(crops.device)))
        return out


class NativeHead(_Handle):
    _destroy = 'mimamo_head_destroy'

    def __init__(self, state_dict, num_phase=12):
        super().__init__()
        _native.require_cuda('Two_Stream_RNN')
        table, n, keep = _tensor_table(state_dict)
        _native.check(_native.lib().mimamo_head_create(table, n, num_phase, ctypes.byref(self.handle)))

        self.channels = 2 * num_phase

    def forward(self, phase_0, phase_1, rgb):
        bs, nf = rgb.shape[0], rgb.shape[1]
        for t in (phase_0, phase_1, rgb):
            assert t.is_cuda and t.dtype == torch.float32, 'head inputs must be float32 CUDA tensors'
        c = self.channels
        assert tuple(phase_0.shape) == (bs, nf, c, 48, 48) and tuple(phase_1.shape) == (bs, nf, c, 24, 24) \
            and rgb.shape[2] == 2048, 'unexpected head input shapes'
        phase_0, phase_1, rgb = phase_0.contiguous(), phase_1.contiguous(), rgb.contiguous()
        out = torch.empty((bs, nf, 2), dtype=torch.float32, device=rgb.device)
        lib = _native.lib()
        need = ctypes.c_size_t(0)
        _native.check(lib.mimamo_head_workspace_bytes(self.handle, bs, nf, ctypes.byref(need)))
        ws = self.workspace(need.value, rgb.device)
        _native.check(lib.mimamo_head_forward(self.handle, _native.dptr(phase_0), _native.dptr(phase_1), _native.dptr(rgb),
                                              bs, nf, _native.dptr(out), _native.dptr(ws), ws.numel(),
                                              _native.stream_ptr(rgb.device)))
        return out

    def forward_nhwc16(self, phase0, cat, rgb):
        """phase0 f16 (bs*nf, 48, 48, C), cat f16 (bs*nf, 24, 24, 128) as Phase_Difference_Extractor.phasenet_operands
        returns them (cat[..., :64] is overwritten), rgb f32 (bs, nf, 2048) -> (bs, nf, 2)."""
        bs, nf = rgb.shape[0], rgb.shape[1]
        c = self.channels
        assert phase0.is_cuda and phase0.dtype == torch.float16 and tuple(phase0.shape) == (bs * nf, 48, 48, c) and phase0.is_contiguous()
        assert cat.is_cuda and cat.dtype == torch.float16 and tuple(cat.shape) == (bs * nf, 24, 24, 128) and cat.is_contiguous()
        assert rgb.is_cuda and rgb.dtype == torch.float32 and rgb.shape[2] == 2048
        rgb = rgb.contiguous()
        out = torch.empty((bs, nf, 2), dtype=torch.float32, device=rgb.device)
        lib = _native.lib()
        need = ctypes.c_size_t(0)
        _native.check(lib.mimamo_head_workspace_bytes(self.handle, bs, nf, ctypes.byref(need)))
        ws = self.workspace(need.value, rgb.device)
        _native.check(lib.mimamo_head_forward_nhwc16(self.handle, _native.dptr(phase0), c, _native.dptr(cat), _native.dptr(rgb),
                                                     bs, nf, _native.dptr(out), _native.dptr(ws), ws.numel(),
                                                     _native.stream_ptr(rgb.device)))
        return out


class NativeMLP(_Handle):
    """MLP.forward on its own (mimamo_mlp_*): state_dict keys `mlp.{1,2,5,6,...}`."""
    _destroy = 'mimamo_mlp_destroy'

    def __init__(self, state_dict):
        super().__init__()
        _native.require_cuda('MLP')
        table, n, keep = _tensor_table(state_dict)
        _native.check(_native.lib().mimamo_mlp_create(table, n, ctypes.byref(self.handle)))
        self.in_features = _native.lib().mimamo_mlp_in_features(self.handle)

    def forward(self, x):
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[1] == self.in_features, \
            'expected a float32 CUDA matrix with %d columns' % self.in_features
        x = x.contiguous()
        rows = x.shape[0]
        out = torch.empty((rows, 256), dtype=torch.float32, device=x.device)
        lib = _native.lib()
        need = ctypes.c_size_t(0)
        _native.check(lib.mimamo_mlp_workspace_bytes(self.handle, rows, ctypes.byref(need)))
        ws = self.workspace(max(need.value, 8), x.device)
        _native.check(lib.mimamo_mlp_forward(self.handle, _native.dptr(x), rows, _native.dptr(out), _native.dptr(ws),
                                             ws.numel(), _native.stream_ptr(x.device)))
        return out


class NativePhaseNet(_Handle):
    """PhaseNet.forward on its own (mimamo_phasenet_*): keys `conv_net.*`, `fc.*`, `classifier.*`."""
    _destroy = 'mimamo_phasenet_destroy'

    def __init__(self, state_dict, num_channels, input_size=48):
        super().__init__()
        _native.require_cuda('PhaseNet')
        table, n, keep = _tensor_table(state_dict)
        _native.check(_native.lib().mimamo_phasenet_create(table, n, input_size, num_channels, ctypes.byref(self.handle)))
        self.channels = num_channels
        self.size = input_size

    def forward(self, level0, level1, feature):
        rows = level0.shape[0]
        for t in (level0, level1):
            assert t.is_cuda and t.dtype == torch.float32, 'PhaseNet inputs must be float32 CUDA tensors'
        assert tuple(level0.shape) == (rows, self.channels, self.size, self.size) and \
            tuple(level1.shape) == (rows, self.channels, self.size // 2, self.size // 2), 'unexpected PhaseNet input shapes'
        level0, level1 = level0.contiguous(), level1.contiguous()
        out = torch.empty((rows, 256 if feature else 1), dtype=torch.float32, device=level0.device)
        lib = _native.lib()
        need = ctypes.c_size_t(0)
        _native.check(lib.mimamo_phasenet_workspace_bytes(self.handle, rows, ctypes.byref(need)))
        ws = self.workspace(need.value, level0.device)
        _native.check(lib.mimamo_phasenet_forward(self.handle, _native.dptr(level0), _native.dptr(level1), rows,
                                                  1 if feature else 0, _native.dptr(out), _native.dptr(ws), ws.numel(),
                                                  _native.stream_ptr(level0.device)))
        return out
