"""Import-time compatibility shim that lets the UNMODIFIED reference `api/` run on
py3.12 / torch 2.11 / numpy 2.3.

TEST INFRASTRUCTURE ONLY.  Used by `oracle/make_golden.py` (in the build
container, where /root/reference exists) to pin the oracle and to generate the
fixtures under `tests/golden/`.  Nothing on the product path imports this.

No reference source is copied or edited: the shim only
  * restores removed numpy / torch spellings (`np.complex`, `torch.rfft`,
    `torch.ifft`, callable `torch.fft`) -- used at
    api/steerable/SCFpyr_PyTorch.py:64-65,110,122,133,171,236,242,250,276;
  * stubs `matplotlib` (imported at api/steerable/utils.py:22, absent here);
  * loads api/utils/phase_utils.py with the single py2-era token
    `cuda(async=True)` (line 107, dead function) respelled so the file parses.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("MIMAMO_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "api", "steerable"))


class _CallableFFT(types.ModuleType):
    """`torch.fft` as the reference knew it: a function AND (today) a module."""

    def __init__(self, real_module):
        super().__init__(real_module.__name__)
        self.__dict__.update(real_module.__dict__)

    def __call__(self, x, signal_ndim=2, normalized=False):
        assert signal_ndim == 2 and not normalized
        return torch.view_as_real(torch.fft.fft2(torch.view_as_complex(x.contiguous())))


def _legacy_rfft(x, signal_ndim=2, normalized=False, onesided=True):
    assert signal_ndim == 2 and not normalized and not onesided
    return torch.view_as_real(torch.fft.fft2(x))


def _legacy_ifft(x, signal_ndim=2, normalized=False):
    assert signal_ndim == 2 and not normalized
    return torch.view_as_real(torch.fft.ifft2(torch.view_as_complex(x.contiguous())))


_installed = False


def install():
    """Idempotently patch the interpreter and put the reference api/ on sys.path."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if not hasattr(np, "complex"):
        np.complex = complex
    if not callable(torch.fft):
        fft_mod = _CallableFFT(torch.fft)
        torch.fft = fft_mod
        sys.modules["torch.fft"] = fft_mod
    torch.rfft = _legacy_rfft
    torch.ifft = _legacy_ifft
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    api = os.path.join(REFERENCE_ROOT, "api")
    if api not in sys.path:
        sys.path.insert(0, api)
    # utils.phase_utils: one token respelled, then exec'd under its real name.
    import utils  # the reference's namespace package api/utils
    src_path = os.path.join(api, "utils", "phase_utils.py")
    with open(src_path) as fh:
        text = fh.read().replace("cuda(async=True)", "cuda(non_blocking=True)")
    mod = types.ModuleType("utils.phase_utils")
    mod.__file__ = src_path
    exec(compile(text, src_path, "exec"), mod.__dict__)
    sys.modules["utils.phase_utils"] = mod
    utils.phase_utils = mod
    _installed = True


def load_training_utils():
    """Aff-wild-exps/utils.py (home of Steerable_Pyramid_Phase, the training-side superset of the extractor) exec'd
    under the shim: its two `cuda(async=True)` tokens (:253,275, functions not used here) are respelled so the file
    parses, mpl_toolkits is stubbed, the blur kernel is cast to float32 as on the CUDA branch, and `Tensor.cuda()` (extract_phase's return_both ends in `.cuda()`, :406) is a
    no-op on this GPU-less container.  No other change; `steerable` resolves to the reference's own api/steerable."""
    install()
    for name in ("mpl_toolkits", "mpl_toolkits.mplot3d"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    for name, attr in (("matplotlib", "cm"), ("matplotlib", "pyplot")):
        if not hasattr(sys.modules[name], attr):
            setattr(sys.modules[name], attr, sys.modules.get("matplotlib." + attr, types.ModuleType(attr)))
    src_path = os.path.join(REFERENCE_ROOT, "Aff-wild-exps", "utils.py")
    with open(src_path) as fh:
        text = fh.read().replace("cuda(async=True)", "cuda(non_blocking=True)")
    mod = types.ModuleType("affwild_utils")
    mod.__file__ = src_path
    prev = torch.get_default_dtype()
    exec(compile(text, src_path, "exec"), mod.__dict__)
    torch.set_default_dtype(prev)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        # amplitude_based_gaussian_blur (:245-257) casts its filters to float32 only on its CUDA branch (the only one the
        # training code ever took); on this CPU-only container the float64 numpy kernel would reach F.conv2d next to
        # float32 inputs and raise.  Hand it the same float32 values the CUDA branch computes with.
        make_kernel = mod.gaussian_kernel
        mod.gaussian_kernel = lambda *a, **k: make_kernel(*a, **k).astype(np.float32)
    return mod


def load():
    """Return the reference's hot-path symbols (unmodified code objects)."""
    install()
    prev = torch.get_default_dtype()
    from phase_difference_extractor import Phase_Difference_Extractor
    from steerable.SCFpyr_PyTorch import SCFpyr_PyTorch
    from steerable.SCFpyr_NumPy import SCFpyr_NumPy
    import steerable.math_utils as math_utils
    import mimamo_net
    phase_utils = sys.modules["utils.phase_utils"]
    torch.set_default_dtype(prev)
    return types.SimpleNamespace(
        Phase_Difference_Extractor=Phase_Difference_Extractor,
        SCFpyr_PyTorch=SCFpyr_PyTorch, SCFpyr_NumPy=SCFpyr_NumPy,
        math_utils=math_utils, phase_utils=phase_utils,
        Two_Stream_RNN=mimamo_net.Two_Stream_RNN, mimamo_net=mimamo_net)
