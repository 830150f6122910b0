"""Device selection helper with the reference's semantics (api/steerable/utils.py:34-50).
The image / visualisation helpers of the reference module are out of scope."""
import torch


def get_device(device=None):
    """The reference hard-codes 'cuda:0' (api/steerable/utils.py:34); with one process per GPU the
    default here is the process's CURRENT device (cuda:0 unless torch.cuda.set_device was called)."""
    if device is None:
        device = 'cuda:%d' % torch.cuda.current_device() if torch.cuda.device_count() > 0 else 'cuda:0'
    assert isinstance(device, str)
    if 'cuda' in device:
        if torch.cuda.device_count() > 0:
            return torch.device(device)
        print('No CUDA devices found, falling back to CPU')
    # The reference falls back to a CPU torch.fft path here.  This build has no CPU compute
    # path: objects can still be constructed (host-side tables, argument checks), but every
    # compute call raises until a CUDA device is present.
    return torch.device('cpu')
