"""Device plumbing: the library handle, HBM buffers (torch tensors used as plain device memory) and
host<->device copies.  No arithmetic happens here and none happens through torch ops: every floating
point operation on the hot path is a kernel of ``libb200npc.so`` reached through :mod:`tenpy_b200._lib`.
"""
# Copyright (C) 2026 tenpy_b200 authors. Apache-2.0.

import os

import numpy as np
import torch

from ._lib import DeviceLib, B200Error, DOT_SCRATCH

__all__ = ['get_lib', 'use_library', 'empty', 'zeros', 'to_device', 'to_host', 'scalar_out', 'read_scalar',
           'B200Error']

_state = {'lib': None, 'scratch': None, 'out': None}


def get_lib():
    """Return the process-wide :class:`DeviceLib`; loads ``libb200npc.so`` on first use.

    Raises :class:`B200Error` when the extension is not built or no GPU is visible (no CPU fallback)."""
    lib = _state['lib']
    if lib is None:
        lib = _state['lib'] = DeviceLib()
    return lib


def use_library(lib):
    """Install an already constructed library object (e.g. bound to another device)."""
    _state['lib'] = lib
    _state['scratch'] = None
    _state['out'] = None
    # plans and index records cached on interned layouts live in the memory of the previous library's device
    import sys
    lay = sys.modules.get('tenpy_b200.linalg._layout')
    if lay is not None:
        lay._INTERN.clear()
    npc = sys.modules.get('tenpy_b200.linalg.np_conserved')
    if npc is not None:
        npc._PLAN_CACHE.clear()
        npc._EMPTY_LAYOUTS.clear()
        npc._SMALL_DEV_CACHE.clear()
    return lib


def device():
    return get_lib().device


def empty(n):
    return torch.empty(int(n), dtype=torch.float64, device=get_lib().device)


def zeros(n):
    return torch.zeros(int(n), dtype=torch.float64, device=get_lib().device)


PIN_MAX_BYTES = 1 << 20
_BLOCKING_H2D = bool(os.environ.get('B200_BLOCKING_H2D'))     # A/B switch: the blocking uploads of rounds 1 / 2


def to_device(a, pin=False):
    """numpy array (float64 / int64 / int32) -> device tensor, WITHOUT synchronising the stream.

    A blocking ``tensor.to(device)`` from pageable memory ends in a stream synchronisation (ATen's copy kernel), and the
    engine uploads ~10 small arrays per bond update (index records, Schmidt values for `scale_axis`, ...): each one made the
    host wait for everything the GPU had queued, i.e. the host could never run ahead of the device across such a call.
    Small arrays go through torch's caching pinned-memory allocator (the copy is asynchronous, the pinned block is recycled
    after the copy has run); large ones (MPS tensors at set-up) are copied from pageable memory without the final
    synchronisation (the driver stages the source before `cudaMemcpyAsync` returns).  Stream order makes every later kernel
    see the data."""
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a)
    dev = get_lib().device
    if dev.type == 'cpu':
        return t.clone()
    if t.numel() == 0:
        return torch.empty(t.shape, dtype=t.dtype, device=dev)
    if _BLOCKING_H2D:
        return t.to(dev)
    try:
        src = t.pin_memory() if (pin or t.numel() * t.element_size() <= PIN_MAX_BYTES) else t
        return src.to(dev, non_blocking=True)
    except RuntimeError:          # no pinned memory to be had: the blocking copy is always available
        return t.to(dev)


def to_host(t):
    """device tensor -> numpy array (synchronises)."""
    return t.detach().cpu().numpy()


def dot_scratch():
    if _state['scratch'] is None:
        _state['scratch'] = empty(DOT_SCRATCH)
    return _state['scratch']


def scalar_out():
    if _state['out'] is None:
        _state['out'] = empty(8)
    return _state['out']


def read_scalar(t):
    """read element 0 of a device tensor (synchronises the stream)."""
    return float(t[:1].cpu().numpy()[0])
