"""dev tool (not a test, not part of the product): run bench.py's GPU arm on the numpy TEST DOUBLE of the device
library with stubbed CUDA timing, to catch host-side errors in bench.py on a box without a GPU.  Numbers printed by
this dry run are meaningless.

    python tests/dev_bench_dryrun.py --L 12 --chi 16 --steps 1 --warmup 1
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


class _Event:
    def __init__(self, enable_timing=True):
        self.t = None

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


def main():
    import fake_device
    from tenpy_b200 import _lib

    real = _lib.DeviceLib

    class DryLib(fake_device.FakeDeviceLib):
        profile = None

        def tdot_plan(self, *args):
            return fake_device._FakePlan(real.tdot_plan(self, *args))

        def profile_summary(self):
            return {'gemm': (1, 1.0), 'svd': (1, 0.5)}

        def profile_detail(self):
            return {'gemm': [(1.0, (2.e6, 3, 2))], 'svd': [(0.5, None)]}

        def kernel_launch_count(self, reset=False):
            return 1

    _lib.DeviceLib = DryLib
    torch.cuda.set_device = lambda i: None
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.Event = _Event
    torch.cuda.profiler.start = lambda: None
    torch.cuda.profiler.stop = lambda: None
    torch.Tensor.pin_memory = lambda self: self
    # N > 1 (under torchrun): gloo instead of nccl, CPU tensors
    import torch.distributed as dist
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend=None, **kw: real_init('gloo')
    import bench
    bench.main()


if __name__ == '__main__':
    main()
