"""Small host utilities (reference: sdf-net/lib/utils.py)."""
import time

import torch


def setparam(args, param, paramstr):
    """kwargs-over-argparse bridge, same resolution order as the reference (utils.py:33-38):
    an explicit keyword wins; otherwise fall back to `args.<paramstr>`; otherwise None."""
    from_args = getattr(args, paramstr, None)
    return param if (param is not None or from_args is None) else from_args


class PerfTimer:
    """Checkpoint timer (reference: utils.py:95-131).  CUDA events are created lazily and only
    when active on a CUDA box, so constructing one never needs a driver (the reference's
    constructor does, which is why it cannot be imported on a CPU-only host)."""

    def __init__(self, activate=False):
        self.activate = activate
        self.counter = 0
        self._cpu = time.process_time()
        self._ev = None
        if activate and torch.cuda.is_available():
            self._ev = torch.cuda.Event(enable_timing=True)
            self._ev.record()

    def reset(self):
        self.__init__(self.activate)

    def check(self, name=None):
        if not self.activate:
            return None
        cpu = time.process_time() - self._cpu
        gpu = float("nan")
        if self._ev is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            torch.cuda.synchronize()
            gpu = self._ev.elapsed_time(end) / 1e3
            self._ev = end
        label = name if name else self.counter
        print(f"CPU Checkpoint {label}: {cpu:.3e} s")
        print(f"GPU Checkpoint {label}: {gpu:.3e} s")
        self._cpu = time.process_time()
        self.counter += 1
        return cpu, gpu
