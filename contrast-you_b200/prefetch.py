"""Double-buffered host -> device staging for the losses' inputs.

The reference feeds its criteria from tensors that are already on the device (the U-Net produced them); a caller that
holds the criteria's inputs in pinned host memory (bench.py's end-to-end leg, an offline evaluation over stored
embeddings / probability maps) would otherwise pay the PCIe copy in front of every step.  ``HostPrefetcher`` overlaps it:
two sets of device buffers, a copy stream, and events in both directions — while step k computes on one set, step k+1's
inputs are copied into the other.  One copy per step, every copy ordered behind the last reader of the set it overwrites.
"""
from typing import Sequence

import torch

__all__ = ["HostPrefetcher"]


class HostPrefetcher:
    def __init__(self, host_tensors: Sequence[torch.Tensor], device):
        assert all(t.is_pinned() for t in host_tensors), "pinned host memory is what makes the copies asynchronous"
        self.host = list(host_tensors)
        self.device = torch.device(device)
        self.sets = [[torch.empty_like(t, device=self.device) for t in self.host] for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]        # copy of the set has landed
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]     # the compute stream is past the set's readers
        self.cur = 0
        self.last = None
        for e in self.consumed:
            e.record(torch.cuda.current_stream(self.device))
        self._issue(0)

    def _issue(self, s: int):
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[s])
            for dst, src in zip(self.sets[s], self.host):
                dst.copy_(src, non_blocking=True)
            self.ready[s].record(self.copy_stream)

    def next(self):
        """device tensors of the current step (valid until the call after next); starts the copy for the following step.
        ``host_tensors`` may be refilled by the caller once this returns and the previous copy's event has been waited on
        — bench.py keeps them fixed."""
        main = torch.cuda.current_stream(self.device)
        s = self.cur
        if self.last is not None:
            self.consumed[self.last].record(main)      # everything enqueued so far read the previous set
        main.wait_event(self.ready[s])
        self.last = s
        self.cur = s ^ 1
        self._issue(self.cur)
        return self.sets[s]
