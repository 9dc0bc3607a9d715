"""Host <-> device plumbing around a training step: staged input copies and lagged result reads.

A CRD step on B200 is ~0.5 ms of device work.  Driven the way the reference's KD loop is written
(``KD/common/base_class.py:346-405``: copy the batch, forward, backward, ``loss.item()``) the device idles while the host
copies the next batch and while it waits for the loss it just asked for.  ``StepPipeline`` removes both gaps without
changing what a step computes:

* ``stage(*host_tensors)`` enqueues the host->device copies of the NEXT step's inputs on a copy stream (double-buffered
  device slots, pinned host tensors), so they overlap the current step's kernels; ``take()`` hands the oldest staged
  batch to the compute stream (a device-side wait, no host synchronisation);
* ``publish(loss)`` enqueues a device->host copy of a step's scalar result into pinned memory behind the step's
  kernels, and ``collect()`` returns the OLDEST published value, waiting only for that step -- so the host can read the
  loss of step i after it has launched step i+1 (every step's loss is still read, one step late).

``GraphedStep`` goes one step further for loops whose shapes never change: the WHOLE step -- forward and backward, every
launch of it -- is captured once in a CUDA graph on static input tensors and replayed per batch, so the host's share of a
step shrinks to a few copies and one ``cudaGraphLaunch`` (the CRD step is ~0.1 ms of device work per rank when its banks
are sharded over 8 GPUs: the per-step Python / autograd / ctypes path is several times that).

Both are plumbing (streams, events, pinned buffers, graph capture) and know nothing about CRD.
"""
from __future__ import annotations

from collections import deque

import torch


class StepPipeline:
    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("StepPipeline needs a CUDA device: this package has no CPU fallback")
        self.depth = int(depth)
        self.copy_stream = torch.cuda.Stream(self.device)
        self._slots = [None] * (self.depth + 1)   # per slot: (signature, [device buffers], free_event or None)
        self._next = 0
        self._staged = deque()                    # (slot index, ready_event)
        self._in_use = None                       # slot consumed by the step being enqueued
        self._results = deque()                   # (pinned 1-element tensor, event)
        self._pinned_pool = {}
        self._d2h_stream = None

    # -- inputs ------------------------------------------------------------------------------------------------
    def stage(self, *host_tensors: torch.Tensor) -> None:
        """Enqueue the H2D copies of one batch (pinned host tensors) on the copy stream; returns at once."""
        if len(self._staged) >= self.depth:
            raise RuntimeError("StepPipeline.stage: take() the staged batches first (depth exceeded)")
        k = self._next
        self._next = (self._next + 1) % len(self._slots)
        sig = tuple((tuple(t.shape), t.dtype) for t in host_tensors)
        slot = self._slots[k]
        if slot is None or slot[0] != sig:
            bufs = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host_tensors]
            free = slot[2] if slot is not None else None
            slot = self._slots[k] = [sig, bufs, free]
        with torch.cuda.stream(self.copy_stream):
            if slot[2] is not None:
                self.copy_stream.wait_event(slot[2])   # the step that last read this slot has been enqueued and must finish
            for b, t in zip(slot[1], host_tensors):
                b.copy_(t, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        self._staged.append((k, ready))

    def take(self) -> list[torch.Tensor]:
        """Device tensors of the oldest staged batch; the current stream waits for their copies on the device."""
        if not self._staged:
            raise RuntimeError("StepPipeline.take: nothing staged")
        cur = torch.cuda.current_stream(self.device)
        if self._in_use is not None:   # everything enqueued so far has read the previous slot: it is free after this point
            ev = torch.cuda.Event()
            ev.record(cur)
            self._slots[self._in_use][2] = ev
        k, ready = self._staged.popleft()
        cur.wait_event(ready)
        self._in_use = k
        return [b.detach() for b in self._slots[k][1]]

    # -- results -----------------------------------------------------------------------------------------------
    def publish(self, value: torch.Tensor) -> None:
        """Enqueue a D2H copy of a step's result (a scalar loss, or a whole feature tensor) into pinned memory behind the
        kernels that produce it.  Scalars ride on the compute stream; larger tensors are copied on a second stream so
        that the next step's kernels do not wait for the PCIe transfer."""
        value = value.detach()
        key = (tuple(value.shape), value.dtype)
        pool = self._pinned_pool.setdefault(key, [])
        pin = pool.pop() if pool else torch.empty(value.shape, dtype=value.dtype).pin_memory()
        cur = torch.cuda.current_stream(self.device)
        if value.numel() <= 1024:
            pin.copy_(value, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cur)
        else:
            if self._d2h_stream is None:
                self._d2h_stream = torch.cuda.Stream(self.device)
            done = torch.cuda.Event()
            done.record(cur)
            with torch.cuda.stream(self._d2h_stream):
                self._d2h_stream.wait_event(done)
                pin.copy_(value, non_blocking=True)
                value.record_stream(self._d2h_stream)
                ev = torch.cuda.Event()
                ev.record(self._d2h_stream)
        self._results.append((pin, ev, key))

    def pending(self) -> int:
        return len(self._results)

    def collect(self):
        """The oldest published result (a float for one-element results, else the pinned host tensor, valid until the
        next publish of that shape); blocks only until THAT step has finished."""
        pin, ev, key = self._results.popleft()
        ev.synchronize()
        self._pinned_pool[key].append(pin)
        return float(pin.reshape(-1)[0]) if pin.numel() == 1 else pin


class GraphedStep:
    """One training step -- ``fn(*inputs)`` must run forward AND backward and return the loss tensor -- captured in a CUDA
    graph on static device inputs and driven from pinned host batches::

        crit.contrast.device_sampler_offset()            # fresh negatives on every replay (CRDLoss(f_s, f_t, idx) only)
        def fwd_bwd(f_s, f_t, idx):
            loss = crit(f_s, f_t, idx); loss.backward(); return loss
        step = GraphedStep(fwd_bwd, (f_s_host, f_t_host, idx_host), device, grad_inputs=(0,),
                           zero_grad=lambda: crit.zero_grad(set_to_none=True))
        step.stage(*first_batch)
        for batch in loader:
            step.run()                      # device: staged batch -> static inputs, replay; loss copied to pinned memory
            step.stage(*next_batch)         # H2D of the next batch overlaps this step's kernels
            optimizer.step()                # parameter .grad tensors are static: every replay overwrites them
            loss = step.collect()           # the oldest unread loss (one step late when called after the next run())

    Rules of graph capture apply: shapes and dtypes are fixed; ``fn`` must not synchronise, read device values on the host
    or allocate outside the captured allocator; do NOT call ``zero_grad(set_to_none=True)`` after construction (the graph
    writes the gradient tensors that existed at capture; a replay overwrites, it does not accumulate).  ``grad_inputs``
    names the inputs that need ``.grad`` (read it from ``step.static[i].grad`` after ``run()``); ``zero_grad`` is a callable
    that drops the parameters' gradients (``lambda: model.zero_grad(set_to_none=True)``), called before every warm-up pass
    and before capture so that the captured backward allocates the gradient tensors itself.  Anything that must happen
    eagerly before each replay (e.g. filling an extra static tensor) goes into ``before_replay(*static_inputs)``."""

    def __init__(self, fn, example_inputs, device, grad_inputs=(), before_replay=None, zero_grad=None, warmup: int = 3,
                 depth: int = 2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GraphedStep needs a CUDA device: this package has no CPU fallback")
        self.fn, self.before_replay = fn, before_replay
        self.pipe = StepPipeline(self.device, depth)
        self.static = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in example_inputs]
        for s_, t in zip(self.static, example_inputs):
            s_.copy_(t)
        for i in grad_inputs:
            self.static[i].requires_grad_()
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):       # warm-up outside capture: workspaces, one-time attributes, lazy allocations
            for _ in range(max(warmup, 1)):
                self._drop_grads(zero_grad, grad_inputs)
                if before_replay is not None:
                    before_replay(*self.static)
                fn(*self.static)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        self._drop_grads(zero_grad, grad_inputs)   # the captured backward allocates the gradient tensors it will overwrite
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.loss = fn(*self.static)
        self.replays = 0

    def _drop_grads(self, zero_grad, grad_inputs):
        if zero_grad is not None:
            zero_grad()
        for i in grad_inputs:
            self.static[i].grad = None

    def stage(self, *host_tensors) -> None:
        """Enqueue the host->device copies of the NEXT batch on the copy stream."""
        self.pipe.stage(*host_tensors)

    def run(self) -> None:
        """Staged batch -> static inputs (device-side copies), eager ``before_replay``, graph replay, loss -> pinned memory."""
        for s_, t in zip(self.static, self.pipe.take()):
            s_.detach().copy_(t, non_blocking=True)
        if self.before_replay is not None:
            self.before_replay(*self.static)
        self.graph.replay()
        self.replays += 1
        self.pipe.publish(self.loss)

    def pending(self) -> int:
        return self.pipe.pending()

    def collect(self) -> float:
        return self.pipe.collect()
