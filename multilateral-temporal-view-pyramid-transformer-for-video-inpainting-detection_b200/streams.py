"""Branch-level concurrency for the forward: independent branches of the model (the three temporal views inside a
stage, the frequency branch, the decoder's pyramid inputs) are enqueued on side streams that fork from and re-join
the caller's stream.  Under `torch.cuda.graph` capture the fork/join pattern becomes parallel branches of the graph,
so the ~750 small dependent kernels of one forward no longer form a single serial chain (at batch 1 that chain alone
costs ~6 ms of launch + drain latency on a B200).

Rules that keep this safe with PyTorch's caching allocator (which reuses a freed block on its *allocation* stream
without waiting for other streams):
  * a tensor produced on one lane and consumed on another is handed over with `Region.publish()` / `Region.need()`
    (CUDA event) and kept alive by the region until the join, so its block cannot be recycled while a consumer on a
    different stream is still pending;
  * everything produced inside a region is only used by the caller after the region's join.
Regions nest: only the outermost one forks and joins, inner ones (e.g. a three-view block called from the encoder)
just pick lanes and hand results over with events (`publish_tensor` / `need_tensor` / `wait_lanes` / `need_main`).  Wrapping
`encoder(x)` and `decoder(...)` in ONE outer region (mumpy_b200.forward does) therefore lets the decoder's pyramid and
frequency branches start as soon as their stage features exist and overlap the encoder's tail (stage 3 and the 12 small
global blocks).  No torch computation happens here; streams and events only.
"""
import contextlib
import os

import torch

N_LANES = 4
# CUDA stream priority per lane (lower = scheduled first; stream capture records it on the graph's kernel nodes).  Measured on
# B200 at batch 32: ANY unequal assignment is slower than equal priorities -- view 3's lane (60 % of the encoder's work) high:
# 1995 -> 1873 clips/s; the light lanes high instead: 1987 -> 1856 -- so the default is equal; MUMPY_LANE_PRIO overrides.
LANE_PRIORITY = [int(v) for v in os.environ.get("MUMPY_LANE_PRIO", "0,0,0,0").split(",")]
_lanes = {}            # device index -> [torch.cuda.Stream]
_active = {}           # device index -> Region
enabled = os.environ.get("MUMPY_STREAMS", "1") != "0"      # MUMPY_STREAMS=0: one serial chain (A/B measurements)


def set_enabled(flag: bool):
    """False runs every branch on the caller's stream (single serial chain)."""
    global enabled
    enabled = bool(flag)


class Region:
    def __init__(self, device):
        self.device = device
        self.keep = []
        self.events = {}
        self.parallel = enabled
        if self.parallel:
            idx = device.index if device.index is not None else torch.cuda.current_device()
            if idx not in _lanes:
                _lanes[idx] = [torch.cuda.Stream(device=idx, priority=LANE_PRIORITY[i]) for i in range(N_LANES)]
            self.lanes = _lanes[idx]
            self.main = torch.cuda.current_stream(idx)
            for s in self.lanes:
                s.wait_stream(self.main)

    @contextlib.contextmanager
    def lane(self, i):
        """Work issued inside runs on side stream i (or on the caller's stream when concurrency is off)."""
        if not self.parallel:
            yield
            return
        with torch.cuda.stream(self.lanes[i % N_LANES]):
            yield

    def publish(self, key, *tensors):
        """Called on the producing lane right after `tensors` were enqueued: marks them ready and keeps them alive."""
        self.keep.extend(t for t in tensors if t is not None)
        if self.parallel:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self.events[key] = ev

    def need(self, key):
        """Called on the consuming lane before the first use of what `key` published."""
        if self.parallel:
            torch.cuda.current_stream().wait_event(self.events[key])

    def publish_tensor(self, *tensors):
        """Like publish(), keyed by the tensors' storage: called on the lane that produced them, lets any other lane (or the
        caller's stream) `need_tensor()` them later without knowing who produced them (views share the key)."""
        for t in tensors:
            if t is not None:
                self.publish(("t", t.data_ptr()), t)

    def need_tensor(self, *tensors):
        """Waits (on the current stream) for every tensor that was `publish_tensor`-ed in this region; others are assumed to be
        ordered already (produced on the current stream, or before the region forked)."""
        if not self.parallel:
            return
        cur = torch.cuda.current_stream()
        for t in tensors:
            ev = self.events.get(("t", t.data_ptr())) if t is not None else None
            if ev is not None:
                cur.wait_event(ev)

    def wait_lanes(self, lanes=None):
        """The current stream waits for everything enqueued so far on the given lanes (default: all) -- a one-way join that
        leaves the lanes running; used where a nested region must hand results back to the caller's stream."""
        if not self.parallel:
            return
        cur = torch.cuda.current_stream()
        for i in (range(N_LANES) if lanes is None else lanes):
            if self.lanes[i % N_LANES] != cur:
                ev = torch.cuda.Event()
                ev.record(self.lanes[i % N_LANES])
                cur.wait_event(ev)

    def need_main(self):
        """The current lane waits for everything enqueued so far on the caller's stream (results produced there after the fork)."""
        if not self.parallel:
            return
        cur = torch.cuda.current_stream()
        if cur != self.main:
            ev = torch.cuda.Event()
            ev.record(self.main)
            cur.wait_event(ev)

    def hold(self, *tensors):
        """Keeps tensors alive until the join (inputs of work that is still pending on a lane)."""
        self.keep.extend(t for t in tensors if t is not None)

    def join(self):
        if self.parallel:
            for s in self.lanes:
                self.main.wait_stream(s)
        self.keep.clear()
        self.events.clear()


@contextlib.contextmanager
def region(device):
    """Outermost use forks the lanes from the current stream and joins them on exit; nested uses share the region."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    cur = _active.get(idx)
    if cur is not None:
        yield cur
        return
    reg = Region(torch.device("cuda", idx))
    _active[idx] = reg
    try:
        yield reg
    finally:
        del _active[idx]
        reg.join()
