"""Several batches in flight on one GPU.

A Monte-Carlo campaign (``test/disturbRejComp.py:74-100``, ``test/success_rates_test.py:64-75`` scaled up) is a stream of
independent batches.  One batch ends with a tail: the whole-loop team kernel is a persistent grid whose CTAs leave as the
lane queue runs dry, and a few lanes need 10-100x the average work (runs of 4000-iteration solves that no input feature
predicts), so the last tenth of a launch keeps a handful of SMs busy.  ``BatchPipeline`` keeps
``depth`` engines (handles, each with its own CUDA stream and device state) of ONE problem family and drives each from its
own host thread: the next batch's CTAs take over the SMs the previous batch's tail has released, and host<->device copies
of one batch overlap the kernels of the other.  Results are exactly those of a single engine (lanes are independent and
each engine is deterministic); only the schedule changes.

ctypes releases the GIL for the duration of a library call, so plain Python threads are enough.
"""
from __future__ import annotations

import threading
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Callable, Optional

from .engine import Engine
from .problem import Problem


class BatchPipeline:
    def __init__(self, problem: Problem, device: int = 0, depth: int = 2, pin_outputs: bool = False, lanes: Optional[int] = None):
        assert depth >= 1
        self.problem, self.device, self.depth = problem, int(device), int(depth)
        self._free = []
        self._lock = threading.Lock()
        self.engines = [Engine(problem, device, pin_outputs=pin_outputs) for _ in range(depth)]
        if lanes:
            for e in self.engines:
                e.batch_alloc(lanes)
        self._free = list(self.engines)
        self._pool = ThreadPoolExecutor(max_workers=depth, thread_name_prefix="mpcb-pipe")

    def _run(self, fn: Callable, args, kw):
        with self._lock:
            eng = self._free.pop()
        try:
            return fn(eng, *args, **kw)
        finally:
            with self._lock:
                self._free.append(eng)

    def submit(self, fn, *args, **kw) -> Future:
        """``fn``: an ``Engine`` method name (``"simulate_discrete"``, ``"simulate_continuous"``) or a callable ``fn(engine, *args)``.
        The call runs on whichever engine is free; at most ``depth`` run at once, the rest queue in submission order."""
        if isinstance(fn, str):
            name = fn
            fn = lambda eng, *a, **k: getattr(eng, name)(*a, **k)      # noqa: E731
        return self._pool.submit(self._run, fn, args, kw)

    def map_discrete(self, batches, nsteps: int, record=()):
        """``batches``: iterable of ``(x0[4, B], noise[R, 2, B] | None)`` -> list of ``BatchSimRun`` in the same order."""
        futs = [self.submit("simulate_discrete", x0, nz, nsteps, record) for x0, nz in batches]
        return [f.result() for f in futs]

    def counters(self) -> dict:
        out = {}
        for e in self.engines:
            for k, v in e.counters().items():
                out[k] = out.get(k, 0) + v
        return out

    def close(self):
        self._pool.shutdown(wait=True)
        for e in self.engines:
            e.close()
        self.engines = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
