"""Wall-clock accounting under named labels (same interface as the reference's
`auxilliary/logging.py:11-60`: ``PerformanceLog(label)`` as context manager or decorator and
``log_summary()``), written for the engine so that the five reference buckets -- timestep,
pressure_solve, tentative_velocity_solve, bdm_projection, unsplit_solve -- keep their names."""

from __future__ import annotations

import functools
import time
from collections import defaultdict

import numpy as np

__all__ = ["PerformanceLog", "log_summary"]


class PerformanceLog:
    #: label -> list of elapsed seconds, shared by all instances
    records: dict = defaultdict(list)
    #: optional callable run before every clock read (e.g. torch.cuda.synchronize) so that the
    #: asynchronous GPU work is attributed to the right bucket
    sync = None

    def __init__(self, label: str):
        self.label = label
        self._t0 = None

    def __enter__(self):
        if PerformanceLog.sync is not None:
            PerformanceLog.sync()
        self._t0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        if PerformanceLog.sync is not None:
            PerformanceLog.sync()
        PerformanceLog.records[self.label].append(time.perf_counter() - self._t0)
        return False

    def __call__(self, fn):
        @functools.wraps(fn)
        def wrapped(*args, **kwargs):
            with PerformanceLog(self.label):
                return fn(*args, **kwargs)

        return wrapped

    @classmethod
    def reset(cls):
        cls.records.clear()


def log_summary(file=None):
    """the table the reference prints (`logging.py:34-60`): one right-aligned row per label with calls, total,
    mean and standard deviation in seconds, largest total first; nothing at all when no timer ran.  The text is
    byte-identical to the reference's for the same timings (`tests/test_golden.py`, golden made by running the
    reference's own module).  Returns the rows."""
    rows = []
    for label, ts in PerformanceLog.records.items():
        a = np.asarray(ts)
        rows.append((label, int(a.size), float(a.sum()), float(a.mean()), float(a.std())))
    if not rows:
        return rows
    rows.sort(key=lambda r: -r[2])
    cols = ("timer", "ncall", "total", "avg", "std")
    print("%32s : %6s    %10s %10s %10s" % cols, file=file)
    print("-" * 77, file=file)
    for label, n, tot, avg, std in rows:
        print("%32s : %6d    %10.4e %10.4e %10.4e" % (label, n, tot, avg, std), file=file)
    return rows
