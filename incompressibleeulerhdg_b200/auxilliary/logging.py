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
    """table of label / calls / total / mean / std, largest total first"""
    rows = []
    for label, ts in PerformanceLog.records.items():
        a = np.asarray(ts)
        rows.append((label, a.size, a.sum(), a.mean(), a.std()))
    rows.sort(key=lambda r: -r[2])
    width = max([len(r[0]) for r in rows] + [5])
    print(f"{'label':<{width}}  {'ncall':>6}  {'total[s]':>10}  {'avg[s]':>10}  {'std[s]':>10}", file=file)
    print("-" * (width + 44), file=file)
    for label, n, tot, avg, std in rows:
        print(f"{label:<{width}}  {n:6d}  {tot:10.4f}  {avg:10.4e}  {std:10.4e}", file=file)
    return rows
