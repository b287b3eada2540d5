"""Wall-clock stage timers with the reference's stage names (`retrieve`, `retrieve_faiss`,
`retrieve_faiss_ts`, `retrieve_bm25`, metric `retrieved_chunks`) — the hooks EnsembleRetriever.invoke
uses (/root/reference/src/utils/ensembleRetriever.py:50,63,135,138,185,188,229,231).  Thread-safe."""
from __future__ import annotations

import functools
import statistics
import threading
import time
from collections import defaultdict


class StageProfiler:
    def __init__(self):
        self._lock = threading.Lock()
        self._open: dict[tuple[int, str], float] = {}
        self.samples: dict[str, list[float]] = defaultdict(list)
        self.metrics: dict[str, list[float]] = defaultdict(list)

    def start(self, name: str) -> None:
        self._open[(threading.get_ident(), name)] = time.perf_counter()

    def end(self, name: str) -> None:
        t0 = self._open.pop((threading.get_ident(), name), None)
        if t0 is not None:
            with self._lock:
                self.samples[name].append(time.perf_counter() - t0)

    def add_metric(self, name: str, value: float) -> None:
        with self._lock:
            self.metrics[name].append(float(value))

    def profile_function(self, name: str | None = None):
        def deco(fn):
            label = name or fn.__name__

            @functools.wraps(fn)
            def wrapper(*a, **kw):
                t0 = time.perf_counter()
                try:
                    return fn(*a, **kw)
                finally:
                    with self._lock:
                        self.samples[label].append(time.perf_counter() - t0)
            return wrapper
        return deco

    def summary(self) -> dict:
        out = {}
        with self._lock:
            for k, v in self.samples.items():
                if v:
                    s = sorted(v)
                    out[k] = {"count": len(v), "min": s[0], "max": s[-1], "mean": statistics.fmean(v),
                              "p50": s[len(s) // 2], "p95": s[min(len(s) - 1, int(0.95 * len(s)))],
                              "p99": s[min(len(s) - 1, int(0.99 * len(s)))]}
        return out


profiler = StageProfiler()
