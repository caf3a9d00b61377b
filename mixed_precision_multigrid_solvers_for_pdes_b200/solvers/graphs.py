"""CUDA-graph replay of fixed kernel sequences (one multigrid cycle step).

A step is ~3 launches per level, most of them microseconds long on the coarse levels, plus (multi-GPU) a
grouped NCCL send/recv per pass: launch-bound on the host.  Each (step kind, buffer-role state) is run eagerly
once (warms lazily allocated workspaces and NCCL communicators), captured on its second use and replayed
afterwards.  The fused passes are out of place and swap the roles of `u`/`tmp`; the roles after a step are
recorded at capture time and re-applied after every replay."""
from __future__ import annotations

from typing import Any, Callable, Dict

import gc
import os

import torch

_DEBUG = bool(os.environ.get("MGB200_GRAPH_DEBUG"))


class GraphCache:
    def __init__(self, state_fn: Callable[[], Any], snapshot_fn: Callable[[], Any], restore_fn: Callable[[Any], None],
                 enabled: bool = True):
        self.state_fn, self.snapshot_fn, self.restore_fn = state_fn, snapshot_fn, restore_fn
        self.enabled = enabled
        self.entries: Dict[Any, Any] = {}

    @property
    def captured(self) -> int:
        return sum(1 for v in self.entries.values() if isinstance(v, tuple))

    def run(self, name: str, launch: Callable[[], None]) -> None:
        if not self.enabled:
            launch()
            return
        key = (name, self.state_fn())
        entry = self.entries.get(key)
        if _DEBUG:
            print(f"[graph] {name} key={hash(key) & 0xffff:04x} len={len(key[1])} "
                  f"{'replay' if isinstance(entry, tuple) else ('capture' if entry == 'warm' else 'eager')}", flush=True)
        if entry is None:  # first visit: eager
            launch()
            self.entries[key] = "warm"
            return
        if entry == "warm":  # second visit: capture (capture does not execute), then replay
            before = self.snapshot_fn()
            g = torch.cuda.CUDAGraph()
            # No cyclic garbage collection while capturing: collecting a dead solver destroys ITS CUDA graphs
            # (graph-exec destruction, frees into its private pool), and such calls invalidate a capture in progress
            # ("operation failed due to a previous error during capture").  Switching the collector off for the
            # capture window is enough; a full gc.collect() here cost ~100 ms per capture (10 captures per heat run).
            gc_was_enabled = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(g):
                    launch()
            finally:
                if gc_was_enabled:
                    gc.enable()
            after = self.snapshot_fn()
            self.restore_fn(before)
            entry = self.entries[key] = (g, after)
        g, after = entry
        g.replay()
        self.restore_fn(after)


def prime(step_fn: Callable[[], None], cache: GraphCache, solves_done: Callable[[], int], max_solves: int = 8) -> int:
    """Untimed set-up for benchmarks: run whole solves until two consecutive solves add no new graph state
    (every step then replays).  Returns the number of solves used."""
    if not cache.enabled:
        return 0
    stable, first = 0, solves_done()
    while stable < 2 and solves_done() - first < max_solves:
        before, s0 = (len(cache.entries), cache.captured), solves_done()
        while solves_done() == s0:
            step_fn()
        stable = stable + 1 if (len(cache.entries), cache.captured) == before else 0
    return solves_done() - first
