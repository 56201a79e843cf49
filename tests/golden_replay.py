"""Replay a golden case's write sequence through any HippocampalFormation-like object."""
from __future__ import annotations

import os

import numpy as np
import torch

import cases as C

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(case_name: str):
    return np.load(os.path.join(GOLDEN_DIR, f"{case_name}.npz"))


class SeedFeeder:
    """Hands out the recorded randperm prefixes in the order the reference drew them."""

    def __init__(self, gold):
        self.seeds = [torch.from_numpy(gold[f"seeds_{j}"]) for j in range(int(gold["n_rebuilds"]))]
        self.at = gold["seeds_at"].tolist()
        self.pos = 0

    def next(self, memory_count: int) -> torch.Tensor:
        assert self.pos < len(self.seeds), "more rebuilds than the reference performed"
        assert self.at[self.pos] == memory_count, (self.at[self.pos], memory_count)
        s = self.seeds[self.pos]
        self.pos += 1
        return s


def replay_writes(case: C.Case, gold, target, set_time, create, rebuild, decay, set_location=None):
    """Drive `target` through the case's writes.

    set_time(t); create(i, row_tensor, next_seeds_fn); rebuild(seeds); decay(rate).
    `create` must call next_seeds_fn(memory_count) iff the write triggers a rebuild.
    """
    rows = C.make_rows(case)
    locs = C.make_locations(case)
    feeder = SeedFeeder(gold)
    for i in range(case.n):
        set_time(C.insert_time(case, i))
        if locs is not None and set_location is not None:
            set_location(torch.from_numpy(locs[i]))
        create(i, torch.from_numpy(rows[i]), feeder.next)
        if case.decay_every and (i + 1) % case.decay_every == 0:
            decay(0.05)
    if case.final_rebuild:
        rebuild(feeder.next(min(case.n, case.max_memories)))
    assert feeder.pos == len(feeder.seeds), "fewer rebuilds than the reference performed"
    set_time(C.query_time(case))
    return rows, C.make_queries(case, rows), locs
