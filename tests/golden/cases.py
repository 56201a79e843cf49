"""Seeded synthetic cases shared by the golden generator and the parity tests.

Inputs come from ``numpy.random.default_rng`` (PCG64: bit-stable across
machines and numpy versions), never from torch's generator, so the GPU box can
regenerate exactly the inputs the reference saw in the build container.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

T0 = 1.79e9  # epoch seconds, the magnitude that exposes the fp32-timestamp quirk


@dataclass(frozen=True)
class Case:
    name: str
    n: int                 # memories inserted
    d: int                 # feature dim
    centroids_k: int
    interval: int          # centroids_update_interval
    max_memories: int
    n_queries: int
    k: int
    kind: str              # "gauss" | "clustered" | "two_blobs"
    seed: int
    dt: float = 0.37       # seconds between inserts on the fake clock
    n_clusters: int = 16
    sigma: float = 0.05
    decay_every: int = 0   # call decay_memories(0.05) every this many inserts
    with_location: bool = False
    final_rebuild: bool = True


CASES = [
    # tests/test_hippocampal_index.py:13-51 scenario (d=4, 4 centroids, rebuild every insert)
    Case("idx_d4_n20", n=20, d=4, centroids_k=4, interval=1, max_memories=100, n_queries=4, k=5,
         kind="two_blobs", seed=11),
    # tests/test_hippocampal_formation.py:61-79 scenario (5 rows d=64, exact path, with locations)
    Case("bank_d64_n5", n=5, d=64, centroids_k=256, interval=512, max_memories=1000, n_queries=5, k=3,
         kind="gauss", seed=12, with_location=True, final_rebuild=False),
    # bank overflow: the reference's full-bank slot quirk (row 0 forever), hippocampal.py:200-202
    Case("full_d16_n40", n=40, d=16, centroids_k=4, interval=8, max_memories=32, n_queries=4, k=5,
         kind="clustered", seed=13, n_clusters=4, sigma=0.2),
    # mid-size IVF with decay and several time buckets
    Case("ivf_d64_n3000", n=3000, d=64, centroids_k=32, interval=512, max_memories=4096, n_queries=16, k=10,
         kind="clustered", seed=14, n_clusters=24, sigma=0.15, dt=1.9, decay_every=700),
    # BASELINE config 1: 10k x 768 fp32, 256 centroids, k=10 (iid Gaussian = worst case for IVF)
    Case("c1_d768_n10000", n=10000, d=768, centroids_k=256, interval=512, max_memories=10000, n_queries=32,
         k=10, kind="gauss", seed=1234, dt=0.05),
    # same scale, clustered rows (the meaningful IVF benchmark distribution, SURVEY 8d "K")
    Case("c1k_d768_n6000", n=6000, d=768, centroids_k=64, interval=512, max_memories=8192, n_queries=32,
         k=10, kind="clustered", seed=77, n_clusters=48, sigma=0.05, dt=0.9),
]

CASE_BY_NAME = {c.name: c for c in CASES}


def make_rows(case: Case) -> np.ndarray:
    rng = np.random.default_rng(case.seed)
    if case.kind == "gauss":
        return rng.standard_normal((case.n, case.d), dtype=np.float32)
    if case.kind == "two_blobs":
        # cluster A near e0 first, then cluster B near e1 (test_hippocampal_index.py:28-39)
        x = 0.01 * rng.standard_normal((case.n, case.d), dtype=np.float32)
        half = case.n // 2
        x[:half, 0] += 1.0
        x[half:, 1] += 1.0
        return x
    if case.kind == "clustered":
        centres = rng.standard_normal((case.n_clusters, case.d), dtype=np.float32)
        centres /= np.linalg.norm(centres, axis=1, keepdims=True)
        which = rng.integers(0, case.n_clusters, size=case.n)
        return (centres[which] + case.sigma * rng.standard_normal((case.n, case.d), dtype=np.float32)).astype(np.float32)
    raise ValueError(case.kind)


def make_queries(case: Case, rows: np.ndarray) -> np.ndarray:
    """queries = stored row + 0.1 * N(0,1)."""
    rng = np.random.default_rng(case.seed + 4321)
    pick = rng.integers(0, min(case.n, case.max_memories), size=case.n_queries)
    noise = 0.1 * rng.standard_normal((case.n_queries, case.d), dtype=np.float32)
    scale = np.float32(1.0 if case.kind != "clustered" else case.sigma * 2)
    return (rows[pick] + scale * noise).astype(np.float32)


def make_locations(case: Case) -> Optional[np.ndarray]:
    if not case.with_location:
        return None
    rng = np.random.default_rng(case.seed + 99)
    return (5.0 * rng.standard_normal((case.n, 2))).astype(np.float32)


def insert_time(case: Case, i: int) -> float:
    return T0 + case.dt * i


def query_time(case: Case) -> float:
    return T0 + case.dt * case.n + 50.0
