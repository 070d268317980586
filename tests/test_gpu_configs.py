"""BASELINE.json configs[2] (Plummer sphere: Domain::sync + halo discovery) and configs[4] (32-bit Morton / float
neighbour-search stress) through tools/run_configs.py at reduced sizes; the full-size runs are recorded in
profiles/r1_configs_2_and_4.jsonl.  The checks are size independent: sorted keys, particle conservation, bucket limit,
no halo flag inside the own range, neighbour lists equal to brute force in the reference's float expression."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_plummer_sync_and_halo_discovery():
    import run_configs

    out = run_configs.config_plummer(1 << 21, 64)
    assert all(out["checks"].values()), out
    assert out["max_leaf_level"] >= 9, "the Plummer tree must be deeper than a uniform one of the same size"


def test_morton_float_neighbor_stress():
    import run_configs

    out = run_configs.config_morton(1 << 20, 300, 384, 64, 32)
    assert all(out["checks"].values()), out
    assert 250 < out["mean_neighbors"] < 330
