"""helpers shared by the tests: synthetic particle sets and a python OctreeMaker (tree/cs_util.hpp:136-197)"""
import numpy as np

from _libs import KEYS, MAXLEVEL


def node_range(kt, level):
    return 1 << (3 * (MAXLEVEL[kt] - level))


class OctreeMaker:
    """builds cornerstone leaf arrays by repeated subdivision, like the reference's test fixture generator"""

    def __init__(self, kt):
        self.kt = kt
        self.leaves = {(0, 0)}  # (startKey, level)

    def divide(self, *path):
        key = 0
        for lvl, digit in enumerate(path, start=1):
            key += digit * node_range(self.kt, lvl)
        level = len(path)
        assert (key, level) in self.leaves, "node to divide must be a leaf"
        self.leaves.remove((key, level))
        for s in range(8):
            self.leaves.add((key + s * node_range(self.kt, level + 1), level + 1))
        return self

    def make_tree(self):
        ks = sorted(k for k, _ in self.leaves) + [node_range(self.kt, 0)]
        return np.array(ks, dtype=KEYS[self.kt])


def uniform_particles(n, T, seed=42, lo=0.0, hi=1.0):
    rng = np.random.default_rng(seed)
    x = (lo + (hi - lo) * rng.random(n)).astype(T)
    y = (lo + (hi - lo) * rng.random(n)).astype(T)
    z = (lo + (hi - lo) * rng.random(n)).astype(T)
    if T == np.float32:  # float rounding may produce hi exactly; keep inside [lo, hi)
        for a in (x, y, z):
            np.clip(a, lo, np.nextafter(T(hi), T(lo)), out=a)
    return x, y, z


def gaussian_particles(n, T, seed=42, lo=-1.0, hi=1.0):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(3):
        a = rng.normal(0.5 * (lo + hi), 0.15 * (hi - lo), n)
        a = np.clip(a, lo, np.nextafter(hi, lo)).astype(T)
        np.clip(a, T(lo), np.nextafter(T(hi), T(lo)), out=a)
        out.append(a)
    return out


def plummer_particles(n, T, seed=42):
    """Plummer sphere (clustered, deep tree), truncated radius; test/coord_samples/plummer.hpp is the model"""
    rng = np.random.default_rng(seed)
    m = rng.random(n) * 0.999
    r = 1.0 / np.sqrt(m ** (-2.0 / 3.0) - 1.0)
    cz = 2 * rng.random(n) - 1
    phi = 2 * np.pi * rng.random(n)
    s = np.sqrt(1 - cz * cz)
    return (r * s * np.cos(phi)).astype(T), (r * s * np.sin(phi)).astype(T), (r * cz).astype(T)


def const_h(n, ng, T, volume=1.0):
    return np.full(n, 0.5 * np.cbrt(3.0 * ng * volume / (4 * np.pi * n)), dtype=T)
