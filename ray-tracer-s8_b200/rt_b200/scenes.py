"""Deterministic synthetic scenes in the layout of include/rt_b200.h.

The stream is the reference's own `splitmix64`
(ray-tracer-slave/local-dependencies/bvh/src/testbase.rs:321-327), state starting at `scene_seed`;
U() = (next >> 40) * 2^-24.  Arithmetic in float64, stored as float32.  Eleven draws per sphere, in
this order: cx, cy, cz, r, albedo r/g/b, k, roughness-u, e, emission-u.  The generator is input
tooling: the scene is handed to the oracle and the GPU path as data.
"""
from __future__ import annotations

import numpy as np

SPHERE_DTYPE = np.dtype(
    [("center", "<f4", 3), ("radius", "<f4"), ("albedo", "<f4", 3), ("roughness", "<f4"), ("emission", "<f4")]
)
TRIANGLE_DTYPE = np.dtype(
    [("a", "<f4", 3), ("b", "<f4", 3), ("c", "<f4", 3), ("albedo", "<f4", 3), ("roughness", "<f4"), ("emission", "<f4")]
)

_M64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, state: int = 0):
        self.state = state & _M64

    def next(self) -> int:
        self.state = (self.state + 0x9E3779B97F4A7C15) & _M64
        z = self.state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def u(self) -> float:
        return (self.next() >> 40) * (2.0 ** -24)


def synthetic_spheres(n: int, scene_seed: int = 0) -> np.ndarray:
    """`n` spheres in front of the reference camera (origin, looking down -z, fov 90 degrees)."""
    g = SplitMix64(scene_seed)
    out = np.zeros(n, dtype=SPHERE_DTYPE)
    rscale = 0.35 * (256.0 / n) ** (1.0 / 3.0)
    for i in range(n):
        cx = -8.0 + 16.0 * g.u()
        cy = -4.5 + 9.0 * g.u()
        cz = -20.0 + 16.0 * g.u()
        r = rscale * (0.5 + 0.5 * g.u())
        alb = [0.2 + 0.75 * g.u() for _ in range(3)]
        k, ru, e, eu = g.u(), g.u(), g.u(), g.u()
        rough = 0.0 if k < 0.5 else (ru if k < 0.8 else 1.0)   # reference: roughness 1 = mirror
        emis = (1.0 + 3.0 * eu) if e < 0.05 else 0.0
        out[i] = ((cx, cy, cz), r, tuple(alb), rough, emis)
    return out


def ground_plane() -> np.ndarray:
    """"+ plane": two triangles at y = -5 (a plane is two big triangles, t-range is [0.001, 1000))."""
    out = np.zeros(2, dtype=TRIANGLE_DTYPE)
    out[0] = ((-100, -5, -100), (-100, -5, 100), (100, -5, 100), (0.5, 0.5, 0.5), 0.1, 0.0)
    out[1] = ((-100, -5, -100), (100, -5, 100), (100, -5, -100), (0.5, 0.5, 0.5), 0.1, 0.0)
    return out


def no_triangles() -> np.ndarray:
    return np.zeros(0, dtype=TRIANGLE_DTYPE)


def no_spheres() -> np.ndarray:
    return np.zeros(0, dtype=SPHERE_DTYPE)


# Named workloads = BASELINE.json configs as concrete inputs
CONFIGS = {
    "C1": dict(width=640, height=480, spp=1, max_bounces=10, n_spheres=64, plane=False),
    "C2": dict(width=1920, height=1080, spp=1, max_bounces=5, n_spheres=256, plane=False),
    "C3": dict(width=3840, height=2160, spp=16, max_bounces=5, n_spheres=1024, plane=True),
    "C4": dict(width=7680, height=4320, spp=4, max_bounces=8, n_spheres=1024, plane=True),
}


def config_scene(name: str, scene_seed: int = 0):
    c = CONFIGS[name]
    return synthetic_spheres(c["n_spheres"], scene_seed), (ground_plane() if c["plane"] else no_triangles())
