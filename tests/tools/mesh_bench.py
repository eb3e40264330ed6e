"""Dev tool (GPU box): a tessellated height field (2*q*q triangles) + some spheres, parity on a band against the oracle
and kernel time at 1080p.  usage: mesh_bench.py [q=200] [spp=4]"""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ray-tracer-s8_b200"))
import numpy as np
import rt_b200 as rt
from rt_b200 import scenes
from oracle import oracle as O

q = int(sys.argv[1]) if len(sys.argv) > 1 else 200
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
xs = np.linspace(-12, 12, q + 1, dtype=np.float32)
zs = np.linspace(-24, -3, q + 1, dtype=np.float32)
X, Z = np.meshgrid(xs, zs, indexing="ij")
Y = (-3.0 + 0.6 * np.sin(X * 0.9) * np.cos(Z * 0.7)).astype(np.float32)
P = np.stack([X, Y, Z], axis=-1)
tr = np.zeros(2 * q * q, dtype=scenes.TRIANGLE_DTYPE)
a, b, c, d = P[:-1, :-1].reshape(-1, 3), P[1:, :-1].reshape(-1, 3), P[1:, 1:].reshape(-1, 3), P[:-1, 1:].reshape(-1, 3)
tr["a"][0::2], tr["b"][0::2], tr["c"][0::2] = a, b, c
tr["a"][1::2], tr["b"][1::2], tr["c"][1::2] = a, c, d
tr["albedo"] = (0.6, 0.55, 0.5)
tr["roughness"] = 0.7
sp = scenes.synthetic_spheres(64, 7)
ctx = rt.Context(0)
t0 = time.perf_counter()
sc = ctx.scene(sp, tr)
t1 = time.perf_counter()
w, h = 1920, 1080
p = rt.make_params(w, h, spp=spp, max_bounces=5, intersector=2)
ms = []
for _ in range(3):
    img, st = ctx.render_frame(sc, p, want_stats=True)
    ms.append(st["kernel_ms"])
pc = rt.make_params(w, h, spp=spp, max_bounces=5, intersector=2, collect_counters=True)
_, cst = ctx.render_frame(sc, pc, want_stats=True)
print(f"triangles={len(tr)} scene_create_ms={(t1 - t0) * 1e3:.1f} kernel_ms={min(ms):.3f} rays={st['rays']} "
      f"Mrays/s={st['rays'] / min(ms) / 1e3:.0f} slab/ray={cst['slab_tests'] / st['rays']:.1f} tri_tests/ray={cst['tri_tests'] / st['rays']:.2f} "
      f"smem={st['scene_in_smem']} threads={st['cta_threads']}")
# parity on one band of a small frame
ref, _ = O.render_rows(sp, tr, O.make_params(320, 180, 6, 3, 2, 5, 0), want_stats=True)
band = ctx.render_division(sc, rt.make_params(320, 180, divisions=6, division_no=3, spp=2, max_bounces=5, intersector=2))
print("band ndiff", int((band != ref).sum()))
