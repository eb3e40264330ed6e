"""Dev tool (GPU box): one 1080p, 1 spp, depth-5 frame of an N-sphere scene with a chosen intersector, for ncu.
usage: prof_n.py N isect [reps]"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ray-tracer-s8_b200"))
import rt_b200 as rt
from rt_b200 import scenes

n, isect = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = rt.Context(0)
sc = ctx.scene(scenes.synthetic_spheres(n), None)
p = rt.make_params(1920, 1080, spp=1, max_bounces=5, intersector=isect)
for _ in range(reps):
    img, st = ctx.render_frame(sc, p, want_stats=True)
    print(n, isect, st["kernel_ms"], st["rays"])
