"""Dev tool (GPU box): a few launches of one workload/intersector, for ncu.  usage: prof_one.py C3 2 [reps]"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ray-tracer-s8_b200"))
import rt_b200 as rt
from rt_b200 import scenes

name, isect = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = scenes.CONFIGS[name]
sp, tr = scenes.config_scene(name)
ctx = rt.Context(0)
sc = ctx.scene(sp, tr).wait_ready()
p = rt.make_params(cfg["width"], cfg["height"], spp=cfg["spp"], max_bounces=cfg["max_bounces"], intersector=isect)
for _ in range(reps):
    img, st = ctx.render_frame(sc, p, want_stats=True)
    print(name, isect, st["kernel_ms"], st["rays"])
