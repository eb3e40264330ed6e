import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ray-tracer-s8_b200"))
import rt_b200 as rt
from rt_b200 import scenes
cfg = scenes.CONFIGS["C3"]
sp, tr = scenes.config_scene("C3")
ctx = rt.Context(0)
sc = ctx.scene(sp, tr)
p = rt.make_params(cfg["width"], cfg["height"], spp=cfg["spp"], max_bounces=cfg["max_bounces"], intersector=2)
print(ctx.trace_bench(sc, p, 240_000_000, False, False))
