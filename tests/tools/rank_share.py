"""Dev tool (GPU box): time one rank's share of the C3 frame (tile_ranks = N emulated on one GPU)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ray-tracer-s8_b200"))
import numpy as np, rt_b200 as rt
from rt_b200 import scenes
cfg = scenes.CONFIGS["C3"]; sp, tr = scenes.config_scene("C3")
ctx = rt.Context(0); sc = ctx.scene(sp, tr)
p = rt.make_params(cfg["width"], cfg["height"], spp=cfg["spp"], max_bounces=cfg["max_bounces"])
dev, _ = ctx.frame_alloc(cfg["width"] * cfg["height"] * 3)
for ranks in (1, 2, 4, 8):
    best = {}
    for r in range(min(ranks, 2)):
        ms = [ctx.render_tiles_device(sc, p, r, ranks, dev, sync=True, want_stats=True)["kernel_ms"] for _ in range(4)]
        best[r] = min(ms)
    print(os.environ.get("RT_B200_TILE_ORDER", "bottomup"), "ranks", ranks, {k: round(v, 3) for k, v in best.items()}, "ideal", round(51.6 / ranks, 3))
