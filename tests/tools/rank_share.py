"""Dev tool (GPU box): time every rank's share of a frame (tile_ranks = N emulated on one GPU): what each GPU of an
N-GPU run would do, minus NVLink.  usage: rank_share.py [C3] [ranks ...]"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ray-tracer-s8_b200"))
import numpy as np, rt_b200 as rt
from rt_b200 import scenes
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
rank_counts = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]
cfg = scenes.CONFIGS[name]; sp, tr = scenes.config_scene(name)
ctx = rt.Context(0); sc = ctx.scene(sp, tr).wait_ready()
p = rt.make_params(cfg["width"], cfg["height"], spp=cfg["spp"], max_bounces=cfg["max_bounces"])
dev, _ = ctx.frame_alloc(cfg["width"] * cfg["height"] * 3)
one = None
for ranks in rank_counts:
    best, rays = [], []
    for r in range(ranks):
        runs = [ctx.render_tiles_device(sc, p, r, ranks, dev, sync=True, want_stats=True) for _ in range(3)]
        best.append(min(x["kernel_ms"] for x in runs)); rays.append(runs[0]["rays"])
    if ranks == 1:
        one = best[0]
    tag = " ".join(f"{k}={os.environ[k]}" for k in sorted(os.environ) if k.startswith("RT_B200_"))
    print(f"{name} ranks={ranks} kernel_ms max={max(best):.3f} mean={np.mean(best):.3f} min={min(best):.3f} "
          f"ideal={(one or best[0] * ranks) / ranks:.3f} eff={(one or 0) / ranks / max(best):.3f} "
          f"rays max/mean={max(rays) / np.mean(rays):.4f} [{tag}]", flush=True)
