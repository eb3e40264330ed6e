"""Dev tool (GPU box): BASELINE config 5 — primitive-count sweep 64 → 65,536 spheres at 1080p, 1 spp, depth 5.
Prints kernel time for both intersectors (brute only up to 4096) and parity of a crop against the oracle."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ray-tracer-s8_b200"))
import numpy as np
import rt_b200 as rt
from rt_b200 import scenes
from oracle import oracle as O

ctx = rt.Context(0)
W, H = 1920, 1080
for n in [int(x) for x in os.environ.get("C5_NS", "64,128,256,512,1024,2048,4096,8192,16384,32768,65536").split(",")]:
    sp = scenes.synthetic_spheres(n)
    t0 = time.time(); sc = ctx.scene(sp, None); t_scene = (time.time() - t0) * 1e3
    row = f"N={n:6d} scene_create_ms={t_scene:7.2f}"
    for isect in (1, 2):
        if isect == 1 and n > 4096:
            continue
        p = rt.make_params(W, H, spp=1, max_bounces=5, intersector=isect)
        ms = []
        for _ in range(3):
            img, st = ctx.render_frame(sc, p, want_stats=True)
            ms.append(st["kernel_ms"])
        row += f" | isect={isect} ms={min(ms):8.3f} Mrays/s={st['rays']/min(ms)/1e3:8.0f} smem={st['scene_in_smem']}"
    # parity on one of the controller's 20 bands
    p = rt.make_params(W, H, divisions=20, division_no=10, spp=1, max_bounces=5, intersector=2)
    band = ctx.render_division(sc, p)
    ref, _ = O.render_rows(sp, None, O.make_params(W, H, 20, 10, 1, 5))
    row += f" | band10 ndiff={(band != ref).sum()}"
    print(row, flush=True)
    sc.close()
