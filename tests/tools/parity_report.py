"""Dev tool (GPU box): render cases on the GPU and with the oracle, print mismatch statistics."""
import os, sys, time, json
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ray-tracer-s8_b200"))
import numpy as np
import rt_b200 as rt
from rt_b200 import scenes
from oracle import oracle as O


def compare(a, b):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    pix_bad1 = (d.max(axis=-1) > 1).mean()
    pix_ne = (d.max(axis=-1) > 0).mean()
    mse = float((d.astype(np.float64) ** 2).mean())
    psnr = float("inf") if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
    return dict(pix_gt1=float(pix_bad1), pix_ne=float(pix_ne), psnr=psnr, maxdiff=int(d.max()))


def main():
    ctx = rt.Context(0)
    print(ctx.device_info())
    cases = [
        ("C1-like 320x240 n64 spp2 mb10", 320, 240, 2, 10, scenes.synthetic_spheres(64), None),
        ("n256 480x270 spp2 mb5", 480, 270, 2, 5, scenes.synthetic_spheres(256), None),
        ("n1024+plane 480x270 spp4 mb5", 480, 270, 4, 5, scenes.synthetic_spheres(1024), scenes.ground_plane()),
        ("n1 sphere", 128, 96, 2, 5, scenes.synthetic_spheres(1), None),
        ("plane only", 128, 96, 2, 5, None, scenes.ground_plane()),
        ("odd size 101x67 n32", 101, 67, 3, 4, scenes.synthetic_spheres(32, 7), scenes.ground_plane()),
    ]
    for name, w, h, spp, mb, sp, tr in cases:
        t0 = time.time()
        ref, ost = O.render_frame(sp, tr, w, h, spp, mb, want_stats=True)
        t_or = time.time() - t0
        ref_brute, _ = O.render_frame(sp, tr, w, h, spp, mb, mode=1)
        sc = ctx.scene(sp, tr)
        for isect in (rt.INTERSECT_BRUTE, rt.INTERSECT_BVH):
            p = rt.make_params(w, h, spp=spp, max_bounces=mb, intersector=isect, collect_counters=True)
            img, st = ctx.render_frame(sc, p, want_stats=True)
            c = compare(img, ref)
            cb = compare(img, ref_brute)
            print(f"{name:34s} isect={isect} vs_oracle={c} vs_brute_ne={cb['pix_ne']:.2e} rays gpu={st['rays']} oracle={ost['rays']} "
                  f"kernel_ms={st['kernel_ms']:.3f} oracle_s={t_or:.2f} eff={st['active_lane_iters']/max(1,st['total_lane_iters']):.3f}")
            print("    ", {k: st[k] for k in ('slab_tests','sphere_tests','sphere_exact','tri_tests','hits','shades','emissive','sky')})
        sc.close()
    tf, ms = ctx.measure_fp32_peak()
    print("fp32 peak TFLOP/s", tf, "ms", ms)
    # timing of C2 and a C3 crop
    for name in ("C2", "C3"):
        cfg = scenes.CONFIGS[name]
        sp, tr = scenes.config_scene(name)
        sc = ctx.scene(sp, tr)
        for isect in (rt.INTERSECT_BRUTE, rt.INTERSECT_BVH):
            if name == "C3" and isect == rt.INTERSECT_BRUTE:
                continue
            p = rt.make_params(cfg["width"], cfg["height"], spp=cfg["spp"], max_bounces=cfg["max_bounces"], intersector=isect)
            for _ in range(3):
                img, st = ctx.render_frame(sc, p, want_stats=True)
            print(name, "isect", isect, "kernel_ms", st["kernel_ms"], "total_ms", st["total_ms"], "rays", st["rays"],
                  "Mrays/s", st["rays"] / st["kernel_ms"] / 1e3)
        sc.close()


if __name__ == "__main__":
    main()
