"""Golden frames for the render path.

The reference (Rust) cannot be built or run in this environment and has no fixtures for this path
(SURVEY.md §8c), so these vectors are produced by the ORACLE (oracle/rt_oracle.cpp) and committed as
regression pins: `python tests/golden/make_golden.py` regenerates them.  Both the oracle (CPU tests) and the
CUDA path (GPU tests) are compared against the committed files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: scene + params (all seeds explicit)
    "spheres64_96x72_spp2_mb10": dict(n=64, scene_seed=0, plane=False, w=96, h=72, spp=2, mb=10, seed=0),
    "spheres256_plane_128x72_spp3_mb5": dict(n=256, scene_seed=0, plane=True, w=128, h=72, spp=3, mb=5, seed=7),
    "plane_only_64x48_spp2_mb3": dict(n=0, scene_seed=0, plane=True, w=64, h=48, spp=2, mb=3, seed=1),
    "odd_61x37_spp1_mb2": dict(n=20, scene_seed=4, plane=True, w=61, h=37, spp=1, mb=2, seed=123456789),
}


def scene_of(scenes, case):
    sp = scenes.synthetic_spheres(case["n"], case["scene_seed"]) if case["n"] else None
    tr = scenes.ground_plane() if case["plane"] else None
    return sp, tr


def render_case(O, scenes, case):
    sp, tr = scene_of(scenes, case)
    img, _ = O.render_frame(sp, tr, case["w"], case["h"], case["spp"], case["mb"], seed=case["seed"])
    return img


def load(name):
    return np.load(os.path.join(HERE, name + ".npy"))


if __name__ == "__main__":
    root = os.path.abspath(os.path.join(HERE, "..", ".."))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "ray-tracer-s8_b200"))
    from oracle import oracle as O
    from rt_b200 import scenes

    for name, case in CASES.items():
        img = render_case(O, scenes, case)
        np.save(os.path.join(HERE, name + ".npy"), img)
        print(name, img.shape, int(img.sum()))
