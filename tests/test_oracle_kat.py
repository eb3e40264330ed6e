"""CPU: pin the oracle against every known-answer vector available for the path.

Pinned by the reference's own tests: BVH build/traverse/slab test (21-unit-box KAT, doc-test).
Pinned by published vectors: xoshiro256++ and SplitMix64.
Everything else on the path has no reference test ("parity unpinned"): covered by self-consistency checks.
"""
import math

import numpy as np
import pytest

from conftest import frame_compare


def test_splitmix64_published_vectors(O):
    # Vigna's splitmix64.c, seed 1234567 (also Rosetta Code "Pseudo-random numbers/Splitmix64")
    assert O.splitmix64_stream(1234567, 5) == [
        6457827717110365317, 3203168211198807973, 9817491932198370423, 4593380528125082431, 16408922859458223821]
    assert O.splitmix64_stream(0, 1)[0] == 0xE220A8397B1DCDAF


def test_xoshiro256plusplus_published_vector(O):
    # rand_xoshiro's reference test for Xoshiro256PlusPlus: state [1,2,3,4]
    out, _ = O.xoshiro_next_u64([1, 2, 3, 4], 10)
    assert out == [41943041, 58720359, 3588806011781223, 3591011842654386, 9228616714210784205,
                   9973669472204895162, 14011001112246962877, 12406186145184390807, 15849039046786891736,
                   10450023813501588000]


def test_seed_from_u64_is_four_splitmix_outputs(O):
    for seed in (0, 1, 42, 2**64 - 1):
        assert O.seed_from_u64(seed) == O.splitmix64_stream(seed, 4)


def test_uniform_float_grid(O):
    v = O.rng_floats(7, 0, 4096)
    assert v.min() >= 0.0 and v.max() < 1.0
    assert np.all(v * 2.0**23 == np.round(v * 2.0**23))          # 2^-23 grid
    g = O.rng_floats(7, 1, 4096)
    assert np.array_equal(v, g)                                   # gen_range(0..1) = value*1+0
    u = O.rng_floats(7, 2, 4096)
    assert np.array_equal(u, (v * np.float32(2.0) + np.float32(-1.0)).astype(np.float32))


def test_unit_disc_and_sphere(O):
    d = O.unit_disc(3, 5000)
    assert np.all(d[:, 0] ** 2 + d[:, 1] ** 2 <= 1.0 + 1e-6)
    s = O.unit_sphere(3, 5000)
    assert np.allclose(np.linalg.norm(s.astype(np.float64), axis=1), 1.0, atol=2e-6)
    assert abs(s.mean()) < 0.03                                   # surface of the unit sphere, centred


def test_21_unit_box_kat(O):
    # bvh/src/testbase.rs:92-99,127-166 — exact hit-id sets
    ids = list(range(-10, 11))
    boxes = [[x - 0.5, -0.5, -0.5, x + 0.5, 0.5, 0.5] for x in ids]
    hit, n_nodes, shape_node = O.bvh_traverse_boxes(boxes, (-1000, 0, 0), (1, 0, 0))
    assert sorted(ids[i] for i in hit) == ids
    assert n_nodes == 2 * 21 - 1
    hit, _, _ = O.bvh_traverse_boxes(boxes, (0, -1000, 0), (0, 1, 0))
    assert [ids[i] for i in hit] == [0]
    hit, _, _ = O.bvh_traverse_boxes(boxes, (6, 0.5, 0), (-2, -1, 0))
    assert sorted(ids[i] for i in hit) == [4, 5, 6]
    # bvh_impl.rs:742-768: every shape is a leaf exactly once
    assert len(set(shape_node.tolist())) == 21


def test_slab_doc_test_and_properties(O):
    # ray.rs:160-168
    assert O.ray_intersects_aabb((0, 0, 0), (1, 0, 0), (99.9, -1, -1), (100.1, 1, 1))
    # ray.rs:377-450: a ray aimed at the centre of a box hits it; aimed away from it misses (origin outside)
    rng = np.random.default_rng(0)
    for _ in range(300):
        lo = rng.uniform(-100, 100, 3)
        hi = lo + rng.uniform(0.1, 50, 3)
        o = rng.uniform(-300, 300, 3)
        c = (lo + hi) / 2
        if np.all((o > lo) & (o < hi)):
            continue
        assert O.ray_intersects_aabb(o, c - o, lo, hi)
        assert not O.ray_intersects_aabb(o, o - c, lo, hi)


def test_find_roots_quadratic(O):
    assert O.find_roots_quadratic(1, -3, 2) == [1.0, 2.0]
    assert O.find_roots_quadratic(1, 2, 1) == [-1.0]
    assert O.find_roots_quadratic(1, 0, 1) == []
    r = O.find_roots_quadratic(1, -1e4, 1)                        # stable pair: small root keeps precision
    assert r[0] == pytest.approx(1e-4, rel=1e-6) and r[1] == pytest.approx(1e4, rel=1e-6)
    rng = np.random.default_rng(1)
    for _ in range(200):
        x1, x2 = sorted(rng.uniform(-50, 50, 2))
        r = O.find_roots_quadratic(1.0, -(x1 + x2), x1 * x2)
        assert len(r) == 2 and r[0] <= r[1]
        assert r[0] == pytest.approx(x1, abs=2e-3) and r[1] == pytest.approx(x2, abs=2e-3)


def test_as_u8_cast(O):
    assert O.f32_as_u8(255.999) == 255 and O.f32_as_u8(300.0) == 255 and O.f32_as_u8(-3.0) == 0
    assert O.f32_as_u8(float("nan")) == 0 and O.f32_as_u8(0.999) == 0 and O.f32_as_u8(1.0) == 1
    assert O.f32_as_u8(float("inf")) == 255


def _far_sphere(scenes):
    sp = np.zeros(1, dtype=scenes.SPHERE_DTYPE)
    sp[0] = ((0, 0, 50), 1.0, (0.5, 0.5, 0.5), 0.0, 0.0)          # behind the camera: every ray misses
    return sp


def test_sky_only_frame_closed_form(O, rt):
    # miss → t = dir.y*0.5+1 in [0.5,1.5]; r=g=0.3+0.7t, b=0.8+0.2t; t>1 saturates the u8 cast (main.rs:135-144)
    img, st = O.render_frame(_far_sphere(rt.scenes), None, 64, 48, 4, 3, want_stats=True)
    assert st["rays"] == 64 * 48 * 4 and st["sky"] == st["rays"] and st["shades_sphere"] == 0
    assert np.all(img[:20] == 255)                                # upper rows: dir.y > 0 → all channels > 1
    assert np.all(img[..., 0] == img[..., 1]) and np.all(img[..., 2] >= img[..., 0])
    bottom = img[-1, :, 0].astype(int)
    assert bottom.min() >= int(255.999 * math.sqrt(0.3 + 0.7 * 0.5)) - 1


def test_single_emissive_sphere_silhouette(O, rt):
    sc = rt.scenes
    sp = np.zeros(1, dtype=sc.SPHERE_DTYPE)
    sp[0] = ((0, 0, -5), 1.0, (1.0, 0.0, 0.0), 0.0, 1.0)          # emission*albedo = pure red
    h = w = 200
    img, _ = O.render_frame(sp, None, w, h, 1, 3, aperture=1e-6)
    red = (img[..., 0] == 255) & (img[..., 1] == 0) & (img[..., 2] == 0)
    # tangent cone: tan = r/sqrt(d^2-r^2); image plane half-height 1 ↔ (h-1)/2 pixels
    r_px = (1.0 / math.sqrt(24.0)) * (h - 1) / 2.0
    assert red.sum() == pytest.approx(math.pi * r_px**2, rel=0.03)
    ys, xs = np.nonzero(red)
    assert abs(xs.mean() - (w - 1) / 2) < 1.0 and abs(ys.mean() - (h - 1) / 2) < 1.0


def test_depth_convention(O, rt):
    # max_bounces = D ⇒ at most D+1 queries per sample (main.rs:76,109-111); inside a closed mirror box of
    # spheres nothing escapes, so use counts: rays <= primary*(D+1) and depth_exhausted counted
    sp, tr = rt.scenes.synthetic_spheres(64), rt.scenes.ground_plane()
    for D in (1, 3):
        _, st = O.render_frame(sp, tr, 48, 32, 2, D, want_stats=True)
        assert st["primary"] == 48 * 32 * 2
        assert st["primary"] <= st["rays"] <= st["primary"] * (D + 1)
        assert st["sky"] + st["emissive"] + st["depth_exhausted"] == st["primary"]


def test_reference_bvh_equals_brute_force(O, rt):
    # the BVH is a conservative cull: nearest hit must equal brute force
    sp, tr = rt.scenes.synthetic_spheres(200, 11), rt.scenes.ground_plane()
    a, sa = O.render_frame(sp, tr, 96, 54, 3, 5, mode=0, want_stats=True)
    b, sb = O.render_frame(sp, tr, 96, 54, 3, 5, mode=1, want_stats=True)
    assert np.array_equal(a, b) and sa["rays"] == sb["rays"]
    assert sa["sphere_tests"] < sb["sphere_tests"]


def test_nearest_hit_random_rays_bvh_vs_brute(O, rt):
    sp, tr = rt.scenes.synthetic_spheres(64, 5), rt.scenes.ground_plane()
    rng = np.random.default_rng(2)
    hits = 0
    for _ in range(300):
        o = rng.uniform(-3, 3, 3)
        d = rng.normal(size=3)
        a = O.nearest_hit(sp, tr, o, d, mode=0)
        b = O.nearest_hit(sp, tr, o, d, mode=1)
        assert (a is None) == (b is None)
        if a is not None:
            hits += 1
            assert np.array_equal(a, b)
    assert hits > 30


def test_division_bands_tile_the_frame(O, rt):
    # band d covers rows [d*h/div, (d+1)*h/div), band 0 = top (main.rs:66-68); per-pixel streams make the
    # image independent of the partition
    sp = rt.scenes.synthetic_spheres(32, 9)
    full, _ = O.render_frame(sp, None, 40, 30, 2, 4)
    bands = []
    for d in range(5):
        p = O.make_params(40, 30, 5, d, 2, 4)
        img, _ = O.render_rows(sp, None, p)
        bands.append(img)
    assert np.array_equal(np.concatenate(bands, axis=0), full)


def test_thread_count_does_not_change_the_image(O, rt):
    sp = rt.scenes.synthetic_spheres(32, 9)
    a, _ = O.render_frame(sp, None, 40, 30, 2, 4, threads=1)
    b, _ = O.render_frame(sp, None, 40, 30, 2, 4, threads=5)
    assert np.array_equal(a, b)


def test_defaults_are_reference_literals(O, rt):
    sp = rt.scenes.synthetic_spheres(8, 1)
    a, st = O.render_frame(sp, None, 16, 12, 0, 0, want_stats=True)        # spp 100, max_bounces 10
    assert st["primary"] == 16 * 12 * 100
    b, _ = O.render_frame(sp, None, 16, 12, 100, 10, aperture=0.1, focus_distance=1.0,
                          field_of_view=float(np.float32(np.pi) / np.float32(2)), focal_length=1.0)
    assert np.array_equal(a, b)


def test_golden_frames(O, rt):
    import golden.make_golden as G

    for name, case in G.CASES.items():
        got = G.render_case(O, rt.scenes, case)
        want = G.load(name)
        assert np.array_equal(got, want), (name, frame_compare(got, want))


def test_oracle_rejects_what_the_reference_cannot_render(O, rt):
    sp = rt.scenes.synthetic_spheres(4)
    with pytest.raises(RuntimeError):
        O.render_rows(None, None, O.make_params(8, 8, 1, 0, 1, 1))           # empty world: build never returns
    with pytest.raises(RuntimeError):
        O.render_rows(sp, None, O.make_params(8, 9, 2, 0, 1, 1))             # height % divisions != 0
