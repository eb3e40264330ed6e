"""CPU: the C-ABI library loads and exports what include/rt_b200.h declares; host-side logic
(BVH build, scenes, wire types, tile partition, gloo world_size-2 assembly).  No compute calls."""
import json
import os
import re
import socket

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_exports_every_declared_symbol(rt):
    from rt_b200 import _abi

    L = _abi.lib()
    header = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_abi.EXPORTS), declared ^ set(_abi.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.rt_abi_version() == 2


def test_struct_layouts_match_header(rt):
    import ctypes as C

    from rt_b200 import _abi, scenes

    assert scenes.SPHERE_DTYPE.itemsize == 36 and scenes.TRIANGLE_DTYPE.itemsize == 56
    sizes = (C.c_size_t * 4)()
    _abi.lib().rt_struct_sizes(sizes)
    assert list(sizes) == [36, 56, C.sizeof(_abi.RtParams), C.sizeof(_abi.RtStats)]


def test_no_device_fails_loudly(rt):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(rt.RtError) as e:
        rt.Context(0)
    assert e.value.status == -3 and "no CPU fallback" in e.value.message


def test_host_bvh_matches_reference_leaf_order(O, rt):
    from rt_b200 import api, scenes

    rng = np.random.default_rng(5)
    for n, plane in [(1, False), (2, False), (3, True), (64, False), (257, True), (1024, True)]:
        sp = scenes.synthetic_spheres(n, n)
        tr = scenes.ground_plane() if plane else None
        total = n + (2 if plane else 0)
        for wi in (None, rng.permutation(total).astype(np.uint32)):
            rank, n_nodes, depth = api.bvh_build_host(sp, tr, wi)
            want, want_depth = O.leaf_order(sp, tr, wi)
            assert np.array_equal(rank, want)
            assert n_nodes == 2 * total - 1 and depth == want_depth
            assert sorted(rank.tolist()) == list(range(total))
    # coincident centroids → the "split the list in half" branch (bvh_impl.rs:277-290)
    sp = scenes.synthetic_spheres(37, 5)
    sp["center"] = (0, 0, -5)
    assert np.array_equal(api.bvh_build_host(sp, None)[0], O.leaf_order(sp, None)[0])


def test_host_bvh_two_shape_shortcut_and_tiny_extents(O, rt):
    """The host builder answers two-shape nodes in closed form (half of all nodes); it must agree with the bucket
    machinery of bvh_impl.rs:229-364 (as restated by the oracle) when centroids tie, nearly tie (extent around the
    crate's EPSILON = 1e-5) or sit on a coarse lattice."""
    from rt_b200 import api, scenes

    rng = np.random.default_rng(11)
    for trial in range(120):
        n = int(rng.integers(2, 40))
        sp = scenes.synthetic_spheres(n, trial)
        kind = trial % 4
        if kind == 0:    # coarse lattice: many equal coordinates
            sp["center"] = np.round(sp["center"] * 0.5) * 2.0
        elif kind == 1:  # clusters with extents around EPSILON
            base = rng.uniform(-5, 5, size=(1, 3)) + np.array([0, 0, -10.0])
            sp["center"] = (base + rng.uniform(-1, 1, size=(n, 3)) * 10.0 ** rng.uniform(-7, -3)).astype(np.float32)
        elif kind == 2:  # pairs of identical spheres
            sp["center"][1::2] = sp["center"][0::2][: len(sp["center"][1::2])]
        sp["radius"] = np.maximum(sp["radius"], 1e-3).astype(np.float32)
        got = api.bvh_build_host(sp, None)
        want = O.leaf_order(sp, None)
        assert np.array_equal(got[0], want[0]), (trial, kind, n)
        assert got[2] == want[1]


def test_host_bvh_threaded_build_matches(O, rt):
    """Above 2048 shapes per child the host builders hand subtrees to their own threads (node slots are known in
    advance: a subtree of n shapes owns n - 1 consecutive pre-order slots); the result must not depend on that."""
    from rt_b200 import api, scenes

    sp = scenes.synthetic_spheres(20000, 3)
    tr = scenes.ground_plane()
    got = api.bvh_build_host(sp, tr)
    want = O.leaf_order(sp, tr)
    assert np.array_equal(got[0], want[0])
    assert got[1] == 2 * 20002 - 1 and got[2] == want[1]


def test_host_bvh_errors(rt):
    from rt_b200 import api, scenes

    with pytest.raises(rt.RtError) as e:
        api.bvh_build_host(None, None)
    assert e.value.status == -2                                   # empty scene
    sp = scenes.synthetic_spheres(4)
    with pytest.raises(rt.RtError) as e:
        api.bvh_build_host(sp, None, np.array([0, 1, 1, 2], dtype=np.uint32))
    assert e.value.status == -1                                   # not a permutation
    sp["center"][1] = (np.nan, 0, 0)
    with pytest.raises(rt.RtError) as e:
        api.bvh_build_host(sp, None)
    assert e.value.status == -5


def test_synthetic_scene_generator_is_pinned(rt):
    from rt_b200 import scenes

    g = scenes.SplitMix64(0)
    assert g.next() == 0xE220A8397B1DCDAF                          # same stream as testbase.rs:321-327
    a, b = scenes.synthetic_spheres(256), scenes.synthetic_spheres(256)
    assert a.tobytes() == b.tobytes()
    assert np.all(a["center"][:, 2] <= -4) and np.all(a["center"][:, 2] >= -20)
    assert np.all((a["radius"] >= 0.175 - 1e-6) & (a["radius"] <= 0.35 + 1e-6))
    assert 0 < (a["emission"] > 0).sum() < 40 and set(np.unique(a["roughness"] == 1.0)) == {False, True}
    # first sphere of the default stream, pinned
    s0 = scenes.synthetic_spheres(1)[0]
    g = scenes.SplitMix64(0)
    u = [g.u() for _ in range(4)]
    assert s0["center"][0] == np.float32(-8 + 16 * u[0]) and s0["radius"] == np.float32(0.35 * 256 ** (1 / 3) * (0.5 + 0.5 * u[3]))
    assert scenes.synthetic_spheres(8, 1).tobytes() != scenes.synthetic_spheres(8, 2).tobytes()
    for name, c in scenes.CONFIGS.items():
        sp, tr = scenes.config_scene(name)
        assert len(sp) == c["n_spheres"] and len(tr) == (2 if c["plane"] else 0)


def test_wire_roundtrip_and_shapes(rt):
    from rt_b200 import scenes, wire

    sp, tr = scenes.synthetic_spheres(5, 3), scenes.ground_plane()
    wi = np.array([3, 0, 6, 1, 4, 2, 5], dtype=np.uint32)          # interleave spheres and triangles
    meta = wire.RenderMeta(1080, 1920, 20, "123e4567-e89b-12d3-a456-426614174000")
    text = wire.render_info_to_json(sp, tr, meta, 7, wi)
    o = json.loads(text)
    assert list(o.keys()) == ["world", "render_meta", "division_no"]
    assert list(o["render_meta"].keys()) == ["height", "width", "divisions", "id"]
    assert [list(x.keys())[0] for x in o["world"]] == ["Sphere", "Sphere", "Triangle", "Sphere", "Sphere", "Triangle", "Sphere"]
    assert list(o["world"][0]["Sphere"].keys()) == ["radius", "center", "node_index", "p_albedo_at", "p_roughness_at", "p_emission_at"]
    assert list(o["world"][2]["Triangle"].keys()) == ["a", "b", "c", "node_index", "p_albedo_at", "p_roughness_at", "p_emission_at"]
    info = wire.parse_render_info(text)
    assert info.division_no == 7 and info.render_meta == meta
    assert info.world.spheres.tobytes() != b"" and len(info.world) == 7
    # spheres come back in world order; map back through world_index and compare bit-for-bit
    back = {int(p): info.world.spheres[i] for i, p in enumerate(info.world.world_index[:5])}
    for i in range(5):
        assert back[int(wi[i])].tobytes() == sp[i].tobytes()
    backt = {int(p): info.world.triangles[i] for i, p in enumerate(info.world.world_index[5:])}
    for j in range(2):
        assert backt[int(wi[5 + j])].tobytes() == tr[j].tobytes()


def test_image_slice_json(rt):
    from rt_b200 import wire

    sl = wire.ImageSlice(3, np.array([0, 1, 255, 17], dtype=np.uint8), "123e4567-e89b-12d3-a456-426614174000")
    text = sl.to_json()
    assert text == '{"division_no":3,"image":[0,1,255,17],"id":"123e4567-e89b-12d3-a456-426614174000"}'
    back = wire.ImageSlice.from_json(text)
    assert back.division_no == 3 and back.image.tolist() == [0, 1, 255, 17] and back.id == sl.id


@pytest.mark.parametrize("body", [
    b"not json", b"[]", b'{"world":[],"division_no":0}',
    b'{"world":[{"Cube":{}}],"render_meta":{"height":1,"width":1,"divisions":1,"id":"123e4567-e89b-12d3-a456-426614174000"},"division_no":0}',
    b'{"world":[{"Sphere":{"radius":1}}],"render_meta":{"height":1,"width":1,"divisions":1,"id":"123e4567-e89b-12d3-a456-426614174000"},"division_no":0}',
    b'{"world":[],"render_meta":{"height":-1,"width":1,"divisions":1,"id":"123e4567-e89b-12d3-a456-426614174000"},"division_no":0}',
    b'{"world":[],"render_meta":{"height":1,"width":1,"divisions":1,"id":"nope"},"division_no":0}',
])
def test_wire_rejects_malformed_bodies(rt, body):
    from rt_b200 import wire

    with pytest.raises(wire.WireError):
        wire.parse_render_info(body)


def test_tile_partition_is_exact_cover(rt):
    from rt_b200 import multi

    for (w, h) in [(64, 32), (101, 67), (1920, 1080)]:
        tx, ty = multi.tile_grid(w, h)
        for ranks in (1, 2, 3, 8):
            owners = np.array([multi.tile_owner(t, ranks) for t in range(min(tx * ty, 5000))])
            assert owners.min() >= 0 and owners.max() < ranks
            # kernel's ticket map: rank r, ticket k → tile k*ranks + (r+k) % ranks
            for r in range(ranks):
                mine = multi.tiles_of_rank(w, h, r, ranks)
                k = np.arange(len(mine) + 2)
                g = k * ranks + (r + k) % ranks
                assert np.array_equal(np.sort(g[g < tx * ty]), mine)
            om = multi.owner_map(w, h, ranks)
            assert om.shape == (h, w)
            counts = np.bincount(om.reshape(-1), minlength=ranks)
            assert counts.sum() == w * h and counts.min() > 0.8 * counts.max() - 64


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, w, h, q):
    import torch
    import torch.distributed as dist

    from rt_b200 import multi

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    om = multi.owner_map(w, h, world)
    # stand-in for the render kernel: each rank fills ITS tiles of a zeroed frame with a pixel-dependent value
    yy, xx = np.mgrid[0:h, 0:w]
    truth = ((yy * 31 + xx * 7) % 251 + 1).astype(np.uint8)
    local = np.where(om == rank, truth, 0).astype(np.uint8)
    frame = torch.from_numpy(np.repeat(local[..., None], 3, axis=2).copy().reshape(-1))
    multi.assemble_reduce(frame, 0)
    if rank == 0:
        q.put(bool(np.array_equal(frame.numpy().reshape(h, w, 3)[..., 0], truth)))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_frame_assembly(rt):
    """N>1 path on CPU: disjoint per-rank tiles + reduce(MAX) on rank 0 reproduce the whole frame."""
    import torch.multiprocessing as mp

    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_gloo_worker, args=(r, 2, port, 101, 67, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


# ---------------------------------------------------------------------------------------------------
# "next" rows of SURVEY §8f: OBJ ingest (f4) and the controller's job state machine (f2), host side only
# ---------------------------------------------------------------------------------------------------
def _icosphere(subdiv=1, radius=1.0, centre=(0.0, 0.0, -4.0)):
    t = (1 + 5 ** 0.5) / 2
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7),
         (9, 8, 1)]
    v = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache, nf = {}, []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[k] = len(v) - 1
            return cache[k]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return [tuple(np.array(centre) + radius * p) for p in v], f


def mesh_scene_obj():
    """An 80-triangle icosphere over a two-triangle floor, as OBJ ++ MTL bytes (two materials, two models)."""
    from rt_b200 import obj

    vs, fs = _icosphere(1, 1.2, (0.3, 0.1, -4.0))
    body = obj.write_obj(vs, fs, "ball")
    n = len(vs)
    floor = "usemtl floor\n" + "".join(f"v {x} -1.5 {z}\n" for x, z in [(-30, -40), (-30, 10), (30, 10), (30, -40)])
    floor += f"f {n + 1} {n + 2} {n + 3}\nf {n + 1} {n + 3} {n + 4}\n"
    o = body + floor.encode()
    m = b"newmtl ball\nKd 0.9 0.3 0.2\nNs 600\nnewmtl floor\nKd 0.5 0.5 0.55\nNs 50\n"
    return o + m, len(o)


def test_obj_ingest_matches_reference_rules(rt):
    from rt_b200 import obj

    data, n = mesh_scene_obj()
    tr = obj.build_world(data, n)
    assert len(tr) == 82
    assert np.allclose(tr["albedo"][0], (0.9, 0.3, 0.2)) and tr["roughness"][0] == np.float32(600) / np.float32(1000)
    assert np.allclose(tr["albedo"][-1], (0.5, 0.5, 0.55)) and tr["roughness"][-1] == np.float32(0.05)
    assert np.all(tr["emission"] == 0)
    # face index forms: v, v/vt/vn, negative (relative)
    o = b"usemtl m\nv 0 0 -3\nv 1 0 -3\nv 0 1 -3\nv 1 1 -3\nf 1 2 3\nf 2/1/1 4/2/2 -2//3\n"
    m = b"newmtl m\nKd 0.8 0.2 0.1\nNs 250\n"
    t = obj.build_world(o + m, len(o))
    assert t["a"].tolist() == [[0, 0, -3], [1, 0, -3]] and t["c"][1].tolist() == [0, 1, -3]
    # a quad is NOT triangulated: the reference cuts the flat index list into triples (obj.rs:23-26)
    q = b"usemtl m\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 2 2 2\nv 3 3 3\nf 1 2 3 4\nf 5 6 1\n"
    t = obj.build_world(q + m, len(q))
    assert len(t) == 2 and t["a"][1].tolist() == [0, 1, 0] and t["b"][1].tolist() == [2, 2, 2]
    with pytest.raises(obj.ObjError):
        obj.build_world(b"v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n", 36)       # no material: reference unwraps None
    with pytest.raises(obj.ObjError):
        obj.build_world(o + m, len(o) + len(m) + 5)


class _FakeWorker:
    """Stands in for rt_b200.slave.Worker on a CPU box: a slice whose bytes encode (division, position)."""

    def render(self, info):
        from rt_b200 import wire

        m = info.render_meta
        rows = m.height // m.divisions
        img = np.full((rows, m.width, 3), info.division_no * 10, dtype=np.uint8)
        return wire.ImageSlice(info.division_no, img.reshape(-1), m.id)


def test_controller_job_state_machine(rt):
    from PIL import Image
    import io

    from rt_b200 import controller, wire

    data, n = mesh_scene_obj()
    c = controller.Controller(worker=_FakeWorker(), width=64, height=40, divisions=4)
    assert c.poll("nope") == (b"Invalid Uuid", False)
    assert c.poll("123e4567-e89b-12d3-a456-426614174000") == (b"No such job", False)
    job = c.upload(data, n)
    jpeg, is_img = c.poll(job)
    assert is_img and jpeg[:2] == b"\xff\xd8"
    img = np.array(Image.open(io.BytesIO(jpeg)))
    assert img.shape == (40, 64, 3) and abs(int(img[35, 5, 0]) - 30) <= 2 and abs(int(img[2, 5, 0]) - 0) <= 2
    assert c.poll(job) == (b"No such job", False)                    # jobs are removed after the first poll
    # partial results: the reference's progress text
    c2 = controller.Controller(worker=None, width=64, height=40, divisions=4)
    meta = wire.RenderMeta(40, 64, 4, "123e4567-e89b-12d3-a456-426614174000")
    c2.jobs[meta.id] = {"meta": meta, "result": {}}
    c2.result(wire.ImageSlice(2, np.zeros(10 * 64 * 3, np.uint8), meta.id))
    assert c2.result(wire.ImageSlice(0, np.zeros(10 * 64 * 3, np.uint8), "00000000-0000-4000-8000-000000000009")) == controller.SAVED_TEXT
    assert c2.poll(meta.id) == (b"Job not finished yet 1/4", False)
