"""ctypes loader for the CPU oracle (oracle/rt_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librt_oracle.so")

SPHERE_DTYPE = np.dtype(
    [("center", "<f4", 3), ("radius", "<f4"), ("albedo", "<f4", 3), ("roughness", "<f4"), ("emission", "<f4")]
)
TRIANGLE_DTYPE = np.dtype(
    [("a", "<f4", 3), ("b", "<f4", 3), ("c", "<f4", 3), ("albedo", "<f4", 3), ("roughness", "<f4"), ("emission", "<f4")]
)


class OrcParams(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("divisions", C.c_uint32), ("division_no", C.c_uint32),
        ("spp", C.c_uint32), ("max_bounces", C.c_uint32), ("seed", C.c_uint64),
        ("cam_origin", C.c_float * 3),
        ("aperture", C.c_float), ("focus_distance", C.c_float), ("field_of_view", C.c_float),
        ("focal_length", C.c_float),
    ]


class OrcStats(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64), ("primary", C.c_uint64), ("aabb_tests", C.c_uint64),
        ("sphere_tests", C.c_uint64), ("sphere_hits", C.c_uint64), ("tri_tests", C.c_uint64),
        ("tri_exit", C.c_uint64 * 4), ("tri_hits", C.c_uint64),
        ("shades_sphere", C.c_uint64), ("shades_tri", C.c_uint64), ("emissive", C.c_uint64),
        ("sky", C.c_uint64), ("depth_exhausted", C.c_uint64), ("rng_draws", C.c_uint64),
        ("build_ms", C.c_double), ("render_ms", C.c_double),
    ]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (g++ -O2 -ffp-contract=off)."""
    src = os.path.join(HERE, "rt_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        L.orc_render_rows.argtypes = [vp, u32, vp, u32, vp, C.POINTER(OrcParams), u32, u32, vp, C.c_size_t, i32, i32, C.POINTER(OrcStats)]
        L.orc_render_rows.restype = i32
        L.orc_render_division.argtypes = [vp, u32, vp, u32, vp, C.POINTER(OrcParams), vp, C.c_size_t, i32, i32, C.POINTER(OrcStats)]
        L.orc_render_division.restype = i32
        L.orc_splitmix64.argtypes = [C.POINTER(u64)]
        L.orc_splitmix64.restype = u64
        L.orc_seed_from_u64.argtypes = [u64, C.POINTER(u64)]
        L.orc_xoshiro_next_u64.argtypes = [C.POINTER(u64), u32, C.POINTER(u64)]
        L.orc_rng_floats.argtypes = [u64, i32, u32, vp]
        L.orc_unit_disc.argtypes = [u64, u32, vp]
        L.orc_unit_sphere.argtypes = [u64, u32, vp]
        L.orc_find_roots_quadratic.argtypes = [C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float)]
        L.orc_find_roots_quadratic.restype = i32
        L.orc_ray_intersects_aabb.argtypes = [vp, vp, vp, vp]
        L.orc_ray_intersects_aabb.restype = i32
        L.orc_bvh_traverse_boxes.argtypes = [vp, u32, vp, vp, vp, C.POINTER(u32), vp]
        L.orc_bvh_traverse_boxes.restype = i32
        L.orc_bvh_leaf_order.argtypes = [vp, u32, vp, u32, vp, vp, C.POINTER(u32)]
        L.orc_bvh_leaf_order.restype = i32
        L.orc_get_ray.argtypes = [C.POINTER(OrcParams), u32, u32, u64, vp]
        L.orc_nearest_hit.argtypes = [vp, u32, vp, u32, vp, vp, vp, i32, vp]
        L.orc_nearest_hit.restype = i32
        L.orc_trace_pixel.argtypes = [vp, u32, vp, u32, vp, C.POINTER(OrcParams), u32, u32, i32, vp, u32]
        L.orc_trace_pixel.restype = i32
        L.orc_f32_as_u8.argtypes = [C.c_float]
        L.orc_f32_as_u8.restype = C.c_uint8
        L.orc_hardware_threads.restype = i32
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None or a.size == 0 else a.ctypes.data_as(C.c_void_p)


def _f3(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(3))


def make_params(width, height, divisions=1, division_no=0, spp=0, max_bounces=0, seed=0,
                cam_origin=(0.0, 0.0, 0.0), aperture=0.0, focus_distance=0.0, field_of_view=0.0,
                focal_length=0.0) -> OrcParams:
    p = OrcParams()
    p.width, p.height, p.divisions, p.division_no = width, height, divisions, division_no
    p.spp, p.max_bounces, p.seed = spp, max_bounces, seed
    p.cam_origin[:] = [float(c) for c in cam_origin]
    p.aperture, p.focus_distance, p.field_of_view, p.focal_length = aperture, focus_distance, field_of_view, focal_length
    return p


def _scene_args(spheres, triangles, world_index):
    spheres = np.ascontiguousarray(spheres if spheres is not None else np.zeros(0, SPHERE_DTYPE), dtype=SPHERE_DTYPE)
    triangles = np.ascontiguousarray(triangles if triangles is not None else np.zeros(0, TRIANGLE_DTYPE), dtype=TRIANGLE_DTYPE)
    wi = None if world_index is None else np.ascontiguousarray(world_index, dtype=np.uint32)
    return spheres, triangles, wi


def render_rows(spheres, triangles, params: OrcParams, row0=None, row1=None, world_index=None, mode=0,
                threads=0, want_stats=False):
    """Render rows [row0,row1) of the division band → (uint8 array (rows, width, 3), stats|None)."""
    spheres, triangles, wi = _scene_args(spheres, triangles, world_index)
    div = params.divisions or 1
    band_h = params.height // div
    row0 = 0 if row0 is None else row0
    row1 = band_h if row1 is None else row1
    out = np.zeros((row1 - row0, params.width, 3), dtype=np.uint8)
    st = OrcStats() if want_stats else None
    rc = lib().orc_render_rows(_ptr(spheres), len(spheres), _ptr(triangles), len(triangles), _ptr(wi),
                               C.byref(params), row0, row1, _ptr(out) if out.size else out.ctypes.data_as(C.c_void_p),
                               out.size, mode, threads, C.byref(st) if st is not None else None)
    if rc != 0:
        raise RuntimeError(f"orc_render_rows failed: {rc}")
    return out, (st.as_dict() if st is not None else None)


def render_frame(spheres, triangles, width, height, spp, max_bounces, seed=0, world_index=None, mode=0,
                 threads=0, want_stats=False, **cam):
    p = make_params(width, height, 1, 0, spp, max_bounces, seed, **cam)
    return render_rows(spheres, triangles, p, None, None, world_index, mode, threads, want_stats)


def leaf_order(spheres, triangles, world_index=None):
    spheres, triangles, wi = _scene_args(spheres, triangles, world_index)
    n = len(spheres) + len(triangles)
    rank = np.zeros(n, dtype=np.uint32)
    depth = C.c_uint32(0)
    rc = lib().orc_bvh_leaf_order(_ptr(spheres), len(spheres), _ptr(triangles), len(triangles), _ptr(wi),
                                  _ptr(rank), C.byref(depth))
    if rc != 0:
        raise RuntimeError(f"orc_bvh_leaf_order failed: {rc}")
    return rank, depth.value


def nearest_hit(spheres, triangles, origin, direction, world_index=None, mode=0):
    spheres, triangles, wi = _scene_args(spheres, triangles, world_index)
    out = np.zeros(8, dtype=np.float32)
    o, d = _f3(origin), _f3(direction)
    rc = lib().orc_nearest_hit(_ptr(spheres), len(spheres), _ptr(triangles), len(triangles), _ptr(wi),
                               _ptr(o), _ptr(d), mode, _ptr(out))
    if rc < 0:
        raise RuntimeError(f"orc_nearest_hit failed: {rc}")
    return (None if rc == 0 else out)


def trace_pixel(spheres, triangles, params: OrcParams, x, y_global, world_index=None, mode=0, max_rays=4096):
    """Every nearest-hit query of one pixel: array (n, 7) = origin, direction, winner world position or -1."""
    spheres, triangles, wi = _scene_args(spheres, triangles, world_index)
    log = np.zeros((max_rays, 7), dtype=np.float32)
    n = lib().orc_trace_pixel(_ptr(spheres), len(spheres), _ptr(triangles), len(triangles), _ptr(wi), C.byref(params),
                              x, y_global, mode, _ptr(log), max_rays)
    if n < 0:
        raise RuntimeError(f"orc_trace_pixel failed: {n}")
    return log[:min(n, max_rays)].copy()


def bvh_traverse_boxes(boxes, origin, direction):
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 6)
    n = len(boxes)
    out = np.zeros(max(n, 1), dtype=np.uint32)
    shape_node = np.zeros(max(n, 1), dtype=np.uint32)
    nn = C.c_uint32(0)
    o, d = _f3(origin), _f3(direction)
    rc = lib().orc_bvh_traverse_boxes(_ptr(boxes), n, _ptr(o), _ptr(d), _ptr(out), C.byref(nn), _ptr(shape_node))
    if rc < 0:
        raise RuntimeError(f"orc_bvh_traverse_boxes failed: {rc}")
    return out[:rc].copy(), nn.value, shape_node[:n].copy()


def ray_intersects_aabb(origin, direction, bmin, bmax) -> bool:
    o, d, a, b = _f3(origin), _f3(direction), _f3(bmin), _f3(bmax)
    return bool(lib().orc_ray_intersects_aabb(_ptr(o), _ptr(d), _ptr(a), _ptr(b)))


def find_roots_quadratic(a2, a1, a0):
    out = (C.c_float * 2)()
    n = lib().orc_find_roots_quadratic(a2, a1, a0, out)
    return [out[i] for i in range(n)]


def splitmix64_stream(state: int, n: int):
    s = C.c_uint64(state)
    return [lib().orc_splitmix64(C.byref(s)) for _ in range(n)]


def seed_from_u64(seed: int):
    st = (C.c_uint64 * 4)()
    lib().orc_seed_from_u64(seed, st)
    return list(st)


def xoshiro_next_u64(state, n):
    st = (C.c_uint64 * 4)(*state)
    out = (C.c_uint64 * n)()
    lib().orc_xoshiro_next_u64(st, n, out)
    return list(out), list(st)


def rng_floats(seed, kind, n):
    out = np.zeros(n, dtype=np.float32)
    lib().orc_rng_floats(seed, kind, n, _ptr(out))
    return out


def unit_disc(seed, n):
    out = np.zeros((n, 2), dtype=np.float32)
    lib().orc_unit_disc(seed, n, _ptr(out))
    return out


def unit_sphere(seed, n):
    out = np.zeros((n, 3), dtype=np.float32)
    lib().orc_unit_sphere(seed, n, _ptr(out))
    return out


def get_ray(params: OrcParams, x, y_cam, stream_seed):
    out = np.zeros(6, dtype=np.float32)
    lib().orc_get_ray(C.byref(params), x, y_cam, stream_seed, _ptr(out))
    return out[:3].copy(), out[3:].copy()


def f32_as_u8(f) -> int:
    return int(lib().orc_f32_as_u8(C.c_float(f)))


def hardware_threads() -> int:
    return int(lib().orc_hardware_threads())
