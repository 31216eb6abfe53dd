"""ctypes binding of the CPU oracle (oracle/libb2icp_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under icpslam_b200/ does.  PARITY UNPINNED — see oracle/b2icp_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libb2icp_oracle.so")


class Params(C.Structure):
    """Mirror of b2icp_params (include/b2icp.h)."""
    _fields_ = [
        ("mode", C.c_int32),
        ("max_iterations", C.c_int32),
        ("transformation_epsilon", C.c_double),
        ("max_correspondence_distance", C.c_double),
        ("euclidean_fitness_epsilon", C.c_double),
        ("rotation_epsilon", C.c_double),
        ("gicp_epsilon", C.c_double),
        ("k_correspondences", C.c_int32),
        ("max_inner_iterations", C.c_int32),
        ("device", C.c_int32),
        ("profile", C.c_int32),
        ("grid_cell", C.c_float),
        ("reserved", C.c_int32 * 5),
    ]


class Result(C.Structure):
    """Mirror of b2icp_result (include/b2icp.h)."""
    _fields_ = [
        ("T", C.c_double * 16),
        ("converged", C.c_int32),
        ("iterations", C.c_int32),
        ("n_corr_last", C.c_int32),
        ("status_detail", C.c_int32),
        ("mse_last", C.c_double),
        ("fitness", C.c_double),
    ]

    def matrix(self) -> np.ndarray:
        return np.array(list(self.T), dtype=np.float64).reshape(4, 4)


class StageMs(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("build", "covariances", "nn", "solve", "transform", "total")]


MODE_P2P_SVD = 0
MODE_GICP_BFGS = 1


def default_params(preset: str = "odometer", mode: int = MODE_P2P_SVD) -> Params:
    """The reference's constant blocks: include/icpslam/icp_odometer.h:62-65 (10 iterations) and
    include/icpslam/octree_mapper.h:53-56 (30 iterations); GICP internals = PCL defaults."""
    p = Params()
    p.mode = mode
    p.max_iterations = 10 if preset == "odometer" else 30
    p.transformation_epsilon = 1e-6
    p.max_correspondence_distance = 1.0
    p.euclidean_fitness_epsilon = -1.7976931348623157e308
    p.rotation_epsilon = 2e-3
    p.gicp_epsilon = 1e-3
    p.k_correspondences = 20
    p.max_inner_iterations = 20
    p.device = 0
    p.profile = 0
    p.grid_cell = 0.0
    return p


def build(force: bool = False) -> str:
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("b2icp_oracle.cpp", "gicp_oracle.cpp", "octree_oracle.cpp", "b2icp_oracle.h")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs if os.path.exists(s))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        L.b2o_set_threads.argtypes = [C.c_int]
        L.b2o_nn_brute.argtypes = [fp, C.c_size_t, fp, C.c_size_t, ip, fp]
        L.b2o_kdtree_build.argtypes = [fp, C.c_size_t]
        L.b2o_kdtree_build.restype = C.c_void_p
        L.b2o_kdtree_free.argtypes = [C.c_void_p]
        L.b2o_kdtree_free.restype = None
        L.b2o_kdtree_nn.argtypes = [C.c_void_p, fp, C.c_size_t, ip, fp]
        L.b2o_kdtree_knn.argtypes = [C.c_void_p, fp, C.c_size_t, C.c_int, ip, fp]
        L.b2o_transform_cloud_d.argtypes = [fp, C.c_size_t, dp, fp]
        L.b2o_transform_cloud_f.argtypes = [fp, C.c_size_t, fp, fp]
        L.b2o_covariances.argtypes = [fp, C.c_size_t, C.c_int, C.c_double, dp]
        L.b2o_umeyama.argtypes = [fp, fp, C.c_size_t, dp]
        L.b2o_svd3.argtypes = [dp, dp, dp, dp]
        L.b2o_svd3_cov.argtypes = [dp, dp, dp, dp]
        L.b2o_align.argtypes = [C.POINTER(Params), fp, C.c_size_t, fp, C.c_size_t, fp, C.POINTER(Result), fp,
                                C.c_int, ip, fp, C.POINTER(StageMs)]
        L.b2o_fitness.argtypes = [fp, C.c_size_t, fp, C.c_size_t, fp, C.c_double, dp]
        L.b2o_octree_create.restype = C.c_void_p
        L.b2o_octree_create.argtypes = [C.c_double]
        L.b2o_octree_free.argtypes = [C.c_void_p]
        L.b2o_octree_free.restype = None
        L.b2o_octree_add_points.argtypes = [C.c_void_p, fp, C.c_size_t]
        L.b2o_octree_add_points.restype = C.c_size_t
        L.b2o_octree_size.argtypes = [C.c_void_p]
        L.b2o_octree_size.restype = C.c_size_t
        L.b2o_octree_points.argtypes = [C.c_void_p, fp]
        L.b2o_octree_points.restype = None
        L.b2o_octree_box.argtypes = [C.c_void_p, dp, C.POINTER(C.c_int)]
        L.b2o_octree_box.restype = None
        L.b2o_octree_approx_nearest.argtypes = [C.c_void_p, fp, C.c_size_t, C.c_int, ip]
        L.b2o_octree_approx_nearest.restype = None
        L.b2o_gicp_trace.argtypes = [dp, C.c_size_t]
        L.b2o_gicp_trace.restype = None
        L.b2o_gicp_trace_count.restype = C.c_size_t
        for f in ("b2o_pose_compose", "b2o_pose_inverse", "b2o_pose_from_matrix"):
            getattr(L, f).restype = None
        L.b2o_pose_compose.argtypes = [dp, dp, dp]
        L.b2o_pose_inverse.argtypes = [dp, dp]
        L.b2o_pose_from_matrix.argtypes = [dp, dp]
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _cloud(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4, "clouds are float32[N,4]"
    return a


def set_threads(n: int) -> int:
    return lib().b2o_set_threads(n)


def max_threads() -> int:
    return lib().b2o_get_max_threads()


def nn_brute(tgt, q):
    tgt, q = _cloud(tgt), _cloud(q)
    idx = np.empty(len(q), np.int32)
    d2 = np.empty(len(q), np.float32)
    rc = lib().b2o_nn_brute(_f(tgt), len(tgt), _f(q), len(q), _i(idx), _f(d2))
    assert rc == 0, rc
    return idx, d2


class KdTree:
    def __init__(self, tgt):
        self.tgt = _cloud(tgt)
        self.h = lib().b2o_kdtree_build(_f(self.tgt), len(self.tgt))

    def __del__(self):
        if getattr(self, "h", None):
            lib().b2o_kdtree_free(self.h)
            self.h = None

    def nn(self, q):
        q = _cloud(q)
        idx = np.empty(len(q), np.int32)
        d2 = np.empty(len(q), np.float32)
        rc = lib().b2o_kdtree_nn(self.h, _f(q), len(q), _i(idx), _f(d2))
        assert rc == 0, rc
        return idx, d2

    def knn(self, q, k):
        q = _cloud(q)
        idx = np.empty((len(q), k), np.int32)
        d2 = np.empty((len(q), k), np.float32)
        rc = lib().b2o_kdtree_knn(self.h, _f(q), len(q), k, _i(idx), _f(d2))
        if rc != 0:
            raise RuntimeError(f"b2o_kdtree_knn rc={rc}")
        return idx, d2


def transform_cloud(cloud, T, double: bool = True):
    cloud = _cloud(cloud)
    out = np.empty_like(cloud)
    if double:
        T = np.ascontiguousarray(T, dtype=np.float64)
        rc = lib().b2o_transform_cloud_d(_f(cloud), len(cloud), _d(T), _f(out))
    else:
        T = np.ascontiguousarray(T, dtype=np.float32)
        rc = lib().b2o_transform_cloud_f(_f(cloud), len(cloud), _f(T), _f(out))
    assert rc == 0, rc
    return out


def covariances(cloud, k=20, gicp_epsilon=1e-3):
    cloud = _cloud(cloud)
    out = np.empty((len(cloud), 3, 3), np.float64)
    rc = lib().b2o_covariances(_f(cloud), len(cloud), k, gicp_epsilon, _d(out))
    if rc != 0:
        raise RuntimeError(f"b2o_covariances rc={rc}")
    return out


def umeyama(src, dst):
    src, dst = _cloud(src), _cloud(dst)
    T = np.empty(16, np.float64)
    rc = lib().b2o_umeyama(_f(src), _f(dst), len(src), _d(T))
    assert rc == 0, rc
    return T.reshape(4, 4)


def svd3(A, cov=False):
    """One-sided Jacobi SVD of the restatement; cov=True: the pair-skip threshold of the GICP covariances (1e-15)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    U = np.empty((3, 3))
    s = np.empty(3)
    V = np.empty((3, 3))
    (lib().b2o_svd3_cov if cov else lib().b2o_svd3)(_d(A), _d(U), _d(s), _d(V))
    return U, s, V


def align(params: Params, src, tgt, guess=None, want_aligned=False, record_iter=None):
    """Returns dict(rc, result, T, aligned, corr_idx, corr_d2, stages)."""
    src, tgt = _cloud(src), _cloud(tgt)
    res = Result()
    st = StageMs()
    aligned = np.empty_like(src) if want_aligned else None
    ci = np.full(len(src), -2, np.int32) if record_iter is not None else None
    cd = np.zeros(len(src), np.float32) if record_iter is not None else None
    g = None
    if guess is not None:
        g = np.ascontiguousarray(guess, dtype=np.float32).reshape(16)
    rc = lib().b2o_align(C.byref(params), _f(src), len(src), _f(tgt), len(tgt), _f(g) if g is not None else None,
                         C.byref(res), _f(aligned) if aligned is not None else None,
                         -2 if record_iter is None else int(record_iter), _i(ci) if ci is not None else None,
                         _f(cd) if cd is not None else None, C.byref(st))
    return dict(rc=rc, result=res, T=res.matrix(), aligned=aligned, corr_idx=ci, corr_d2=cd,
                stages={n: getattr(st, n) for n, _ in StageMs._fields_}, iterations=res.iterations,
                converged=res.converged, n_corr=res.n_corr_last, mse=res.mse_last)


def align_gicp_traced(params: Params, src, tgt, cap: int = 20000):
    """align() in GICP mode with the evaluation trace switched on: returns (align dict, trace[n, 8]) where every row
    is one cost-functor evaluation: x[0..5], f, |gradient| (NaN = not asked for)."""
    buf = np.full((cap, 8), np.nan, np.float64)
    lib().b2o_gicp_trace(_d(buf), cap)
    try:
        out = align(params, src, tgt)
        n = int(lib().b2o_gicp_trace_count())
    finally:
        lib().b2o_gicp_trace(None, 0)
    return out, buf[:n].copy()


def fitness(src, tgt, T, max_range=1.7976931348623157e308):
    src, tgt = _cloud(src), _cloud(tgt)
    T = np.ascontiguousarray(T, dtype=np.float32).reshape(16)
    out = C.c_double()
    rc = lib().b2o_fitness(_f(src), len(src), _f(tgt), len(tgt), _f(T), max_range, C.byref(out))
    assert rc == 0, rc
    return out.value


def voxel_filter(cloud, leaf: float) -> np.ndarray:
    cloud = _cloud(cloud)
    out = np.empty_like(cloud)
    n_out = C.c_size_t()
    L = lib()
    L.b2o_voxel_filter.argtypes = [C.POINTER(C.c_float), C.c_size_t, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_size_t)]
    rc = L.b2o_voxel_filter(_f(cloud), len(cloud), leaf, _f(out), C.byref(n_out))
    if rc != 0:
        raise RuntimeError(f"b2o_voxel_filter rc={rc}")
    return out[: n_out.value].copy()


def map_insert(map_cloud, cloud, resolution: float) -> np.ndarray:
    """Points of `cloud` that OctreeMapper::addPointsToMap appends to a map holding `map_cloud`."""
    cloud = _cloud(cloud)
    m = _cloud(map_cloud) if map_cloud is not None and len(map_cloud) else np.zeros((0, 4), np.float32)
    out = np.empty_like(cloud)
    n_out = C.c_size_t()
    L = lib()
    fp = C.POINTER(C.c_float)
    L.b2o_map_insert.argtypes = [fp, C.c_size_t, fp, C.c_size_t, C.c_double, fp, C.POINTER(C.c_size_t)]
    rc = L.b2o_map_insert(_f(m) if len(m) else None, len(m), _f(cloud), len(cloud), resolution, _f(out), C.byref(n_out))
    if rc != 0:
        raise RuntimeError(f"b2o_map_insert rc={rc}")
    return out[: n_out.value].copy()


class CompatOctree:
    """pcl::octree::OctreePointCloudSearch as OctreeMapper uses it (oracle/octree_oracle.cpp):
    resetMap / addPointsToMap / approxNearestNeighbors with PCL's first-point-anchored, root-growing lattice and its
    greedy centre-distance descent.  key_rule 0 = PCL 1.8.x literal, 1 = the chosen child's key is handed down."""

    def __init__(self, resolution: float):
        self._t = lib().b2o_octree_create(float(resolution))
        if not self._t:
            raise ValueError("resolution must be > 0")

    def __del__(self):
        if getattr(self, "_t", None):
            lib().b2o_octree_free(self._t)
            self._t = None

    def add_points(self, cloud) -> int:
        cloud = _cloud(cloud)
        return int(lib().b2o_octree_add_points(self._t, _f(cloud), len(cloud))) if len(cloud) else 0

    def size(self) -> int:
        return int(lib().b2o_octree_size(self._t))

    def points(self) -> np.ndarray:
        out = np.zeros((self.size(), 4), np.float32)
        if len(out):
            lib().b2o_octree_points(self._t, _f(out))
        return out

    def box(self):
        mn = np.zeros(3)
        depth = C.c_int()
        lib().b2o_octree_box(self._t, _d(mn), C.byref(depth))
        return mn, depth.value

    def approx_nearest(self, q, key_rule: int = 0) -> np.ndarray:
        q = _cloud(q)
        idx = np.full(len(q), -1, np.int32)
        if len(q):
            lib().b2o_octree_approx_nearest(self._t, _f(q), len(q), key_rule, _i(idx))
        return idx


def pose_compose(a, b):
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    o = np.empty(7)
    lib().b2o_pose_compose(_d(a), _d(b), _d(o))
    return o


def pose_inverse(a):
    a = np.ascontiguousarray(a, np.float64)
    o = np.empty(7)
    lib().b2o_pose_inverse(_d(a), _d(o))
    return o


def pose_from_matrix(T):
    T = np.ascontiguousarray(T, np.float64).reshape(16)
    o = np.empty(7)
    lib().b2o_pose_from_matrix(_d(T), _d(o))
    return o
