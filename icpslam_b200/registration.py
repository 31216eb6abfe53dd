"""ctypes mirror of the registration call surface the reference uses, on top of libb2icp.so.

The reference drives a PCL registration object with exactly these calls
(reference src/icpslam/icp_odometer.cpp:188-201, src/icpslam/octree_mapper.cpp:104-117):

    icp.setMaximumIterations / setTransformationEpsilon / setMaxCorrespondenceDistance /
    setRANSACIterations / setInputSource / setInputTarget / align / getFinalTransformation /
    hasConverged / getFitnessScore

``Registration`` keeps those names and meanings (Python test harness and bench use it; the C++
shims under icpslam_b200/csrc/shims keep the IcpOdometer / OctreeMapper classes themselves).
Every method ends in a call of the C ABI declared in include/b2icp.h — there is no other compute
path: if libb2icp.so is missing or no CUDA device is present this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

MODE_P2P_SVD = 0
MODE_GICP_BFGS = 1
PRESET_ODOMETER = 0
PRESET_MAPPER = 1

STATUS = {
    0: "OK", -1: "INVALID_ARG", -2: "EMPTY_CLOUD", -3: "TOO_FEW_POINTS", -4: "NOT_ENOUGH_CORRESPONDENCES",
    -5: "SOLVER_FAILED", -6: "NONFINITE_INPUT", -7: "CUDA", -8: "NO_TARGET", -9: "NO_SOURCE", -10: "NOT_ALIGNED",
}


class B2icpError(RuntimeError):
    def __init__(self, code: int, msg: str = ""):
        self.code = code
        super().__init__(f"b2icp status {code} ({STATUS.get(code, '?')}): {msg}")


class Params(C.Structure):
    """b2icp_params (include/b2icp.h)."""
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
    """b2icp_result (include/b2icp.h)."""
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


class Timing(C.Structure):
    """b2icp_timing (include/b2icp.h)."""
    _fields_ = [
        ("nn_sweep_launches", C.c_int32),
        ("reserved", C.c_int32),
        ("nn_sweep_ms", C.c_double),
        ("build_ms", C.c_double),
        ("total_ms", C.c_double),
        ("kernel_launches", C.c_int64),
        ("nn_searches", C.c_uint64),
    ]


# b2icp_record (include/b2icp.h): the 96-byte per-scan record of the device record sink
RECORD_DTYPE = np.dtype([("T", np.float32, (16,)), ("converged", np.int32), ("iterations", np.int32),
                         ("n_corr_last", np.int32), ("status_detail", np.int32), ("mse_last", np.float64),
                         ("fitness", np.float64)])
assert RECORD_DTYPE.itemsize == 96

EXPORTS = [
    "b2icp_default_params", "b2icp_create", "b2icp_destroy", "b2icp_set_params", "b2icp_set_target",
    "b2icp_set_source", "b2icp_set_target_device", "b2icp_set_source_device", "b2icp_promote_source_to_target",
    "b2icp_align", "b2icp_fitness", "b2icp_get_correspondences", "b2icp_nn_search", "b2icp_nn_search_device",
    "b2icp_transform_cloud", "b2icp_transform_cloud_f", "b2icp_align_batch", "b2icp_align_batch_device",
    "b2icp_set_stream", "b2icp_compute_covariances", "b2icp_voxel_filter", "b2icp_get_timing",
    "b2icp_align_batch_submit", "b2icp_align_batch_submit_device", "b2icp_align_batch_wait",
    "b2icp_set_record_sink", "b2icp_record_sink_count",
    "b2icp_map_reset", "b2icp_map_reset_octree", "b2icp_map_insert", "b2icp_map_insert_device", "b2icp_map_size", "b2icp_map_download",
    "b2icp_map_nearest", "b2icp_set_target_map", "b2icp_mapper_register", "b2icp_mapper_grow",
    "b2icp_pointcloud2_to_xyzw", "b2icp_get_grid_info", "b2icp_host_alloc", "b2icp_host_free", "b2icp_last_error", "b2icp_status_string",
    "b2icp_version",
]

_lib = None


def load_library() -> C.CDLL:
    """dlopen icpslam_b200/lib/libb2icp.so.  Raises if it has not been built — no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    # B2ICP_LIB selects another build of the SAME sources (kernel tuning variants under lib/variants/)
    path = os.environ.get("B2ICP_LIB") or _build.LIB_PATH
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the CUDA extension is the only compute path)")
    L = C.CDLL(path)
    fp, dp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p
    L.b2icp_default_params.argtypes = [C.POINTER(Params), C.c_int]
    L.b2icp_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.b2icp_destroy.argtypes = [vp]
    L.b2icp_set_params.argtypes = [vp, C.POINTER(Params)]
    L.b2icp_set_target.argtypes = [vp, vp, C.c_size_t]
    L.b2icp_set_source.argtypes = [vp, vp, C.c_size_t]
    L.b2icp_set_target_device.argtypes = [vp, vp, C.c_size_t]
    L.b2icp_set_source_device.argtypes = [vp, vp, C.c_size_t]
    L.b2icp_promote_source_to_target.argtypes = [vp]
    L.b2icp_align.argtypes = [vp, fp, C.POINTER(Result), vp]
    L.b2icp_fitness.argtypes = [vp, C.c_double, dp]
    L.b2icp_get_correspondences.argtypes = [vp, ip, fp]
    L.b2icp_nn_search.argtypes = [vp, vp, C.c_size_t, vp, vp]
    L.b2icp_nn_search_device.argtypes = [vp, vp, C.c_size_t, vp, vp]
    L.b2icp_transform_cloud.argtypes = [vp, vp, C.c_size_t, dp, vp]
    L.b2icp_transform_cloud_f.argtypes = [vp, vp, C.c_size_t, fp, vp]
    L.b2icp_align_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(vp), C.POINTER(C.c_size_t),
                                    C.c_size_t, C.c_int, C.POINTER(Result)]
    L.b2icp_align_batch_device.argtypes = L.b2icp_align_batch.argtypes
    L.b2icp_set_stream.argtypes = [vp, vp]
    L.b2icp_compute_covariances.argtypes = [vp, vp, C.c_size_t, dp]
    L.b2icp_voxel_filter.argtypes = [vp, vp, C.c_size_t, C.c_float, vp, C.POINTER(C.c_size_t)]
    szp = C.POINTER(C.c_size_t)
    L.b2icp_align_batch_submit.argtypes = [vp, vp, vp, C.c_size_t, C.c_int]
    L.b2icp_align_batch_submit_device.argtypes = [vp, vp, vp, C.c_size_t, C.c_int]
    L.b2icp_align_batch_wait.argtypes = [vp, vp, C.c_size_t, szp]
    L.b2icp_set_record_sink.argtypes = [vp, vp, C.c_size_t]
    L.b2icp_record_sink_count.argtypes = [vp, szp]
    L.b2icp_map_reset.argtypes = [vp, C.c_double]
    L.b2icp_map_reset_octree.argtypes = [vp, C.c_double]
    L.b2icp_map_insert.argtypes = [vp, vp, C.c_size_t, szp]
    L.b2icp_map_insert_device.argtypes = [vp, vp, C.c_size_t, szp]
    L.b2icp_map_size.argtypes = [vp, szp]
    L.b2icp_map_download.argtypes = [vp, vp, C.c_size_t, szp]
    L.b2icp_map_nearest.argtypes = [vp, vp, C.c_size_t, vp, vp, szp]
    L.b2icp_set_target_map.argtypes = [vp]
    L.b2icp_mapper_register.argtypes = [vp, vp, C.c_size_t, vp, vp, C.POINTER(Result)]
    L.b2icp_mapper_grow.argtypes = [vp, vp, C.c_size_t, vp, szp]
    L.b2icp_pointcloud2_to_xyzw.argtypes = [vp, vp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                            C.c_uint32, C.c_uint32, C.c_int, vp]
    L.b2icp_get_timing.argtypes = [vp, C.POINTER(Timing)]
    L.b2icp_get_grid_info.argtypes = [vp, fp, ip, dp]
    L.b2icp_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.b2icp_host_free.argtypes = [vp]
    L.b2icp_last_error.argtypes = [vp]
    L.b2icp_last_error.restype = C.c_char_p
    L.b2icp_status_string.argtypes = [C.c_int]
    L.b2icp_status_string.restype = C.c_char_p
    _lib = L
    return L


def default_params(preset: int = PRESET_ODOMETER, mode: int = MODE_P2P_SVD, **over) -> Params:
    p = Params()
    rc = load_library().b2icp_default_params(C.byref(p), preset)
    if rc:
        raise B2icpError(rc)
    p.mode = mode
    for k, v in over.items():
        setattr(p, k, v)
    return p


def _cloud(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError("clouds are float32[N,4] arrays {x,y,z,w} (pcl::PointXYZ layout)")
    return a


def _ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """numpy array backed by page-locked memory from b2icp_host_alloc (kept alive by the array)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    rc = load_library().b2icp_host_alloc(n, C.byref(p))
    if rc:
        raise B2icpError(rc, "b2icp_host_alloc")
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.ctypes.data] = p
    return arr


_PINNED: dict[int, C.c_void_p] = {}


class Registration:
    """pcl::IterativeClosestPoint / GeneralizedIterativeClosestPoint look-alike over one b2icp handle."""

    def __init__(self, preset: int = PRESET_ODOMETER, mode: int = MODE_P2P_SVD, device: int = 0, **over):
        self._L = load_library()
        self.params = default_params(preset, mode, device=device, **over)
        self._h = C.c_void_p()
        rc = self._L.b2icp_create(C.byref(self.params), C.byref(self._h))
        if rc:
            raise B2icpError(rc, "b2icp_create (is a CUDA device visible? there is no CPU fallback)")
        self._result = Result()
        self._n_source = 0
        self._fitness = None

    # -- lifetime --------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.b2icp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str = ""):
        if rc:
            msg = self._L.b2icp_last_error(self._h)
            raise B2icpError(rc, f"{what}: {msg.decode() if msg else ''}")

    def _push(self):
        self._check(self._L.b2icp_set_params(self._h, C.byref(self.params)), "set_params")

    # -- the PCL setters the reference calls ------------------------------------------------------
    def setMaximumIterations(self, n: int):
        self.params.max_iterations = int(n)
        self._push()

    def setTransformationEpsilon(self, eps: float):
        self.params.transformation_epsilon = float(eps)
        self._push()

    def setMaxCorrespondenceDistance(self, d: float):
        self.params.max_correspondence_distance = float(d)
        self._push()

    def setEuclideanFitnessEpsilon(self, eps: float):
        self.params.euclidean_fitness_epsilon = float(eps)
        self._push()

    def setRANSACIterations(self, n: int):
        """Accepted and ignored, like PCL with zero rejectors (icp_odometer.cpp:192)."""

    def setProfile(self, on: bool):
        self.params.profile = 1 if on else 0
        self._push()

    def setInputSource(self, cloud):
        c = _cloud(cloud)
        self._check(self._L.b2icp_set_source(self._h, _ptr(c), len(c)), "set_source")
        self._n_source = len(c)

    def setInputTarget(self, cloud):
        c = _cloud(cloud)
        self._check(self._L.b2icp_set_target(self._h, _ptr(c), len(c)), "set_target")

    def setInputSourceDevice(self, dev_ptr: int, n: int):
        self._check(self._L.b2icp_set_source_device(self._h, C.c_void_p(dev_ptr), n), "set_source_device")
        self._n_source = n

    def setInputTargetDevice(self, dev_ptr: int, n: int):
        self._check(self._L.b2icp_set_target_device(self._h, C.c_void_p(dev_ptr), n), "set_target_device")

    def promoteSourceToTarget(self):
        """`*prev_cloud_ = *curr_cloud_` (icp_odometer.cpp:209) without leaving the device."""
        self._check(self._L.b2icp_promote_source_to_target(self._h), "promote_source_to_target")

    # -- align and its getters --------------------------------------------------------------------
    def align(self, guess=None, want_aligned: bool = False, raise_on_fail: bool = True):
        g = None
        if guess is not None:
            g = np.ascontiguousarray(guess, dtype=np.float32).reshape(16)
        out = np.empty((self._n_source, 4), np.float32) if want_aligned else None
        rc = self._L.b2icp_align(self._h, g.ctypes.data_as(C.POINTER(C.c_float)) if g is not None else None,
                                 C.byref(self._result), _ptr(out) if out is not None else None)
        self._fitness = None
        self.last_status = rc
        if rc and raise_on_fail:
            self._check(rc, "align")
        return out

    def getFinalTransformation(self) -> np.ndarray:
        return self._result.matrix()

    def hasConverged(self) -> bool:
        return bool(self._result.converged)

    def getFitnessScore(self, max_range: float = 1.7976931348623157e308) -> float:
        out = C.c_double()
        self._check(self._L.b2icp_fitness(self._h, max_range, C.byref(out)), "fitness")
        return out.value

    @property
    def iterations(self) -> int:
        return self._result.iterations

    @property
    def result(self) -> Result:
        return self._result

    def getCorrespondences(self):
        idx = np.empty(self._n_source, np.int32)
        d2 = np.empty(self._n_source, np.float32)
        self._check(self._L.b2icp_get_correspondences(self._h, idx.ctypes.data_as(C.POINTER(C.c_int32)),
                                                      d2.ctypes.data_as(C.POINTER(C.c_float))), "get_correspondences")
        return idx, d2

    # -- stand-alone kernels ----------------------------------------------------------------------
    def nearestKSearch1(self, queries):
        """KdTreeFLANN::nearestKSearch(k=1) of every query against the current target."""
        q = _cloud(queries)
        idx = np.empty(len(q), np.int32)
        d2 = np.empty(len(q), np.float32)
        self._check(self._L.b2icp_nn_search(self._h, _ptr(q), len(q), _ptr(idx), _ptr(d2)), "nn_search")
        return idx, d2

    def nearestKSearch1Device(self, q_ptr: int, n: int, idx_ptr: int, d2_ptr: int):
        self._check(self._L.b2icp_nn_search_device(self._h, C.c_void_p(q_ptr), n, C.c_void_p(idx_ptr),
                                                   C.c_void_p(d2_ptr)), "nn_search_device")

    def transformPointCloud(self, cloud, T, double: bool = True):
        c = _cloud(cloud)
        out = np.empty_like(c)
        if double:
            T = np.ascontiguousarray(T, np.float64).reshape(16)
            rc = self._L.b2icp_transform_cloud(self._h, _ptr(c), len(c), T.ctypes.data_as(C.POINTER(C.c_double)), _ptr(out))
        else:
            T = np.ascontiguousarray(T, np.float32).reshape(16)
            rc = self._L.b2icp_transform_cloud_f(self._h, _ptr(c), len(c), T.ctypes.data_as(C.POINTER(C.c_float)), _ptr(out))
        self._check(rc, "transform_cloud")
        return out

    def alignBatch(self, sources, targets=None, with_fitness: bool = False):
        """b2icp_align_batch.  targets=None: every source against the current target (resident map);
        targets[i] None (i > 0): sources[i-1] is the target of pair i (consecutive sweeps)."""
        n = len(sources)
        srcs = [_cloud(s) for s in sources]
        sp = (C.c_void_p * n)(*[s.ctypes.data for s in srcs])
        sn = (C.c_size_t * n)(*[len(s) for s in srcs])
        res = (Result * n)()
        if targets is None:
            tp = tn = None
        else:
            tgts = [None if targets[i] is None else _cloud(targets[i]) for i in range(n)]
            tp = (C.c_void_p * n)(*[(t.ctypes.data if t is not None else None) for t in tgts])
            tn = (C.c_size_t * n)(*[(len(t) if t is not None else 0) for t in tgts])
        rc = self._L.b2icp_align_batch(self._h, sp, sn, tp, tn, n, 1 if with_fitness else 0, res)
        self._n_source = len(srcs[0]) if srcs else 0
        return rc, list(res)

    def alignBatchSubmit(self, sources, with_fitness: bool = False) -> int:
        """b2icp_align_batch_submit: enqueue up to 32 host clouds (page-locked ones overlap best) against the
        current target and return at once; at most B2ICP_MAX_IN_FLIGHT (8) batches in flight.  Pair with alignBatchWait()."""
        n = len(sources)
        srcs = [_cloud(s) for s in sources]
        sp = (C.c_void_p * n)(*[s.ctypes.data for s in srcs])
        sn = (C.c_size_t * n)(*[len(s) for s in srcs])
        rc = self._L.b2icp_align_batch_submit(self._h, sp, sn, n, 1 if with_fitness else 0)
        if rc == 0:
            self._inflight = getattr(self, "_inflight", []) + [(srcs, n)]  # keep the host buffers alive
        return rc

    def alignBatchSubmitDevice(self, src_ptrs, n_src, with_fitness: bool = False) -> int:
        """b2icp_align_batch_submit_device: same, the clouds already live in device memory."""
        n = len(src_ptrs)
        sp = (C.c_void_p * n)(*src_ptrs)
        sn = (C.c_size_t * n)(*n_src)
        rc = self._L.b2icp_align_batch_submit_device(self._h, sp, sn, n, 1 if with_fitness else 0)
        if rc == 0:
            self._inflight = getattr(self, "_inflight", []) + [(None, n)]
        return rc

    def alignBatchWait(self):
        """b2icp_align_batch_wait: (rc, results) of the oldest batch in flight."""
        pend = getattr(self, "_inflight", [])
        if not pend:
            raise B2icpError(-1, "alignBatchWait: no batch in flight")
        _, n = pend[0]
        res = (Result * n)()
        got = C.c_size_t()
        rc = self._L.b2icp_align_batch_wait(self._h, res, n, C.byref(got))
        self._inflight = pend[1:]
        return rc, list(res)[: got.value]

    def fromROSMsg(self, data: bytes, width: int, height: int, point_step: int, row_step: int, offsets=(0, 4, 8),
                   is_bigendian: bool = False) -> np.ndarray:
        """pcl::fromROSMsg for PointXYZ (icp_odometer.cpp:168,173): the payload of a sensor_msgs/PointCloud2 ->
        float32[width * height, 4]."""
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        out = np.empty((width * height, 4), np.float32)
        self._check(self._L.b2icp_pointcloud2_to_xyzw(self._h, _ptr(buf) if len(buf) else None, len(buf), width, height, point_step,
                                                      row_step, offsets[0], offsets[1], offsets[2], 1 if is_bigendian else 0,
                                                      _ptr(out) if len(out) else None), "pointcloud2_to_xyzw")
        return out

    def setRecordSink(self, device_ptr, capacity: int) -> None:
        """b2icp_set_record_sink: streamed batches append one 96-byte b2icp_record per scan at `device_ptr`
        (device memory, `capacity` records); None switches the sink off."""
        self._check(self._L.b2icp_set_record_sink(self._h, device_ptr, capacity if device_ptr else 0), "set_record_sink")

    def recordSinkCount(self) -> int:
        n = C.c_size_t()
        self._check(self._L.b2icp_record_sink_count(self._h, C.byref(n)), "record_sink_count")
        return int(n.value)

    def alignBatchDevice(self, src_ptrs, n_src, with_fitness: bool = False):
        """b2icp_align_batch_device against the current target; src_ptrs are device addresses."""
        n = len(src_ptrs)
        sp = (C.c_void_p * n)(*src_ptrs)
        sn = (C.c_size_t * n)(*n_src)
        res = (Result * n)()
        rc = self._L.b2icp_align_batch_device(self._h, sp, sn, None, None, n, 1 if with_fitness else 0, res)
        return rc, list(res)

    def computeCovariances(self, cloud) -> np.ndarray:
        """GICP::computeCovariances of a cloud: float64[N,3,3]."""
        c = _cloud(cloud)
        out = np.empty((len(c), 3, 3), np.float64)
        self._check(self._L.b2icp_compute_covariances(self._h, _ptr(c), len(c), out.ctypes.data_as(C.POINTER(C.c_double))),
                    "compute_covariances")
        return out

    def voxelFilterCloud(self, cloud, leaf: float) -> np.ndarray:
        """IcpOdometer::voxelFilterCloud (pcl::VoxelGrid with a cubic leaf)."""
        c = _cloud(cloud)
        out = np.empty_like(c)
        n_out = C.c_size_t()
        self._check(self._L.b2icp_voxel_filter(self._h, _ptr(c), len(c), leaf, _ptr(out), C.byref(n_out)), "voxel_filter")
        return out[: n_out.value].copy()

    # ---- OctreeMapper's point map (reference src/icpslam/octree_mapper.cpp:56-90), device-resident
    def resetMap(self, resolution: float, pcl_octree: bool = False):
        """OctreeMapper::resetMap.  pcl_octree=True: PCL-compatible mode (b2icp_map_reset_octree): the lattice is
        anchored on the first point, the root box grows as PCL's does, approxNearestNeighbors is PCL's greedy
        octree descent instead of the exact nearest neighbour."""
        fn = self._L.b2icp_map_reset_octree if pcl_octree else self._L.b2icp_map_reset
        self._check(fn(self._h, float(resolution)), "map_reset")

    def addPointsToMap(self, cloud) -> int:
        """OctreeMapper::addPointsToMap: one point per voxel, first come wins; returns the points added."""
        c = _cloud(cloud)
        n_added = C.c_size_t()
        self._check(self._L.b2icp_map_insert(self._h, _ptr(c), len(c), C.byref(n_added)), "map_insert")
        return n_added.value

    def mapSize(self) -> int:
        n = C.c_size_t()
        self._check(self._L.b2icp_map_size(self._h, C.byref(n)), "map_size")
        return n.value

    def mapCloud(self) -> np.ndarray:
        n = self.mapSize()
        out = np.empty((n, 4), np.float32)
        got = C.c_size_t()
        self._check(self._L.b2icp_map_download(self._h, _ptr(out) if n else None, n, C.byref(got)), "map_download")
        return out[: got.value]

    def approxNearestNeighbors(self, cloud):
        """OctreeMapper::approxNearestNeighbors with the exact nearest neighbour: (indices into the map,
        nn_cloud = the map point of every query that has one, in query order)."""
        c = _cloud(cloud)
        idx = np.empty(len(c), np.int32)
        nn = np.empty((len(c), 4), np.float32)
        n_nn = C.c_size_t()
        self._check(self._L.b2icp_map_nearest(self._h, _ptr(c), len(c), _ptr(idx), _ptr(nn), C.byref(n_nn)), "map_nearest")
        return idx, nn[: n_nn.value].copy()

    def mapperRegister(self, cloud, T_raw, T_raw_inv):
        """b2icp_mapper_register: the registration half of OctreeMapper::refineTransformAndGrowMap, device-resident."""
        c = _cloud(cloud)
        a = np.ascontiguousarray(T_raw, np.float32)
        b = np.ascontiguousarray(T_raw_inv, np.float32)
        rc = self._L.b2icp_mapper_register(self._h, _ptr(c), len(c), _ptr(a), _ptr(b), C.byref(self._result))
        self._n_source = len(c)
        self._check(rc, "mapper_register")
        return self._result

    def mapperGrow(self, T, cloud=None) -> int:
        """b2icp_mapper_grow: transform the retained (or given) scan by T and add it to the map."""
        t = np.ascontiguousarray(T, np.float32)
        n_added = C.c_size_t()
        if cloud is None:
            rc = self._L.b2icp_mapper_grow(self._h, None, 0, _ptr(t), C.byref(n_added))
        else:
            c = _cloud(cloud)
            rc = self._L.b2icp_mapper_grow(self._h, _ptr(c), len(c), _ptr(t), C.byref(n_added))
        self._check(rc, "mapper_grow")
        return n_added.value

    def setInputTargetFromMap(self):
        """The map becomes the registration target without leaving the device."""
        self._check(self._L.b2icp_set_target_map(self._h), "set_target_map")

    def setStream(self, cuda_stream: int):
        self._check(self._L.b2icp_set_stream(self._h, C.c_void_p(cuda_stream)), "set_stream")

    def timing(self) -> Timing:
        t = Timing()
        self._check(self._L.b2icp_get_timing(self._h, C.byref(t)), "get_timing")
        return t

    def gridInfo(self):
        cell = C.c_float()
        dims = (C.c_int32 * 3)()
        occ = C.c_double()
        self._check(self._L.b2icp_get_grid_info(self._h, C.byref(cell), dims, C.byref(occ)), "get_grid_info")
        return dict(cell=cell.value, dims=tuple(dims), occupancy=occ.value)
