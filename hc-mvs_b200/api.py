"""ctypes binding of the C ABI in include/hcmvs_b200.h (libhcmvs_b200.so, CUDA sm_100a).

No fallback: loading fails loudly when the CUDA library is missing, and ``Context`` raises when
``hcmvs_create`` cannot find a CUDA device.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhcmvs_b200.so")

EXPORTS = [
    "hcmvs_default_params", "hcmvs_last_error", "hcmvs_create", "hcmvs_destroy", "hcmvs_set_params", "hcmvs_sync",
    "hcmvs_set_view", "hcmvs_set_neighbors", "hcmvs_set_neighbor_image", "hcmvs_init_depthmap", "hcmvs_init_depthmap_triangles", "hcmvs_download_depthmap_begin", "hcmvs_download_depthmap_wait", "hcmvs_set_depthmap", "hcmvs_get_depthmap",
    "hcmvs_set_prior", "hcmvs_set_coarse_estimate", "hcmvs_get_coarse_estimate", "hcmvs_snapshot_maps", "hcmvs_restore_snapshot", "hcmvs_remove_small_segments", "hcmvs_gap_interpolation", "hcmvs_set_view_remote", "hcmvs_comm_allgather_host", "hcmvs_pin_host_memory", "hcmvs_unpin_host_memory", "hcmvs_get_gradient_map", "hcmvs_score_depthmap", "hcmvs_estimate_depthmap", "hcmvs_estimate_depthmap_rows", "hcmvs_end_depthmap", "hcmvs_score_hypotheses",
    "hcmvs_filter_depthmap", "hcmvs_commit_filtered", "hcmvs_set_fuse_priority", "hcmvs_fuse_depthmaps",
    "hcmvs_free_pointcloud", "hcmvs_get_fused_device", "hcmvs_get_fused_support", "hcmvs_estimate_point_colors", "hcmvs_estimate_point_normals", "hcmvs_pointcloud_filter", "hcmvs_download_fused", "hcmvs_download_fused_pinned", "hcmvs_get_depthmap_device", "hcmvs_set_depth_range", "hcmvs_alloc_depthmap",
    "hcmvs_comm_unique_id", "hcmvs_comm_init", "hcmvs_exchange_maps", "hcmvs_exchange_wait", "hcmvs_export_maps_d", "hcmvs_import_maps_d", "hcmvs_get_timers", "hcmvs_reset_timers", "hcmvs_stream",
]


class Params(C.Structure):
    _fields_ = [
        ("nNumViews", C.c_uint32), ("nMaxViews", C.c_uint32), ("nMinViews", C.c_uint32), ("nMinViewsTrustPoint", C.c_uint32),
        ("nMinViewsFuse", C.c_uint32), ("nMinViewsFilter", C.c_uint32), ("nMinViewsFilterAdjust", C.c_uint32),
        ("bFilterAdjust", C.c_int32),
        ("fNCCThresholdKeep", C.c_float),
        ("nEstimationIters", C.c_uint32), ("nEstimationIters_external", C.c_uint32), ("nRandomIters", C.c_uint32),
        ("fRandomDepthRatio", C.c_float), ("fRandomAngle1Range", C.c_float), ("fRandomAngle2Range", C.c_float),
        ("fRandomSmoothDepth", C.c_float), ("fRandomSmoothNormal", C.c_float), ("fRandomSmoothBonus", C.c_float),
        ("fDescriptorMinMagnitudeThreshold", C.c_float),
        ("fDepthDiffThreshold", C.c_float), ("fNormalDiffThreshold", C.c_float), ("depthweight", C.c_float), ("normalweight", C.c_float),
        ("adapthalfwin", C.c_int32), ("propagatehalfwin", C.c_int32), ("propagatestep", C.c_int32), ("photo2geo", C.c_int32),
        ("photometric_flow", C.c_float), ("para_prior", C.c_float), ("fsigmaPrior", C.c_float),
        ("rb_far_reach", C.c_int32), ("rb_prop_dirs", C.c_int32), ("sampler", C.c_int32), ("viewspread", C.c_int32),
    ]


class PointCloudC(C.Structure):
    _fields_ = [
        ("n_points", C.c_uint64), ("points", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)),
        ("colors", C.POINTER(C.c_uint8)), ("view_offsets", C.POINTER(C.c_uint32)), ("views", C.POINTER(C.c_uint32)),
        ("weights", C.POINTER(C.c_float)),
    ]


class Timers(C.Structure):
    _fields_ = [
        ("ms_score", C.c_double), ("ms_sweeps", C.c_double), ("ms_end", C.c_double), ("ms_prep", C.c_double),
        ("ms_filter", C.c_double), ("ms_fuse", C.c_double),
        ("n_hypotheses", C.c_uint64), ("n_pixel_iters", C.c_uint64), ("n_view_scores", C.c_uint64), ("n_smooth_terms", C.c_uint64),
        ("n_launches", C.c_uint32), ("n_fuse_rounds", C.c_uint64), ("n_window_walks", C.c_uint64), ("ms_exchange", C.c_double),
        ("filter_bytes", C.c_uint64), ("fuse_seeds", C.c_uint64), ("fuse_probes", C.c_uint64),
    ]


class HcmvsError(RuntimeError):
    pass


_lib = None


def load():
    """Load libhcmvs_b200.so (raises ImportError if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                          "hcmvs_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u32, i32, f32 = C.c_void_p, C.c_uint32, C.c_int, C.c_float
    L.hcmvs_last_error.restype = C.c_char_p
    L.hcmvs_create.restype = vp
    L.hcmvs_create.argtypes = [i32, C.POINTER(Params)]
    L.hcmvs_destroy.argtypes = [vp]
    L.hcmvs_default_params.argtypes = [C.POINTER(Params)]
    L.hcmvs_set_params.argtypes = [vp, C.POINTER(Params)]
    L.hcmvs_sync.argtypes = [vp]
    L.hcmvs_set_view.argtypes = [vp, u32, i32, i32, vp, vp, vp, vp, vp]
    L.hcmvs_set_neighbors.argtypes = [vp, u32, vp, vp, i32, i32]
    L.hcmvs_set_neighbor_image.argtypes = [vp, u32, i32, i32, i32, vp, vp]
    L.hcmvs_init_depthmap.argtypes = [vp, u32, vp, vp, f32, f32]
    fp = C.POINTER(C.c_float)
    L.hcmvs_download_depthmap_begin.argtypes = [vp, u32, C.c_int]
    L.hcmvs_download_depthmap_wait.argtypes = [vp, C.c_int, C.POINTER(fp), C.POINTER(fp), C.POINTER(fp), fp, fp]
    L.hcmvs_init_depthmap_triangles.argtypes = [vp, u32, vp, C.c_int, vp, C.c_int, f32, f32]
    L.hcmvs_set_depthmap.argtypes = [vp, u32, vp, vp, vp, f32, f32]
    L.hcmvs_get_depthmap.argtypes = [vp, u32, vp, vp, vp, C.POINTER(f32), C.POINTER(f32)]
    L.hcmvs_set_prior.argtypes = [vp, u32, vp]
    L.hcmvs_get_gradient_map.argtypes = [vp, u32, vp]
    L.hcmvs_set_coarse_estimate.argtypes = [vp, u32, i32, i32, vp, vp]
    L.hcmvs_get_coarse_estimate.argtypes = [vp, u32, vp, vp]
    L.hcmvs_snapshot_maps.argtypes = [vp]
    L.hcmvs_restore_snapshot.argtypes = [vp]
    L.hcmvs_remove_small_segments.argtypes = [vp, C.c_uint32, C.c_uint, C.POINTER(C.c_uint64)]
    L.hcmvs_gap_interpolation.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.c_uint, C.POINTER(C.c_uint64)]
    L.hcmvs_score_depthmap.argtypes = [vp, u32, i32, C.c_uint64]
    L.hcmvs_estimate_depthmap.argtypes = [vp, u32, i32, C.c_uint64]
    L.hcmvs_estimate_depthmap_rows.argtypes = [vp, u32, i32, C.c_uint64, i32, i32]
    L.hcmvs_end_depthmap.argtypes = [vp, u32]
    L.hcmvs_score_hypotheses.argtypes = [vp, u32, vp, vp, i32, vp]
    L.hcmvs_filter_depthmap.argtypes = [vp, u32, vp, i32, i32, vp, vp]
    L.hcmvs_commit_filtered.argtypes = [vp]
    L.hcmvs_set_fuse_priority.argtypes = [vp, u32, f32]
    L.hcmvs_fuse_depthmaps.argtypes = [vp, i32, i32, C.POINTER(PointCloudC)]
    L.hcmvs_free_pointcloud.argtypes = [C.POINTER(PointCloudC)]
    L.hcmvs_get_fused_support.argtypes = [vp, u32, vp, vp]
    L.hcmvs_download_fused_pinned.argtypes = [vp, C.POINTER(PointCloudC)]
    L.hcmvs_estimate_point_colors.argtypes = [vp, C.c_uint64, vp, vp, vp, vp]
    L.hcmvs_pointcloud_filter.argtypes = [vp, C.c_uint64, vp, vp, vp, vp, vp]
    L.hcmvs_estimate_point_normals.argtypes = [vp, C.c_uint64, vp, vp, vp, C.c_int, vp]
    L.hcmvs_get_fused_device.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)] + [C.POINTER(vp)] * 6
    L.hcmvs_get_depthmap_device.argtypes = [vp, u32, C.POINTER(vp), C.POINTER(vp), C.POINTER(f32), C.POINTER(f32)]
    L.hcmvs_set_depth_range.argtypes = [vp, u32, f32, f32]
    L.hcmvs_alloc_depthmap.argtypes = [vp, u32]
    L.hcmvs_comm_unique_id.argtypes = [vp]
    L.hcmvs_comm_init.argtypes = [vp, vp, i32, i32]
    L.hcmvs_exchange_maps.argtypes = [vp, vp, u32, i32]
    L.hcmvs_exchange_wait.argtypes = [vp]
    L.hcmvs_export_maps_d.argtypes = [vp, u32, vp, vp]
    L.hcmvs_import_maps_d.argtypes = [vp, u32, vp, vp, f32, f32]
    L.hcmvs_get_timers.argtypes = [vp, C.POINTER(Timers)]
    L.hcmvs_reset_timers.argtypes = [vp]
    L.hcmvs_stream.restype = vp
    L.hcmvs_stream.argtypes = [vp]
    _lib = L
    return L


def comm_unique_id():
    """NCCL unique id (bytes) for Context.comm_init — create on rank 0 and hand to every rank."""
    buf = (C.c_char * 128)()
    if load().hcmvs_comm_unique_id(buf) != 0:
        raise HcmvsError(load().hcmvs_last_error().decode())
    return bytes(buf.raw)


def default_params(**over):
    p = Params()
    load().hcmvs_default_params(C.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One device context == the reference's ``DepthMapsData`` object (libs/MVS/SceneDensify.h:49-88)."""

    def __init__(self, device=0, params=None, **over):
        self.L = load()
        self.params = params if params is not None else default_params(**over)
        self.h = self.L.hcmvs_create(int(device), C.byref(self.params))
        if not self.h:
            raise HcmvsError(self.L.hcmvs_last_error().decode())
        self.sizes = {}

    def _ck(self, r):
        if r != 0:
            raise HcmvsError(f"[{r}] " + self.L.hcmvs_last_error().decode())

    def set_params(self, **over):
        for k, v in over.items():
            if not hasattr(self.params, k):
                raise KeyError(k)
            setattr(self.params, k, v)
        self._ck(self.L.hcmvs_set_params(self.h, C.byref(self.params)))

    def set_view(self, view, K, R, Cc, gray, bgr=None):
        gray = np.ascontiguousarray(gray, np.float32)
        h, w = gray.shape
        if bgr is not None:
            bgr = np.ascontiguousarray(bgr, np.uint8)
        K = np.ascontiguousarray(K, np.float64).reshape(9); R = np.ascontiguousarray(R, np.float64).reshape(9); Cc = np.ascontiguousarray(Cc, np.float64).reshape(3)
        self._ck(self.L.hcmvs_set_view(self.h, view, w, h, _p(K), _p(R), _p(Cc), _p(gray), _p(bgr)))
        self.sizes[view] = (h, w)

    def set_neighbors(self, ref, ids, n_match, scores=None):
        ids = np.ascontiguousarray(ids, np.uint32)
        sc = np.ascontiguousarray(scores, np.float32) if scores is not None else None
        self._ck(self.L.hcmvs_set_neighbors(self.h, ref, _p(ids), _p(sc), int(n_match), len(ids)))

    def set_neighbor_image(self, ref, slot, K, gray):
        """ViewData::ScaleImage: matching view `slot` of `ref` is matched against this rescaled image / intrinsics (None: its own image)."""
        if gray is None:
            self._ck(self.L.hcmvs_set_neighbor_image(self.h, ref, slot, 0, 0, None, None))
            return
        gray = np.ascontiguousarray(gray, np.float32); K = np.ascontiguousarray(K, np.float64)
        h, w = gray.shape
        self._ck(self.L.hcmvs_set_neighbor_image(self.h, ref, slot, w, h, _p(K), _p(gray)))

    def init_depthmap(self, ref, depth0, normal0, dmin, dmax):
        depth0 = np.ascontiguousarray(depth0, np.float32)
        normal0 = np.ascontiguousarray(normal0, np.float32) if normal0 is not None else None
        self._ck(self.L.hcmvs_init_depthmap(self.h, ref, _p(depth0), _p(normal0), dmin, dmax))

    def init_depthmap_triangles(self, ref, vertices, tris, dmin, dmax):
        """TriangulatePoints2DepthMap on the device: vertices (n, 3) f64 (x, y, depth), tris (m, 3) u32 counter-clockwise."""
        vertices = np.ascontiguousarray(vertices, np.float64); tris = np.ascontiguousarray(tris, np.uint32)
        self._ck(self.L.hcmvs_init_depthmap_triangles(self.h, ref, _p(vertices), len(vertices), _p(tris), len(tris), dmin, dmax))

    def download_begin(self, view, slot):
        """Queue the asynchronous read-back of a view's maps into page-locked slot `slot` (returns at once)."""
        self._ck(self.L.hcmvs_download_depthmap_begin(self.h, view, slot))
        self._dl = getattr(self, "_dl", {}); self._dl[slot] = self.sizes[view]

    def download_wait(self, slot):
        """Block until the slot's copy has landed -> (depth, normal, conf, dMin, dMax) as copies of the page-locked buffers."""
        fp = C.POINTER(C.c_float)
        d, n, c = fp(), fp(), fp(); lo, hi = C.c_float(), C.c_float()
        self._ck(self.L.hcmvs_download_depthmap_wait(self.h, slot, C.byref(d), C.byref(n), C.byref(c), C.byref(lo), C.byref(hi)))
        h, w = self._dl[slot]
        return (np.ctypeslib.as_array(d, (h, w)).copy(), np.ctypeslib.as_array(n, (h, w, 3)).copy(), np.ctypeslib.as_array(c, (h, w)).copy(), lo.value, hi.value)

    def set_depthmap(self, view, depth, normal, conf, dmin, dmax):
        depth = np.ascontiguousarray(depth, np.float32)
        normal = np.ascontiguousarray(normal, np.float32) if normal is not None else None
        conf = np.ascontiguousarray(conf, np.float32) if conf is not None else None
        self._ck(self.L.hcmvs_set_depthmap(self.h, view, _p(depth), _p(normal), _p(conf), dmin, dmax))

    def get_depthmap(self, view):
        h, w = self.sizes[view]
        d = np.zeros((h, w), np.float32); n = np.zeros((h, w, 3), np.float32); c = np.zeros((h, w), np.float32)
        a = C.c_float(); b = C.c_float()
        self._ck(self.L.hcmvs_get_depthmap(self.h, view, _p(d), _p(n), _p(c), C.byref(a), C.byref(b)))
        return d, n, c, a.value, b.value

    def gradient_map(self, view):
        h, w = self.sizes[view]
        g = np.zeros((h, w), np.uint8)
        self._ck(self.L.hcmvs_get_gradient_map(self.h, view, _p(g)))
        return g

    def set_prior(self, ref, prior):
        prior = np.ascontiguousarray(prior, np.float32) if prior is not None else None
        self._ck(self.L.hcmvs_set_prior(self.h, ref, _p(prior)))

    def set_coarse_estimate(self, view, depth, normal):
        """restore tree hand-off: coarse (depth, normal) maps of the previous level; None removes them."""
        if depth is None:
            self._ck(self.L.hcmvs_set_coarse_estimate(self.h, view, 0, 0, None, None))
            return
        depth = np.ascontiguousarray(depth, np.float32); normal = np.ascontiguousarray(normal, np.float32)
        hc, wc = depth.shape
        self._ck(self.L.hcmvs_set_coarse_estimate(self.h, view, wc, hc, _p(depth), _p(normal)))

    def get_coarse_estimate(self, view):
        h, w = self.sizes[view]
        d = np.zeros((h, w), np.float32); n = np.zeros((h, w, 3), np.float32)
        self._ck(self.L.hcmvs_get_coarse_estimate(self.h, view, _p(d), _p(n)))
        return d, n

    def snapshot_maps(self):
        self._ck(self.L.hcmvs_snapshot_maps(self.h))

    def remove_small_segments(self, view, speckle_size=100):
        """DepthMapsData::RemoveSmallSegments (stock OpenMVS speckle filter) on the view's maps; returns the number of zeroed pixels."""
        n = C.c_uint64()
        self._ck(self.L.hcmvs_remove_small_segments(self.h, int(view), int(speckle_size), C.byref(n)))
        return int(n.value)

    def gap_interpolation(self, depth, normal=None, conf=None, gap_size=7):
        """Small-gap branch of DepthMapsData::GapInterpolation on copies of the given maps -> (depth, normal, conf, n_filled)."""
        d = np.ascontiguousarray(depth, np.float32).copy(); h, w = d.shape
        nn = None if normal is None else np.ascontiguousarray(normal, np.float32).copy()
        cc = None if conf is None else np.ascontiguousarray(conf, np.float32).copy()
        n = C.c_uint64()
        self._ck(self.L.hcmvs_gap_interpolation(self.h, w, h, _p(d), _p(nn), _p(cc), int(gap_size), C.byref(n)))
        return d, nn, cc, int(n.value)

    def restore_snapshot(self):
        self._ck(self.L.hcmvs_restore_snapshot(self.h))

    def score_depthmap(self, ref, it_external=0, seed=1):
        self._ck(self.L.hcmvs_score_depthmap(self.h, ref, it_external, seed))

    def estimate_depthmap(self, ref, it_external=0, seed=1):
        self._ck(self.L.hcmvs_estimate_depthmap(self.h, ref, it_external, seed))

    def estimate_depthmap_rows(self, ref, row_begin, row_end, it_external=0, seed=1):
        """Rows [row_begin, row_end) of the view only (one band of a view split between GPUs): bit-identical to those rows of the full estimate."""
        self._ck(self.L.hcmvs_estimate_depthmap_rows(self.h, ref, it_external, seed, row_begin, row_end))

    def end_depthmap(self, ref):
        self._ck(self.L.hcmvs_end_depthmap(self.h, ref))

    def score_hypotheses(self, ref, depth, normal, smooth_mode=0):
        h, w = self.sizes[ref]
        depth = np.ascontiguousarray(depth, np.float32); normal = np.ascontiguousarray(normal, np.float32)
        out = np.zeros((h, w), np.float32)
        self._ck(self.L.hcmvs_score_hypotheses(self.h, ref, _p(depth), _p(normal), smooth_mode, _p(out)))
        return out

    def filter_depthmap(self, ref, nb_idx, adjust=True, download=True):
        h, w = self.sizes[ref]
        nb = np.ascontiguousarray(nb_idx, np.uint32)
        d = np.zeros((h, w), np.float32) if download else None
        c = np.zeros((h, w), np.float32) if download else None
        self._ck(self.L.hcmvs_filter_depthmap(self.h, ref, _p(nb), len(nb), int(adjust), _p(d), _p(c)))
        return d, c

    def commit_filtered(self):
        self._ck(self.L.hcmvs_commit_filtered(self.h))

    def set_fuse_priority(self, view, score):
        self._ck(self.L.hcmvs_set_fuse_priority(self.h, view, float(score)))

    def fused_support(self, view):
        """depthMap_fuse / normalMap_fuse of the fork's RemoveSmallSegments: the view's estimate where it joined a fused point."""
        h, w = self.sizes[view]
        d = np.zeros((h, w), np.float32); n = np.zeros((h, w, 3), np.float32)
        self._ck(self.L.hcmvs_get_fused_support(self.h, view, _p(d), _p(n)))
        return d, n

    def estimate_point_colors(self, points=None, view_offsets=None, views=None):
        """MVS::EstimatePointColors: colours of the given points (host CSR view lists), or of the device-resident fused cloud."""
        if points is None:  # the fused cloud resident on the device: recoloured in place, colours returned
            n = C.c_uint64()
            self._ck(self.L.hcmvs_get_fused_device(self.h, C.byref(n), None, None, None, None, None, None, None))
            out = np.zeros((int(n.value), 3), np.uint8)
            self._ck(self.L.hcmvs_estimate_point_colors(self.h, 0, None, None, None, _p(out)))
            return out
        points = np.ascontiguousarray(points, np.float32); view_offsets = np.ascontiguousarray(view_offsets, np.uint32); views = np.ascontiguousarray(views, np.uint32)
        out = np.zeros((len(points), 3), np.uint8)
        self._ck(self.L.hcmvs_estimate_point_colors(self.h, len(points), _p(points), _p(view_offsets), _p(views), _p(out)))
        return out

    def estimate_point_normals(self, points=None, view_offsets=None, views=None, num_neighbors=16):
        """MVS::EstimatePointNormals (k-NN PCA, oriented to the first view) for host points or the resident fused cloud."""
        if points is None:
            n = C.c_uint64()
            self._ck(self.L.hcmvs_get_fused_device(self.h, C.byref(n), None, None, None, None, None, None, None))
            out = np.zeros((int(n.value), 3), np.float32)
            self._ck(self.L.hcmvs_estimate_point_normals(self.h, 0, None, None, None, num_neighbors, _p(out)))
            return out
        points = np.ascontiguousarray(points, np.float32); view_offsets = np.ascontiguousarray(view_offsets, np.uint32); views = np.ascontiguousarray(views, np.uint32)
        out = np.zeros((len(points), 3), np.float32)
        self._ck(self.L.hcmvs_estimate_point_normals(self.h, len(points), _p(points), _p(view_offsets), _p(views), num_neighbors, _p(out)))
        return out

    def pointcloud_filter(self, points=None, view_offsets=None, views=None):
        """Scene::PointCloudFilter's visibility votes -> (visibility int32 per point, stats) for host points or the resident fused cloud."""
        stats = np.zeros(3, np.uint64)
        if points is None:
            n = C.c_uint64()
            self._ck(self.L.hcmvs_get_fused_device(self.h, C.byref(n), None, None, None, None, None, None, None))
            vis = np.zeros(int(n.value), np.int32)
            self._ck(self.L.hcmvs_pointcloud_filter(self.h, 0, None, None, None, _p(vis), _p(stats)))
            return vis, stats
        points = np.ascontiguousarray(points, np.float32); view_offsets = np.ascontiguousarray(view_offsets, np.uint32); views = np.ascontiguousarray(views, np.uint32)
        vis = np.zeros(len(points), np.int32)
        self._ck(self.L.hcmvs_pointcloud_filter(self.h, len(points), _p(points), _p(view_offsets), _p(views), _p(vis), _p(stats)))
        return vis, stats

    def download_fused_pinned(self):
        """Copy the device-resident fused cloud into the context's page-locked arena; returns (n_points, bytes that crossed PCIe)."""
        pc = PointCloudC()
        self._ck(self.L.hcmvs_download_fused_pinned(self.h, C.byref(pc)))
        n = int(pc.n_points)
        m = int(pc.view_offsets[n]) if n else 0
        return n, n * (12 + (12 if pc.normals else 0) + (3 if pc.colors else 0) + 4) + 4 + m * 8

    def fused_counts(self):
        """(n_points, n_view_refs) of the fused cloud that lives on the device."""
        n = C.c_uint64(); m = C.c_uint64()
        self._ck(self.L.hcmvs_get_fused_device(self.h, C.byref(n), C.byref(m), None, None, None, None, None, None))
        return int(n.value), int(m.value)

    def fuse_depthmaps_device(self, color=True, normal=True):
        """FuseDepthMaps leaving the cloud in HBM; returns (n_points, n_view_refs)."""
        self._ck(self.L.hcmvs_fuse_depthmaps(self.h, int(color), int(normal), None))
        n = C.c_uint64(); m = C.c_uint64()
        self._ck(self.L.hcmvs_get_fused_device(self.h, C.byref(n), C.byref(m), None, None, None, None, None, None))
        return int(n.value), int(m.value)

    def fuse_depthmaps(self, color=True, normal=True):
        pc = PointCloudC()
        self._ck(self.L.hcmvs_fuse_depthmaps(self.h, int(color), int(normal), C.byref(pc)))
        n = int(pc.n_points)
        out = dict(xyz=np.zeros((0, 3), np.float32), normals=None, colors=None, n_views=np.zeros(0, np.int32),
                   views=np.zeros(0, np.uint32), weights=np.zeros(0, np.float32))
        if n:
            out["xyz"] = np.ctypeslib.as_array(pc.points, (n, 3)).copy()
            off = np.ctypeslib.as_array(pc.view_offsets, (n + 1,)).copy()
            m = int(off[-1])
            out["n_views"] = np.diff(off.astype(np.int64)).astype(np.int32)
            out["views"] = np.ctypeslib.as_array(pc.views, (m,)).copy()
            out["weights"] = np.ctypeslib.as_array(pc.weights, (m,)).copy()
            if pc.normals:
                out["normals"] = np.ctypeslib.as_array(pc.normals, (n, 3)).copy()
            if pc.colors:
                out["colors"] = np.ctypeslib.as_array(pc.colors, (n, 3)).copy()
        self.L.hcmvs_free_pointcloud(C.byref(pc))
        return out

    def depthmap_device(self, view):
        dn = C.c_void_p(); cf = C.c_void_p(); a = C.c_float(); b = C.c_float()
        self._ck(self.L.hcmvs_get_depthmap_device(self.h, view, C.byref(dn), C.byref(cf), C.byref(a), C.byref(b)))
        return dn.value, cf.value, a.value, b.value

    def set_depth_range(self, view, dmin, dmax):
        self._ck(self.L.hcmvs_set_depth_range(self.h, view, dmin, dmax))

    def export_maps_d(self, view, dn_ptr, conf_ptr):
        self._ck(self.L.hcmvs_export_maps_d(self.h, view, dn_ptr, conf_ptr))

    def import_maps_d(self, view, dn_ptr, conf_ptr, dmin, dmax):
        self._ck(self.L.hcmvs_import_maps_d(self.h, view, dn_ptr, conf_ptr, dmin, dmax))

    def alloc_depthmap(self, view):
        self._ck(self.L.hcmvs_alloc_depthmap(self.h, view))

    def sync(self):
        self._ck(self.L.hcmvs_sync(self.h))

    def comm_init(self, id_bytes, rank, world):
        """Join the NCCL communicator of the map exchange (collective); id_bytes from comm_unique_id() on rank 0."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(id_bytes))
        self._ck(self.L.hcmvs_comm_init(self.h, buf, int(rank), int(world)))

    def exchange_maps(self, owner, what=0, overlap=False):
        """Broadcast every view's maps in place from its owning rank (what = 0 estimated maps, 1 pending filter output);
        overlap=True runs it on the communication stream (join with exchange_wait)."""
        owner = np.ascontiguousarray(owner, np.int32)
        self._ck(self.L.hcmvs_exchange_maps(self.h, _p(owner), len(owner), int(what) | (0x100 if overlap else 0)))

    def exchange_wait(self):
        self._ck(self.L.hcmvs_exchange_wait(self.h))

    def timers(self):
        t = Timers()
        self._ck(self.L.hcmvs_get_timers(self.h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in Timers._fields_}

    def reset_timers(self):
        self._ck(self.L.hcmvs_reset_timers(self.h))

    def stream(self):
        return self.L.hcmvs_stream(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.hcmvs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
