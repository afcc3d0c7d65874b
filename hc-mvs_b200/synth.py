"""ctypes binding of the synthetic scene generator (include/hcmvs_synth.h)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class SynthCfg(C.Structure):
    _fields_ = [
        ("n_views", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
        ("focal", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
        ("surface", C.c_int32), ("n_bumps", C.c_int32),
        ("plane_a", C.c_double), ("plane_b", C.c_double),
        ("bump_sigma_min", C.c_double), ("bump_sigma_max", C.c_double), ("bump_height", C.c_double),
        ("layout", C.c_int32),
        ("cam_distance", C.c_double), ("cam_radius", C.c_double), ("cam_step_deg", C.c_double),
        ("extent_x", C.c_double), ("extent_y", C.c_double), ("tex_wavelength", C.c_double),
        ("n_sparse", C.c_int32), ("seed", C.c_uint64),
    ]


def _load():
    path = os.path.join(_HERE, "libhcmvs_synth.so")  # stand-alone: loads nothing of the CUDA product
    if not os.path.exists(path):
        raise ImportError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(path)
    lib.hcmvs_synth_create.restype = C.c_void_p
    lib.hcmvs_synth_create.argtypes = [C.POINTER(SynthCfg)]
    lib.hcmvs_synth_destroy.argtypes = [C.c_void_p]
    lib.hcmvs_synth_preset.argtypes = [C.c_int, C.c_double, C.c_int, C.POINTER(SynthCfg)]
    lib.hcmvs_synth_camera.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hcmvs_synth_render.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.hcmvs_synth_sparse_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.hcmvs_synth_sparse.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hcmvs_synth_height.restype = C.c_double
    lib.hcmvs_synth_height.argtypes = [C.c_void_p, C.c_double, C.c_double]
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class SynthScene:
    """A seeded synthetic scene: cameras, sparse cloud, on-demand rendering."""

    def __init__(self, config=1, scale=1.0, n_views=0, cfg=None):
        L = lib()
        if cfg is None:
            cfg = SynthCfg()
            if L.hcmvs_synth_preset(config, float(scale), int(n_views), C.byref(cfg)) != 0:
                raise ValueError("bad synthetic preset")
        self.cfg = cfg
        self._h = L.hcmvs_synth_create(C.byref(cfg))
        if not self._h:
            raise RuntimeError("hcmvs_synth_create failed")
        self.n_views, self.width, self.height = cfg.n_views, cfg.width, cfg.height
        self.K, self.R, self.Cc = [], [], []
        for i in range(self.n_views):
            K = np.zeros(9); R = np.zeros(9); Cc = np.zeros(3)
            L.hcmvs_synth_camera(self._h, i, _p(K), _p(R), _p(Cc))
            self.K.append(K); self.R.append(R); self.Cc.append(Cc)
        n = C.c_int(); m = C.c_int()
        L.hcmvs_synth_sparse_size(self._h, C.byref(n), C.byref(m))
        self.sparse_xyz = np.zeros((n.value, 3), np.float32)
        self.sparse_off = np.zeros(n.value + 1, np.int32)
        self.sparse_views = np.zeros(m.value, np.uint32)
        L.hcmvs_synth_sparse(self._h, _p(self.sparse_xyz), _p(self.sparse_off), _p(self.sparse_views))

    def render(self, view, want_bgr=True, want_depth=True, want_normal=True, threads=0):
        H, W = self.height, self.width
        bgr = np.zeros((H, W, 3), np.uint8) if want_bgr else None
        depth = np.zeros((H, W), np.float32) if want_depth else None
        normal = np.zeros((H, W, 3), np.float32) if want_normal else None
        if lib().hcmvs_synth_render(self._h, view, _p(bgr), _p(depth), _p(normal), threads) != 0:
            raise RuntimeError("render failed")
        return bgr, depth, normal

    def height_at(self, x, y):
        return lib().hcmvs_synth_height(self._h, float(x), float(y))

    def close(self):
        if self._h:
            lib().hcmvs_synth_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def c5_maps(gt, seed):
    """SURVEY §8(d) C5 recipe (fusion stress): GT depth x (1 + N(0, 0.002)), 2 % outliers U(dMin, dMax), normals rotated by U(0, 5 deg),
    conf U(0.5, 1). gt = (depth, normal) of one view -> (depth, normal, conf, dMin, dMax)."""
    rng = np.random.default_rng(seed)
    d, n = gt
    valid = d > 0
    lo, hi = float(d[valid].min()), float(d[valid].max())
    depth = (d * (1 + 0.002 * rng.standard_normal(d.shape))).astype(np.float32)
    out = rng.uniform(size=d.shape) < 0.02
    depth[out] = rng.uniform(lo, hi, int(out.sum())).astype(np.float32)
    depth[~valid] = 0
    ang = np.deg2rad(rng.uniform(0, 5, d.shape))
    axis = rng.standard_normal(d.shape + (3,))
    t = np.cross(n.astype(np.float64), axis)
    t /= np.maximum(np.linalg.norm(t, axis=2, keepdims=True), 1e-12)
    nn = np.cos(ang)[..., None] * n + np.sin(ang)[..., None] * t
    nn /= np.maximum(np.linalg.norm(nn, axis=2, keepdims=True), 1e-12)
    conf = rng.uniform(0.5, 1.0, d.shape).astype(np.float32)
    conf[depth == 0] = 0
    return depth, nn.astype(np.float32), conf, lo * 0.5, hi * 2.0


def frame_neighbors(i, n_views, reach=6):
    """<= 12 neighbours by frame distance, nearest first (C5: no sparse cloud, no view selection)."""
    ids = []
    for k in range(1, reach + 1):
        for j in (i - k, i + k):
            if 0 <= j < n_views:
                ids.append(j)
    return np.array(ids[:12], np.uint32)
