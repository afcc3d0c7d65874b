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
