"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so) — test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(ROOT, "oracle", f) for f in ("hcmvs_oracle.cpp", "oracle_capi.cpp", "hcmvs_oracle.hpp")]
        if not os.path.exists(_SO) or any(os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
            build()
        L = C.CDLL(_SO)
        L.orc_scene_create.restype = C.c_void_p
        L.orc_sample.restype = C.c_float
        vp = C.c_void_p
        L.orc_scene_destroy.argtypes = [vp]
        L.orc_set_param.argtypes = [vp, C.c_char_p, C.c_double]
        L.orc_add_image.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
        L.orc_get_gray.argtypes = [vp, C.c_int, vp]
        L.orc_set_sparse.argtypes = [vp, C.c_int, vp, vp, vp]
        L.orc_select_views.argtypes = [vp, C.c_int]
        L.orc_init_views.argtypes = [vp, C.c_int, C.c_int]
        L.orc_get_neighbors.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int]
        L.orc_get_match_views.argtypes = [vp, C.c_int, vp, C.c_int]
        L.orc_get_points.argtypes = [vp, C.c_int, vp, C.c_int]
        L.orc_set_neighbors.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.c_int]
        L.orc_init_depth_sparse.argtypes = [vp, C.c_int]
        L.orc_set_neighbor_image.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
        L.orc_set_depthmap.argtypes = [vp, C.c_int, vp, vp, vp, C.c_float, C.c_float]
        L.orc_get_depthmap.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.orc_set_prior.argtypes = [vp, C.c_int, vp]
        L.orc_get_gramap.argtypes = [vp, C.c_int, vp]
        L.orc_set_coarse.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.c_int]
        L.orc_get_coarse.argtypes = [vp, C.c_int, vp, vp]
        L.orc_snapshot_maps.argtypes = [vp]
        L.orc_resize_area_up.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int]
        L.orc_score_depthmap.argtypes = [vp, C.c_int, C.c_int, C.c_uint64, C.c_int]
        L.orc_estimate_depthmap.argtypes = [vp, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.orc_score_hypotheses.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp]
        L.orc_end_depthmap.argtypes = [vp, C.c_int]
        L.orc_filter_depthmap.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, vp]
        L.orc_remove_small_segments.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_uint, C.c_float]
        L.orc_gap_interpolation.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_uint, C.c_float]
        L.orc_fuse.argtypes = [vp, C.c_int, C.c_int]
        L.orc_fuse_get.argtypes = [vp, vp, vp, vp, vp]
        L.orc_fuse_get_views.argtypes = [vp, vp, vp]
        L.orc_median3.argtypes = [vp, C.c_int, C.c_int]
        L.orc_gramap.argtypes = [vp, C.c_int, C.c_int, vp]
        L.orc_togray.argtypes = [vp, C.c_int, C.c_int, vp]
        L.orc_zigzag.argtypes = [C.c_int, C.c_int, C.c_int, vp]
        L.orc_philox.argtypes = [vp, vp, vp]
        L.orc_sample.argtypes = [vp, C.c_int, C.c_int, C.c_float, C.c_float]
        L.orc_dir2normal.argtypes = [C.c_float, C.c_float, vp]
        L.orc_normal2dir.argtypes = [vp, vp]
        _lib = L
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def resize_area_up(src, dw, dh):
    """cv::resize(src, (dw, dh), interpolation=INTER_AREA) for an enlargement, f32, 1 or 3 channels (oracle restatement)."""
    src = np.ascontiguousarray(src, np.float32)
    sh, sw = src.shape[:2]; cn = 1 if src.ndim == 2 else src.shape[2]
    dst = np.zeros((dh, dw) if src.ndim == 2 else (dh, dw, cn), np.float32)
    lib().orc_resize_area_up(_p(src), sw, sh, cn, _p(dst), dw, dh)
    return dst


class OracleScene:
    """Thin object wrapper: images + sparse cloud + per-view DepthData, driven like DepthMapsData."""

    def __init__(self, **params):
        self.L = lib()
        self.h = self.L.orc_scene_create()
        self.sizes = []
        self.set_params(**params)

    def set_params(self, **params):
        for k, v in params.items():
            if self.L.orc_set_param(self.h, k.encode(), float(v)) != 0:
                raise KeyError(k)

    def add_image(self, K, R, Cc, bgr=None, gray=None):
        if bgr is not None:
            h, w = bgr.shape[:2]
            bgr = np.ascontiguousarray(bgr, np.uint8)
        else:
            h, w = gray.shape
        if gray is not None:
            gray = np.ascontiguousarray(gray, np.float32)
        K = np.ascontiguousarray(K, np.float64); R = np.ascontiguousarray(R, np.float64); Cc = np.ascontiguousarray(Cc, np.float64)
        i = self.L.orc_add_image(self.h, w, h, _p(K), _p(R), _p(Cc), _p(bgr), _p(gray))
        assert i >= 0
        self.sizes.append((h, w))
        return i

    def gray(self, i):
        h, w = self.sizes[i]
        out = np.zeros((h, w), np.float32)
        self.L.orc_get_gray(self.h, i, _p(out))
        return out

    def set_sparse(self, xyz, off, views):
        xyz = np.ascontiguousarray(xyz, np.float32); off = np.ascontiguousarray(off, np.int32); views = np.ascontiguousarray(views, np.uint32)
        self.L.orc_set_sparse(self.h, len(xyz), _p(xyz), _p(off), _p(views))

    def select_views(self, i):
        return self.L.orc_select_views(self.h, i)

    def init_views(self, i, num):
        return self.L.orc_init_views(self.h, i, num)

    def neighbors(self, i, which=1, cap=64):
        ids = np.zeros(cap, np.uint32); pts = np.zeros(cap, np.uint32)
        sc = np.zeros(cap, np.float32); an = np.zeros(cap, np.float32); ar = np.zeros(cap, np.float32); s = np.zeros(cap, np.float32)
        n = self.L.orc_get_neighbors(self.h, i, which, _p(ids), _p(pts), _p(sc), _p(an), _p(ar), _p(s), cap)
        n = min(n, cap)
        return dict(ids=ids[:n], points=pts[:n], scale=sc[:n], angle=an[:n], area=ar[:n], score=s[:n])

    def match_views(self, i, cap=64):
        ids = np.zeros(cap, np.uint32)
        n = self.L.orc_get_match_views(self.h, i, _p(ids), cap)
        return ids[:n]

    def set_neighbors(self, i, ids, n_match, scores=None):
        ids = np.ascontiguousarray(ids, np.uint32)
        sc = np.ascontiguousarray(scores, np.float32) if scores is not None else None
        self.L.orc_set_neighbors(self.h, i, _p(ids), _p(sc), n_match, len(ids))

    def set_neighbor_image(self, i, slot, K, gray):
        """matching view `slot` of view i uses this rescaled gray image and intrinsics (None restores the scene image)."""
        if gray is None:
            assert self.L.orc_set_neighbor_image(self.h, i, slot, 0, 0, None, None) == 0
            return
        gray = np.ascontiguousarray(gray, np.float32); K = np.ascontiguousarray(K, np.float64)
        h, w = gray.shape
        assert self.L.orc_set_neighbor_image(self.h, i, slot, w, h, _p(K), _p(gray)) == 0

    def init_depth_sparse(self, i):
        self.L.orc_init_depth_sparse(self.h, i)

    def set_depthmap(self, i, depth, normal, conf, dmin, dmax):
        depth = np.ascontiguousarray(depth, np.float32)
        normal = np.ascontiguousarray(normal, np.float32) if normal is not None else None
        conf = np.ascontiguousarray(conf, np.float32) if conf is not None else None
        self.L.orc_set_depthmap(self.h, i, _p(depth), _p(normal), _p(conf), dmin, dmax)

    def get_depthmap(self, i):
        h, w = self.sizes[i]
        d = np.zeros((h, w), np.float32); n = np.zeros((h, w, 3), np.float32); c = np.zeros((h, w), np.float32); mm = np.zeros(2, np.float32)
        self.L.orc_get_depthmap(self.h, i, _p(d), _p(n), _p(c), _p(mm))
        return d, n, c, float(mm[0]), float(mm[1])

    def set_prior(self, i, prior):
        """DepthData::depthMapPrior (DepthMap.cpp:941-955); None removes it."""
        prior = np.ascontiguousarray(prior, np.float32) if prior is not None else None
        assert self.L.orc_set_prior(self.h, i, _p(prior)) == 0

    def set_coarse(self, i, depth, normal):
        """restore tree: coarse maps of the previous level -> resized to the view, widen [dMin,dMax) (None clears)."""
        if depth is None:
            self.L.orc_set_coarse(self.h, i, None, None, 0, 0)
            return
        depth = np.ascontiguousarray(depth, np.float32); normal = np.ascontiguousarray(normal, np.float32)
        hc, wc = depth.shape
        assert self.L.orc_set_coarse(self.h, i, _p(depth), _p(normal), wc, hc) == 0

    def get_coarse(self, i):
        h, w = self.sizes[i]
        d = np.zeros((h, w), np.float32); n = np.zeros((h, w, 3), np.float32)
        assert self.L.orc_get_coarse(self.h, i, _p(d), _p(n)) == 0
        return d, n

    def snapshot_maps(self):
        self.L.orc_snapshot_maps(self.h)

    def gramap(self, i):
        h, w = self.sizes[i]
        g = np.zeros((h, w), np.uint8)
        self.L.orc_get_gramap(self.h, i, _p(g))
        return g

    def score_depthmap(self, i, it_external=0, seed=1, threads=1):
        self.L.orc_score_depthmap(self.h, i, it_external, seed, threads)

    def estimate(self, i, it_external=0, seed=1, threads=1, mode=0, far_reach=11, run_end=True):
        st = np.zeros(5)
        r = self.L.orc_estimate_depthmap(self.h, i, it_external, seed, threads, mode, far_reach, int(run_end), _p(st))
        assert r == 0
        return dict(sec_score=st[0], sec_sweeps=st[1], sec_end=st[2], n_hyp=st[3], n_pixel_iters=st[4])

    def score_hypotheses(self, i, depth, normal, smooth_mode=0):
        h, w = self.sizes[i]
        depth = np.ascontiguousarray(depth, np.float32); normal = np.ascontiguousarray(normal, np.float32)
        out = np.zeros((h, w), np.float32)
        self.L.orc_score_hypotheses(self.h, i, _p(depth), _p(normal), smooth_mode, _p(out))
        return out

    def end_depthmap(self, i):
        self.L.orc_end_depthmap(self.h, i)

    def filter(self, i, nb_idx, adjust=True):
        h, w = self.sizes[i]
        nb = np.ascontiguousarray(nb_idx, np.uint32)
        d = np.zeros((h, w), np.float32); c = np.zeros((h, w), np.float32)
        r = self.L.orc_filter_depthmap(self.h, i, _p(nb), len(nb), int(adjust), _p(d), _p(c))
        return (d, c) if r == 0 else None

    def fuse(self, color=True, normal=True):
        n = self.L.orc_fuse(self.h, int(color), int(normal))
        xyz = np.zeros((n, 3), np.float32); nrm = np.zeros((n, 3), np.float32); col = np.zeros((n, 3), np.uint8); nv = np.zeros(n, np.int32)
        self.L.orc_fuse_get(self.h, _p(xyz), _p(nrm) if normal else None, _p(col) if color else None, _p(nv))
        tot = int(nv.sum())
        views = np.zeros(tot, np.uint32); wts = np.zeros(tot, np.float32)
        self.L.orc_fuse_get_views(self.h, _p(views), _p(wts))
        return dict(xyz=xyz, normals=nrm, colors=col, n_views=nv, views=views, weights=wts)

    def close(self):
        if self.h:
            self.L.orc_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def scene_from_synth(syn, views=None, threads=0, **params):
    """Build an OracleScene from a hcmvs_b200.synth.SynthScene; returns (oracle_scene, gt) with gt[i] = (depth, normal)."""
    osc = OracleScene(**params)
    gt = []
    for i in range(syn.n_views):
        bgr, d, n = syn.render(i, threads=threads)
        osc.add_image(syn.K[i], syn.R[i], syn.Cc[i], bgr=bgr)
        gt.append((d, n))
    osc.set_sparse(syn.sparse_xyz, syn.sparse_off, syn.sparse_views)
    return osc, gt
