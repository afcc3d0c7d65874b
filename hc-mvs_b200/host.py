"""ctypes binding of the host-side mirror of MVS::Scene / MVS::DepthMapsData (include/hcmvs_host.h).

``HostScene.dense_reconstruction`` is the public end-to-end call: Scene::DenseReconstruction
(libs/MVS/SceneDensify.cpp:3532-3574) driven through the C ABI with host buffers.
"""
import ctypes as C
import os

import numpy as np

from . import api

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        api.load()  # libhcmvs_host.so links the CUDA library
        path = os.path.join(_HERE, "libhcmvs_host.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} missing: run __graft_entry__.build()")
        L = C.CDLL(path)
        vp, i32 = C.c_void_p, C.c_int
        L.hcmvs_host_scene_create.restype = vp
        L.hcmvs_host_scene_destroy.argtypes = [vp]
        L.hcmvs_host_last_error.restype = C.c_char_p
        L.hcmvs_host_last_error.argtypes = [vp]
        L.hcmvs_host_add_image.argtypes = [vp, i32, i32, vp, vp, vp, vp, C.c_char_p]
        L.hcmvs_host_set_sparse.argtypes = [vp, i32, vp, vp, vp]
        L.hcmvs_host_select_views.argtypes = [vp, C.POINTER(api.Params), i32]
        L.hcmvs_host_select_views_mt.argtypes = [vp, C.POINTER(api.Params), i32, i32]
        L.hcmvs_host_get_neighbors.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, i32]
        L.hcmvs_host_get_gray.argtypes = [vp, i32, vp]
        L.hcmvs_host_scale_image.argtypes = [vp, i32, i32, C.c_float, vp, C.POINTER(i32), C.POINTER(i32), vp, vp]
        L.hcmvs_host_init_depth.argtypes = [vp, i32, vp, vp]
        L.hcmvs_host_dense_reconstruction.argtypes = [vp, vp, C.POINTER(api.Params), C.c_uint64, i32, C.c_char_p, vp]
        L.hcmvs_host_dense_reconstruction_distributed.argtypes = [vp, vp, C.POINTER(api.Params), C.c_uint64, i32, i32, i32, vp]
        L.hcmvs_host_pin_images.argtypes = [vp, i32]
        L.hcmvs_host_dist_prepare.argtypes = [vp, vp, C.POINTER(api.Params), i32, i32]
        L.hcmvs_host_dist_upload_initial.argtypes = [vp]
        L.hcmvs_host_dist_info.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
        L.hcmvs_host_dist_run.argtypes = [vp, C.c_uint64, i32, i32, vp]
        L.hcmvs_host_shard_plan.argtypes = [vp, i32, vp, i32, i32, i32, vp, vp, vp, C.POINTER(i32), C.POINTER(i32)]
        L.hcmvs_host_cloud_size.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.hcmvs_host_cloud_get.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.hcmvs_host_cloud_save_ply.argtypes = [vp, C.c_char_p]
        L.hcmvs_host_write_dmap.argtypes = [C.c_char_p, C.c_char_p, vp, i32, i32, i32, vp, vp, vp, C.c_float, C.c_float, i32, i32, vp, vp, vp]
        L.hcmvs_host_read_dmap_header.argtypes = [C.c_char_p, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
        L.hcmvs_host_read_dmap.argtypes = [C.c_char_p, vp, vp, vp, vp, vp, vp, vp, vp]
        L.hcmvs_host_scene_load_mvs.argtypes = [vp, C.c_char_p, i32]
        L.hcmvs_host_scene_save_mvs.argtypes = [vp, C.c_char_p, i32, i32]
        L.hcmvs_host_num_images.argtypes = [vp]
        L.hcmvs_host_pointcloud_filter.argtypes = [vp, vp, i32]
        L.hcmvs_host_pointcloud_filter.restype = C.c_long
        L.hcmvs_host_cloud_remove_by_visibility.argtypes = [vp, vp, i32]
        L.hcmvs_host_cloud_remove_by_visibility.restype = C.c_long
        L.hcmvs_host_cloud_set.argtypes = [vp, C.c_uint64, vp, vp, vp, vp, vp, vp]
        L.hcmvs_host_delaunay.argtypes = [vp, i32, vp, i32]
        L.hcmvs_host_triangulate_init.argtypes = [vp, i32, i32, vp, i32, vp, i32, C.POINTER(i32), C.POINTER(i32), vp]
        L.hcmvs_host_scene_reload_images.argtypes = [vp, C.c_uint, C.c_uint, C.c_uint]
        L.hcmvs_host_resize_area_bgr.argtypes = [vp, i32, i32, i32, i32, vp]
        L.hcmvs_host_get_image_info.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_uint32), vp, vp, vp, C.c_char_p, i32]
        L.hcmvs_host_get_image_bgr.argtypes = [vp, i32, vp]
        L.hcmvs_host_get_sparse.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), vp, vp, vp, vp]
        L.hcmvs_host_save_depthmap.argtypes = [C.c_char_p, vp, i32, i32]
        L.hcmvs_host_save_normalmap.argtypes = [C.c_char_p, vp, i32, i32]
        L.hcmvs_host_load_depthmap.argtypes = [C.c_char_p, vp, C.POINTER(i32), C.POINTER(i32)]
        L.hcmvs_host_load_normalmap.argtypes = [C.c_char_p, vp, C.POINTER(i32), C.POINTER(i32)]
        L.hcmvs_host_load_image.argtypes = [C.c_char_p, C.POINTER(i32), C.POINTER(i32), vp]
        _lib = L
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class HostScene:
    def __init__(self):
        self.L = lib()
        self.h = self.L.hcmvs_host_scene_create()
        self.sizes = []

    @classmethod
    def from_synth(cls, syn, images=None):
        """images: list (or dict) of BGR arrays per view; an entry that is None (or missing from the dict) adds the camera only — the
        pixels of that image live on another rank (dense_reconstruction_distributed)."""
        s = cls()
        for i in range(syn.n_views):
            if images is None:
                bgr = syn.render(i, want_depth=False, want_normal=False)[0]
            else:
                bgr = images.get(i) if isinstance(images, dict) else images[i]
            s.add_image(syn.K[i], syn.R[i], syn.Cc[i], bgr, name=f"{i:05d}.png", size=(syn.height, syn.width))
        s.set_sparse(syn.sparse_xyz, syn.sparse_off, syn.sparse_views)
        return s

    def add_image(self, K, R, Cc, bgr, name="", size=None):
        if bgr is None:
            h, w = size
        else:
            bgr = np.ascontiguousarray(bgr, np.uint8)
            h, w = bgr.shape[:2]
        K = np.ascontiguousarray(K, np.float64); R = np.ascontiguousarray(R, np.float64); Cc = np.ascontiguousarray(Cc, np.float64)
        i = self.L.hcmvs_host_add_image(self.h, w, h, _p(K), _p(R), _p(Cc), _p(bgr), name.encode())
        if i < 0:
            raise RuntimeError("hcmvs_host_add_image failed")
        self.sizes.append((h, w))
        return i

    def set_sparse(self, xyz, off, views):
        xyz = np.ascontiguousarray(xyz, np.float32); off = np.ascontiguousarray(off, np.int32); views = np.ascontiguousarray(views, np.uint32)
        self.L.hcmvs_host_set_sparse(self.h, len(xyz), _p(xyz), _p(off), _p(views))

    def select_views(self, params, idx, threads=1):
        return self.L.hcmvs_host_select_views_mt(self.h, C.byref(params), idx, threads)

    def neighbors(self, idx, which=1, cap=64):
        ids = np.zeros(cap, np.uint32); pts = np.zeros(cap, np.uint32)
        sc = np.zeros(cap, np.float32); an = np.zeros(cap, np.float32); ar = np.zeros(cap, np.float32); s = np.zeros(cap, np.float32)
        n = min(max(self.L.hcmvs_host_get_neighbors(self.h, idx, which, _p(ids), _p(pts), _p(sc), _p(an), _p(ar), _p(s), cap), 0), cap)
        return dict(ids=ids[:n], points=pts[:n], scale=sc[:n], angle=an[:n], area=ar[:n], score=s[:n])

    def init_depth(self, idx):
        """Sparse-point initial depth map of a selected view -> (depth, dMin, dMax)."""
        h, w = self.sizes[idx]
        d = np.zeros((h, w), np.float32); mm = np.zeros(2, np.float32)
        if self.L.hcmvs_host_init_depth(self.h, idx, _p(d), _p(mm)) != 0:
            raise RuntimeError("view not selected")
        return d, float(mm[0]), float(mm[1])

    def triangulate_init(self, idx, add_corners=True):
        """TriangulatePointsDelaunay of a selected view -> (vertices (n,3) f64, tris (m,3) u32, dMin, dMax) (raw bounds)."""
        nv, nt = C.c_int(), C.c_int()
        if self.L.hcmvs_host_triangulate_init(self.h, idx, int(add_corners), None, 0, None, 0, C.byref(nv), C.byref(nt), None) != 0:
            raise RuntimeError("triangulate_init failed")
        v = np.empty((nv.value, 3), np.float64); t = np.empty((nt.value, 3), np.uint32); mm = np.zeros(2, np.float32)
        self.L.hcmvs_host_triangulate_init(self.h, idx, int(add_corners), _p(v), nv.value, _p(t), nt.value, None, None, _p(mm))
        return v, t, float(mm[0]), float(mm[1])

    def gray(self, idx):
        h, w = self.sizes[idx]
        g = np.zeros((h, w), np.float32)
        self.L.hcmvs_host_get_gray(self.h, idx, _p(g))
        return g

    def dense_reconstruction(self, ctx, seed=1, run_filter=True, dmap_dir=None):
        """Scene::DenseReconstruction: select views, upload, estimate all depth maps, (filter,) fuse. Returns stats."""
        st = np.zeros(8)
        r = self.L.hcmvs_host_dense_reconstruction(self.h, ctx.h, C.byref(ctx.params), seed, int(run_filter),
                                                   dmap_dir.encode() if dmap_dir else None, _p(st))
        if r != 0:
            raise RuntimeError(self.L.hcmvs_host_last_error(self.h).decode())
        for i, sz in enumerate(self.sizes):
            ctx.sizes.setdefault(i, sz)
        return dict(sec_select=st[0], sec_upload=st[1], sec_estimate=st[2], sec_filter=st[3], sec_fuse=st[4],
                    h2d_bytes=int(st[5]), d2h_bytes=int(st[6]), n_points=int(st[7]))

    def dense_reconstruction_distributed(self, ctx, rank, world, seed=1, run_filter=True):
        """Scene::DenseReconstruction over `world` GPUs (one process per GPU; ctx joined the communicator). Collective."""
        st = np.zeros(8)
        r = self.L.hcmvs_host_dense_reconstruction_distributed(self.h, ctx.h, C.byref(ctx.params), seed, int(run_filter), int(rank), int(world), _p(st))
        if r != 0:
            raise RuntimeError(self.L.hcmvs_host_last_error(self.h).decode())
        for i, sz in enumerate(self.sizes):
            ctx.sizes.setdefault(i, sz)
        return dict(sec_select=st[0], sec_upload=st[1], sec_estimate=st[2], sec_filter=st[3], sec_fuse=st[4],
                    h2d_bytes=int(st[5]), d2h_bytes=int(st[6]), n_points=int(st[7]))

    def _ckd(self, r):
        if r != 0:
            raise RuntimeError(self.L.hcmvs_host_last_error(self.h).decode())

    def pin_images(self, pin=True):
        """Page-lock the images' pixel buffers (uploads at PCIe rate); returns the number of pinned buffers."""
        return self.L.hcmvs_host_pin_images(self.h, int(pin))

    def dist_prepare(self, ctx, rank, world):
        self._ckd(self.L.hcmvs_host_dist_prepare(self.h, ctx.h, C.byref(ctx.params), int(rank), int(world)))
        self._rank = int(rank)
        for i, sz in enumerate(self.sizes):
            ctx.sizes.setdefault(i, sz)

    def dist_info(self):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._ckd(self.L.hcmvs_host_dist_info(self.h, self._rank, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(n_valid=a.value, n_mine_whole=b.value, n_split=c.value, whole_rounds=d.value)

    def dist_upload_initial(self):
        self._ckd(self.L.hcmvs_host_dist_upload_initial(self.h))

    def dist_run(self, seed=1, run_filter=True, download=False):
        st = np.zeros(8)
        self._ckd(self.L.hcmvs_host_dist_run(self.h, seed, int(run_filter), int(download), _p(st)))
        return dict(sec_select=st[0], sec_upload=st[1], sec_estimate=st[2], sec_filter=st[3], sec_fuse=st[4],
                    h2d_bytes=int(st[5]), d2h_bytes=int(st[6]), n_points=int(st[7]))

    def cloud(self):
        n = C.c_uint64(); m = C.c_uint64()
        self.L.hcmvs_host_cloud_size(self.h, C.byref(n), C.byref(m))
        n, m = n.value, m.value
        xyz = np.zeros((n, 3), np.float32); nrm = np.zeros((n, 3), np.float32); col = np.zeros((n, 3), np.uint8)
        off = np.zeros(n + 1, np.uint32); views = np.zeros(m, np.uint32); wts = np.zeros(m, np.float32)
        if n:
            self.L.hcmvs_host_cloud_get(self.h, _p(xyz), _p(nrm), _p(col), _p(off), _p(views), _p(wts))
        return dict(xyz=xyz, normals=nrm, colors=col, n_views=np.diff(off.astype(np.int64)).astype(np.int32), views=views, weights=wts)

    # ---- MVSI project files (Scene::LoadInterface / SaveInterface, libs/MVS/Scene.cpp:62-286)
    @classmethod
    def load_mvs(cls, path, load_images=True):
        s = cls()
        if s.L.hcmvs_host_scene_load_mvs(s.h, str(path).encode(), int(load_images)) != 0:
            raise RuntimeError("LoadInterface: " + s.L.hcmvs_host_last_error(s.h).decode())
        s.sizes = [(im["height"], im["width"]) for im in (s.image_info(i) for i in range(s.num_images()))]
        return s

    def save_mvs(self, path, version=-1, dense=False):
        if self.L.hcmvs_host_scene_save_mvs(self.h, str(path).encode(), version, int(dense)) != 0:
            raise RuntimeError("SaveInterface: " + self.L.hcmvs_host_last_error(self.h).decode())

    def reload_images(self, resolution_level, min_resolution=640, max_resolution=3200):
        """--resolution-level of DensifyPointCloud: shrink the images (INTER_AREA) and update the cameras."""
        if self.L.hcmvs_host_scene_reload_images(self.h, resolution_level, min_resolution, max_resolution) != 0:
            raise RuntimeError("ReloadImages: " + self.L.hcmvs_host_last_error(self.h).decode())
        self.sizes = [(im["height"], im["width"]) for im in (self.image_info(i) for i in range(self.num_images()))]

    def num_images(self):
        return self.L.hcmvs_host_num_images(self.h)

    def image_info(self, idx):
        w, h, cal, iid = C.c_int(), C.c_int(), C.c_int(), C.c_uint32()
        K = np.zeros(9); R = np.zeros(9); Cc = np.zeros(3)
        name = C.create_string_buffer(1024)
        if self.L.hcmvs_host_get_image_info(self.h, idx, C.byref(w), C.byref(h), C.byref(cal), C.byref(iid), _p(K), _p(R), _p(Cc), name, 1024) != 0:
            raise IndexError(idx)
        return dict(width=w.value, height=h.value, calibrated=bool(cal.value), id=iid.value, K=K, R=R, C=Cc, name=name.value.decode())

    def image_bgr(self, idx):
        h, w = self.sizes[idx]
        out = np.empty((h, w, 3), np.uint8)
        if self.L.hcmvs_host_get_image_bgr(self.h, idx, _p(out)) != 0:
            return None
        return out

    def sparse(self):
        n, m = C.c_uint64(), C.c_uint64()
        self.L.hcmvs_host_get_sparse(self.h, C.byref(n), C.byref(m), None, None, None, None)
        xyz = np.empty((n.value, 3), np.float32); off = np.empty(n.value + 1, np.int32)
        ids = np.empty(m.value, np.uint32); wts = np.empty(m.value, np.float32)
        self.L.hcmvs_host_get_sparse(self.h, None, None, _p(xyz), _p(off), _p(ids), _p(wts))
        return xyz, off, ids, wts

    def pointcloud_filter(self, ctx, th_remove):
        """Scene::PointCloudFilter on the dense cloud; returns the number of removed points."""
        r = self.L.hcmvs_host_pointcloud_filter(self.h, ctx.h, th_remove)
        if r < 0:
            raise RuntimeError(self.L.hcmvs_host_last_error(self.h).decode())
        return r

    def set_cloud(self, xyz, view_offsets, views, normals=None, colors=None, weights=None):
        a = lambda x, t: None if x is None else np.ascontiguousarray(x, t)
        xyz, view_offsets, views = a(xyz, np.float32), a(view_offsets, np.uint32), a(views, np.uint32)
        if self.L.hcmvs_host_cloud_set(self.h, len(xyz), _p(xyz), _p(a(normals, np.float32)), _p(a(colors, np.uint8)), _p(view_offsets), _p(views), _p(a(weights, np.float32))) != 0:
            raise ValueError("bad cloud")

    def remove_by_visibility(self, visibility, th_remove):
        return self.L.hcmvs_host_cloud_remove_by_visibility(self.h, _p(np.ascontiguousarray(visibility, np.int32)), th_remove)

    def save_ply(self, path):
        if self.L.hcmvs_host_cloud_save_ply(self.h, path.encode()) != 0:
            raise RuntimeError("PointCloud::Save failed")

    def close(self):
        if self.h:
            self.L.hcmvs_host_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_plan(valid_views, n_scored, world, split_rows=True):
    """The multi-GPU schedule of DenseReconstructionDistributed (host only): dict(order, owner_whole, owner_filter, whole_rounds, n_split)."""
    valid = np.ascontiguousarray(valid_views, np.uint32); ns = np.ascontiguousarray(n_scored, np.uint32)
    order = np.zeros(len(valid), np.uint32); ow = np.zeros(len(ns), np.int32); of = np.zeros(len(ns), np.int32)
    wr, nsp = C.c_int(), C.c_int()
    if lib().hcmvs_host_shard_plan(_p(valid), len(valid), _p(ns), len(ns), int(world), int(split_rows), _p(order), _p(ow), _p(of), C.byref(wr), C.byref(nsp)) != 0:
        raise ValueError("bad plan arguments")
    return dict(order=order, owner_whole=ow, owner_filter=of, whole_rounds=wr.value, n_split=nsp.value)


def load_image(path):
    """The project loader's image decoder (BMP / PNG / binary PNM -> BGR u8)."""
    w, h = C.c_int(), C.c_int()
    if lib().hcmvs_host_load_image(str(path).encode(), C.byref(w), C.byref(h), None) != 0:
        raise RuntimeError(f"cannot read the header of {path}")
    out = np.empty((h.value, w.value, 3), np.uint8)
    if lib().hcmvs_host_load_image(str(path).encode(), C.byref(w), C.byref(h), _p(out)) != 0:
        raise RuntimeError(f"cannot decode {path}")
    return out


def delaunay(xy):
    """Delaunay triangulation of (n, 2) f64 points -> (m, 3) u32 faces (counter-clockwise, smallest index first, sorted)."""
    xy = np.ascontiguousarray(xy, np.float64)
    m = lib().hcmvs_host_delaunay(_p(xy), len(xy), None, 0)
    if m < 0:
        raise ValueError("degenerate point set")
    t = np.empty((m, 3), np.uint32)
    lib().hcmvs_host_delaunay(_p(xy), len(xy), _p(t), m)
    return t


def resize_area_bgr(bgr, dsize):
    """cv::resize(bgr, dsize, interpolation=INTER_AREA) as Image::ResizeImage applies it (shrinking, 8-bit BGR); dsize = (w, h)."""
    bgr = np.ascontiguousarray(bgr, np.uint8)
    out = np.empty((dsize[1], dsize[0], 3), np.uint8)
    if lib().hcmvs_host_resize_area_bgr(_p(bgr), bgr.shape[1], bgr.shape[0], dsize[0], dsize[1], _p(out)) != 0:
        raise ValueError("resize_area_bgr: only shrinking is supported")
    return out


def scale_image(gray, scale, K=None):
    """ViewData::ScaleImage on an f32 gray image -> (scaled image, K of the new resolution) or None when |scale-1| < 0.15."""
    gray = np.ascontiguousarray(gray, np.float32); sh, sw = gray.shape
    dw = C.c_int(); dh = C.c_int()
    if lib().hcmvs_host_scale_image(_p(gray), sw, sh, float(scale), None, C.byref(dw), C.byref(dh), None, None) != 0:
        return None
    out = np.zeros((dh.value, dw.value), np.float32)
    Kin = np.ascontiguousarray(K, np.float64) if K is not None else None
    Kout = np.zeros(9) if K is not None else None
    lib().hcmvs_host_scale_image(_p(gray), sw, sh, float(scale), _p(out), C.byref(dw), C.byref(dh), _p(Kin), _p(Kout))
    return out, Kout


def write_dmap(path, image_name, ids, image_size, K, R, Cc, dmin, dmax, depth, normal=None, conf=None):
    """MVS::ExportDepthDataRaw (libs/MVS/DepthMap.cpp:2781-2846)."""
    depth = np.ascontiguousarray(depth, np.float32); h, w = depth.shape
    normal = np.ascontiguousarray(normal, np.float32) if normal is not None else None
    conf = np.ascontiguousarray(conf, np.float32) if conf is not None else None
    ids = np.ascontiguousarray(ids, np.uint32)
    K = np.ascontiguousarray(K, np.float64); R = np.ascontiguousarray(R, np.float64); Cc = np.ascontiguousarray(Cc, np.float64)
    r = lib().hcmvs_host_write_dmap(path.encode(), image_name.encode(), _p(ids), len(ids), image_size[0], image_size[1],
                                    _p(K), _p(R), _p(Cc), dmin, dmax, w, h, _p(depth), _p(normal), _p(conf))
    if r != 0:
        raise IOError(path)


def read_dmap(path):
    """MVS::ImportDepthDataRaw (libs/MVS/DepthMap.cpp:2848-2925)."""
    L = lib()
    w = C.c_int(); h = C.c_int(); n = C.c_int(); hn = C.c_int(); hc = C.c_int()
    if L.hcmvs_host_read_dmap_header(path.encode(), C.byref(w), C.byref(h), C.byref(n), C.byref(hn), C.byref(hc)) != 0:
        raise IOError(path)
    ids = np.zeros(n.value, np.uint32); K = np.zeros(9); R = np.zeros(9); Cc = np.zeros(3); mm = np.zeros(2, np.float32)
    depth = np.zeros((h.value, w.value), np.float32)
    normal = np.zeros((h.value, w.value, 3), np.float32) if hn.value else None
    conf = np.zeros((h.value, w.value), np.float32) if hc.value else None
    L.hcmvs_host_read_dmap(path.encode(), _p(ids), _p(K), _p(R), _p(Cc), _p(mm), _p(depth), _p(normal), _p(conf))
    return dict(ids=ids, K=K, R=R, C=Cc, dmin=float(mm[0]), dmax=float(mm[1]), depth=depth, normal=normal, conf=conf)


def save_depthmap(path, depth):
    """MVS::SaveDepthMap: depthmap/depthNNNN.dmap (zlib-compressed Boost binary archive of TImage<float>)."""
    d = np.ascontiguousarray(depth, np.float32)
    if lib().hcmvs_host_save_depthmap(str(path).encode(), _p(d), d.shape[1], d.shape[0]) != 0:
        raise RuntimeError("SaveDepthMap failed")


def save_normalmap(path, normal):
    n = np.ascontiguousarray(normal, np.float32)
    if lib().hcmvs_host_save_normalmap(str(path).encode(), _p(n), n.shape[1], n.shape[0]) != 0:
        raise RuntimeError("SaveNormalMap failed")


def load_depthmap(path):
    w, h = C.c_int(), C.c_int()
    if lib().hcmvs_host_load_depthmap(str(path).encode(), None, C.byref(w), C.byref(h)) != 0:
        raise RuntimeError("LoadDepthMap failed")
    d = np.zeros((h.value, w.value), np.float32)
    lib().hcmvs_host_load_depthmap(str(path).encode(), _p(d), C.byref(w), C.byref(h))
    return d


def load_normalmap(path):
    w, h = C.c_int(), C.c_int()
    if lib().hcmvs_host_load_normalmap(str(path).encode(), None, C.byref(w), C.byref(h)) != 0:
        raise RuntimeError("LoadNormalMap failed")
    n = np.zeros((h.value, w.value, 3), np.float32)
    lib().hcmvs_host_load_normalmap(str(path).encode(), _p(n), C.byref(w), C.byref(h))
    return n
