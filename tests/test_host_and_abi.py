"""CPU tests of the host-side mirror (view selection, image prep, file formats) and of the C-ABI surface."""
import ctypes as C
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import common
import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hcmvs_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(built):
    """The shared libraries load without a GPU and export every entry point include/*.h declares."""
    from hcmvs_b200 import api, host, synth
    L = api.load()
    for name in _declared("hcmvs_b200.h"):
        assert hasattr(L, name), f"libhcmvs_b200.so lacks {name}"
    assert set(api.EXPORTS) <= set(_declared("hcmvs_b200.h"))
    H = host.lib()
    for name in _declared("hcmvs_host.h"):
        assert hasattr(H, name), f"libhcmvs_host.so lacks {name}"
    S = synth.lib()
    for name in _declared("hcmvs_synth.h"):
        assert hasattr(S, name), f"libhcmvs_host.so lacks {name}"
    # only sm_100a code is embedded (no multi-arch fatbin, no PTX fallback for other GPUs)
    out = subprocess.run(["cuobjdump", "-lelf", api.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_headers_are_plain_c_and_link_from_c(built, tmp_path):
    """The drop-in boundary is a C ABI: every header under include/ compiles as strict C99 and a C program links and calls it."""
    src = tmp_path / "abi_check.c"
    src.write_text('#include "hcmvs_b200.h"\n#include "hcmvs_host.h"\n#include "hcmvs_synth.h"\n'
                   'int main(void) { hcmvs_params p; hcmvs_default_params(&p); return (p.nNumViews == 5 && p.nMinViewsTrustPoint == 2 && hcmvs_sync(0) < 0) ? 0 : 1; }\n')
    inc, libdir = os.path.join(ROOT, "include"), os.path.join(ROOT, "hc-mvs_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I" + inc, str(src)], check=True)
    exe = tmp_path / "abi_check"
    subprocess.run(["gcc", "-std=c99", "-I" + inc, str(src), "-L" + libdir, "-lhcmvs_b200", "-Wl,-rpath," + libdir, "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_no_cpu_fallback_and_error_convention(built):
    """Without a CUDA device creation fails with a message (bool-return convention -> status codes); defaults mirror OPTDENSE."""
    import torch
    from hcmvs_b200 import api
    p = api.default_params()
    assert (p.nNumViews, p.nMaxViews, p.nMinViewsFuse, p.nRandomIters) == (5, 12, 2, 6)
    assert abs(p.fNCCThresholdKeep - 0.55) < 1e-7 and abs(p.fRandomDepthRatio - 0.003) < 1e-9 and p.adapthalfwin == 5
    if torch.cuda.is_available():
        pytest.skip("GPU present: the negative path is covered on the CPU box")
    with pytest.raises(api.HcmvsError, match="no CUDA device"):
        api.Context(0)
    L = api.load()
    assert L.hcmvs_sync(None) < 0 and L.hcmvs_set_neighbors(None, 0, None, None, 0, 0) < 0


def test_host_view_selection_and_image_prep_bit_exact(built):
    """Scene::SelectNeighborViews / FilterNeighborViews (Scene.cpp:545-678) in the product host code == oracle, bit for bit."""
    from hcmvs_b200 import api, host
    for cfg, scale, nv in ((1, 0.25, 0), (2, 0.125, 0), (4, 0.125, 30)):
        syn, osc, gt, imgs, ok = common.make_scene(cfg, scale, nv)
        hs = host.HostScene.from_synth(syn, imgs)
        P = api.default_params(nMinViewsTrustPoint=1)
        for i in range(syn.n_views):
            # every third view with the work inside the image spread over 5 threads (DenseReconstruction's first view): same bits
            r = hs.select_views(P, i, threads=5 if i % 3 == 0 else 1)
            assert (r > 0) == ok[i]
            for which in (0, 1):
                a, b = hs.neighbors(i, which), osc.neighbors(i, which)
                for k in a:
                    assert np.array_equal(a[k], b[k]), (cfg, i, which, k)
            assert np.array_equal(hs.gray(i), osc.gray(i))
            if ok[i]:
                osc.init_depth_sparse(i)
                d0, _, _, lo, hi = osc.get_depthmap(i)
                d1, lo1, hi1 = hs.init_depth(i)
                assert np.array_equal(d0, d1) and lo == lo1 and hi == hi1


def test_view_selection_is_independent_of_the_thread_count(built):
    """Scene::SelectNeighborViews with the work inside one image spread over threads (DenseReconstruction's first view, the one the GPU
    waits for): per-point terms in parallel, f32 sums serially in cloud order -> the same bits for 1, 2, 3, 7 and 16 threads."""
    from hcmvs_b200 import api, host
    syn, osc, gt, imgs, ok = common.make_scene(2, 0.125, 0)
    hs = host.HostScene.from_synth(syn, imgs)
    P = api.default_params(nMinViewsTrustPoint=1)
    for i in (0, 7, 24, 48):
        ref = None
        for th in (1, 2, 3, 7, 16):
            r = hs.select_views(P, i, threads=th)
            cur = (r, {w: {k: v.copy() for k, v in hs.neighbors(i, w).items()} for w in (0, 1)}, hs.init_depth(i) if r > 0 else None)
            if ref is None:
                ref = cur
                continue
            assert cur[0] == ref[0]
            for w in (0, 1):
                for k in ref[1][w]:
                    assert np.array_equal(cur[1][w][k], ref[1][w][k]), (i, th, w, k)
            if ref[2] is not None:
                assert np.array_equal(cur[2][0], ref[2][0]) and cur[2][1:] == ref[2][1:]


def test_dmap_roundtrip_and_layout(built, tmp_path):
    """Raw 'DR' depth-data file (Interface.h:634-652, DepthMap.cpp:2781-2925): 28-byte header, name, ids, K R C, maps."""
    from hcmvs_b200 import host
    rng = np.random.default_rng(0)
    h, w = 13, 17
    d = rng.uniform(1, 5, (h, w)).astype(np.float32); n = rng.standard_normal((h, w, 3)).astype(np.float32); c = rng.uniform(0, 1, (h, w)).astype(np.float32)
    K = np.array([100, 0, 8, 0, 100, 6, 0, 0, 1.0]); R = np.eye(3).ravel(); Cc = np.array([1.0, 2, 3])
    path = str(tmp_path / "depth0007.dmap")
    host.write_dmap(path, "00007.png", [7, 3, 9], (w, h), K, R, Cc, 0.5, 9.5, d, n, c)
    raw = open(path, "rb").read()
    name, typ, pad, iw, ih, dw, dh, dmin, dmax = struct.unpack("<HBBIIIIff", raw[:28])
    assert (name, typ, iw, ih, dw, dh) == (0x5244, 7, w, h, w, h) and raw[:2] == b"DR" and (dmin, dmax) == (0.5, 9.5)
    assert struct.unpack("<H", raw[28:30])[0] == 9 and raw[30:39] == b"00007.png"
    assert struct.unpack("<I3I", raw[39:55]) == (3, 7, 3, 9)
    assert len(raw) == 55 + 21 * 8 + h * w * 4 * 5
    back = host.read_dmap(path)
    assert np.array_equal(back["depth"], d) and np.array_equal(back["normal"], n) and np.array_equal(back["conf"], c)
    assert list(back["ids"]) == [7, 3, 9] and np.array_equal(back["K"], K) and back["dmin"] == 0.5
    host.write_dmap(path, "x", [1], (w, h), K, R, Cc, 1, 2, d)                 # depth only
    back = host.read_dmap(path)
    assert back["normal"] is None and back["conf"] is None


def test_shard_plan_covers_every_view_once():
    from hcmvs_b200 import shard
    valid = [0, 1, 2, 4, 5, 7, 8, 9, 11]
    nall = {v: (v * 7) % 5 + 3 for v in valid}
    for world in (1, 2, 4, 8):
        plan = shard.make_plan(valid, nall, world)
        owned = [v for r in range(world) for v in plan.views_of(r)]
        assert sorted(owned) == valid
        assert max(len(plan.views_of(r)) for r in range(world)) - min(len(plan.views_of(r)) for r in range(world)) <= 1
        assert [nall[v] for v in plan.order] == sorted((nall[v] for v in valid), reverse=True)
        loc = plan.location()
        assert len({loc[v] for v in valid}) == len(valid) and all(s < plan.slots for _, s in loc.values())


def test_shard_plan_row_split_of_the_incomplete_round():
    """49 views on 8 ranks: 6 whole rounds + 1 view that every rank estimates a band of; bands tile the rows exactly."""
    from hcmvs_b200 import shard
    valid = list(range(49)); nall = {v: 12 - v % 5 for v in valid}
    for world in (2, 4, 8):
        plan = shard.make_plan(valid, nall, world, split_rows=True)
        assert plan.whole_rounds() == 49 // world and len(plan.split_views()) == 49 % world
        whole = [v for r in range(world) for v in plan.whole_views_of(r)]
        assert sorted(whole + plan.split_views()) == valid and len({len(plan.whole_views_of(r)) for r in range(world)}) == 1
        rows = [plan.rows_of(r, 1201) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == 1201 and all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
        o = plan.split_owner_array(49)
        assert (o == shard.OWNER_SPLIT_ROWS).sum() == len(plan.split_views()) and len(plan.round_owner_arrays(49)) == plan.whole_rounds()
        assert all((a >= 0).sum() == world for a in plan.round_owner_arrays(49))
    assert not shard.make_plan(list(range(48)), {v: 1 for v in range(48)}, 8, split_rows=True).split_views()   # nothing left over
    assert not shard.make_plan(valid, nall, 1, split_rows=True).split_views()


def test_scale_image_matches_opencv(built):
    """ViewData::ScaleImage (DepthMap.h:232-238) = cv::resize INTER_AREA (shrinking) / INTER_CUBIC (enlarging) on the f32 gray image.
    The host restatement follows OpenCV's scalar code: bit-equal to cv2 for the general area decimation; the vectorised kernels cv2
    runs for integer factors and for the cubic taps associate the sums differently (<= 1e-6 on [0,1] images)."""
    cv2 = pytest.importorskip("cv2")
    from hcmvs_b200 import host
    rng = np.random.default_rng(0)
    img = cv2.GaussianBlur(rng.uniform(0, 1, (240, 320)).astype(np.float32), (0, 0), 2.0)
    K = np.array([600.0, 0, 159.5, 0, 600.0, 119.5, 0, 0, 1])
    for sc, tol in ((0.7, 0.0), (0.83, 0.0), (0.6180339, 0.0), (0.5, 2e-7), (0.25, 2e-7), (1.2, 1e-6), (1.5, 1e-6), (2.0, 1e-6), (1.37, 1e-6)):
        got, Ks = host.scale_image(img, sc, K)
        f = np.float32(sc).item()
        want = cv2.resize(img, None, fx=f, fy=f, interpolation=cv2.INTER_CUBIC if sc > 1 else cv2.INTER_AREA)
        assert got.shape == want.shape, (sc, got.shape, want.shape)
        assert np.abs(got - want).max() <= tol, (sc, np.abs(got - want).max())
        s = max(got.shape) / 320.0                                    # Image::GetCamera: K scales with max(width, height)
        assert np.allclose(Ks[[0, 4, 2, 5]], K[[0, 4, 2, 5]] * s, rtol=1e-7) and Ks[8] == 1
    for sc in (1.0, 0.9, 1.14):                                       # |scale-1| < 0.15: the reference keeps the image
        assert host.scale_image(img, sc) is None


def test_pointcloud_filter_removal_order_and_ply_writer(built, tmp_path):
    """Scene::PointCloudFilter's removal (RFOREACH + RemovePoint: the last point moves into the hole) against a direct simulation, on
    random votes; then PointCloud::Save: binary little-endian PLY, x y z float32, red green blue uint8 (stored B G R), nx ny nz."""
    from hcmvs_b200 import host
    rng = np.random.default_rng(4)
    n = 5000
    xyz = rng.standard_normal((n, 3)).astype(np.float32); nrm = rng.standard_normal((n, 3)).astype(np.float32); col = rng.integers(0, 256, (n, 3)).astype(np.uint8)
    counts = rng.integers(1, 6, n); off = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
    views = rng.integers(0, 30, int(off[-1])).astype(np.uint32); wts = rng.uniform(0, 1, int(off[-1])).astype(np.float32)
    vis = rng.integers(-3, 4, n).astype(np.int32)
    vis[-1] = -3; vis[0] = -3                                                # the last and the first point go too
    hs = host.HostScene()
    hs.set_cloud(xyz, off, views, nrm, col, wts)
    removed = hs.remove_by_visibility(vis, -1)
    ids = list(range(n))                                                       # the reference's loop, literally
    for i in range(n - 1, -1, -1):
        if vis[i] <= -1:
            last = ids.pop()
            if i < len(ids):
                ids[i] = last
    ids = np.asarray(ids)
    after = hs.cloud()
    assert removed == n - len(ids) == int((vis <= -1).sum())
    assert np.array_equal(after["xyz"], xyz[ids]) and np.array_equal(after["normals"], nrm[ids]) and np.array_equal(after["colors"], col[ids])
    assert np.array_equal(after["n_views"], counts[ids])
    assert np.array_equal(after["views"], np.concatenate([views[off[i]:off[i + 1]] for i in ids])) and np.array_equal(after["weights"], np.concatenate([wts[off[i]:off[i + 1]] for i in ids]))
    assert hs.remove_by_visibility(np.zeros(len(ids), np.int32), -1) == 0     # nothing below the threshold: untouched
    path = tmp_path / "cloud.ply"
    hs.save_ply(str(path))
    head, body = path.read_bytes().split(b"end_header\n", 1)
    assert head.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % len(ids)) and head.count(b"property") == 9
    rec = np.frombuffer(body, dtype=np.dtype([("p", "<f4", 3), ("rgb", "u1", 3), ("n", "<f4", 3)]))
    assert len(rec) == len(ids) and np.array_equal(rec["p"], xyz[ids]) and np.array_equal(rec["rgb"], col[ids][:, ::-1]) and np.array_equal(rec["n"], nrm[ids])
    hs.close()


def test_boost_archive_dmap_round_trip_and_documented_layout(tmp_path):
    """depthmap/*.dmap / normalmap/*.dmap of the fork (MVS::SaveDepthMap / SaveNormalMap, DepthMap.cpp:2368-2393): round trip through the
    host reader, and the bytes re-derived independently here from the documented Boost binary_oarchive layout (zlib + struct). Byte parity
    with a real Boost build cannot be checked in this image (no Boost)."""
    import struct, zlib
    from hcmvs_b200 import host
    rng = np.random.default_rng(3)
    d = rng.uniform(0, 9, (7, 11)).astype(np.float32); d[2, 3] = 0
    n = rng.standard_normal((7, 11, 3)).astype(np.float32)
    fd, fn = tmp_path / "depth0000.dmap", tmp_path / "normal0000.dmap"
    host.save_depthmap(fd, d); host.save_normalmap(fn, n)
    assert np.array_equal(host.load_depthmap(fd), d) and np.array_equal(host.load_normalmap(fn), n)
    head = struct.pack("<Q", 22) + b"serialization::archive" + struct.pack("<HBBBBi", 17, 4, 8, 4, 8, 1)
    cls = b"\x00" + struct.pack("<I", 0)                                    # tracking_type + version_type of a class seen for the first time
    want_d = head + cls * 3 + struct.pack("<ii", 11, 7) + d.tobytes()           # TImage : TDMatrix : cv::Mat_<float>, cols, rows, one raw block
    want_n = head + cls * 3 + struct.pack("<ii", 11, 7) + cls * 2 + n.tobytes() # ... elements one by one; the first carries TPoint3 / cv::Point3_
    raw_d, raw_n = fd.read_bytes(), fn.read_bytes()
    assert raw_d[:2] == b"\x78\x01"                                          # zlib stream, best_speed
    assert zlib.decompress(raw_d) == want_d and zlib.decompress(raw_n) == want_n
    # a file written by another zlib level / implementation still loads; garbage and truncated files are refused
    (tmp_path / "other.dmap").write_bytes(zlib.compress(want_d, 9))
    assert np.array_equal(host.load_depthmap(tmp_path / "other.dmap"), d)
    (tmp_path / "bad.dmap").write_bytes(raw_d[:len(raw_d) // 2])
    with pytest.raises(RuntimeError):
        host.load_depthmap(tmp_path / "bad.dmap")
    (tmp_path / "bad2.dmap").write_bytes(zlib.compress(b"x" * 64))
    with pytest.raises(RuntimeError):
        host.load_depthmap(tmp_path / "bad2.dmap")
