"""MVSI project files and the raw .dmap header, pinned against the REFERENCE's own code.

oracle/_ref/mvsi_ref_tool is compiled (oracle/Makefile) from /root/reference/frame_main/libs/MVS/Interface.h where it lies —
MVS::Interface, ARCHIVE::SerializeSave/SerializeLoad, Platform::GetFullK/GetPose, HeaderDepthDataRaw. These CPU tests make the
product's reader/writer (hc-mvs_b200/host/mvsi.cpp, densify.cpp) exchange files with it: byte-identical writes, field-identical
reads, the same absolute cameras. Skipped when the tool is not built (no /root/reference at build time)."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "oracle", "_ref", "mvsi_ref_tool")
needs_ref = pytest.mark.skipif(not os.path.exists(TOOL), reason="oracle/_ref/mvsi_ref_tool not built (reference tree absent)")
NO_ID = 0xFFFFFFFF


# ---------------------------------------------------------------- the neutral "flat" dump the tool reads / writes
def _s(b):
    return struct.pack("<I", len(b)) + b


def pack_flat(sc):
    o = [struct.pack("<I", len(sc["platforms"]))]
    for p in sc["platforms"]:
        o += [_s(p["name"]), struct.pack("<I", len(p["cameras"]))]
        for c in p["cameras"]:
            o += [_s(c["name"]), _s(c["band"]), struct.pack("<II", c["w"], c["h"]), np.asarray(c["K"], "<f8").tobytes(),
                  np.asarray(c["R"], "<f8").tobytes(), np.asarray(c["C"], "<f8").tobytes()]
        o.append(struct.pack("<I", len(p["poses"])))
        for q in p["poses"]:
            o += [np.asarray(q["R"], "<f8").tobytes(), np.asarray(q["C"], "<f8").tobytes()]
    o.append(struct.pack("<I", len(sc["images"])))
    for im in sc["images"]:
        o += [_s(im["name"]), _s(im["mask"]), struct.pack("<IIII", im["platform"], im["camera"], im["pose"], im["id"])]
    o.append(struct.pack("<I", len(sc["vertices"])))
    for X, views in sc["vertices"]:
        o += [np.asarray(X, "<f4").tobytes(), struct.pack("<I", len(views))]
        o += [struct.pack("<If", i, c) for i, c in views]
    o += [struct.pack("<I", len(sc["normals"])), np.asarray(sc["normals"], "<f4").tobytes()]
    o += [struct.pack("<I", len(sc["colors"])), np.asarray(sc["colors"], np.uint8).tobytes()]
    o.append(struct.pack("<I", len(sc["lines"])))
    for a, b, views in sc["lines"]:
        o += [np.asarray(a, "<f4").tobytes(), np.asarray(b, "<f4").tobytes(), struct.pack("<I", len(views))]
        o += [struct.pack("<If", i, c) for i, c in views]
    o += [struct.pack("<I", 0), struct.pack("<I", 0)]  # linesNormal, linesColor
    o.append(np.asarray(sc["transform"], "<f8").tobytes())
    return b"".join(o)


def _rot(rng):
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    return q * np.sign(np.linalg.det(q))


def make_scene(rng, n_img=5, n_pts=40, with_resolution=True, weights=True, image_size=(64, 48)):
    w, h = image_size
    cams = []
    for k in range(2):
        f = 70.0 + 13.7 * k
        K = np.array([[f, 0.0, w * 0.5 - 0.3 * k], [0, f * 1.01, h * 0.5 + 0.2], [0, 0, 1]])
        if not with_resolution:
            K = K.copy(); K[:2] /= max(w, h)
        cams.append(dict(name=b"cam%d" % k, band=b"RGB" if k else b"", w=w if with_resolution else 0, h=h if with_resolution else 0,
                         K=K, R=_rot(rng) if k else np.eye(3), C=rng.standard_normal(3) * 0.01 * k))
    poses = [dict(R=_rot(rng), C=rng.standard_normal(3)) for _ in range(n_img)]
    images = [dict(name=b"images/%05d.png" % i, mask=b"", platform=0, camera=i % 2, pose=i, id=i) for i in range(n_img)]
    images.append(dict(name=b"images/uncalibrated.png", mask=b"m.png", platform=NO_ID, camera=NO_ID, pose=NO_ID, id=NO_ID))
    verts = []
    for _ in range(n_pts):
        ids = rng.permutation(n_img)[: rng.integers(2, n_img + 1)]          # unsorted on purpose: the loader sorts by image id
        verts.append((rng.standard_normal(3).astype(np.float32), [(int(i), float(np.float32(rng.uniform(0.1, 1))) if weights else 0.0) for i in ids]))
    return dict(platforms=[dict(name=b"rig", cameras=cams, poses=poses)], images=images, vertices=verts,
                normals=rng.standard_normal((n_pts, 3)).astype(np.float32), colors=rng.integers(0, 256, (n_pts, 3)).astype(np.uint8),
                lines=[(np.zeros(3), np.ones(3), [(0, 0.5), (1, 0.25)])], transform=np.arange(16, dtype=np.float64).reshape(4, 4) / 7)


def tool(*args):
    return subprocess.run([TOOL, *map(str, args)], check=True, capture_output=True, text=True).stdout


# ---------------------------------------------------------------- tests
@needs_ref
@pytest.mark.parametrize("version", [5, 4, 3, 2, 1, 0])
def test_reader_matches_reference_writer(built, tmp_path, version):
    """.mvs written by the reference's SerializeSave (every stream version) -> the product's LoadInterface."""
    from hcmvs_b200 import host
    rng = np.random.default_rng(version)
    sc = make_scene(rng, with_resolution=version > 0)  # version 0 has no width/height fields
    (tmp_path / "flat.bin").write_bytes(pack_flat(sc))
    mvs = tmp_path / "scene.mvs"
    tool("from-flat", tmp_path / "flat.bin", mvs, version)
    if version == 0:
        # header-less first format: cameras carry no resolution, K is normalised, the size comes from the image header (Scene.cpp:155-158)
        (tmp_path / "images").mkdir()
        for im in sc["images"][:-1]:
            with open(tmp_path / im["name"].decode(), "wb") as f:
                f.write(b"P6\n64 48\n255\n" + bytes(64 * 48 * 3))
        hs = host.HostScene.load_mvs(mvs, load_images=False)
        for i, im in enumerate(sc["images"][:-1]):
            info, cam = hs.image_info(i), sc["platforms"][0]["cameras"][im["camera"]]
            assert (info["width"], info["height"]) == (64, 48)
            assert np.array_equal(info["K"][[0, 2, 4, 5]], (cam["K"] * np.float32(64.0)).ravel()[[0, 2, 4, 5]])
        return
    hs = host.HostScene.load_mvs(mvs, load_images=False)
    assert hs.num_images() == len(sc["images"])
    # absolute cameras: Interface::Platform::GetFullK / Interface::GetPose of the reference vs Scene::LoadInterface + UpdateCamera
    tool("cams", mvs, tmp_path / "cams.bin")
    rec = np.frombuffer((tmp_path / "cams.bin").read_bytes(), dtype=np.dtype([("i", "<u4"), ("K", "<f8", 9), ("R", "<f8", 9), ("C", "<f8", 3)]))
    assert len(rec) == len(sc["images"]) - 1
    for r in rec:
        info = hs.image_info(int(r["i"]))
        assert info["calibrated"] and (info["width"], info["height"]) == (64, 48)
        assert np.array_equal(info["R"], r["R"]) and np.array_equal(info["C"], r["C"])     # same f64 products, same order
        # K: the reference path normalises by max(w,h) on load and scales back (Scene.cpp:80-88, Camera.h:167-180): <= 1 ulp from GetFullK,
        # and the skew entry is dropped
        assert np.allclose(info["K"][[0, 2, 4, 5]], r["K"][[0, 2, 4, 5]], rtol=3e-16, atol=0) and info["K"][1] == 0 and info["K"][8] == 1
    last = hs.image_info(len(sc["images"]) - 1)
    assert not last["calibrated"] and last["name"].endswith("images/uncalibrated.png")
    assert hs.image_info(2)["id"] == (2 if version > 2 else 2)  # ID defaults to the index before v3
    # sparse points: views sorted by image id, weights follow
    xyz, off, ids, wts = hs.sparse()
    assert len(xyz) == len(sc["vertices"])
    for k, (X, views) in enumerate(sc["vertices"]):
        assert np.array_equal(xyz[k], X)
        want = sorted(views)
        assert list(ids[off[k]:off[k + 1]]) == [i for i, _ in want]
        assert np.array_equal(wts[off[k]:off[k + 1]], np.array([c for _, c in want], np.float32))


@needs_ref
@pytest.mark.parametrize("version", [5, 3, 1])
def test_writer_is_byte_identical_to_reference(built, tmp_path, version):
    """Load a reference-written project, write it back with the product's SaveMVSI path, and compare with what the reference
    writes for the same content: read our file with the reference's SerializeLoad, re-save it with SerializeSave -> same bytes."""
    from hcmvs_b200 import host
    rng = np.random.default_rng(10 + version)
    sc = make_scene(rng)
    (tmp_path / "flat.bin").write_bytes(pack_flat(sc))
    tool("from-flat", tmp_path / "flat.bin", tmp_path / "ref.mvs", version)
    hs = host.HostScene.load_mvs(tmp_path / "ref.mvs", load_images=False)
    ours = tmp_path / "ours.mvs"
    hs.save_mvs(ours, version=version)
    assert tool("to-flat", ours, tmp_path / "ours.flat").strip() == str(version)       # the reference parses our file ...
    tool("from-flat", tmp_path / "ours.flat", tmp_path / "ours_by_ref.mvs", version)    # ... and writes the same content itself
    assert ours.read_bytes() == (tmp_path / "ours_by_ref.mvs").read_bytes()
    # and the content survives: cameras and points of a second load are identical
    hs2 = host.HostScene.load_mvs(ours, load_images=False)
    for i in range(hs.num_images()):
        a, b = hs.image_info(i), hs2.image_info(i)
        assert a["calibrated"] == b["calibrated"]
        if a["calibrated"]:
            assert np.array_equal(a["R"], b["R"]) and np.array_equal(a["C"], b["C"]) and np.allclose(a["K"], b["K"], rtol=4e-16, atol=0)
    for x, y in zip(hs.sparse(), hs2.sparse()):
        assert np.array_equal(x, y)


@needs_ref
def test_dmap_header_matches_reference_struct(built, tmp_path):
    """The raw 'DR' header: layout facts printed from the reference's struct, our file read INTO that struct, and a header
    written FROM that struct parsed by our reader."""
    from hcmvs_b200 import host
    size, magic, has_d, has_n, has_c, *offs = map(int, tool("dmap-layout").split())
    assert (size, magic, has_d, has_n, has_c) == (28, 0x5244, 1, 2, 4) and offs == [2, 4, 8, 12, 16, 20, 24]
    rng = np.random.default_rng(3)
    h, w = 9, 14
    d = rng.uniform(1, 5, (h, w)).astype(np.float32); n = rng.standard_normal((h, w, 3)).astype(np.float32); c = rng.uniform(0, 1, (h, w)).astype(np.float32)
    K = np.array([100, 0, 8, 0, 100, 6, 0, 0, 1.0]); R = np.eye(3).ravel(); Cc = np.array([1.0, 2, 3])
    path = tmp_path / "depth0001.dmap"
    host.write_dmap(str(path), "00001.png", [1, 0, 2], (w + 2, h + 1), K, R, Cc, 0.75, 8.5, d, n, c)
    ok, typ, iw, ih, dw, dh, dmin, dmax = tool("dmap-header", path).split()
    assert (int(ok), int(typ), int(iw), int(ih), int(dw), int(dh), float(dmin), float(dmax)) == (1, 7, w + 2, h + 1, w, h, 0.75, 8.5)
    # reference-written header + our body layout -> our reader
    ref = tmp_path / "ref.dmap"
    tool("dmap-write-header", ref, 5, w, h, w, h, 0.5, 4.0)
    body = struct.pack("<H", 3) + b"a.b" + struct.pack("<II", 1, 7) + K.astype("<f8").tobytes() + R.astype("<f8").tobytes() + Cc.astype("<f8").tobytes()
    with open(ref, "ab") as f:
        f.write(body + d.tobytes() + c.tobytes())
    back = host.read_dmap(str(ref))
    assert back["normal"] is None and np.array_equal(back["depth"], d) and np.array_equal(back["conf"], c)
    assert list(back["ids"]) == [7] and (back["dmin"], back["dmax"]) == (0.5, 4.0)


@needs_ref
def test_scene_with_images_and_selection(built, tmp_path):
    """A synthetic scene written as a real project (reference-written .mvs + PNG/BMP/PPM image files) loads into the same host Scene
    as the in-memory path: same pixels, same gray images, same cameras (<= 1 ulp in K), bit-identical neighbour-view selection."""
    cv2 = pytest.importorskip("cv2")
    import common
    from hcmvs_b200 import api, host
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.25)
    hs0 = host.HostScene.from_synth(syn, imgs)
    (tmp_path / "images").mkdir()
    exts = ["png", "bmp", "ppm"]
    cams, poses, images = [], [], []
    for i in range(syn.n_views):
        h, w = imgs[i].shape[:2]
        name = f"images/{i:05d}.{exts[i % 3]}"
        if exts[i % 3] == "ppm":
            with open(tmp_path / name, "wb") as f:
                f.write(b"P6\n# synthetic\n%d %d\n255\n" % (w, h) + imgs[i][:, :, ::-1].tobytes())
        else:
            assert cv2.imwrite(str(tmp_path / name), imgs[i])
        assert np.array_equal(host.load_image(tmp_path / name), imgs[i])
        cams.append(dict(name=b"c", band=b"", w=w, h=h, K=np.asarray(syn.K[i]).reshape(3, 3), R=np.eye(3), C=np.zeros(3)))
        poses.append(dict(R=np.asarray(syn.R[i]).reshape(3, 3), C=np.asarray(syn.Cc[i])))
        images.append(dict(name=name.encode(), mask=b"", platform=0, camera=i, pose=i, id=i))
    off = np.asarray(syn.sparse_off)
    verts = [(np.asarray(syn.sparse_xyz[k], np.float32), [(int(v), 0.0) for v in syn.sparse_views[off[k]:off[k + 1]]]) for k in range(len(off) - 1)]
    sc = dict(platforms=[dict(name=b"p", cameras=cams, poses=poses)], images=images, vertices=verts, normals=np.zeros((0, 3)),
              colors=np.zeros((0, 3), np.uint8), lines=[], transform=np.eye(4))
    (tmp_path / "flat.bin").write_bytes(pack_flat(sc))
    tool("from-flat", tmp_path / "flat.bin", tmp_path / "scene.mvs", 5)
    hs = host.HostScene.load_mvs(tmp_path / "scene.mvs")
    assert hs.num_images() == syn.n_views
    P = api.default_params(nMinViewsTrustPoint=1)
    for i in range(syn.n_views):
        assert np.array_equal(hs.image_bgr(i), imgs[i]) and np.array_equal(hs.gray(i), hs0.gray(i))
        info = hs.image_info(i)
        assert np.array_equal(info["R"], np.asarray(syn.R[i]).ravel()) and np.array_equal(info["C"], np.asarray(syn.Cc[i]))
        assert np.allclose(info["K"], np.asarray(syn.K[i]).ravel(), rtol=3e-16, atol=0)
        assert (hs.select_views(P, i) > 0) == (hs0.select_views(P, i) > 0) == bool(ok[i])
        for which in (0, 1):
            a, b = hs.neighbors(i, which), hs0.neighbors(i, which)
            assert np.array_equal(a["ids"], b["ids"]) and np.array_equal(a["points"], b["points"])
            for k in ("scale", "angle", "area", "score"):     # K differs by <= 1 ulp after the normalise / scale-back round trip
                assert np.allclose(a[k], b[k], rtol=1e-5, atol=0), (i, which, k)


def test_image_decoders_reject_what_they_cannot_read(built, tmp_path):
    from hcmvs_b200 import host
    (tmp_path / "x.jpg").write_bytes(b"\xff\xd8\xff\xe0" + b"\0" * 64)
    with pytest.raises(RuntimeError):
        host.load_image(tmp_path / "x.jpg")
    with pytest.raises(RuntimeError):
        host.HostScene.load_mvs(tmp_path / "missing.mvs")


def test_resize_area_bgr_matches_opencv(built):
    """Image::ResizeImage = cv::resize(INTER_AREA) on the 8-bit colour image: the host restatement is bit-equal to cv2 for the
    2x2 fast path ((sum+2)>>2), other integer factors (round-half-even of sum/area) and the general DecimateAlpha path."""
    cv2 = pytest.importorskip("cv2")
    from hcmvs_b200 import host
    rng = np.random.default_rng(0)
    for sw, sh, dw, dh in [(320, 240, 160, 120), (320, 240, 80, 60), (321, 241, 160, 120), (200, 140, 67, 47), (384, 216, 48, 27),
                           (33, 17, 16, 8), (320, 240, 213, 160), (64, 48, 64, 48)]:
        for img in (rng.integers(0, 256, (sh, sw, 3)).astype(np.uint8), cv2.GaussianBlur(rng.integers(0, 256, (sh, sw, 3)).astype(np.uint8), (0, 0), 2.0)):
            assert np.array_equal(host.resize_area_bgr(img, (dw, dh)), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA)), (sw, sh, dw, dh)
    with pytest.raises(ValueError):
        host.resize_area_bgr(np.zeros((10, 10, 3), np.uint8), (20, 20))


@needs_ref
def test_resolution_level_reload(built, tmp_path):
    """--resolution-level: Scene::ComputeDepthMaps reloads every image at max(w,h) >> level (computeMaxResolution, ResizeImage) and
    rebuilds its camera from the normalised intrinsics (UpdateCamera)."""
    cv2 = pytest.importorskip("cv2")
    from hcmvs_b200 import host
    rng = np.random.default_rng(5)
    w, h = 322, 241                      # odd sizes: the integer size arithmetic and the general area path
    sc = make_scene(rng, n_img=3, image_size=(w, h))
    sc["images"] = sc["images"][:-1]
    (tmp_path / "images").mkdir()
    imgs = []
    for im in sc["images"]:
        px = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3)).astype(np.uint8), (0, 0), 1.0)
        imgs.append(px)
        assert cv2.imwrite(str(tmp_path / im["name"].decode()), px)
    (tmp_path / "flat.bin").write_bytes(pack_flat(sc))
    tool("from-flat", tmp_path / "flat.bin", tmp_path / "scene.mvs", 5)
    for level, min_res, want_max in ((1, 100, 161), (2, 50, 80), (3, 100, 161), (0, 100, 322)):
        # level 3: 322 >> 3 = 40 < min 100 -> the level is lowered until the size is >= min: 322 >> 1 = 161
        hs = host.HostScene.load_mvs(tmp_path / "scene.mvs")
        hs.reload_images(level, min_res, 3200)
        for i, im in enumerate(sc["images"]):
            info = hs.image_info(i)
            nh = h * want_max // w if want_max < w else h
            assert (info["width"], info["height"]) == (want_max, nh)
            want_px = imgs[i] if want_max == w else cv2.resize(imgs[i], (want_max, nh), interpolation=cv2.INTER_AREA)
            assert np.array_equal(hs.image_bgr(i), want_px)
            cam = sc["platforms"][0]["cameras"][im["camera"]]
            Kn = cam["K"] * (1.0 / float(np.float32(max(w, h))))             # Scene.cpp:80-88
            Kw = Kn * float(np.float32(max(want_max, nh)))                   # Camera.h:167-180
            assert np.array_equal(info["K"][[0, 2, 4, 5]], Kw.ravel()[[0, 2, 4, 5]])
        hs.close()


def test_corrupt_project_and_image_files_fail_cleanly(built, tmp_path):
    """Truncated / corrupted inputs return an error (the reference's bool-return convention) — no crash, no runaway allocation."""
    cv2 = pytest.importorskip("cv2")
    from hcmvs_b200 import host
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (24, 32, 3)).astype(np.uint8)
    hs = host.HostScene()
    K = np.array([30.0, 0, 16, 0, 30, 12, 0, 0, 1])
    for i in range(3):
        hs.add_image(K, np.eye(3).ravel(), np.array([i * 0.1, 0, 0]), img, name=str(tmp_path / f"{i}.png"))
        assert cv2.imwrite(str(tmp_path / f"{i}.png"), img)
    hs.set_sparse(rng.standard_normal((20, 3)).astype(np.float32), np.arange(0, 42, 2, dtype=np.int32), np.tile([0, 1], 20).astype(np.uint32))
    good = tmp_path / "good.mvs"
    hs.save_mvs(good)
    raw = good.read_bytes()
    assert host.HostScene.load_mvs(good).num_images() == 3
    for cut in (0, 3, 8, 13, 40, len(raw) // 2, len(raw) - 1):
        bad = tmp_path / "bad.mvs"
        bad.write_bytes(raw[:cut])
        with pytest.raises(RuntimeError):
            host.HostScene.load_mvs(bad)
    huge = bytearray(raw); huge[12:20] = (2 ** 62).to_bytes(8, "little")       # platform count far beyond the file size
    (tmp_path / "huge.mvs").write_bytes(bytes(huge))
    with pytest.raises(RuntimeError):
        host.HostScene.load_mvs(tmp_path / "huge.mvs")
    (tmp_path / "v9.mvs").write_bytes(raw[:4] + (9).to_bytes(4, "little") + raw[8:])  # version newer than MVSI_PROJECT_VER
    with pytest.raises(RuntimeError):
        host.HostScene.load_mvs(tmp_path / "v9.mvs")
    png = (tmp_path / "0.png").read_bytes()
    for name, data in (("trunc.png", png[: len(png) // 2]), ("flip.png", png[:60] + bytes([png[60] ^ 0xFF]) + png[61:]), ("tiny.bmp", b"BM" + bytes(20))):
        (tmp_path / name).write_bytes(data)
        try:
            out = host.load_image(tmp_path / name)                             # a flipped byte may still inflate: then the size must hold
            assert out.shape == img.shape
        except RuntimeError:
            pass
