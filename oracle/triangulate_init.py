"""TEST INFRASTRUCTURE ONLY — CPU restatement of MVS::TriangulatePoints2DepthMap (libs/MVS/DepthMap.cpp:1797-1936), the reference's
default depth-map initialisation, independent of the product code: the Delaunay triangulation comes from scipy (Qhull) instead of
the product's Bowyer-Watson (the reference uses CGAL; a Delaunay triangulation of points in general position is unique), the corner
depths and the 28.4 fixed-point rasteriser (libs/Common/Types.inl:2469-2606) are restated in numpy. Only tests/ may import this.

Quirks pinned here (q14-q16, continuing oracle/hcmvs_oracle.hpp):
 q14  CGAL's face order and the rotation of a face's vertices are unspecified; faces are canonical here: counter-clockwise, smallest
      vertex index first, sorted lexicographically, drawn in that order (the rotation only changes the last bits of a plane).
 q15  with bAddCorners the reference never clears the maps (DepthMap.cpp:1895-1898): pixels no face covers, or whose depth comes
      out <= 0, keep uninitialised memory. Defined as depth 0 / normal 0.
 q16  the reference is compiled without FMA contraction on x86-64; every f32 expression below is rounded operation by operation.
"""
import numpy as np

f32 = np.float32
FINV_ZERO = f32(1000000.0)  # INVZERO(float), libs/Common/Types.h:573
INV_ZERO = 1e14             # INVZERO(double), Types.h:555


def canonical_faces(simplices, xy):
    t = np.asarray(simplices, np.int64)
    a, b, c = xy[t[:, 0]], xy[t[:, 1]], xy[t[:, 2]]
    o = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
    t = np.where((o < 0)[:, None], t[:, [0, 2, 1]], t)
    k = np.argmin(t, axis=1)
    t = np.stack([t[np.arange(len(t)), (k + i) % 3] for i in range(3)], 1)
    return t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))]


def project_points(P, xyz):
    """Camera::ProjectPointP3<float> (Camera.h:276-282) then (x/z, y/z, z), DepthMap.cpp:1803-1804."""
    P = np.asarray(P, np.float64).reshape(3, 4)
    X = np.asarray(xyz, np.float32).astype(np.float64)
    q = [(((P[r, 0] * X[:, 0] + P[r, 1] * X[:, 1]) + P[r, 2] * X[:, 2]) + P[r, 3]).astype(np.float32) for r in range(3)]
    return np.stack([(q[0] / q[2]).astype(np.float64), (q[1] / q[2]).astype(np.float64), q[2].astype(np.float64)], 1)


def _i2c(v, K):
    fx, fy, cx, cy = K[0], K[4], K[2], K[5]
    return np.array([(v[0] - cx) * v[2] / fx, (v[1] - cy) * v[2] / fy, v[2]])


def triangulate(P, K, width, height, xyz, avg_depth, add_corners=True):
    """TriangulatePointsDelaunay (DepthMap.cpp:1797-1876) -> vertices (n,3) f64, faces (m,3), dMin, dMax (raw f32 bounds)."""
    from scipy.spatial import Delaunay
    K = np.asarray(K, np.float64).ravel()
    v = project_points(P, xyz)
    d_min, d_max = f32(v[:, 2].min()), f32(v[:, 2].max())
    n_pts = len(v)
    if add_corners:
        v = np.vstack([v, [[0, 0, avg_depth], [width, 0, avg_depth], [0, height, avg_depth], [width, height, avg_depth]]])
    # CGAL's insert() keeps the first of two points with equal (x, y): triangulate the unique ones, indices of the first occurrences
    _, first = np.unique(v[:, :2], axis=0, return_index=True)
    first = np.sort(first)
    faces = canonical_faces(first[Delaunay(v[first, :2]).simplices], v[:, :2])
    if not add_corners:
        return v, faces, d_min, d_max
    corners = list(range(n_pts, n_pts + 4))
    edge = {}
    for fi, (a, b, c) in enumerate(faces):
        for k, (e0, e1) in enumerate(((b, c), (c, a), (a, b))):       # the edge opposite vertex slot k
            edge[(e0, e1)] = fi
    for vc in corners:
        A = v[vc].copy()
        ray = _i2c(A, K)
        nrm = np.sqrt(ray @ ray)
        ray = ray * (1.0 / nrm if nrm else 0.0)
        top = []                                                     # (score, depth), score descending — cList::StoreTop<3>
        for fi, face in enumerate(faces):
            if vc not in face:
                continue
            k = list(face).index(vc)
            e0, e1 = face[(k + 1) % 3], face[(k + 2) % 3]
            fc = edge.get((e1, e0))                                   # the face BEHIND this one, across the edge opposite the corner
            if fc is None or any(c in faces[fc] for c in corners):
                continue
            B = v[faces[fc]]
            c0, c1, c2 = (_i2c(b, K) for b in B)
            N = np.cross(c1 - c0, c2 - c0)
            nn = np.sqrt(N @ N)
            if nn > 0:
                N = N / nn
            Vd = N @ ray
            t_hit = INV_ZERO if Vd == 0 else (N @ c0) / Vd            # Ray3d(0, dir).IntersectsDist(plane) = -D / (n.dir), D = -n.c0
            z = ray[2] * t_hit
            if not z > 0:
                continue
            pos_b = (B[0, :2] + B[1, :2] + B[2, :2]) / 3.0
            fd = f32(np.sqrt(((pos_b - A[:2]) ** 2).sum()))
            score = FINV_ZERO if fd == 0 else f32(1) / fd
            depth = min(max(f32(z), d_min), d_max)
            pos = 0
            while pos < len(top) and top[pos][0] > score:
                pos += 1
            if pos < len(top):
                if len(top) >= 3:
                    top.pop()
                top.insert(pos, (score, depth))
            elif len(top) < 3:
                top.append((score, depth))
        if len(top) != 3:
            continue
        s = [t[0] for t in top]; d = [t[1] for t in top]
        inv = f32(1) / f32(f32(s[0] + s[1]) + s[2])
        w = [f32(x * inv) for x in s]
        v[vc, 2] = float(f32(f32(f32(d[0] * w[0]) + f32(d[1] * w[1])) + f32(d[2] * w[2])))
    return v, faces, d_min, d_max


def _round16(x):
    return np.floor(f32(16) * f32(x) + f32(0.5)).astype(np.int64)     # ROUND2INT(T(16) * v)


def rasterize(vertices, faces, K, width, height):
    """The face loop of TriangulatePoints2DepthMap (DepthMap.cpp:1914-1935) with TImage::RasterizeTriangle -> depth (H,W), normal (H,W,3)."""
    K = np.asarray(K, np.float64).ravel()
    fx, fy, cx, cy = K[0], K[4], K[2], K[5]
    depth = np.zeros((height, width), np.float32)
    normal = np.zeros((height, width, 3), np.float32)
    V = np.asarray(vertices, np.float64)
    for face in faces:
        I = V[face].astype(np.float32)                                                             # Point3f i0, i1, i2
        c = np.stack([((I[:, 0].astype(np.float64) - cx) * I[:, 2].astype(np.float64) / fx).astype(np.float32),
                      ((I[:, 1].astype(np.float64) - cy) * I[:, 2].astype(np.float64) / fy).astype(np.float32), I[:, 2]], 1)
        e1, e2 = c[1] - c[0], c[2] - c[0]
        n = np.array([f32(e2[1] * e1[2]) - f32(e2[2] * e1[1]), f32(e2[2] * e1[0]) - f32(e2[0] * e1[2]), f32(e2[0] * e1[1]) - f32(e2[1] * e1[0])], np.float32)
        nd = n.astype(np.float64)
        nv = np.sqrt((nd[0] * nd[0] + nd[1] * nd[1]) + nd[2] * nd[2])
        nrm = (nd * (1.0 / nv if nv else 0.0)).astype(np.float32)                                   # cv::normalize(Vec3f)
        d0 = f32(f32(f32(nrm[0] * c[0, 0]) + f32(nrm[1] * c[0, 1])) + f32(nrm[2] * c[0, 2]))
        plane = (nrm * (FINV_ZERO if d0 == 0 else f32(1) / d0)).astype(np.float32)
        X1, Y1, X2, Y2, X3, Y3 = (int(_round16(I[2, 0])), int(_round16(I[2, 1])), int(_round16(I[1, 0])), int(_round16(I[1, 1])),
                                  int(_round16(I[0, 0])), int(_round16(I[0, 1])))                   # RasterizeTriangle(i2, i1, i0)
        DX12, DX23, DX31, DY12, DY23, DY31 = X1 - X2, X2 - X3, X3 - X1, Y1 - Y2, Y2 - Y3, Y3 - Y1
        minx, maxx = (min(X1, X2, X3) + 0xF) >> 4, (max(X1, X2, X3) + 0xF) >> 4
        miny, maxy = (min(Y1, Y2, Y3) + 0xF) >> 4, (max(Y1, Y2, Y3) + 0xF) >> 4
        C1, C2, C3 = DY12 * X1 - DX12 * Y1, DY23 * X2 - DX23 * Y2, DY31 * X3 - DX31 * Y3
        C1 += DY12 < 0 or (DY12 == 0 and DX12 > 0)
        C2 += DY23 < 0 or (DY23 == 0 and DX23 > 0)
        C3 += DY31 < 0 or (DY31 == 0 and DX31 > 0)
        x0, x1, y0, y1 = max(minx, 0), min(maxx, width), max(miny, 0), min(maxy, height)          # depthMap.isInside(pt)
        if x0 >= x1 or y0 >= y1:
            continue
        xs, ys = np.arange(x0, x1, dtype=np.int64), np.arange(y0, y1, dtype=np.int64)
        fxp, fyp = (xs << 4)[None, :], (ys << 4)[:, None]
        inside = (C1 + DX12 * fyp - DY12 * fxp > 0) & (C2 + DX23 * fyp - DY23 * fxp > 0) & (C3 + DX31 * fyp - DY31 * fxp > 0)
        if not inside.any():
            continue
        Xc = ((xs.astype(np.float32).astype(np.float64) - cx) / fx).astype(np.float32)[None, :]      # TransformPointI2C(Point2f(pt))
        Yc = ((ys.astype(np.float32).astype(np.float64) - cy) / fy).astype(np.float32)[:, None]
        d = ((plane[0] * Xc).astype(np.float32) + (plane[1] * Yc).astype(np.float32)).astype(np.float32) + plane[2]
        with np.errstate(divide="ignore"):
            z = np.where(d == 0, FINV_ZERO, f32(1) / d).astype(np.float32)
        write = inside & (z > 0)
        sub_d, sub_n = depth[y0:y1, x0:x1], normal[y0:y1, x0:x1]
        sub_d[write] = z[write]
        sub_n[write] = nrm
    return depth, normal
