"""TEST INFRASTRUCTURE ONLY — numpy restatement of Scene::PointCloudFilter's visibility vote (libs/MVS/SceneDensify.cpp:4189-4320),
brute force over all point pairs. Only tests/ may import this.

For every point X and every view v listed for it: Collector::Init (:4221-4229) builds the cone (origin Cast<float>(C_v), direction
(X - origin)/|X - origin|, half-angle float(ComputeFOV(0)/width), height 1.02*|X - origin|); every cloud point p with
TConeIntersect::Classify(p) == VISIBLE (libs/Common/Ray.inl:985-1002) and !IsDepthSimilar(distance, t, 0.01) gets
+#views(p) when t > distance, else -#views(X) (:4233-4246).

Quirk q17: Eigen's fixed-size reductions associate a 3-vector as x0 + (x1 + x2) (redux_novec_unroller halves); COS(angle) is glibc's
cosf, taken here (and in the kernel's host code) as the f64 cosine rounded to f32. Every f32 operation is rounded on its own (no FMA)."""
import numpy as np

f32 = np.float32


def _sq3(a, b, c):
    return (a * a).astype(f32) + ((b * b).astype(f32) + (c * c).astype(f32)).astype(f32)


def visibility(cams, points, view_offsets, views):
    """cams[v] = dict(C=(3,) f64, K=(9,) f64, width=int); CSR view lists -> int32 visibility per point."""
    pts = np.asarray(points, f32)
    nv = np.diff(np.asarray(view_offsets).astype(np.int64)).astype(np.int32)
    vis = np.zeros(len(pts), np.int64)
    cone = {}
    for v, cam in cams.items():
        angle = f32(2.0 * np.arctan(float(cam["width"]) / (cam["K"][0] * 2.0)) / float(cam["width"]))
        ca = f32(np.cos(np.float64(angle)))
        cone[v] = (np.asarray(cam["C"], np.float64).astype(f32), f32(ca * ca))
    for i in range(len(pts)):
        for v in views[view_offsets[i]:view_offsets[i + 1]]:
            o, cos2 = cone[int(v)]
            D = (pts[i] - o).astype(f32)
            dist = f32(np.sqrt(_sq3(D[0:1], D[1:2], D[2:3])[0]))
            d = (D / dist).astype(f32)
            max_h = f32(dist * f32(1.02))
            E = (pts - o).astype(f32)
            t = ((d[0] * E[:, 0]).astype(f32) + ((d[1] * E[:, 1]).astype(f32) + (d[2] * E[:, 2]).astype(f32)).astype(f32)).astype(f32)
            ok = ~(np.abs(t) < f32(0.0001)) & ~(t < 0) & ~(t > max_h)
            ok &= (t * t).astype(f32) > (cos2 * _sq3(E[:, 0], E[:, 1], E[:, 2])).astype(f32)
            ok &= ~((np.abs((dist - t).astype(f32)) / dist).astype(f32) < f32(0.01))
            behind = ok & (t > dist)
            vis[behind] += nv[behind]
            vis[ok & ~behind] -= nv[i]
    return vis.astype(np.int32)
