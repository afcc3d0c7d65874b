"""TEST INFRASTRUCTURE ONLY — numpy restatement of MVS::EstimatePointColors (libs/MVS/DepthMap.cpp:2125-2161). Only tests/ may import this.

Per point: among its views that hold an image, the one with the smallest Camera::PointDepth (f64, strict '>' so the first minimum
wins); Camera::ProjectPointP<float> (Camera.h:283-288); TImage::isInsideWithBorder<float,1> (Common/Types.h:1632-1635) else white;
TImage<Pixel8U>::sample (Common/Types.inl:2248-2258) with TPixel<uint8_t>::operator*(float) / operator+ (Common/Types.h:1930-1936):
every product and every sum of the bilinear kernel is truncated to uint8 before the next operation."""
import numpy as np

f32 = np.float32


def _u8mul(c, v):
    return (f32(v) * c.astype(np.float32)).astype(np.float32).astype(np.int32).astype(np.uint8)      # (TYPE)(v*c), truncation


def estimate_point_colors(P_list, images, points, view_offsets, views):
    """P_list[v]: 3x4 f64 projection of view v; images[v]: (H, W, 3) u8 or None; CSR view lists -> (n, 3) u8."""
    pts = np.asarray(points, np.float32)
    out = np.full((len(pts), 3), 255, np.uint8)
    for i, X in enumerate(pts):
        Xd = X.astype(np.float64)
        best, bv = float(np.float32(3.402823466e38)), -1
        for v in views[view_offsets[i]:view_offsets[i + 1]]:
            if images[v] is None:
                continue
            P = P_list[v]
            dist = ((P[2, 0] * Xd[0] + P[2, 1] * Xd[1]) + P[2, 2] * Xd[2]) + P[2, 3]
            if best > dist:
                best, bv = dist, int(v)
        if bv < 0:
            continue
        P, img = P_list[bv], images[bv]
        q = [f32(((P[r, 0] * Xd[0] + P[r, 1] * Xd[1]) + P[r, 2] * Xd[2]) + P[r, 3]) for r in range(3)]
        inv = f32(1000000.0) if q[2] == 0 else f32(1) / q[2]
        px, py = f32(q[0] * inv), f32(q[1] * inv)
        h, w = img.shape[:2]
        if not (px >= 1 and py >= 1 and px <= f32(w - 2) and py <= f32(h - 2)):
            continue
        lx, ly = int(px), int(py)
        x = f32(px - f32(lx)); x1 = f32(f32(1) - x); y = f32(py - f32(ly)); y1 = f32(f32(1) - y)
        top = (_u8mul(img[ly, lx], x1) + _u8mul(img[ly, lx + 1], x)).astype(np.uint8)
        bot = (_u8mul(img[ly + 1, lx], x1) + _u8mul(img[ly + 1, lx + 1], x)).astype(np.uint8)
        out[i] = (_u8mul(top, y1) + _u8mul(bot, y)).astype(np.uint8)
    return out
