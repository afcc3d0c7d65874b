// Sparse-point triangulation for the depth-map initialisation — the host half of MVS::TriangulatePoints2DepthMap
// (libs/MVS/DepthMap.cpp:1797-1936, the reference's DEFAULT init: nMinViewsTrustPoint >= 2 -> DepthMapsData::InitDepthMap,
// SceneDensify.cpp:514-525). The reference triangulates on the host too (CGAL::Delaunay_triangulation_2 over the projected points);
// CGAL is a third-party dependency absent from /root/reference, so the Delaunay triangulation is built here from its definition
// (empty-circumcircle property, unique for points in general position) with an incremental Bowyer-Watson insertion over ghost
// triangles; tests cross-check the triangle set against scipy.spatial.Delaunay (Qhull). The per-pixel work — plane / ray
// intersection for every pixel of every triangle — runs on the GPU (hcmvs_init_depthmap_triangles, csrc/api.cu).
#include "densify.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace hcmvs_host {

namespace {
typedef long double LD;
constexpr int GHOST = -1;

struct Delaunay2D {
	struct Tri { int v[3]; int n[3]; bool alive; };
	const std::vector<double>& xy; // 2 per vertex
	std::vector<Tri> tris;
	std::vector<int> freeList;
	int last = -1; // a live real triangle to start the walk from
	// scratch
	std::vector<int> cavity, stack, byStart, mark;
	int stamp = 0;

	explicit Delaunay2D(const std::vector<double>& pts) : xy(pts) { byStart.assign(pts.size()/2+1, -1); }

	LD Orient(int a, int b, int c) const { // > 0: c to the left of a -> b
		const LD ax = xy[a*2], ay = xy[a*2+1], bx = xy[b*2], by = xy[b*2+1], cx = xy[c*2], cy = xy[c*2+1];
		return (ax-cx)*(by-cy)-(ay-cy)*(bx-cx);
	}
	bool InCircle(const Tri& t, int p) const {
		int g = -1;
		for (int k=0; k<3; ++k) if (t.v[k] == GHOST) g = k;
		if (g >= 0) {
			// ghost triangle (a, b, ghost) in cyclic order: p conflicts with it when it lies beyond the hull edge a -> b, or on the open edge
			const int a = t.v[(g+1)%3], b = t.v[(g+2)%3];
			const LD o = Orient(a, b, p);
			if (o > 0) return true;
			if (o < 0) return false;
			const LD ax = xy[a*2], ay = xy[a*2+1], bx = xy[b*2], by = xy[b*2+1], px = xy[p*2], py = xy[p*2+1];
			const LD d = (px-ax)*(bx-ax)+(py-ay)*(by-ay), l = (bx-ax)*(bx-ax)+(by-ay)*(by-ay);
			return d > 0 && d < l;
		}
		const LD px = xy[p*2], py = xy[p*2+1];
		const LD ax = xy[t.v[0]*2]-px, ay = xy[t.v[0]*2+1]-py, bx = xy[t.v[1]*2]-px, by = xy[t.v[1]*2+1]-py, cx = xy[t.v[2]*2]-px, cy = xy[t.v[2]*2+1]-py;
		const LD det = (ax*ax+ay*ay)*(bx*cy-cx*by)-(bx*bx+by*by)*(ax*cy-cx*ay)+(cx*cx+cy*cy)*(ax*by-bx*ay);
		return det > 0;
	}
	int NewTri(int a, int b, int c) {
		int id;
		if (!freeList.empty()) { id = freeList.back(); freeList.pop_back(); } else { id = (int)tris.size(); tris.emplace_back(); }
		Tri& t = tris[id]; t.v[0] = a; t.v[1] = b; t.v[2] = c; t.n[0] = t.n[1] = t.n[2] = -1; t.alive = true;
		return id;
	}
	bool IsGhost(int t) const { return tris[t].v[0] == GHOST || tris[t].v[1] == GHOST || tris[t].v[2] == GHOST; }

	// first triangle from three non-collinear vertices + its three ghosts
	void Init(int a, int b, int c) {
		if (Orient(a, b, c) < 0) std::swap(b, c);
		const int t = NewTri(a, b, c);
		const int g0 = NewTri(c, b, GHOST), g1 = NewTri(a, c, GHOST), g2 = NewTri(b, a, GHOST); // across the edges opposite a, b, c
		tris[t].n[0] = g0; tris[t].n[1] = g1; tris[t].n[2] = g2;
		tris[g0].n[2] = t; tris[g1].n[2] = t; tris[g2].n[2] = t;
		// ghosts around the ghost vertex: (c,b,G): edge (b,G) opposite c is shared with the ghost starting at b = g2 (b,a,G); edge (G,c) opposite b with g1 (a,c,G)
		tris[g0].n[0] = g2; tris[g0].n[1] = g1;
		tris[g1].n[0] = g0; tris[g1].n[1] = g2;
		tris[g2].n[0] = g1; tris[g2].n[1] = g0;
		last = t;
	}

	// returns false when p duplicates an existing vertex (CGAL's insert() then returns the old vertex and nothing changes)
	bool Insert(int p) {
		// visibility walk from the last real triangle
		int t = last;
		for (size_t guard=0; guard<tris.size()*4+16; ++guard) {
			const Tri& T = tris[t];
			if (IsGhost(t)) break;
			bool moved = false;
			for (int k=0; k<3 && !moved; ++k) {
				const int e0 = T.v[(k+1)%3], e1 = T.v[(k+2)%3];
				if (Orient(e0, e1, p) < 0) { t = T.n[k]; moved = true; }
			}
			if (!moved) break;
		}
		if (!IsGhost(t)) for (int k=0; k<3; ++k) { const int v = tris[t].v[k]; if (xy[v*2] == xy[p*2] && xy[v*2+1] == xy[p*2+1]) return false; }
		// cavity: every triangle whose circumcircle holds p, grown from the located one
		++stamp; if (mark.size() < tris.size()) mark.resize(tris.size()*2+16, 0);
		cavity.clear(); stack.clear();
		stack.push_back(t); mark[t] = stamp;
		while (!stack.empty()) {
			const int c = stack.back(); stack.pop_back();
			cavity.push_back(c);
			for (int k=0; k<3; ++k) {
				const int nb = tris[c].n[k];
				if (nb < 0 || mark[nb] == stamp) continue;
				if (InCircle(tris[nb], p)) {
					// a duplicate can hide behind a ghost start: check the vertices of every real triangle entering the cavity
					if (!IsGhost(nb)) for (int q=0; q<3; ++q) { const int v = tris[nb].v[q]; if (xy[v*2] == xy[p*2] && xy[v*2+1] == xy[p*2+1]) return false; }
					mark[nb] = stamp; stack.push_back(nb);
				}
			}
		}
		// re-triangulate: one new triangle (e0, e1, p) per boundary edge of the cavity
		struct Edge { int e0, e1, outer; };
		std::vector<Edge> border;
		for (int c: cavity) for (int k=0; k<3; ++k) {
			const int nb = tris[c].n[k];
			if (nb >= 0 && mark[nb] == stamp) continue;
			border.push_back({tris[c].v[(k+1)%3], tris[c].v[(k+2)%3], nb});
		}
		for (int c: cavity) { tris[c].alive = false; freeList.push_back(c); }
		std::vector<int> created; created.reserve(border.size());
		for (const Edge& e: border) {
			const int id = NewTri(e.e0, e.e1, p);
			if (mark.size() <= (size_t)id) mark.resize((size_t)id*2+16, 0);
			mark[id] = 0;
			tris[id].n[2] = e.outer;
			if (e.outer >= 0) { Tri& O = tris[e.outer]; for (int q=0; q<3; ++q) { const int a = O.v[(q+1)%3], b = O.v[(q+2)%3]; if (a == e.e1 && b == e.e0) O.n[q] = id; } }
			byStart[e.e0+1] = id;
			created.push_back(id);
		}
		for (int id: created) {
			const int nx = byStart[tris[id].v[1]+1]; // the fan triangle that starts where this one ends: shares the edge (e1, p)
			tris[id].n[0] = nx; tris[nx].n[1] = id;
		}
		for (int id: created) if (!IsGhost(id)) { last = id; break; }
		return true;
	}
};
} // namespace

// Delaunay triangulation of 2-D points (xy: 2 per point); duplicates of an earlier point are skipped. tris: 3 vertex indices per
// triangle, counter-clockwise in the (x right, y up) sense, each rotated so that its smallest index comes first, sorted; adjacency
// (optional): per triangle and vertex slot k the index of the triangle across the edge opposite vertex k, or -1 on the hull.
bool DelaunayTriangulate(const std::vector<double>& xy, std::vector<uint32_t>& tris, std::vector<int>* adjacency) {
	const int n = (int)(xy.size()/2);
	tris.clear(); if (adjacency) adjacency->clear();
	if (n < 3) return false;
	// insertion order: boustrophedon over a coarse grid (keeps the walk short); the result does not depend on it
	double lo[2] = {DBL_MAX, DBL_MAX}, hi[2] = {-DBL_MAX, -DBL_MAX};
	for (int i=0; i<n; ++i) for (int d=0; d<2; ++d) { lo[d] = std::min(lo[d], xy[i*2+d]); hi[d] = std::max(hi[d], xy[i*2+d]); }
	const int G = std::max(1, (int)std::sqrt((double)n/4.0));
	std::vector<std::pair<uint32_t, int>> order(n);
	for (int i=0; i<n; ++i) {
		const int gx = std::min(G-1, (int)((xy[i*2]-lo[0])/std::max(hi[0]-lo[0], 1e-300)*G)), gy = std::min(G-1, (int)((xy[i*2+1]-lo[1])/std::max(hi[1]-lo[1], 1e-300)*G));
		order[i] = {(uint32_t)(gy*G+((gy&1) ? G-1-gx : gx)), i};
	}
	std::stable_sort(order.begin(), order.end());
	Delaunay2D D(xy);
	// seed: the first point of the order, the next distinct one, the next not collinear with them
	std::vector<char> used(n, 0);
	const int a = order[0].second; int b = -1, c = -1;
	for (int k=1; k<n && b < 0; ++k) { const int i = order[k].second; if (xy[i*2] != xy[a*2] || xy[i*2+1] != xy[a*2+1]) b = i; }
	if (b < 0) return false;
	for (int k=1; k<n && c < 0; ++k) { const int i = order[k].second; if (i != b && D.Orient(a, b, i) != 0) c = i; }
	if (c < 0) return false;
	D.Init(a, b, c); used[a] = used[b] = used[c] = 1;
	for (int k=0; k<n; ++k) { const int i = order[k].second; if (!used[i]) D.Insert(i); }
	// collect the finite faces
	std::vector<int> remap(D.tris.size(), -1);
	struct Face { uint32_t v[3]; int src; };
	std::vector<Face> faces;
	for (size_t t=0; t<D.tris.size(); ++t) {
		const Delaunay2D::Tri& T = D.tris[t];
		if (!T.alive || D.IsGhost((int)t)) continue;
		int r = 0; for (int k=1; k<3; ++k) if (T.v[k] < T.v[r]) r = k;
		faces.push_back({{(uint32_t)T.v[r], (uint32_t)T.v[(r+1)%3], (uint32_t)T.v[(r+2)%3]}, (int)t});
	}
	std::sort(faces.begin(), faces.end(), [](const Face& x, const Face& y) { return std::lexicographical_compare(x.v, x.v+3, y.v, y.v+3); });
	for (size_t f=0; f<faces.size(); ++f) remap[faces[f].src] = (int)f;
	tris.resize(faces.size()*3);
	if (adjacency) adjacency->assign(faces.size()*3, -1);
	for (size_t f=0; f<faces.size(); ++f) {
		const Delaunay2D::Tri& T = D.tris[faces[f].src];
		int r = 0; for (int k=1; k<3; ++k) if (T.v[k] < T.v[r]) r = k;
		for (int k=0; k<3; ++k) {
			tris[f*3+k] = faces[f].v[k];
			if (adjacency) { const int nb = T.n[(r+k)%3]; (*adjacency)[f*3+k] = (nb >= 0 && remap[nb] >= 0) ? remap[nb] : -1; }
		}
	}
	return !tris.empty();
}

// TriangulatePointsDelaunay (DepthMap.cpp:1797-1876): project the view's sparse points (vertex = (x/z, y/z, z)), add the four image
// corners at the view's average depth (bAddCorners), triangulate, then give every corner the depth of the 3 nearest faces around it.
bool TriangulateInit(const Scene& scene, uint32_t idxImage, const std::vector<uint32_t>& points, bool bAddCorners,
	std::vector<double>& vertices, std::vector<uint32_t>& tris, float& dMin, float& dMax)
{
	const Image& im = scene.images[idxImage];
	const Camera& cam = im.camera;
	vertices.clear(); tris.clear();
	dMin = FLT_MAX; dMax = 0.f;
	std::vector<double> xy;
	for (uint32_t ip: points) {
		// Camera::ProjectPointP3<float> (Camera.h:276-282): f64 products, f32 result; then x/z, y/z in f32, widened
		const float* X = &scene.pointcloud.xyz[(size_t)ip*3];
		const double* p = cam.P;
		const float px = (float)(p[0]*X[0]+p[1]*X[1]+p[2]*X[2]+p[3]);
		const float py = (float)(p[4]*X[0]+p[5]*X[1]+p[6]*X[2]+p[7]);
		const float pz = (float)(p[8]*X[0]+p[9]*X[1]+p[10]*X[2]+p[11]);
		vertices.push_back((double)(px/pz)); vertices.push_back((double)(py/pz)); vertices.push_back((double)pz);
		if (dMin > pz) dMin = pz;
		if (dMax < pz) dMax = pz;
	}
	const int nPts = (int)points.size();
	int corner[4] = {-1, -1, -1, -1};
	if (bAddCorners) {
		const double cx[4] = {0, (double)im.width, 0, (double)im.width}, cy[4] = {0, 0, (double)im.height, (double)im.height};
		for (int i=0; i<4; ++i) { corner[i] = nPts+i; vertices.push_back(cx[i]); vertices.push_back(cy[i]); vertices.push_back((double)im.avgDepth); }
	}
	const int nV = (int)(vertices.size()/3);
	xy.resize((size_t)nV*2);
	for (int i=0; i<nV; ++i) { xy[i*2] = vertices[i*3]; xy[i*2+1] = vertices[i*3+1]; }
	std::vector<int> adj;
	if (!DelaunayTriangulate(xy, tris, &adj)) return false;
	if (!bAddCorners) return true;
	// corner depths (:1810-1874): for every face around the corner take the face BEHIND it (across the edge opposite the corner); if
	// that face is finite and touches no corner, intersect the corner's viewing ray with its plane; keep the 3 faces whose centroids
	// are nearest (StoreTop by 1/distance) and average their depths weighted by 1/distance
	const double fx = cam.K[0], fy = cam.K[4], pcx = cam.K[2], pcy = cam.K[5];
	auto I2C = [&](const double* v, double out[3]) { out[0] = (v[0]-pcx)*v[2]/fx; out[1] = (v[1]-pcy)*v[2]/fy; out[2] = v[2]; }; // TransformPointI2C(Point3), Camera.h:307-312
	const size_t nT = tris.size()/3;
	for (int i=0; i<4; ++i) {
		const int vc = corner[i];
		double* A = &vertices[(size_t)vc*3];
		double rayA[3]; I2C(A, rayA);
		{ const double nrm = std::sqrt(rayA[0]*rayA[0]+rayA[1]*rayA[1]+rayA[2]*rayA[2]); const double inv = nrm ? 1./nrm : 0.; for (double& r: rayA) r *= inv; } // normalized()
		float topDepth[3], topScore[3]; int nTop = 0;
		for (size_t t=0; t<nT; ++t) {
			int k = -1; for (int q=0; q<3; ++q) if ((int)tris[t*3+q] == vc) k = q;
			if (k < 0) continue;
			const int fc = adj[t*3+k];
			if (fc < 0) continue; // infinite face
			bool hasCorner = false;
			for (int q=0; q<3; ++q) for (int j=0; j<4; ++j) if ((int)tris[(size_t)fc*3+q] == corner[j]) hasCorner = true;
			if (hasCorner) continue;
			const double* B0 = &vertices[(size_t)tris[(size_t)fc*3]*3]; const double* B1 = &vertices[(size_t)tris[(size_t)fc*3+1]*3]; const double* B2 = &vertices[(size_t)tris[(size_t)fc*3+2]*3];
			double c0[3], c1[3], c2[3]; I2C(B0, c0); I2C(B1, c1); I2C(B2, c2);
			// Planed(p0, p1, p2) (Plane.inl:55-61) and Ray3d(0, dir).Intersects(plane) (Ray.inl:690-703)
			const double e1[3] = {c1[0]-c0[0], c1[1]-c0[1], c1[2]-c0[2]}, e2[3] = {c2[0]-c0[0], c2[1]-c0[1], c2[2]-c0[2]};
			double N[3] = {e1[1]*e2[2]-e1[2]*e2[1], e1[2]*e2[0]-e1[0]*e2[2], e1[0]*e2[1]-e1[1]*e2[0]};
			const double nn = std::sqrt(N[0]*N[0]+N[1]*N[1]+N[2]*N[2]);
			if (nn > 0) { N[0] /= nn; N[1] /= nn; N[2] /= nn; }
			const double Dp = -(N[0]*c0[0]+N[1]*c0[1]+N[2]*c0[2]);
			const double Vd = N[0]*rayA[0]+N[1]*rayA[1]+N[2]*rayA[2];
			const double tHit = Vd == 0 ? 1e+14 : (-Dp)/Vd; // SAFEDIVIDE -> INVZERO(double) = INV_ZERO (Types.h:555, 1222)
			const double z = rayA[2]*tHit;
			if (!(z > 0)) continue;
			const double bx = (B0[0]+B1[0]+B2[0])/3.0, by = (B0[1]+B1[1]+B2[1])/3.0;
			const double dist = std::sqrt((bx-A[0])*(bx-A[0])+(by-A[1])*(by-A[1]));
			const float fd = (float)dist;
			const float score = fd == 0.f ? 1000000.f : 1.f/fd; // INVERT -> FINV_ZERO (Types.h:573)
			const float depth = std::min(std::max((float)z, dMin), dMax); // CLAMP
			// cList::StoreTop<3>: sorted by score descending, inserted before the first element with score <= its own
			int pos = 0; while (pos < nTop && topScore[pos] > score) ++pos;
			if (pos < nTop) { if (nTop >= 3) --nTop; for (int q=nTop; q>pos; --q) { topScore[q] = topScore[q-1]; topDepth[q] = topDepth[q-1]; } topScore[pos] = score; topDepth[pos] = depth; ++nTop; }
			else if (nTop < 3) { topScore[nTop] = score; topDepth[nTop] = depth; ++nTop; }
		}
		if (nTop != 3) continue;
		const float inv = 1.f/((topScore[0]+topScore[1])+topScore[2]);
		const float w0 = topScore[0]*inv, w1 = topScore[1]*inv, w2 = topScore[2]*inv;
		A[2] = (double)((topDepth[0]*w0+topDepth[1]*w1)+topDepth[2]*w2);
	}
	return true;
}

} // namespace hcmvs_host
