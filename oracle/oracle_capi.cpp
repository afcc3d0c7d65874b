// C API over the oracle for ctypes — TEST INFRASTRUCTURE ONLY (see hcmvs_oracle.hpp).
#include "hcmvs_oracle.hpp"
#include <cstring>
#include <string>
#include <algorithm>

using namespace orc;

struct orc_scene {
	Scene scene;
	PointCloud cloud;
};

#include <cmath>
#include <algorithm>
extern "C" {

orc_scene* orc_scene_create() { return new orc_scene(); }
void orc_scene_destroy(orc_scene* s) { delete s; }

int orc_set_param(orc_scene* s, const char* name, double v) {
	Params& P = s->scene.P;
	const std::string n(name);
	#define PU(f) if (n == #f) { P.f = (unsigned)v; return 0; }
	#define PI(f) if (n == #f) { P.f = (int)v; return 0; }
	#define PF(f) if (n == #f) { P.f = (float)v; return 0; }
	PU(nNumViews) PU(nMaxViews) PU(nMinViews) PU(nMinViewsTrustPoint) PU(nMinViewsFuse) PI(viewspread) PU(nMinViewsFilter)
	PU(nMinViewsFilterAdjust) PI(bFilterAdjust) PF(fViewMinScore) PF(fViewMinScoreRatio) PF(fMinArea) PF(fMinAngle)
	PF(fOptimAngle) PF(fMaxAngle) PF(fNCCThresholdKeep) PU(nEstimationIters) PU(nEstimationIters_external)
	PU(nRandomIters) PF(fRandomDepthRatio) PF(fRandomAngle1Range) PF(fRandomAngle2Range) PF(fRandomSmoothDepth)
	PF(fRandomSmoothNormal) PF(fRandomSmoothBonus) PF(fDescriptorMinMagnitudeThreshold) PF(fDepthDiffThreshold)
	PF(fNormalDiffThreshold) PF(depthweight) PF(normalweight) PI(adapthalfwin) PI(propagatehalfwin) PI(propagatestep)
	PI(photo2geo) PF(photometric_flow) PF(para_prior) PF(fsigmaPrior)
	#undef PU
	#undef PI
	#undef PF
	return -1;
}

int orc_add_image(orc_scene* s, int w, int h, const double* K, const double* R, const double* C, const uint8_t* bgr, const float* gray) {
	ImageData im; im.w = w; im.h = h;
	std::memcpy(im.cam.K, K, 72); std::memcpy(im.cam.R, R, 72); std::memcpy(im.cam.C, C, 24);
	im.cam.ComposeP();
	if (bgr) im.bgr.assign(bgr, bgr+(size_t)w*h*3);
	if (gray) { im.gray.w = w; im.gray.h = h; im.gray.d.assign(gray, gray+(size_t)w*h); }
	else if (bgr) ToGray(bgr, w, h, im.gray);
	else return -1;
	s->scene.images.push_back(std::move(im));
	s->scene.arrDepthData.resize(s->scene.images.size());
	s->scene.arrDepthData.back().idxImage = (uint32_t)s->scene.images.size()-1;
	return (int)s->scene.images.size()-1;
}

int orc_get_gray(orc_scene* s, int idx, float* out) {
	const Image32F& g = s->scene.images[idx].gray;
	std::memcpy(out, g.d.data(), g.d.size()*sizeof(float));
	return 0;
}

void orc_set_sparse(orc_scene* s, int n, const float* xyz, const int32_t* offsets, const uint32_t* view_ids) {
	SparseCloud& pc = s->scene.sparse;
	pc.points.resize(n); pc.views.resize(n);
	for (int i=0; i<n; ++i) {
		pc.points[i] = Vec3f{xyz[i*3], xyz[i*3+1], xyz[i*3+2]};
		pc.views[i].assign(view_ids+offsets[i], view_ids+offsets[i+1]);
		std::sort(pc.views[i].begin(), pc.views[i].end());
	}
}

int orc_select_views(orc_scene* s, int idx) { return SelectViews(s->scene, (uint32_t)idx) ? (int)s->scene.arrDepthData[idx].neighbors.size() : -1; }
int orc_init_views(orc_scene* s, int idx, int numNeighbors) { return InitViews(s->scene, (uint32_t)idx, (unsigned)numNeighbors) ? (int)s->scene.arrDepthData[idx].images.size()-1 : -1; }

// which: 0 = all scored neighbours (Image::neighbors), 1 = filtered (DepthData::neighbors)
int orc_get_neighbors(orc_scene* s, int idx, int which, uint32_t* ids, uint32_t* points, float* scale, float* angle, float* area, float* score, int cap) {
	const std::vector<ViewScore>& v = which == 0 ? s->scene.images[idx].neighbors : s->scene.arrDepthData[idx].neighbors;
	const int n = std::min((int)v.size(), cap);
	for (int i=0; i<n; ++i) {
		if (ids) ids[i] = v[i].ID;
		if (points) points[i] = v[i].points;
		if (scale) scale[i] = v[i].scale;
		if (angle) angle[i] = v[i].angle;
		if (area) area[i] = v[i].area;
		if (score) score[i] = v[i].score;
	}
	return (int)v.size();
}
int orc_get_match_views(orc_scene* s, int idx, uint32_t* ids, int cap) {
	const std::vector<uint32_t>& im = s->scene.arrDepthData[idx].images;
	const int n = im.empty() ? 0 : (int)im.size()-1;
	for (int i=0; i<std::min(n, cap); ++i) ids[i] = im[i+1];
	return n;
}
int orc_get_points(orc_scene* s, int idx, uint32_t* ids, int cap) {
	const std::vector<uint32_t>& p = s->scene.arrDepthData[idx].points;
	for (int i=0; i<std::min((int)p.size(), cap); ++i) ids[i] = p[i];
	return (int)p.size();
}
// manual neighbour override (scenes without a sparse cloud): first n_match ids are the matching views
int orc_set_neighbors(orc_scene* s, int idx, const uint32_t* ids, const float* scores, int n_match, int n_all) {
	Scene& sc = s->scene;
	DepthData& dd = sc.arrDepthData[idx];
	dd.idxImage = (uint32_t)idx; dd.neighbors.clear(); dd.images.clear();
	dd.images.push_back((uint32_t)idx);
	for (int i=0; i<n_all; ++i) {
		ViewScore v{ids[i], 0u, 1.f, 0.2f, 1.f, scores ? scores[i] : float(n_all-i)};
		dd.neighbors.push_back(v);
		if (i < n_match) dd.images.push_back(ids[i]);
	}
	sc.images[idx].neighbors = dd.neighbors;
	dd.valid = true;
	return 0;
}

// InitViews with ViewData::ScaleImage (SceneDensify.cpp:370-376, DepthMap.h:232-238): matching view `slot` (0-based) of `idx` uses this
// resized gray image and the K of the new resolution (R, C unchanged). The resize itself is OpenCV's; the tests produce it with cv2.
int orc_set_neighbor_image(orc_scene* s, int idx, int slot, int w, int h, const double* K, const float* gray) {
	Scene& sc = s->scene;
	DepthData& dd = sc.arrDepthData[idx];
	if (slot < 0 || slot+1 >= (int)dd.images.size()) return -1;
	if (dd.scaledImages.size() < dd.images.size()-1) dd.scaledImages.resize(dd.images.size()-1);
	ImageData& im = dd.scaledImages[slot];
	if (!gray) { im = ImageData(); return 0; }
	im = ImageData();
	im.w = w; im.h = h;
	im.cam = sc.images[dd.images[slot+1]].cam;
	std::memcpy(im.cam.K, K, 72);
	im.cam.ComposeP();
	im.gray.w = w; im.gray.h = h; im.gray.d.assign(gray, gray+(size_t)w*h);
	return 0;
}
int orc_init_depth_sparse(orc_scene* s, int idx) { InitDepthMapFromSparse(s->scene, (uint32_t)idx); return 0; }

int orc_set_depthmap(orc_scene* s, int idx, const float* depth, const float* normal, const float* conf, float dMin, float dMax) {
	Scene& sc = s->scene;
	DepthData& dd = sc.arrDepthData[idx];
	const int w = sc.images[idx].w, h = sc.images[idx].h; const size_t n = (size_t)w*h;
	dd.depthMap.w = w; dd.depthMap.h = h; dd.depthMap.d.assign(depth, depth+n);
	dd.normalMap.resize(n);
	for (size_t i=0; i<n; ++i) dd.normalMap[i] = normal ? Vec3f{normal[i*3], normal[i*3+1], normal[i*3+2]} : Vec3f{0,0,0};
	dd.confMap.w = w; dd.confMap.h = h;
	if (conf) dd.confMap.d.assign(conf, conf+n); else dd.confMap.d.assign(n, 0.f);
	dd.dMin = dMin; dd.dMax = dMax;
	if (dd.graMap.d.empty()) {
		if (!sc.images[idx].bgr.empty()) InitGraMap(sc.images[idx].bgr.data(), w, h, dd.graMap);
		else { dd.graMap.w = w; dd.graMap.h = h; dd.graMap.d.assign(n, 0); }
	}
	return 0;
}
int orc_get_depthmap(orc_scene* s, int idx, float* depth, float* normal, float* conf, float* dMinMax) {
	const DepthData& dd = s->scene.arrDepthData[idx];
	const size_t n = dd.depthMap.d.size();
	if (depth) std::memcpy(depth, dd.depthMap.d.data(), n*4);
	if (normal) std::memcpy(normal, dd.normalMap.data(), n*12);
	if (conf) std::memcpy(conf, dd.confMap.d.data(), n*4);
	if (dMinMax) { dMinMax[0] = dd.dMin; dMinMax[1] = dd.dMax; }
	return 0;
}
int orc_set_prior(orc_scene* s, int idx, const float* prior) {
	Scene& sc = s->scene; DepthData& dd = sc.arrDepthData[idx];
	const int w = sc.images[idx].w, h = sc.images[idx].h;
	dd.depthMapPrior.w = w; dd.depthMapPrior.h = h;
	if (prior) dd.depthMapPrior.d.assign(prior, prior+(size_t)w*h); else dd.depthMapPrior.d.clear();
	return 0;
}
int orc_set_coarse(orc_scene* s, int idx, const float* depth, const float* normal, int wc, int hc) {
	if (idx < 0 || idx >= (int)s->scene.images.size()) return -1;
	SetCoarseEstimate(s->scene, (uint32_t)idx, depth, normal, wc, hc); return 0;
}
int orc_get_coarse(orc_scene* s, int idx, float* depth, float* normal) {
	const DepthData& dd = s->scene.arrDepthData[idx];
	if (dd.coarseDepth.d.empty()) return -1;
	if (depth) std::memcpy(depth, dd.coarseDepth.d.data(), dd.coarseDepth.d.size()*4);
	if (normal) std::memcpy(normal, dd.coarseNormal.data(), dd.coarseNormal.size()*12);
	return 0;
}
void orc_snapshot_maps(orc_scene* s) { SnapshotMaps(s->scene); }
void orc_resize_area_up(const float* src, int sw, int sh, int cn, float* dst, int dw, int dh) { ResizeAreaUp(src, sw, sh, cn, dst, dw, dh); }
int orc_get_gramap(orc_scene* s, int idx, uint8_t* out) {
	const Image8U& g = s->scene.arrDepthData[idx].graMap;
	std::memcpy(out, g.d.data(), g.d.size());
	return 0;
}

int orc_score_depthmap(orc_scene* s, int idx, int it_external, uint64_t seed, int nThreads) {
	ScoreDepthMap(s->scene, (uint32_t)idx, it_external, seed, (unsigned)std::max(nThreads, 1)); return 0;
}
// mode 0 = reference serial/threaded raster sweep, 1 = red-black restatement. stats[5] = secScore, secSweeps, secEnd, nHyp, nPixelIters
int orc_estimate_depthmap(orc_scene* s, int idx, int it_external, uint64_t seed, int nThreads, int mode, int farReach, int runEnd, double* stats) {
	EstimateStats st;
	bool ok;
	if (mode == 0) ok = EstimateDepthMap(s->scene, (uint32_t)idx, it_external, seed, (unsigned)std::max(nThreads, 1), &st, runEnd != 0);
	else { RedBlackCfg cfg; cfg.nDirs = (mode == 2 || mode == 3) ? 2 : 4; cfg.blockShare = (mode == 3 || mode == 4) ? 1 : 0; cfg.farReach = farReach > 0 ? farReach : 1; cfg.useFar = farReach > 1; ok = EstimateDepthMapRedBlack(s->scene, (uint32_t)idx, it_external, seed, (unsigned)std::max(nThreads, 1), cfg, &st, runEnd != 0); }
	if (stats) { stats[0] = st.secScore; stats[1] = st.secSweeps; stats[2] = st.secEnd; stats[3] = (double)st.nHypotheses; stats[4] = (double)st.nPixelIters; }
	return ok ? 0 : -1;
}
int orc_score_hypotheses(orc_scene* s, int idx, const float* depth, const float* normal, int smoothMode, float* out) {
	ScoreHypotheses(s->scene, (uint32_t)idx, depth, normal, smoothMode, out); return 0;
}
int orc_end_depthmap(orc_scene* s, int idx) { EndDepthMap(s->scene, (uint32_t)idx); return 0; }

int orc_filter_depthmap(orc_scene* s, int idx, const uint32_t* nbIdx, int n, int bAdjust, float* outDepth, float* outConf) {
	std::vector<uint32_t> nb(nbIdx, nbIdx+n);
	Image32F d, c;
	if (!FilterDepthMap(s->scene, (uint32_t)idx, nb, bAdjust != 0, d, c)) return -1;
	std::memcpy(outDepth, d.d.data(), d.d.size()*4);
	std::memcpy(outConf, c.d.data(), c.d.size()*4);
	return 0;
}

int orc_fuse(orc_scene* s, int estColor, int estNormal) {
	s->cloud = PointCloud();
	FuseDepthMaps(s->scene, s->cloud, estColor != 0, estNormal != 0);
	return (int)s->cloud.points.size();
}
int orc_fuse_get(orc_scene* s, float* xyz, float* normals, uint8_t* colors, int32_t* nViews) {
	const PointCloud& pc = s->cloud;
	if (xyz) std::memcpy(xyz, pc.points.data(), pc.points.size()*12);
	if (normals && !pc.normals.empty()) std::memcpy(normals, pc.normals.data(), pc.normals.size()*12);
	if (colors && !pc.colors.empty()) std::memcpy(colors, pc.colors.data(), pc.colors.size());
	if (nViews) for (size_t i=0; i<pc.pointViews.size(); ++i) nViews[i] = (int32_t)pc.pointViews[i].size();
	return 0;
}
// CSR views/weights of the fused cloud
int orc_fuse_get_views(orc_scene* s, uint32_t* views, float* weights) {
	size_t k = 0;
	for (size_t i=0; i<s->cloud.pointViews.size(); ++i)
		for (size_t j=0; j<s->cloud.pointViews[i].size(); ++j, ++k) { if (views) views[k] = s->cloud.pointViews[i][j]; if (weights) weights[k] = s->cloud.pointWeights[i][j]; }
	return (int)k;
}

// ---- helper hooks for unit tests
void orc_median3(float* img, int w, int h) { Image32F im; im.w = w; im.h = h; im.d.assign(img, img+(size_t)w*h); MedianBlur3(im); std::memcpy(img, im.d.data(), (size_t)w*h*4); }
void orc_gramap(const uint8_t* bgr, int w, int h, uint8_t* out) { Image8U g; InitGraMap(bgr, w, h, g); std::memcpy(out, g.d.data(), g.d.size()); }
void orc_togray(const uint8_t* bgr, int w, int h, float* out) { Image32F g; ToGray(bgr, w, h, g); std::memcpy(out, g.d.data(), g.d.size()*4); }
int  orc_zigzag(int w, int h, int rawStride, uint16_t* out) { std::vector<uint16_t> c; MapMatrix2ZigzagIdx(w, h, c, rawStride); if (out) std::memcpy(out, c.data(), c.size()*2); return (int)(c.size()/2); }
void orc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { Philox4x32_10(ctr, key, out); }
float orc_sample(const float* img, int w, int h, float x, float y) { Image32F im; im.w = w; im.h = h; im.d.assign(img, img+(size_t)w*h); return SampleBilinear(im, x, y); }
void orc_dir2normal(float a, float b, float* out) { Vec3f n; Dir2Normal(a, b, n); out[0] = n.x; out[1] = n.y; out[2] = n.z; }
void orc_normal2dir(const float* n, float* out) { Normal2Dir(Vec3f{n[0], n[1], n[2]}, out[0], out[1]); }


// The stock OpenMVS speckle filter that the fork keeps under `#if 0` (DepthMapsData::RemoveSmallSegments, SceneDensify.cpp:1956-2042),
// restated with its loop order: seeds in COLUMN-major order, breadth-first growth through 4-neighbours that are valid and
// IsDepthSimilar(depth_curr, depth_neighbor, th) — asymmetric (divides by the current pixel's depth), so a segment is what is REACHABLE
// from its seed among the pixels no earlier segment took; segments of fewer than speckle_size pixels are zeroed (depth, normal, conf).
// Returns the number of zeroed pixels that were valid. normal / conf may be NULL.
int orc_remove_small_segments(float* depth, float* normal, float* conf, int w, int h, unsigned speckle_size, float th) {
	std::vector<unsigned char> done((size_t)w*h, 0);
	std::vector<int> seg((size_t)w*h);
	int removed = 0;
	for (int u=0; u<w; ++u) for (int v=0; v<h; ++v) {
		if (done[(size_t)v*w+u]) continue;
		size_t count = 1, curr = 0;
		seg[0] = v*w+u;
		while (curr < count) {
			const int a = seg[curr];
			const int ax = a%w, ay = a/w;
			const float dc = depth[a];
			if (dc > 0) {
				const int nx[4] = {ax-1, ax+1, ax, ax}, ny[4] = {ay, ay, ay-1, ay+1};
				for (int i=0; i<4; ++i) {
					if (nx[i] < 0 || ny[i] < 0 || nx[i] >= w || ny[i] >= h) continue;
					const int b = ny[i]*w+nx[i];
					if (done[b]) continue;
					const float dn = depth[b];
					if (dn > 0 && std::fabs(dc-dn)/dc < th) { seg[count++] = b; done[b] = 1; } // IsDepthSimilar, Util.inl:657-669
				}
			}
			++curr;
			done[a] = 1;
		}
		if (count < speckle_size)
			for (size_t i=0; i<count; ++i) {
				const int a = seg[i];
				if (depth[a] != 0) ++removed;
				depth[a] = 0;
				if (normal) { normal[(size_t)a*3] = normal[(size_t)a*3+1] = normal[(size_t)a*3+2] = 0; }
				if (conf) conf[a] = 0;
			}
	}
	return removed;
}

// The stock small-gap branch of DepthMapsData::GapInterpolation (SceneDensify.cpp:2294-2352 row-wise, :2640-2683 column-wise; the fork's
// large-gap branches read uninitialised variables and are not restated — DESIGN.md §6): runs of <= gap_size invalid pixels between two
// valid pixels whose depths are IsDepthSimilar(first, last, th) are filled with linearly interpolated depths and (through
// Normal2Dir / Dir2Normal) normals; their confidence becomes the smaller of the two ends'. Rows first, then columns (which see the rows'
// result). normal / conf may be NULL. Returns the number of filled pixels.
int orc_gap_interpolation(float* depth, float* normal, float* conf, int w, int h, unsigned gap_size, float th) {
	int filled = 0;
	auto pass = [&](bool rows) {
		const int outer = rows ? h : w, inner = rows ? w : h;
		for (int o=0; o<outer; ++o) {
			unsigned count = 0;
			for (int i=0; i<inner; ++i) {
				auto idx = [&](int k) { return rows ? (size_t)o*w+k : (size_t)k*w+o; };
				const float d1 = depth[idx(i)];
				if (d1 <= 0) { ++count; continue; }
				if (count == 0) continue;
				if (count <= gap_size && (unsigned)i > count) {
					int k = i-(int)count;
					const int first = k-1;
					const float d0 = depth[idx(first)];
					if (std::fabs(d0-d1)/d0 < th) {
						const float diff = (d1-d0)/(float)(count+1);
						float d = d0;
						const float c = conf ? std::min(conf[idx(first)], conf[idx(i)]) : 0.f;
						float p1x = 0, p1y = 0, dx = 0, dy = 0;
						if (normal) {
							float p2x, p2y;
							Normal2Dir(Vec3f{normal[idx(first)*3], normal[idx(first)*3+1], normal[idx(first)*3+2]}, p1x, p1y);
							Normal2Dir(Vec3f{normal[idx(i)*3], normal[idx(i)*3+1], normal[idx(i)*3+2]}, p2x, p2y);
							dx = (p2x-p1x)/(float)(count+1); dy = (p2y-p1y)/(float)(count+1);
						}
						do {
							depth[idx(k)] = (d += diff);
							if (normal) { p1x += dx; p1y += dy; Vec3f n; Dir2Normal(p1x, p1y, n); normal[idx(k)*3] = n.x; normal[idx(k)*3+1] = n.y; normal[idx(k)*3+2] = n.z; }
							if (conf) conf[idx(k)] = c;
							++filled;
						} while (++k < i);
					}
				}
				count = 0;
			}
		}
	};
	pass(true); pass(false);
	return filled;
}

} // extern "C"
