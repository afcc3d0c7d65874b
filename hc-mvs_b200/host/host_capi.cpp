// C handle API over densify.h — see include/hcmvs_host.h.
#include "hcmvs_host.h"
#include "densify.h"
#include "mvsi.h"
#include <cstring>
#include <algorithm>
#include <exception>

using namespace hcmvs_host;

struct hcmvs_host_scene {
	Scene scene;
	std::vector<DepthData> dd;
	std::string err;
	DistributedReconstruction* dist = nullptr; // hcmvs_host_dist_*
	std::vector<void*> pinned;                 // hcmvs_host_pin_images
	void Unpin() { for (void* p: pinned) hcmvs_unpin_host_memory(p); pinned.clear(); }
	~hcmvs_host_scene() { delete dist; Unpin(); }
};

extern "C" {

hcmvs_host_scene* hcmvs_host_scene_create(void) { return new hcmvs_host_scene(); }
void hcmvs_host_scene_destroy(hcmvs_host_scene* s) { delete s; }
const char* hcmvs_host_last_error(hcmvs_host_scene* s) { return s ? s->err.c_str() : "null scene"; }

int hcmvs_host_add_image(hcmvs_host_scene* s, int w, int h, const double K[9], const double R[9], const double C[3], const uint8_t* bgr, const char* name) {
	if (!s || !K || !R || !C || w <= 0 || h <= 0) return -1;
	Image im; im.width = w; im.height = h;
	memcpy(im.camera.K, K, 72); memcpy(im.camera.R, R, 72); memcpy(im.camera.C, C, 24);
	im.camera.ComposeP();
	if (bgr) { // NULL: camera only — the pixels live on another rank (DenseReconstructionDistributed)
		im.bgr.assign(bgr, bgr+(size_t)w*h*3);
		im.gray.resize((size_t)w*h);
		ToGray(bgr, w, h, im.gray.data());
	}
	if (name) im.name = name;
	s->scene.images.push_back(std::move(im));
	return (int)s->scene.images.size()-1;
}

int hcmvs_host_set_sparse(hcmvs_host_scene* s, int n, const float* xyz, const int32_t* offsets, const uint32_t* view_ids) {
	if (!s || n < 0) return -1;
	SparsePoints& pc = s->scene.pointcloud;
	pc.xyz.assign(xyz, xyz+(size_t)n*3);
	pc.views.resize(n);
	for (int i=0; i<n; ++i) { pc.views[i].assign(view_ids+offsets[i], view_ids+offsets[i+1]); std::sort(pc.views[i].begin(), pc.views[i].end()); }
	return 0;
}

int hcmvs_host_select_views(hcmvs_host_scene* s, const hcmvs_params* p, int idx) { return hcmvs_host_select_views_mt(s, p, idx, 1); }
int hcmvs_host_select_views_mt(hcmvs_host_scene* s, const hcmvs_params* p, int idx, int threads) {
	if (!s || !p || idx < 0 || idx >= (int)s->scene.images.size() || threads < 1) return -1;
	DepthMapsData data(s->scene, nullptr, *p);
	if (!data.SelectViews((uint32_t)idx, (unsigned)threads)) return -1;
	if (s->dd.size() != s->scene.images.size()) s->dd.resize(s->scene.images.size());
	s->dd[idx] = data.arrDepthData[idx];
	return (int)s->dd[idx].neighbors.size();
}

int hcmvs_host_get_neighbors(hcmvs_host_scene* s, int idx, int which, uint32_t* ids, uint32_t* points, float* scale, float* angle, float* area, float* score, int cap) {
	if (!s || idx < 0 || idx >= (int)s->scene.images.size()) return -1;
	static const std::vector<ViewScore> empty;
	const std::vector<ViewScore>& v = which == 0 ? s->scene.images[idx].neighbors : (idx < (int)s->dd.size() ? s->dd[idx].neighbors : empty);
	for (int i=0; i<std::min((int)v.size(), cap); ++i) {
		if (ids) ids[i] = v[i].ID;
		if (points) points[i] = v[i].points;
		if (scale) scale[i] = v[i].scale;
		if (angle) angle[i] = v[i].angle;
		if (area) area[i] = v[i].area;
		if (score) score[i] = v[i].score;
	}
	return (int)v.size();
}

int hcmvs_host_init_depth(hcmvs_host_scene* s, int idx, float* depth, float* dminmax) {
	if (!s || idx < 0 || idx >= (int)s->dd.size() || !s->dd[idx].valid) return -1;
	std::vector<float> d; float a, b;
	SparseInitDepth(s->scene, (uint32_t)idx, s->dd[idx].points, d, a, b);
	memcpy(depth, d.data(), d.size()*4);
	dminmax[0] = a; dminmax[1] = b;
	return 0;
}

int hcmvs_host_scale_image(const float* src, int sw, int sh, float scale, float* dst, int* dw, int* dh, const double* K, double* Kout) {
	if (!src || !dw || !dh) return -1;
	std::vector<float> in(src, src+(size_t)sw*sh), out; int w = 0, h = 0;
	if (!ScaleImage(in, sw, sh, scale, out, w, h)) return 1; // |scale-1| < 0.15: not rescaled
	*dw = w; *dh = h;
	if (dst) memcpy(dst, out.data(), out.size()*4);
	if (K && Kout) ScaleK(K, sw, sh, w, h, Kout);
	return 0;
}

int hcmvs_host_get_gray(hcmvs_host_scene* s, int idx, float* gray) {
	if (!s || idx < 0 || idx >= (int)s->scene.images.size()) return -1;
	const Image& im = s->scene.images[idx];
	memcpy(gray, im.gray.data(), im.gray.size()*4);
	return 0;
}

int hcmvs_host_dense_reconstruction(hcmvs_host_scene* s, hcmvs_ctx* ctx, const hcmvs_params* p, uint64_t seed, int run_filter, const char* dmap_dir, double* stats) {
	if (!s || !ctx || !p) return -1;
	DenseReconstructionStats st;
	for (Image& im: s->scene.images) im.neighbors.clear();
	if (!DenseReconstruction(s->scene, ctx, *p, ViewSelectionParams(), seed, run_filter != 0, dmap_dir ? dmap_dir : "", &st, &s->err)) return -2;
	if (stats) {
		stats[0] = st.secSelect; stats[1] = st.secUpload; stats[2] = st.secEstimate; stats[3] = st.secFilter; stats[4] = st.secFuse;
		stats[5] = (double)st.h2dBytes; stats[6] = (double)st.d2hBytes; stats[7] = (double)s->scene.densecloud.size();
	}
	return 0;
}

int hcmvs_host_dense_reconstruction_distributed(hcmvs_host_scene* s, hcmvs_ctx* ctx, const hcmvs_params* p, uint64_t seed, int run_filter, int rank, int world, double* stats) {
	if (!s || !ctx || !p) return -1;
	DenseReconstructionStats st;
	for (Image& im: s->scene.images) im.neighbors.clear();
	if (!DenseReconstructionDistributed(s->scene, ctx, *p, ViewSelectionParams(), seed, run_filter != 0, rank, world, &st, &s->err)) return -2;
	if (stats) {
		stats[0] = st.secSelect; stats[1] = st.secUpload; stats[2] = st.secEstimate; stats[3] = st.secFilter; stats[4] = st.secFuse;
		stats[5] = (double)st.h2dBytes; stats[6] = (double)st.d2hBytes; stats[7] = (double)s->scene.densecloud.size();
	}
	return 0;
}

int hcmvs_host_pin_images(hcmvs_host_scene* s, int pin) {
	// page-lock the pixel buffers of the scene's images (gray + colour): the uploads then run at PCIe rate. The buffers must not be
	// re-allocated while pinned (hcmvs_host_scene_reload_images: unpin first).
	if (!s) return -1;
	s->Unpin();
	if (!pin) return 0;
	for (Image& im: s->scene.images) {
		if (!im.gray.empty() && hcmvs_pin_host_memory(im.gray.data(), im.gray.size()*4) == HCMVS_OK) s->pinned.push_back(im.gray.data());
		if (!im.bgr.empty() && hcmvs_pin_host_memory(im.bgr.data(), im.bgr.size()) == HCMVS_OK) s->pinned.push_back(im.bgr.data());
	}
	return (int)s->pinned.size();
}
int hcmvs_host_dist_prepare(hcmvs_host_scene* s, hcmvs_ctx* ctx, const hcmvs_params* p, int rank, int world) {
	if (!s || !ctx || !p || world < 2 || rank < 0 || rank >= world) return -1;
	delete s->dist;
	s->dist = new DistributedReconstruction(s->scene, ctx, *p, ViewSelectionParams(), rank, world);
	if (!s->dist->Prepare()) { s->err = s->dist->Error(); return -2; }
	return 0;
}
int hcmvs_host_dist_info(hcmvs_host_scene* s, int rank, int* n_valid, int* n_mine_whole, int* n_split, int* whole_rounds) {
	if (!s || !s->dist) return -1;
	const ShardPlan& p = s->dist->Plan();
	if (n_valid) *n_valid = (int)s->dist->ValidViews();
	if (n_mine_whole) *n_mine_whole = (int)p.WholeViewsOf(rank).size();
	if (n_split) *n_split = (int)p.SplitViews().size();
	if (whole_rounds) *whole_rounds = p.wholeRounds;
	return 0;
}
int hcmvs_host_dist_upload_initial(hcmvs_host_scene* s) {
	if (!s || !s->dist) return -1;
	if (!s->dist->UploadInitial()) { s->err = s->dist->Error(); return -2; }
	return 0;
}
int hcmvs_host_dist_run(hcmvs_host_scene* s, uint64_t seed, int run_filter, int download, double* stats) {
	if (!s || !s->dist) return -1;
	const bool ok = s->dist->Run(seed, run_filter != 0, download != 0);
	if (!ok) { s->err = s->dist->Error(); return -2; }
	if (stats) {
		const DenseReconstructionStats& st = s->dist->Stats();
		stats[0] = st.secSelect; stats[1] = st.secUpload; stats[2] = st.secEstimate; stats[3] = st.secFilter; stats[4] = st.secFuse;
		stats[5] = (double)st.h2dBytes; stats[6] = (double)st.d2hBytes; stats[7] = (double)s->scene.densecloud.size();
	}
	return 0;
}

int hcmvs_host_shard_plan(const uint32_t* valid_views, int n_valid, const uint32_t* n_scored, int n_views, int world, int split_rows,
	uint32_t* order, int32_t* owner_whole, int32_t* owner_filter, int* whole_rounds, int* n_split)
{
	if (!valid_views || !n_scored || n_valid < 0 || n_views <= 0 || world < 1) return -1;
	std::vector<uint32_t> v(valid_views, valid_views+n_valid), ns(n_scored, n_scored+n_views);
	for (uint32_t id: v) if (id >= (uint32_t)n_views) return -1;
	const ShardPlan plan = MakeShardPlan(v, ns, world, split_rows != 0);
	if (order) std::copy(plan.order.begin(), plan.order.end(), order);
	if (owner_whole) { // rank that estimates the whole view, HCMVS_OWNER_SPLIT_ROWS for a row-split view, -1 outside the plan
		std::fill(owner_whole, owner_whole+n_views, -1);
		for (int r=0; r<plan.wholeRounds; ++r) { const std::vector<int32_t> o = plan.RoundOwners(r, (size_t)n_views); for (int i=0; i<n_views; ++i) if (o[i] >= 0) owner_whole[i] = o[i]; }
		for (uint32_t id: plan.SplitViews()) owner_whole[id] = HCMVS_OWNER_SPLIT_ROWS;
	}
	if (owner_filter) { const std::vector<int32_t> o = plan.Owners((size_t)n_views); std::copy(o.begin(), o.end(), owner_filter); }
	if (whole_rounds) *whole_rounds = plan.wholeRounds;
	if (n_split) *n_split = (int)plan.SplitViews().size();
	return 0;
}

int hcmvs_host_cloud_size(hcmvs_host_scene* s, uint64_t* n_points, uint64_t* n_view_refs) {
	if (!s) return -1;
	if (n_points) *n_points = s->scene.densecloud.size();
	if (n_view_refs) *n_view_refs = s->scene.densecloud.views.size();
	return 0;
}
int hcmvs_host_cloud_get(hcmvs_host_scene* s, float* xyz, float* normals, uint8_t* colors, uint32_t* view_offsets, uint32_t* views, float* weights) {
	if (!s) return -1;
	const PointCloud& pc = s->scene.densecloud;
	if (xyz) memcpy(xyz, pc.points.data(), pc.points.size()*4);
	if (normals && !pc.normals.empty()) memcpy(normals, pc.normals.data(), pc.normals.size()*4);
	if (colors && !pc.colors.empty()) memcpy(colors, pc.colors.data(), pc.colors.size());
	if (view_offsets) memcpy(view_offsets, pc.viewOffsets.data(), pc.viewOffsets.size()*4);
	if (views) memcpy(views, pc.views.data(), pc.views.size()*4);
	if (weights) memcpy(weights, pc.weights.data(), pc.weights.size()*4);
	return 0;
}
long hcmvs_host_pointcloud_filter(hcmvs_host_scene* s, hcmvs_ctx* ctx, int th_remove) {
	if (!s || !ctx) return -1;
	return s->scene.PointCloudFilter(ctx, th_remove, &s->err);
}
int hcmvs_host_cloud_set(hcmvs_host_scene* s, uint64_t n, const float* xyz, const float* normals, const uint8_t* colors, const uint32_t* view_offsets, const uint32_t* views, const float* weights) {
	if (!s || (n && (!xyz || !view_offsets || !views))) return -1;
	PointCloud& pc = s->scene.densecloud;
	const size_t m = n ? view_offsets[n] : 0;
	pc.points.resize(n*3); if (n) memcpy(pc.points.data(), xyz, n*12);
	pc.viewOffsets.resize(n+1); if (n) memcpy(pc.viewOffsets.data(), view_offsets, (n+1)*4); else pc.viewOffsets.data()[0] = 0;
	pc.views.resize(m); if (m) memcpy(pc.views.data(), views, m*4);
	if (normals) { pc.normals.resize(n*3); if (n) memcpy(pc.normals.data(), normals, n*12); } else pc.normals.clear();
	if (colors) { pc.colors.resize(n*3); if (n) memcpy(pc.colors.data(), colors, n*3); } else pc.colors.clear();
	if (weights) { pc.weights.resize(m); if (m) memcpy(pc.weights.data(), weights, m*4); } else pc.weights.clear();
	return 0;
}
long hcmvs_host_cloud_remove_by_visibility(hcmvs_host_scene* s, const int32_t* visibility, int th_remove) {
	if (!s || !visibility) return -1;
	return RemovePointsByVisibility(s->scene.densecloud, visibility, th_remove);
}
int hcmvs_host_cloud_save_ply(hcmvs_host_scene* s, const char* file) { return (s && file && s->scene.densecloud.Save(file)) ? 0 : -1; }

int hcmvs_host_write_dmap(const char* file, const char* image_name, const uint32_t* ids, int n_ids, int image_w, int image_h,
	const double K[9], const double R[9], const double C[3], float dmin, float dmax, int w, int h, const float* depth, const float* normal, const float* conf)
{
	std::vector<uint32_t> IDs(ids, ids+n_ids);
	return ExportDepthDataRaw(file, image_name ? image_name : "", IDs, image_w, image_h, K, R, C, dmin, dmax, w, h, depth, normal, conf) ? 0 : -1;
}
int hcmvs_host_read_dmap_header(const char* file, int* w, int* h, int* n_ids, int* has_normal, int* has_conf) {
	std::string name; std::vector<uint32_t> IDs; int iw, ih, ww, hh; double K[9], R[9], C[3]; float a, b; std::vector<float> d, n, c;
	if (!ImportDepthDataRaw(file, name, IDs, iw, ih, K, R, C, a, b, ww, hh, d, n, c)) return -1;
	*w = ww; *h = hh; *n_ids = (int)IDs.size(); *has_normal = !n.empty(); *has_conf = !c.empty();
	return 0;
}
int hcmvs_host_read_dmap(const char* file, uint32_t* ids, double K[9], double R[9], double C[3], float* dminmax, float* depth, float* normal, float* conf) {
	std::string name; std::vector<uint32_t> IDs; int iw, ih, ww, hh; float a, b; std::vector<float> d, n, c;
	if (!ImportDepthDataRaw(file, name, IDs, iw, ih, K, R, C, a, b, ww, hh, d, n, c)) return -1;
	if (ids) memcpy(ids, IDs.data(), IDs.size()*4);
	if (dminmax) { dminmax[0] = a; dminmax[1] = b; }
	if (depth) memcpy(depth, d.data(), d.size()*4);
	if (normal && !n.empty()) memcpy(normal, n.data(), n.size()*4);
	if (conf && !c.empty()) memcpy(conf, c.data(), c.size()*4);
	return 0;
}

// ---- triangulated initialisation (triangulate.cpp)
int hcmvs_host_save_depthmap(const char* file, const float* depth, int w, int h) { return file && SaveDepthMap(file, depth, w, h) ? 0 : -1; }
int hcmvs_host_save_normalmap(const char* file, const float* normal, int w, int h) { return file && SaveNormalMap(file, normal, w, h) ? 0 : -1; }
int hcmvs_host_load_depthmap(const char* file, float* depth, int* w, int* h) { // depth == NULL: only the size
	if (!file || !w || !h) return -1;
	std::vector<float> d; if (!LoadDepthMap(file, d, *w, *h)) return -1;
	if (depth) memcpy(depth, d.data(), d.size()*4);
	return 0;
}
int hcmvs_host_load_normalmap(const char* file, float* normal, int* w, int* h) {
	if (!file || !w || !h) return -1;
	std::vector<float> d; if (!LoadNormalMap(file, d, *w, *h)) return -1;
	if (normal) memcpy(normal, d.data(), d.size()*4);
	return 0;
}

int hcmvs_host_delaunay(const double* xy, int n, uint32_t* tris, int cap_tris) {
	if (!xy || n < 3) return -1;
	std::vector<double> pts(xy, xy+(size_t)n*2); std::vector<uint32_t> t;
	if (!DelaunayTriangulate(pts, t)) return -1;
	const int m = (int)(t.size()/3);
	if (tris) memcpy(tris, t.data(), (size_t)std::min(m, cap_tris)*12);
	return m;
}
int hcmvs_host_triangulate_init(hcmvs_host_scene* s, int idx, int add_corners, double* vertices, int cap_vertices, uint32_t* tris, int cap_tris, int* n_vertices, int* n_tris, float* dminmax) {
	if (!s || idx < 0 || idx >= (int)s->dd.size() || !s->dd[idx].valid) return -1;
	std::vector<double> v; std::vector<uint32_t> t; float a, b;
	if (!TriangulateInit(s->scene, (uint32_t)idx, s->dd[idx].points, add_corners != 0, v, t, a, b)) return -1;
	if (n_vertices) *n_vertices = (int)(v.size()/3);
	if (n_tris) *n_tris = (int)(t.size()/3);
	if (dminmax) { dminmax[0] = a; dminmax[1] = b; }
	if (vertices) memcpy(vertices, v.data(), std::min(v.size(), (size_t)cap_vertices*3)*8);
	if (tris) memcpy(tris, t.data(), std::min(t.size(), (size_t)cap_tris*3)*4);
	return 0;
}

// ---- MVSI project files (mvsi.h)
int hcmvs_host_scene_load_mvs(hcmvs_host_scene* s, const char* file, int load_images) {
	if (!s || !file) return -1;
	s->dd.clear();
	try { return s->scene.LoadInterface(file, load_images != 0, &s->err) ? 0 : -1; }
	catch (const std::exception& e) { s->err = std::string("LoadInterface: ")+e.what(); return -1; } // no exception crosses the C boundary
}
int hcmvs_host_scene_save_mvs(hcmvs_host_scene* s, const char* file, int version, int dense) {
	if (!s || !file) return -1;
	if (!s->scene.SaveInterface(file, version, dense != 0)) { s->err = std::string("cannot write '")+file+"'"; return -1; }
	return 0;
}
int hcmvs_host_scene_reload_images(hcmvs_host_scene* s, unsigned resolution_level, unsigned min_resolution, unsigned max_resolution) {
	if (!s) return -1;
	s->dd.clear();
	return s->scene.ReloadImages(resolution_level, min_resolution, max_resolution, &s->err) ? 0 : -1;
}
int hcmvs_host_resize_area_bgr(const uint8_t* src, int sw, int sh, int dw, int dh, uint8_t* dst) { return ResizeAreaBGR(src, sw, sh, dw, dh, dst) ? 0 : -1; }
int hcmvs_host_num_images(hcmvs_host_scene* s) { return s ? (int)s->scene.images.size() : -1; }
int hcmvs_host_get_image_info(hcmvs_host_scene* s, int idx, int* w, int* h, int* calibrated, uint32_t* id, double K[9], double R[9], double C[3], char* name, int name_cap) {
	if (!s || idx < 0 || idx >= (int)s->scene.images.size()) return -1;
	const Image& im = s->scene.images[idx];
	if (w) *w = im.width;
	if (h) *h = im.height;
	if (calibrated) *calibrated = im.calibrated;
	if (id) *id = im.ID;
	if (K) memcpy(K, im.camera.K, 72);
	if (R) memcpy(R, im.camera.R, 72);
	if (C) memcpy(C, im.camera.C, 24);
	if (name && name_cap > 0) { strncpy(name, im.name.c_str(), (size_t)name_cap-1); name[name_cap-1] = 0; }
	return 0;
}
int hcmvs_host_get_image_bgr(hcmvs_host_scene* s, int idx, uint8_t* bgr) {
	if (!s || idx < 0 || idx >= (int)s->scene.images.size() || !bgr) return -1;
	const Image& im = s->scene.images[idx];
	if (im.bgr.empty()) return -1;
	memcpy(bgr, im.bgr.data(), im.bgr.size());
	return 0;
}
int hcmvs_host_get_sparse(hcmvs_host_scene* s, uint64_t* n_points, uint64_t* n_view_refs, float* xyz, int32_t* offsets, uint32_t* view_ids, float* weights) {
	if (!s) return -1;
	const SparsePoints& pc = s->scene.pointcloud;
	uint64_t refs = 0;
	for (size_t i=0; i<pc.size(); ++i) {
		if (offsets) offsets[i] = (int32_t)refs;
		for (size_t k=0; k<pc.views[i].size(); ++k) {
			if (view_ids) view_ids[refs+k] = pc.views[i][k];
			if (weights) weights[refs+k] = pc.weights.empty() ? 0.f : pc.weights[i][k];
		}
		refs += pc.views[i].size();
	}
	if (offsets) offsets[pc.size()] = (int32_t)refs;
	if (xyz && !pc.xyz.empty()) memcpy(xyz, pc.xyz.data(), pc.xyz.size()*4);
	if (n_points) *n_points = pc.size();
	if (n_view_refs) *n_view_refs = refs;
	return 0;
}
int hcmvs_host_load_image(const char* file, int* w, int* h, uint8_t* bgr) {
	if (!file || !w || !h) return -1;
	if (!bgr) return ReadImageSize(file, *w, *h) ? 0 : -1;
	std::vector<uint8_t> px; int iw, ih;
	try { if (!LoadImageBGR(file, iw, ih, px)) return -1; } catch (const std::exception&) { return -1; }
	if (iw != *w || ih != *h) return -2;
	memcpy(bgr, px.data(), px.size());
	return 0;
}

} // extern "C"
