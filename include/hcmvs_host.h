/* C handle API over the host-side mirror (hc-mvs_b200/host/densify.h) — what a Python / C caller drives.
 * The C++ classes in densify.h are the drop-in for C++ callers (same method names as MVS::DepthMapsData). */
#ifndef HCMVS_HOST_H_
#define HCMVS_HOST_H_
#include <stdint.h>
#include "hcmvs_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct hcmvs_host_scene hcmvs_host_scene;

hcmvs_host_scene* hcmvs_host_scene_create(void);
void hcmvs_host_scene_destroy(hcmvs_host_scene* s);
const char* hcmvs_host_last_error(hcmvs_host_scene* s);
/* MVS::Image + camera (libs/MVS/Image.h, Camera.h). bgr: H*W*3 u8. Returns the image index. */
int hcmvs_host_add_image(hcmvs_host_scene* s, int w, int h, const double K[9], const double R[9], const double C[3], const uint8_t* bgr, const char* name);
/* Sparse MVS::PointCloud (points + per-point view lists, CSR). */
int hcmvs_host_set_sparse(hcmvs_host_scene* s, int n, const float* xyz, const int32_t* offsets, const uint32_t* view_ids);
/* DepthMapsData::SelectViews for one image (host only, needs no device). Returns #filtered neighbours or -1. */
int hcmvs_host_select_views(hcmvs_host_scene* s, const hcmvs_params* p, int idx);
/* The same with the work inside the one image spread over `threads` threads (what DenseReconstruction does for the first view, the one
 * the GPU waits for); bit-identical to the one-thread form. */
int hcmvs_host_select_views_mt(hcmvs_host_scene* s, const hcmvs_params* p, int idx, int threads);
/* which: 0 = Image::neighbors (all scored), 1 = DepthData::neighbors (filtered). Returns the list length. */
int hcmvs_host_get_neighbors(hcmvs_host_scene* s, int idx, int which, uint32_t* ids, uint32_t* points, float* scale, float* angle, float* area, float* score, int cap);
/* sparse-point initial depth map + depth range of a selected view (SceneDensify.cpp:783-808) */
int hcmvs_host_init_depth(hcmvs_host_scene* s, int idx, float* depth, float* dminmax);
int hcmvs_host_get_gray(hcmvs_host_scene* s, int idx, float* gray);
/* DepthData::ViewData::ScaleImage (DepthMap.h:232-238) on an f32 gray image + the intrinsics of the new resolution (Image::GetCamera).
 * Call with dst == NULL first to learn dw x dh. Returns 1 (nothing written) when |scale-1| < 0.15, like the reference. */
int hcmvs_host_scale_image(const float* src, int sw, int sh, float scale, float* dst, int* dw, int* dh, const double* K, double* Kout);
/* Scene::DenseReconstruction (SceneDensify.cpp:3532-3574) through the C ABI with HOST buffers.
 * stats[8] = sec select, upload, estimate, filter, fuse, h2d bytes, d2h bytes, #points. dmap_dir may be NULL. */
int hcmvs_host_dense_reconstruction(hcmvs_host_scene* s, hcmvs_ctx* ctx, const hcmvs_params* p, uint64_t seed, int run_filter, const char* dmap_dir, double* stats);
/* Page-lock (pin != 0) or release the pixel buffers of the scene's images, so that DenseReconstruction uploads them at PCIe rate.
 * Returns the number of pinned buffers. Unpin before anything re-allocates the images (hcmvs_host_scene_reload_images). */
int hcmvs_host_pin_images(hcmvs_host_scene* s, int pin);
/* The same on `world` GPUs of one box (no reference counterpart): one process per GPU, ctx joined to the communicator
 * (hcmvs_comm_init). Every rank adds every image's camera; pixels are only needed for images with index % world == rank (bgr may be
 * NULL in hcmvs_host_add_image for the others — they arrive over NVLink). Rank 0 ends up with the fused cloud. */
int hcmvs_host_dense_reconstruction_distributed(hcmvs_host_scene* s, hcmvs_ctx* ctx, const hcmvs_params* p, uint64_t seed, int run_filter, int rank, int world, double* stats);
/* The same job in stages (DistributedReconstruction::Prepare / UploadInitial / Run in densify.h), for callers that time the hot path with
 * its inputs resident in HBM: prepare once per scene, upload_initial before every run (estimation overwrites the initial maps),
 * run = estimation + exchanges + filter + fusion on rank 0 (download != 0: the cloud lands in the scene, else it stays on the device). */
int hcmvs_host_dist_prepare(hcmvs_host_scene* s, hcmvs_ctx* ctx, const hcmvs_params* p, int rank, int world);
int hcmvs_host_dist_info(hcmvs_host_scene* s, int rank, int* n_valid, int* n_mine_whole, int* n_split, int* whole_rounds); /* the plan prepare made */
int hcmvs_host_dist_upload_initial(hcmvs_host_scene* s);
int hcmvs_host_dist_run(hcmvs_host_scene* s, uint64_t seed, int run_filter, int download, double* stats);
/* The schedule it follows, for inspection / tests (host only): order[n_valid] = views in FuseDepthMaps' connection order; per view the
 * rank that estimates it whole (HCMVS_OWNER_SPLIT_ROWS: estimated in row bands by all ranks; -1: not in the plan) and the rank that
 * filters it. Arrays of n_views entries; any output may be NULL. */
int hcmvs_host_shard_plan(const uint32_t* valid_views, int n_valid, const uint32_t* n_scored, int n_views, int world, int split_rows,
                          uint32_t* order, int32_t* owner_whole, int32_t* owner_filter, int* whole_rounds, int* n_split);
/* fused cloud access */
int hcmvs_host_cloud_size(hcmvs_host_scene* s, uint64_t* n_points, uint64_t* n_view_refs);
int hcmvs_host_cloud_get(hcmvs_host_scene* s, float* xyz, float* normals, uint8_t* colors, uint32_t* view_offsets, uint32_t* views, float* weights);
int hcmvs_host_cloud_save_ply(hcmvs_host_scene* s, const char* file);
/* Scene::PointCloudFilter(thRemove) (SceneDensify.cpp:4189-4320) on the scene's dense cloud: device votes, removal in the reference's
 * order. Returns the number of removed points, -1 on error. */
long hcmvs_host_pointcloud_filter(hcmvs_host_scene* s, hcmvs_ctx* ctx, int th_remove);
/* its removal step alone, on caller-provided votes (no device needed), and a setter for the scene's dense cloud (CSR view lists;
 * normals / colors / weights may be NULL) */
long hcmvs_host_cloud_remove_by_visibility(hcmvs_host_scene* s, const int32_t* visibility, int th_remove);
int hcmvs_host_cloud_set(hcmvs_host_scene* s, uint64_t n, const float* xyz, const float* normals, const uint8_t* colors, const uint32_t* view_offsets, const uint32_t* views, const float* weights);
/* raw "DR" depth-data files (MVS::ExportDepthDataRaw / ImportDepthDataRaw, DepthMap.cpp:2781-2925) */
int hcmvs_host_write_dmap(const char* file, const char* image_name, const uint32_t* ids, int n_ids, int image_w, int image_h,
                          const double K[9], const double R[9], const double C[3], float dmin, float dmax, int w, int h,
                          const float* depth, const float* normal, const float* conf);
int hcmvs_host_read_dmap_header(const char* file, int* w, int* h, int* n_ids, int* has_normal, int* has_conf);
int hcmvs_host_read_dmap(const char* file, uint32_t* ids, double K[9], double R[9], double C[3], float* dminmax, float* depth, float* normal, float* conf);

/* The fork's depthmap/depthNNNN.dmap and normalmap/normalNNNN.dmap hand-off files between the pyramid levels of run.sh (MVS::SaveDepthMap /
 * LoadDepthMap / SaveNormalMap / LoadNormalMap, DepthMap.cpp:2368-2393): zlib-compressed Boost binary archives, written / parsed from the
 * documented archive layout (byte parity with a Boost build unverified). load: call with a NULL array first to learn w x h. */
int hcmvs_host_save_depthmap(const char* file, const float* depth, int w, int h);
int hcmvs_host_save_normalmap(const char* file, const float* normal, int w, int h);
int hcmvs_host_load_depthmap(const char* file, float* depth, int* w, int* h);
int hcmvs_host_load_normalmap(const char* file, float* normal, int* w, int* h);

/* Triangulated depth-map initialisation, host half (MVS::TriangulatePointsDelaunay, DepthMap.cpp:1797-1876).
 * hcmvs_host_delaunay: Delaunay triangulation of n 2-D points (what CGAL::Delaunay_triangulation_2 computes for the reference);
 * returns the number of faces, writes up to cap_tris index triples (counter-clockwise, smallest index first, sorted).
 * hcmvs_host_triangulate_init: vertices (x, y, depth) + faces for a view selected with hcmvs_host_select_views; call with NULL arrays
 * first to learn the sizes; dminmax = raw depth bounds of the points (InitDepthMap then widens them by 0.9 / 1.1). */
int hcmvs_host_delaunay(const double* xy, int n, uint32_t* tris, int cap_tris);
int hcmvs_host_triangulate_init(hcmvs_host_scene* s, int idx, int add_corners, double* vertices, int cap_vertices, uint32_t* tris, int cap_tris, int* n_vertices, int* n_tris, float* dminmax);

/* MVSI project files ("scene.mvs": MVS::Interface, libs/MVS/Interface.h:165-619) — Scene::LoadInterface / SaveInterface
 * (libs/MVS/Scene.cpp:62-286). load_images: also decode the image files the project names (BMP / PNG / binary PNM), relative to
 * the project's folder. dense: write the fused cloud as the project's vertices (DensifyPointCloud's scene_dense.mvs) instead of
 * the sparse points. version < 0: MVSI_PROJECT_VER (5). */
int hcmvs_host_scene_load_mvs(hcmvs_host_scene* s, const char* file, int load_images);
int hcmvs_host_scene_save_mvs(hcmvs_host_scene* s, const char* file, int version, int dense);
/* --resolution-level / --min-resolution / --max-resolution of DensifyPointCloud (Scene::ComputeDepthMaps, SceneDensify.cpp:3617-3631):
 * shrink every calibrated image to max(w,h) >> level (Image::ResizeImage, cv::resize INTER_AREA on the 8-bit colour image) and update
 * its camera (Image::UpdateCamera). */
int hcmvs_host_scene_reload_images(hcmvs_host_scene* s, unsigned resolution_level, unsigned min_resolution, unsigned max_resolution);
/* the resize itself: cv::resize(src, dst, Size(dw, dh), 0, 0, INTER_AREA), 8-bit BGR, dw <= sw and dh <= sh */
int hcmvs_host_resize_area_bgr(const uint8_t* src, int sw, int sh, int dw, int dh, uint8_t* dst);
int hcmvs_host_num_images(hcmvs_host_scene* s);
int hcmvs_host_get_image_info(hcmvs_host_scene* s, int idx, int* w, int* h, int* calibrated, uint32_t* id, double K[9], double R[9], double C[3], char* name, int name_cap);
int hcmvs_host_get_image_bgr(hcmvs_host_scene* s, int idx, uint8_t* bgr);
/* sparse cloud as CSR; call with NULL arrays first to learn the sizes */
int hcmvs_host_get_sparse(hcmvs_host_scene* s, uint64_t* n_points, uint64_t* n_view_refs, float* xyz, int32_t* offsets, uint32_t* view_ids, float* weights);
/* image decoder used by the project loader; bgr == NULL: only the size */
int hcmvs_host_load_image(const char* file, int* w, int* h, uint8_t* bgr);

#ifdef __cplusplus
}
#endif
#endif
