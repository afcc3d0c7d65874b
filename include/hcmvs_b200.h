/* hcmvs_b200 — C ABI of the B200-native HC-MVS dense-reconstruction hot path.
 *
 * Drop-in boundary for the reference's de-facto operator surface, `class DepthMapsData`
 * (libs/MVS/SceneDensify.h:49-88) and its three thread procs (SceneDensify.h:73-75); the reference has no
 * FFI of its own. Every entry point below names the reference interface it replaces.
 *
 * Conventions: plain pointers and sizes only; all pointers are HOST memory unless suffixed `_d`; row-major
 * images; status `int` (0 ok, <0 error, text via hcmvs_last_error()) mirroring the reference's bool returns;
 * no exceptions cross the boundary. One context drives ONE CUDA device (one process per GPU); calls on a
 * context are thread-compatible (one thread at a time — the reference serialises EstimateDepthMap with a
 * Semaphore(1), SceneDensify.cpp:3503,3895). There is no CPU fallback: creation fails without a CUDA device.
 */
#ifndef HCMVS_B200_H_
#define HCMVS_B200_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define HCMVS_OK 0
#define HCMVS_ERR_ARG (-1)
#define HCMVS_ERR_CUDA (-2)
#define HCMVS_ERR_STATE (-3)
#define HCMVS_ERR_UNSUPPORTED (-4)
#define HCMVS_MAX_MATCH_VIEWS 16   /* matching neighbours per reference view (reference runs use 5..10) */
#define HCMVS_MAX_FUSE_VIEWS 32    /* neighbours probed during fusion (OPTDENSE::nMaxViews = 12) */

/* The OPTDENSE fields the hot path reads (libs/MVS/DepthMap.cpp:69-143). */
typedef struct hcmvs_params {
	uint32_t nNumViews, nMaxViews, nMinViews, nMinViewsTrustPoint;
	uint32_t nMinViewsFuse, nMinViewsFilter, nMinViewsFilterAdjust;
	int32_t  bFilterAdjust;
	float fNCCThresholdKeep;
	uint32_t nEstimationIters, nEstimationIters_external, nRandomIters;
	float fRandomDepthRatio, fRandomAngle1Range, fRandomAngle2Range;
	float fRandomSmoothDepth, fRandomSmoothNormal, fRandomSmoothBonus;
	float fDescriptorMinMagnitudeThreshold;
	float fDepthDiffThreshold, fNormalDiffThreshold, depthweight, normalweight;
	int32_t adapthalfwin, propagatehalfwin, propagatestep, photo2geo;
	float photometric_flow, para_prior, fsigmaPrior;
	/* B200 reformulation knobs (no reference counterpart) */
	int32_t rb_far_reach;   /* red-black propagation: per direction the best-confidence pixel among odd offsets 1..rb_far_reach */
	int32_t rb_prop_dirs;   /* 2 (default): one source per image axis = the reference's 2 propagation hypotheses per pixel-iteration; 4: one per direction */
	int32_t sampler;        /* 0 = texture gather path (default), 1 = global-memory loads, 2 = texture gathers + shared-memory windows for the sweeps */
	int32_t viewspread;     /* OPTDENSE::viewspread (DepthMap.cpp:102): cross-view propagation at outer iterations >= 1 (DepthMap.cpp:1504-1608); 0 in every shipped run */
} hcmvs_params;

typedef struct hcmvs_ctx hcmvs_ctx;

/* Fused point cloud (libs/MVS/PointCloud.h:49-109). Arrays are allocated by the library; release with hcmvs_free_pointcloud. */
typedef struct hcmvs_pointcloud {
	uint64_t n_points;
	float*   points;      /* n*3 */
	float*   normals;     /* n*3 or NULL */
	uint8_t* colors;      /* n*3 in the image's channel order, or NULL */
	uint32_t* view_offsets; /* n+1: CSR offsets into views/weights */
	uint32_t* views;      /* sorted view ids per point (PointCloud::pointViews) */
	float*   weights;     /* PointCloud::pointWeights */
} hcmvs_pointcloud;

/* Per-stage device times (CUDA events on the context's stream) and work counters since the last reset. */
typedef struct hcmvs_timers {
	double ms_score, ms_sweeps, ms_end, ms_prep, ms_filter, ms_fuse;
	uint64_t n_hypotheses;     /* ScorePixel evaluations inside the sweeps */
	uint64_t n_pixel_iters;    /* pixels processed x PatchMatch iterations inside the sweeps */
	uint64_t n_view_scores;    /* ScorePixelImage evaluations inside the sweeps */
	uint64_t n_smooth_terms;   /* smoothness-neighbour terms evaluated inside the sweeps (sum over hypotheses) */
	uint32_t n_launches;       /* kernels launched */
	uint64_t n_fuse_rounds;    /* reserve/commit rounds of the last hcmvs_fuse_depthmaps */
	uint64_t n_window_walks;   /* sampler 2: warp-level (hypothesis, view) patch walks served from the shared-memory windows */
	double ms_exchange;        /* hcmvs_exchange_maps (NCCL) */
	/* work of the HBM-bound stages, for their rooflines */
	uint64_t filter_bytes;     /* sum over hcmvs_filter_depthmap calls of (24 n + 16) B x reference pixels, n = neighbour maps */
	uint64_t fuse_seeds;       /* seed pixels scanned by the last hcmvs_fuse_depthmaps */
	uint64_t fuse_probes;      /* (seed, neighbour view) probes of the last hcmvs_fuse_depthmaps */
} hcmvs_timers;

void hcmvs_default_params(hcmvs_params* p);              /* OPTDENSE defaults, DepthMap.cpp:69-143 */
const char* hcmvs_last_error(void);

/* DepthMapsData::DepthMapsData(Scene&) — SceneDensify.cpp:158-166 */
hcmvs_ctx* hcmvs_create(int device, const hcmvs_params* p);
void hcmvs_destroy(hcmvs_ctx* ctx);
int  hcmvs_set_params(hcmvs_ctx* ctx, const hcmvs_params* p);
int  hcmvs_sync(hcmvs_ctx* ctx);

/* Scene image + camera upload: what DepthMapsData::InitViews prepares per view (SceneDensify.cpp:336-397):
 * gray = Image::toGray(BGR2GRAY, normalised) in [0,1] (Common/Types.inl:2352-2402); bgr may be NULL
 * (then the gradient map is zero and fused colours are omitted). K must have zero skew. */
int hcmvs_set_view(hcmvs_ctx* ctx, uint32_t view, int W, int H, const double K[9], const double R[9], const double C[3],
                   const float* gray, const uint8_t* bgr);
/* Result of DepthMapsData::SelectViews (SceneDensify.cpp:307-327): ids[0..n_all) sorted by score; the first
 * n_match are the matching views InitViews keeps (SceneDensify.cpp:361-376); all n_all are probed by fusion. */
int hcmvs_set_neighbors(hcmvs_ctx* ctx, uint32_t ref, const uint32_t* ids, const float* scores, int n_match, int n_all);
/* Optional, once per scene on a context that is reused: waits for everything queued on the context, then declares that until the
 * first hcmvs_filter_depthmap / hcmvs_commit_filtered / hcmvs_fuse_depthmaps / snapshot / export call a view's maps are only touched by
 * ITS OWN hcmvs_init_depthmap* / hcmvs_estimate_depthmap* / hcmvs_end_depthmap / hcmvs_download_depthmap_begin calls (what
 * DepthMapsData::EstimateDepthMap does, SceneDensify.cpp:772-1056). The upload of view i+1's initial maps then waits only for that
 * view's own last use instead of for everything queued so far (the estimation of view i), i.e. it overlaps the kernels. Without
 * this call every re-initialisation is conservatively ordered after all queued work. */
int hcmvs_begin_scene(hcmvs_ctx* ctx);
/* it_external==0 initialisation of EstimateDepthMap (SceneDensify.cpp:772-819): caller-provided rough depth
 * (0 = unknown), optional normals, depth range; builds the gradient map (InitGraMap, :581-595) on device. */
int hcmvs_init_depthmap(hcmvs_ctx* ctx, uint32_t ref, const float* depth0, const float* normal0, float dMin, float dMax);
/* The reference's default initialisation (nMinViewsTrustPoint >= 2): DepthMapsData::InitDepthMap -> TriangulatePoints2DepthMap
 * (SceneDensify.cpp:514-525, DepthMap.cpp:1879-1936). The caller triangulates the projected sparse points on the host as the reference
 * does (vertices: n_vertices x (x, y, depth) f64 in pixels; tris: n_tris x 3 indices, counter-clockwise like CGAL faces, drawn in this
 * order); the device rasterises every triangle with the reference's 28.4 fixed-point rasteriser and gives each covered pixel the depth
 * of its viewing ray on the triangle's plane and the plane's normal. Also builds the gradient map like hcmvs_init_depthmap. */
int hcmvs_init_depthmap_triangles(hcmvs_ctx* ctx, uint32_t ref, const double* vertices, int n_vertices, const uint32_t* tris, int n_tris, float dMin, float dMax);
/* Load finished maps (DepthData::Load / IncRef, DepthMap.cpp:231-302) — used before filter/fuse-only runs. */
int hcmvs_set_depthmap(hcmvs_ctx* ctx, uint32_t view, const float* depth, const float* normal, const float* conf, float dMin, float dMax);
int hcmvs_get_depthmap(hcmvs_ctx* ctx, uint32_t view, float* depth, float* normal, float* conf, float* dMin, float* dMax);
/* Asynchronous read-back of a view's maps into page-locked memory owned by the context — what a .dmap writer should be fed from
 * (SaveDepthMap / ExportDepthDataRaw after the last outer iteration, SceneDensify.cpp:3984-3989, DepthMap.cpp:2781-2846).
 * _begin queues, behind the work already queued for the view, an unpack on the compute stream and a device->host copy on a stream of its
 * own into slot `slot` (0 .. HCMVS_DOWNLOAD_SLOTS-1), and returns at once: later kernels overlap the copy. _wait blocks until that copy
 * has landed and returns pointers to depth (H*W), normal (H*W*3) and conf (H*W) inside the slot, valid until the slot is used again.
 * _wait touches nothing but its slot and may be called from another thread than the one driving the context (a writer thread). */
#define HCMVS_DOWNLOAD_SLOTS 4
int hcmvs_download_depthmap_begin(hcmvs_ctx* ctx, uint32_t view, int slot);
int hcmvs_download_depthmap_wait(hcmvs_ctx* ctx, int slot, const float** depth, const float** normal, const float** conf, float* dMin, float* dMax);
/* DepthData::graMap (u8, H*W) built by hcmvs_init_depthmap (InitGraMap, SceneDensify.cpp:581-595). */
int hcmvs_get_gradient_map(hcmvs_ctx* ctx, uint32_t view, uint8_t* gra);
/* Optional plane prior (DepthData::depthMapPrior, DepthMap.cpp:941-955); NULL clears it. */
int hcmvs_set_prior(hcmvs_ctx* ctx, uint32_t ref, const float* prior);

/* DepthMapsData::ScoreDepthMapTmp (PASS A) — SceneDensify.cpp:649-675 */
int hcmvs_score_depthmap(hcmvs_ctx* ctx, uint32_t ref, int it_external, uint64_t seed);
/* DepthMapsData::EstimateDepthMap — SceneDensify.cpp:758-1072: median blur, PASS A, nEstimationIters red-black
 * sweeps (EstimateDepthMapTmp / DepthEstimator::ProcessPixel, DepthMap.cpp:1050-1501) and, on the last outer
 * iteration, EndDepthMapTmp (SceneDensify.cpp:688-744). */
int hcmvs_estimate_depthmap(hcmvs_ctx* ctx, uint32_t ref, int it_external, uint64_t seed);
/* The same for the rows [row_begin, row_end) of the view only: a band plus the halo its dependencies reach (2 x iterations x the
 * propagation reach) is estimated, which reproduces those rows of the full-image result bit for bit. Used to split a view between
 * GPUs (every rank a band, no exchange inside the estimation); rows outside the band hold by-products afterwards. Same preconditions
 * as hcmvs_estimate_depthmap: the initial maps of the whole view must be present. */
int hcmvs_estimate_depthmap_rows(hcmvs_ctx* ctx, uint32_t ref, int it_external, uint64_t seed, int row_begin, int row_end);
/* DepthMapsData::EndDepthMapTmp (PASS C) alone — SceneDensify.cpp:688-744 */
int hcmvs_end_depthmap(hcmvs_ctx* ctx, uint32_t ref);
/* Parity hook: DepthEstimator::ScorePixel (DepthMap.cpp:987-1046) for caller-fixed per-pixel hypotheses.
 * smooth_mode 0: no smoothness neighbours (PASS A semantics); 1: the 4-neighbourhood of the given maps. */
int hcmvs_score_hypotheses(hcmvs_ctx* ctx, uint32_t ref, const float* depth, const float* normal, int smooth_mode, float* score_out);

/* DepthMapsData::FilterDepthMap — SceneDensify.cpp:3006-3259. nb_idx index DepthData::neighbors of `ref`.
 * Outputs (either may be NULL) are the filtered depth / confidence maps the reference writes to
 * depthNNNN.filtered.dmap/.cmap; they also stay pending on the device until hcmvs_commit_filtered. */
int hcmvs_filter_depthmap(hcmvs_ctx* ctx, uint32_t ref, const uint32_t* nb_idx, int n, int bAdjust,
                          float* out_depth, float* out_conf);
/* Replace every filtered view's depth/confidence by its pending filtered maps — the second phase of
 * Scene::DenseReconstructionFilter (SceneDensify.cpp:4146-4178: all views are filtered from the
 * un-filtered maps first, then the .filtered files replace the originals). */
int hcmvs_commit_filtered(hcmvs_ctx* ctx);
/* FuseDepthMaps processes views in decreasing order of scene.images[i].neighbors.GetSize() — the UN-filtered
 * scored-neighbour count of Scene::SelectNeighborViews (SceneDensify.cpp:3286-3303). Defaults to n_all. */
int hcmvs_set_fuse_priority(hcmvs_ctx* ctx, uint32_t view, float score);
/* DepthMapsData::FuseDepthMaps — SceneDensify.cpp:3265-3495 */
int hcmvs_fuse_depthmaps(hcmvs_ctx* ctx, int estimate_color, int estimate_normal, hcmvs_pointcloud* out);
void hcmvs_free_pointcloud(hcmvs_pointcloud* pc);
/* hcmvs_fuse_depthmaps with out == NULL leaves the cloud on the device; this returns its device arrays
 * (points/normals float[3n], colors u8[3n], view_offsets u32[n+1], views u32[m], weights float[m]). */
/* DepthData::ViewData::ScaleImage (DepthMap.h:232-238) as used by InitViews (SceneDensify.cpp:370-376): when a matching view's
 * footprint differs from the reference view's by |scale-1| >= 0.15 the reference matches against a RESIZED copy of that image with
 * the camera of the new resolution. `slot` is the 0-based index into the matching views given to hcmvs_set_neighbors; gray is the
 * resized image (H x W, [0,1]) and K its intrinsics (R, C stay the neighbour's). gray == NULL returns to the neighbour's own image.
 * Only the PatchMatch cost reads it; filter and fusion work on the neighbour's maps at their own resolution, as in the reference. */
int hcmvs_set_neighbor_image(hcmvs_ctx* ctx, uint32_t ref, int slot, int W, int H, const double K[9], const float* gray);

/* restore tree (restore/libs/MVS/SceneDensify.cpp:513-532, DepthMap.cpp:1527-1550): the previous pyramid level's estimate of this
 * view, wc x hc (depth f32, normal 3 x f32, camera frame). The maps are resized on the device to the view's size exactly as
 * cv::resize(.., INTER_AREA) enlarges them, [dMin, dMax) is widened by the resized depths (call after hcmvs_init_depthmap), and
 * on the last PatchMatch iteration of the last outer iteration every pixel scores that estimate as one more hypothesis, kept
 * unless clearly worse (conf > nconf - 0.1). depth == NULL removes it. */
int hcmvs_set_coarse_estimate(hcmvs_ctx* ctx, uint32_t view, int wc, int hc, const float* depth, const float* normal);
/* the resized coarse maps (nresize_depthMap / nresize_normalMap) back on the host, for parity checks; either pointer may be NULL */
int hcmvs_get_coarse_estimate(hcmvs_ctx* ctx, uint32_t view, float* depth, float* normal);
/* Keep a copy of every view's (normal, depth, confidence) maps as "the previous outer iteration": what viewspread
 * (hcmvs_params.viewspread) reads from the neighbour views. Call between outer iterations — after the map exchange when
 * the scene is sharded over GPUs — so that the result does not depend on the order or placement of the views. */
int hcmvs_snapshot_maps(hcmvs_ctx* ctx);
/* the opposite copy: every view's maps become the snapshot again (FuseDepthMaps zeroes occluded depths in place — a caller that fuses the
 * same maps repeatedly restores them on the device instead of uploading them again) */
int hcmvs_restore_snapshot(hcmvs_ctx* ctx);

/* Copy the last fused cloud into caller-owned host arrays (sizes from hcmvs_get_fused_device); any pointer may be NULL. */
int hcmvs_download_fused(hcmvs_ctx* ctx, float* points, float* normals, uint8_t* colors, uint32_t* view_offsets, uint32_t* views, float* weights);
/* Same, into a page-locked host arena OWNED BY THE CONTEXT (PCIe-rate copy, no per-scene pinning): `out` receives pointers that
 * stay valid until the next hcmvs_fuse_depthmaps / hcmvs_download_fused_pinned on this context or hcmvs_destroy.
 * Do NOT pass `out` to hcmvs_free_pointcloud. Replaces the PointCloud& output of FuseDepthMaps (SceneDensify.cpp:3265). */
int hcmvs_download_fused_pinned(hcmvs_ctx* ctx, hcmvs_pointcloud* out);
/* The fork's RemoveSmallSegments (SceneDensify.cpp:1953-2276, live branch): a whole-scene fusion whose only product is, per view,
 * depthMap_fuse / normalMap_fuse = the estimate where the pixel became part of a fused point, 0 elsewhere (:2228-2260) — the input
 * of GapInterpolation. After hcmvs_fuse_depthmaps this returns those two maps for `view` (H*W and H*W*3 floats; either may be NULL). */
int hcmvs_get_fused_support(hcmvs_ctx* ctx, uint32_t view, float* depth_fuse, float* normal_fuse);
/* MVS::EstimatePointColors (DepthMap.cpp:2125-2161; `--estimate-colors 1`, SceneDensify.cpp:3568-3569): per point the colour sampled in
 * the nearest view (Camera::PointDepth) among those that see it and hold a colour image; white when none or when it projects into the
 * 1-pixel border. points == NULL: recolour the fused cloud resident on the device (then `colors` may be NULL too); otherwise n_points
 * host points with their CSR view lists -> colors (3 bytes per point, B G R like the images). */
int hcmvs_estimate_point_colors(hcmvs_ctx* ctx, uint64_t n_points, const float* points, const uint32_t* view_offsets, const uint32_t* views, uint8_t* colors);
/* MVS::EstimatePointNormals (DepthMap.cpp:2221-2269; `--estimate-normals 1`, SceneDensify.cpp:3570-3571): per point the least-variance
 * direction of its num_neighbors (reference default 16) nearest neighbours plus itself — what CGAL::pca_estimate_normals computes —
 * flipped to face the first view that sees the point. points == NULL: the fused cloud resident on the device (its normal buffer is
 * overwritten; `normals` may be NULL); otherwise host points with CSR view lists -> normals (3 floats per point). The neighbour search
 * is bounded (about 25 mean point spacings): an isolated outlier with fewer neighbours inside gets the fit of those it has, or 0. */
int hcmvs_estimate_point_normals(hcmvs_ctx* ctx, uint64_t n_points, const float* points, const uint32_t* view_offsets, const uint32_t* views, int num_neighbors, float* normals);
/* Scene::PointCloudFilter (SceneDensify.cpp:4189-4320; DensifyPointCloud --filter-point-cloud < 0): the visibility vote of every
 * (point, view) viewing cone over the whole cloud. visibility[i] is the reference's signed count; the caller removes the points with
 * visibility <= thRemove (hcmvs_host::Scene::PointCloudFilter does, in the reference's order). points == NULL: the fused cloud
 * resident on the device. stats3 (optional): [0] cones tested against every point (fallback), [1] unused, [2] candidate tests. */
int hcmvs_pointcloud_filter(hcmvs_ctx* ctx, uint64_t n_points, const float* points, const uint32_t* view_offsets, const uint32_t* views, int32_t* visibility, uint64_t* stats3);
int hcmvs_get_fused_device(hcmvs_ctx* ctx, uint64_t* n_points, uint64_t* n_view_refs, void** points_d, void** normals_d, void** colors_d,
                           void** view_offsets_d, void** views_d, void** weights_d);

/* ---- per-view post-filters between estimation and fusion
 * DepthMapsData::RemoveSmallSegments as OpenMVS ships it — the body the fork keeps under `#if 0`, SceneDensify.cpp:1956-2042: segments of
 * 4-connected pixels with IsDepthSimilar depths (threshold fDepthDiffThreshold * 0.7) that hold fewer than speckle_size (OPTDENSE::nSpeckleSize,
 * 100) pixels are zeroed in the view's maps (depth, normal, confidence). Identical to the CPU's order-dependent flood fill. */
int hcmvs_remove_small_segments(hcmvs_ctx* ctx, uint32_t view, unsigned speckle_size, uint64_t* n_removed);
/* The small-gap branch of DepthMapsData::GapInterpolation (SceneDensify.cpp:2294-2352, 2640-2683; OPTDENSE::nIpolGapSize = 7): runs of at
 * most gap_size invalid pixels between two similar depths (threshold fDepthDiffThreshold * 2.5) are interpolated, rows first, then columns.
 * Works on the caller's maps in place (the reference applies it to depthMap_fuse / normalMap_fuse — hcmvs_get_fused_support — and confMap):
 * depth H*W, normal H*W*3 or NULL, conf H*W or NULL. The fork's large-gap branches are undefined behaviour and not built. */
int hcmvs_gap_interpolation(hcmvs_ctx* ctx, int w, int h, float* depth, float* normal, float* conf, unsigned gap_size, uint64_t* n_filled);

/* ---- multi-GPU: one process (one context) per GPU, views sharded by reference view. No reference counterpart: the
 * reference is single-node CPU code. Rank 0 creates an id (hcmvs_comm_unique_id), the host distributes its HCMVS_COMM_ID_BYTES bytes
 * by any means (file, socket, MPI, torch.distributed), every rank calls hcmvs_comm_init (collective). hcmvs_exchange_maps then
 * broadcasts, in place and in one NCCL group, the maps of every view i from rank owner[i] (owner[i] < 0: view not exchanged) to all
 * other ranks, on the context's stream:
 *   HCMVS_EXCHANGE_ESTIMATED  (normal, depth) + confidence of the estimated maps (20 B/px) — before FilterDepthMap / FuseDepthMaps
 *   HCMVS_EXCHANGE_FILTERED   the pending FilterDepthMap output (depth + confidence, 8 B/px); hcmvs_commit_filtered then applies
 *                             it on every rank
 * Every rank passes the same owner list (n_views entries, indexed by view id) and must hold every view's image (hcmvs_set_view). */
#define HCMVS_COMM_ID_BYTES 128
#define HCMVS_EXCHANGE_ESTIMATED 0
#define HCMVS_EXCHANGE_FILTERED 1
#define HCMVS_EXCHANGE_IMAGES 2      /* the gray + colour images: rank owner[i] uploaded view i from its host (hcmvs_set_view), the others only
                                        allocated it (hcmvs_set_view_remote) and receive it over NVLink instead of PCIe — every rank
                                        uploads 1/world of the scene and ends up holding all of it */
#define HCMVS_OWNER_SPLIT_ROWS (-2)  /* owner[i]: view i was estimated in `world` row bands (hcmvs_estimate_depthmap_rows), band r =
                                        rows [r*H/world, (r+1)*H/world) on rank r; HCMVS_EXCHANGE_ESTIMATED broadcasts every band from its rank */
#define HCMVS_EXCHANGE_ASYNC 0x100   /* OR into `what`: run the broadcasts on a communication stream, behind the work queued so far and
                                        overlapping the work queued afterwards (the next view's sweeps); hcmvs_exchange_wait joins */
int hcmvs_comm_unique_id(void* id128);
int hcmvs_comm_init(hcmvs_ctx* ctx, const void* id128, int rank, int world);
int hcmvs_exchange_maps(hcmvs_ctx* ctx, const int32_t* owner, uint32_t n_views, int what);
int hcmvs_exchange_wait(hcmvs_ctx* ctx);   /* the compute stream waits for every asynchronous exchange issued so far */
/* hcmvs_set_view without pixels: camera + buffers of a view whose image another rank uploads (HCMVS_EXCHANGE_IMAGES delivers it). */
int hcmvs_set_view_remote(hcmvs_ctx* ctx, uint32_t view, int W, int H, const double K[9], const double R[9], const double C[3], int has_bgr);
/* All-gather of small host records (view-selection results of a sharded Scene::SelectNeighborViews pass ...): every rank contributes
 * `bytes_per_rank` bytes, `recv` receives world x bytes_per_rank in rank order. Collective; staged through device memory. */
int hcmvs_comm_allgather_host(hcmvs_ctx* ctx, const void* send, void* recv, uint64_t bytes_per_rank);

/* Device pointers of a view's maps for GPU<->GPU exchange by the host plumbing (NCCL / P2P):
 * dn_d = float4 (nx,ny,nz,depth) per pixel, conf_d = float per pixel. */
int hcmvs_get_depthmap_device(hcmvs_ctx* ctx, uint32_t view, void** dn_d, void** conf_d, float* dMin, float* dMax);
int hcmvs_set_depth_range(hcmvs_ctx* ctx, uint32_t view, float dMin, float dMax);
/* Device-to-device copies of a view's maps out of / into caller-owned device buffers (e.g. NCCL send/receive
 * slots): dn_d holds H*W float4 (nx,ny,nz,depth), conf_d holds H*W float. Enqueued on the context stream. */
int hcmvs_export_maps_d(hcmvs_ctx* ctx, uint32_t view, void* dn_d, void* conf_d);
int hcmvs_import_maps_d(hcmvs_ctx* ctx, uint32_t view, const void* dn_d, const void* conf_d, float dMin, float dMax);
/* Allocate (zeroed) device maps for a view without uploading — receive buffers for the exchange. */
int hcmvs_alloc_depthmap(hcmvs_ctx* ctx, uint32_t view);

int hcmvs_get_timers(hcmvs_ctx* ctx, hcmvs_timers* t);
int hcmvs_reset_timers(hcmvs_ctx* ctx);
/* Page-lock / release caller-owned host memory (cudaHostRegister): uploads from registered buffers run at PCIe rate and asynchronously. */
int hcmvs_pin_host_memory(void* p, uint64_t bytes);
int hcmvs_unpin_host_memory(void* p);
/* CUDA stream the context launches on (cudaStream_t as void*), for event timing by the caller. */
void* hcmvs_stream(hcmvs_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
