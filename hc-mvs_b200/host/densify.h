// Host-side mirror of the reference's operator surface over the C ABI (include/hcmvs_b200.h).
//
//   reference                                               here
//   MVS::Scene (libs/MVS/Scene.h:52-116)                    hcmvs_host::Scene
//   MVS::DepthMapsData (libs/MVS/SceneDensify.h:49-88)      hcmvs_host::DepthMapsData  (same method names)
//   Scene::DenseReconstruction (SceneDensify.cpp:3532-3574) hcmvs_host::DenseReconstruction
//
// Everything numeric on the hot path runs in libhcmvs_b200.so (CUDA); this layer keeps what the
// reference keeps on the host: scene bookkeeping, neighbour-view selection (Scene::SelectNeighborViews,
// Scene.cpp:545-662 — host code in the reference too, bit-exact parity required), image preparation,
// the sparse-point depth initialisation, and the .dmap / .ply writers.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "hcmvs_b200.h"

namespace hcmvs_host {

struct Camera {
	double K[9], R[9], C[3], P[12]; // P = K R [I|-C], libs/MVS/Camera.h:46-54
	void ComposeP();
};

struct ViewScore { // MVS::ViewScore = TIndexScore<ViewInfo,float>, libs/MVS/Image.h:52-71
	uint32_t ID, points;
	float scale, angle, area, score;
};

struct Image {
	int width = 0, height = 0;
	Camera camera;
	std::string name;
	std::vector<uint8_t> bgr;    // height*width*3 (may be empty)
	std::vector<float> gray;     // Image::toGray(BGR2GRAY, normalised), Common/Types.inl:2352-2402
	std::vector<ViewScore> neighbors; // all scored neighbours, best first (Image::neighbors)
	float avgDepth = 0;
	uint32_t ID = 0xFFFFFFFFu;   // global image ID of the project file (NO_ID: the index)
	bool calibrated = true;      // Image::IsValid(): poseID != NO_ID (libs/MVS/Image.h)
	double Knorm[4] = {0, 0, 0, 0}; // fx, fy, cx, cy normalised by max(width, height) as the project stores them (Platform::Camera::K)
	bool hasKnorm = false;          // set by LoadInterface; lets a resolution change rebuild K exactly as Image::UpdateCamera does
};

struct SparsePoints { // the part of MVS::PointCloud the densifier reads
	std::vector<float> xyz;                      // 3 per point
	std::vector<std::vector<uint32_t>> views;    // sorted view ids per point
	std::vector<std::vector<float>> weights;     // per-view confidences (empty when the project has none, Scene.cpp:180-181)
	std::vector<float> normals;                  // optional, 3 per point
	std::vector<uint8_t> colors;                 // optional, 3 per point (B,G,R)
	size_t size() const { return views.size(); }
};

// Array that either owns its storage or borrows it (the fused cloud lives in the CUDA context's page-locked arena)
template<typename T>
struct Array {
	Array() = default;
	// an owning array re-points at its own storage when copied or moved; a borrowing one keeps the borrowed pointer
	Array(const Array& o) : p(o.p), n(o.n), own(o.own) { if (!own.empty()) p = own.data(); }
	Array(Array&& o) noexcept : p(o.p), n(o.n), own(std::move(o.own)) { if (!own.empty()) p = own.data(); o.p = nullptr; o.n = 0; }
	Array& operator=(Array o) noexcept { own.swap(o.own); n = o.n; p = own.empty() ? o.p : own.data(); return *this; }
	const T* data() const { return p; }
	T* data() { return own.empty() ? const_cast<T*>(p) : own.data(); }
	size_t size() const { return n; }
	bool empty() const { return n == 0; }
	const T& operator[](size_t i) const { return p[i]; }
	void resize(size_t k) { own.resize(k); p = own.data(); n = k; }
	void borrow(const T* q, size_t k) { own.clear(); own.shrink_to_fit(); p = q; n = k; }
	void clear() { own.clear(); p = nullptr; n = 0; }
private:
	const T* p = nullptr; size_t n = 0; std::vector<T> own;
};

struct PointCloud { // fused output, libs/MVS/PointCloud.h:49-109
	Array<float> points, normals, weights;
	Array<uint8_t> colors;
	Array<uint32_t> viewOffsets, views;
	size_t size() const { return points.size()/3; }
	bool Save(const std::string& fileName) const; // PointCloud::Save, PointCloud.cpp:188-242 (binary little-endian PLY)
};

struct Scene {
	std::vector<Image> images;
	SparsePoints pointcloud;
	PointCloud densecloud;
	unsigned nCalibratedImages() const { unsigned n = 0; for (const Image& im: images) n += im.calibrated; return n; }
	// MVSI project files (mvsi.h): Scene::LoadInterface / SaveInterface, libs/MVS/Scene.cpp:62-286. bDense: export the fused
	// cloud as the project's vertices (what DensifyPointCloud's scene_dense.mvs holds) instead of the sparse points.
	bool LoadInterface(const std::string& fileName, bool bLoadImages = true, std::string* err = nullptr);
	bool SaveInterface(const std::string& fileName, int version = -1, bool bDense = false) const;
	// --resolution-level / --min-resolution / --max-resolution (Scene::ComputeDepthMaps, SceneDensify.cpp:3617-3631): every calibrated
	// image is shrunk to max(w,h) >> level (Image::ResizeImage: cv::resize INTER_AREA on the 8-bit colour image) and its camera updated
	// Scene::PointCloudFilter (SceneDensify.cpp:4189-4320): visibility votes on the device (hcmvs_pointcloud_filter), then the points with
	// visibility <= thRemove are removed in the reference's order (RFOREACH + cList::RemoveAt, which moves the last point into the hole).
	// The context must hold the views' cameras (it does after DenseReconstruction). Returns the number of removed points, -1 on error.
	long PointCloudFilter(hcmvs_ctx* ctx, int thRemove, std::string* err = nullptr);
	bool ReloadImages(unsigned nResolutionLevel, unsigned nMinResolution = 640, unsigned nMaxResolution = 3200, std::string* err = nullptr);
	// Scene::SelectNeighborViews / FilterNeighborViews, libs/MVS/Scene.cpp:545-678
	// nThreads > 1 spreads the work inside ONE image over threads (bit-identical for every thread count): for the view a GPU waits for
	bool SelectNeighborViews(uint32_t ID, std::vector<uint32_t>& points, unsigned nMinViews, unsigned nMinPointViews, float fOptimAngle, unsigned nThreads = 1);
	static bool FilterNeighborViews(std::vector<ViewScore>& neighbors, float fMinArea, float fMinScale, float fMaxScale, float fMinAngle, float fMaxAngle, unsigned nMaxViews);
};

// the removal step of Scene::PointCloudFilter (SceneDensify.cpp:4310-4314): points with visibility <= thRemove, last to first, each hole
// filled by the then-last point (PointCloud::RemovePoint / cList::RemoveAt). Returns the number of removed points.
long RemovePointsByVisibility(PointCloud& pc, const int32_t* visibility, int thRemove);

struct DepthData { // host part of MVS::DepthData, libs/MVS/DepthMap.h:214-347
	std::vector<uint32_t> images;     // [0] = reference, [1..] = matching views
	std::vector<ViewScore> neighbors; // filtered, <= nMaxViews
	std::vector<uint32_t> points;
	float dMin = 0, dMax = 0;
	bool valid = false, uploaded = false;
};

struct ViewSelectionParams { // OPTDENSE fields SelectViews reads, DepthMap.cpp:69-143
	float fViewMinScore = 0.f, fViewMinScoreRatio = 0.3f;
	float fMinArea = 0.01f, fMinAngle = 3.f, fOptimAngle = 10.f, fMaxAngle = 65.f;
};

void ToGray(const uint8_t* bgr, int w, int h, float* gray);
// cv::resize(.., Size(dw, dh), 0, 0, INTER_AREA) on an 8-bit BGR image, shrinking only (Image::ResizeImage, Image.cpp:139-160)
bool ResizeAreaBGR(const uint8_t* src, int sw, int sh, int dw, int dh, uint8_t* dst);
// TImage::computeMaxResolution, Common/Types.inl:2442-2460
unsigned ComputeMaxResolution(unsigned width, unsigned height, unsigned& level, unsigned minImageSize, unsigned maxImageSize);
// DepthData::ViewData::ScaleImage (DepthMap.h:232-238): cv::resize(image, scaled, Size(), scale, scale, scale > 1 ? INTER_CUBIC : INTER_AREA)
// on the f32 gray image, restated from OpenCV's scalar code paths (imgproc/src/resize.cpp; cross-checked with cv2 in the CPU tests).
// Returns false (and leaves dst alone) when |scale-1| < 0.15, like the reference.
bool ScaleImage(const std::vector<float>& src, int sw, int sh, float scale, std::vector<float>& dst, int& dw, int& dh);
// Image::GetCamera(platforms, newSize) (Image.cpp:194-209): the intrinsics of the same camera at another resolution —
// K scales with the normalisation length max(width, height) (Camera.h:105-108, 167-180)
void ScaleK(const double K[9], int w, int h, int newW, int newH, double Kout[9]);
// sparse-point initialisation of a depth map (SceneDensify.cpp:783-808)
void SparseInitDepth(const Scene& scene, uint32_t idxImage, const std::vector<uint32_t>& points, std::vector<float>& depth, float& dMin, float& dMax);

// Delaunay triangulation of 2-D points (xy: 2 per point) — what CGAL::Delaunay_triangulation_2 gives the reference (DepthMap.cpp:1785-1808);
// tris: 3 indices per finite face, counter-clockwise, smallest index first, sorted; adjacency (optional): face across the edge
// opposite each vertex slot, -1 on the hull. Duplicate points are skipped like CGAL's insert().
bool DelaunayTriangulate(const std::vector<double>& xy, std::vector<uint32_t>& tris, std::vector<int>* adjacency = nullptr);
// TriangulatePointsDelaunay (DepthMap.cpp:1797-1876): vertices = (x, y, depth) of the view's sparse points (+ the 4 image corners with
// their interpolated depths when bAddCorners), tris as above, raw depth bounds of the points
bool TriangulateInit(const Scene& scene, uint32_t idxImage, const std::vector<uint32_t>& points, bool bAddCorners,
	std::vector<double>& vertices, std::vector<uint32_t>& tris, float& dMin, float& dMax);

class DepthMapsData {
public:
	DepthMapsData(Scene& scene, hcmvs_ctx* ctx, const hcmvs_params& params, const ViewSelectionParams& vs = ViewSelectionParams());
	bool SelectViews(uint32_t idxImage, unsigned nThreads = 1);            // SceneDensify.cpp:307-327
	bool InitViews(uint32_t idxImage, uint32_t numNeighbors);              // SceneDensify.cpp:336-397
	// SceneDensify.cpp:772-812: nMinViewsTrustPoint < 2 -> sparse points splatted on the host; else (the reference's default) the
	// triangulated sparse cloud (InitDepthMap :514-525 -> TriangulatePoints2DepthMap), triangulated here and rasterised on the device
	bool InitDepthMap(uint32_t idxImage);
	bool EstimateDepthMap(int it_external, uint32_t idxImage, uint64_t seed); // SceneDensify.cpp:758-1072
	bool FilterDepthMap(uint32_t idxImage, const std::vector<uint32_t>& idxNeighbors, bool bAdjust); // SceneDensify.cpp:3006-3259
	bool FuseDepthMaps(PointCloud& pointcloud, bool bEstimateColor, bool bEstimateNormal);           // SceneDensify.cpp:3265-3495
	// the stock OpenMVS post-filters of a view's maps: the speckle filter the fork keeps under `#if 0`
	// (SceneDensify.cpp:1956-2042; nSpeckleSize = OPTDENSE default 100) and the small-gap branch of GapInterpolation (:2294-2352,
	// :2640-2683; nIpolGapSize = 7) on caller-held fuse maps (H*W depth, H*W*3 normal or empty, H*W confidence or empty)
	bool RemoveSmallSegments(uint32_t idxImage, unsigned nSpeckleSize = 100);
	bool GapInterpolation(uint32_t idxImage, std::vector<float>& depthFuse, std::vector<float>& normalFuse, std::vector<float>& conf, unsigned nIpolGapSize = 7);
	// raw "DR" depth-data file, ExportDepthDataRaw, DepthMap.cpp:2781-2846
	bool SaveDepthMapRaw(uint32_t idxImage, const std::string& fileName);
	bool UploadView(uint32_t idxImage);
	std::vector<DepthData> arrDepthData;
	std::string lastError;
private:
	Scene& scene; hcmvs_ctx* ctx; hcmvs_params P; ViewSelectionParams VS;
	bool Fail(const char* what);
};

struct DenseReconstructionStats { double secSelect = 0, secUpload = 0, secEstimate = 0, secFilter = 0, secFuse = 0; uint64_t h2dBytes = 0, d2hBytes = 0; };
// Scene::DenseReconstruction, SceneDensify.cpp:3532-3574 (+ the optional filter stage, :3722-3760)
bool DenseReconstruction(Scene& scene, hcmvs_ctx* ctx, const hcmvs_params& P, const ViewSelectionParams& VS, uint64_t seed, bool runFilter,
	const std::string& dmapDir, DenseReconstructionStats* stats, std::string* err);

// The multi-GPU schedule (no reference counterpart: the reference is single-node CPU code). MakeShardPlan orders the views by
// decreasing nScoredNeighbors, ties by index, and deals them round-robin: view order[k] is estimated in round k / world by rank k % world.
// (DistributedReconstruction passes equal scores, i.e. INDEX order: the plan then needs no view selection, so estimation starts while the
// other views are still being selected.) With splitRows the views of the last, incomplete round are estimated in `world` row
// bands by all ranks (49 views on 8 ranks: 6 whole rounds + 1 view an eighth of which every rank estimates, instead of one rank
// estimating a 7th view while seven wait). Filtering of view order[k] always stays with rank k % world.
struct ShardPlan {
	std::vector<uint32_t> order; int world = 1; int wholeRounds = 0; bool splitRows = false;
	std::vector<uint32_t> SplitViews() const;                 // views of the incomplete round (empty without splitRows)
	std::vector<uint32_t> WholeViewsOf(int rank) const;       // views a rank estimates whole, in round order
	std::vector<uint32_t> ViewsOf(int rank) const;            // views a rank filters
	std::vector<int32_t> RoundOwners(int round, size_t nViews) const;  // owner list of one round for hcmvs_exchange_maps
	std::vector<int32_t> SplitOwners(size_t nViews) const;    // HCMVS_OWNER_SPLIT_ROWS for the split views
	std::vector<int32_t> Owners(size_t nViews, const std::vector<char>* only = nullptr) const;
	void RowsOf(int rank, int height, int& r0, int& r1) const;
};
ShardPlan MakeShardPlan(const std::vector<uint32_t>& validViews, const std::vector<uint32_t>& nScoredNeighbors, int world, bool splitRows);
// Scene::DenseReconstruction on `world` GPUs: one process per GPU calls this with its rank; ctx must have joined the communicator
// (hcmvs_comm_init). Every rank holds every camera and the sparse cloud; image PIXELS are only needed for the images it uploads
// (index % world == rank) — the others arrive over NVLink. Rank 0 returns with scene.densecloud, identical to the single-GPU cloud.
// The same job in stages, for callers that time the hot path with its inputs already resident in HBM (bench.py's `value`):
//   Prepare()        view selection (sharded + all-gathered), image uploads (1/world over PCIe, the rest over NVLink), neighbour lists,
//                    the plan, the initial depth maps on the host
//   UploadInitial()  H2D of the initial maps of the views this rank estimates (estimation overwrites them: once per Run)
//   Run()            estimation rounds + exchanges, FilterDepthMap + exchange + commit, FuseDepthMaps on rank 0
//                    (download: into scene.densecloud; else the cloud stays on the device)
class DistributedReconstruction {
public:
	DistributedReconstruction(Scene& scene, hcmvs_ctx* ctx, const hcmvs_params& P, const ViewSelectionParams& VS, int rank, int world);
	~DistributedReconstruction();
	DistributedReconstruction(const DistributedReconstruction&) = delete;
	DistributedReconstruction& operator=(const DistributedReconstruction&) = delete;
	bool Prepare(bool lazy = false); // lazy (the one-call job): only the images move here; Run() selects views + makes the initial maps on worker threads AHEAD of the estimation
	bool UploadInitial();
	bool Run(uint64_t seed, bool runFilter, bool download);
	const std::string& Error() const;
	const DenseReconstructionStats& Stats() const;
	const ShardPlan& Plan() const;
	size_t ValidViews() const; // views with enough neighbours (known after Prepare(false) / Run)
private:
	struct Impl; Impl* impl;
};
bool DenseReconstructionDistributed(Scene& scene, hcmvs_ctx* ctx, const hcmvs_params& P, const ViewSelectionParams& VS, uint64_t seed, bool runFilter,
	int rank, int world, DenseReconstructionStats* stats, std::string* err);

// depthmap/depthNNNN.dmap, normalmap/normalNNNN.dmap — the fork's hand-off between the pyramid levels of run.sh (MVS::SaveDepthMap /
// LoadDepthMap / SaveNormalMap / LoadNormalMap, DepthMap.cpp:2368-2393): zlib-compressed Boost binary archives of TImage<float> /
// TImage<Point3f>, written / parsed from the documented archive layout (host/boost_dmap.cpp; byte parity with Boost itself unverified)
bool SaveDepthMap(const std::string& fileName, const float* depth, int w, int h);
bool LoadDepthMap(const std::string& fileName, std::vector<float>& depth, int& w, int& h);
bool SaveNormalMap(const std::string& fileName, const float* normal, int w, int h);
bool LoadNormalMap(const std::string& fileName, std::vector<float>& normal, int& w, int& h);

bool ExportDepthDataRaw(const std::string& fileName, const std::string& imageFileName, const std::vector<uint32_t>& IDs, int imageW, int imageH,
	const double K[9], const double R[9], const double C[3], float dMin, float dMax, int w, int h,
	const float* depth, const float* normal, const float* conf);
bool ImportDepthDataRaw(const std::string& fileName, std::string& imageFileName, std::vector<uint32_t>& IDs, int& imageW, int& imageH,
	double K[9], double R[9], double C[3], float& dMin, float& dMax, int& w, int& h,
	std::vector<float>& depth, std::vector<float>& normal, std::vector<float>& conf);

} // namespace hcmvs_host
