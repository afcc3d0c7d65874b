// Internal host-side state of libhcmvs_b200.so (not part of the ABI).
#pragma once
#include "hcmvs_device.cuh"
#include <vector>

enum { ST_SCORE = 0, ST_SWEEPS, ST_END, ST_PREP, ST_FILTER, ST_FUSE, ST_EXCHANGE, ST_COUNT };

struct NbImage { // a matching view's image rescaled for ONE reference view (ViewData::ScaleImage, DepthMap.h:232-238)
	int w = 0, h = 0; double K[9];
	cudaArray_t arr = nullptr; cudaTextureObject_t tex = 0; float* img_d = nullptr;
};

struct View { // one scene image + its DepthData (libs/MVS/DepthMap.h:214-347)
	bool set = false, hasMaps = false;
	int w = 0, h = 0;
	double K[9], R[9], C[3], P[12];
	cudaArray_t arr = nullptr; cudaTextureObject_t tex = 0; // gray image, gather texture
	float* img_d = nullptr;       // gray image, linear
	uint8_t* bgr_d = nullptr;     // colour image (fusion colours, gradient map)
	uint8_t* gra_d = nullptr;     // DepthData::graMap
	bool graValid = false;        // gra_d matches the current image
	float4* dn_d = nullptr;       // (normal.xyz, depth)
	float* conf_d = nullptr;
	float* depth_d = nullptr; bool depthValid = false; // compact copy of dn.w for the filter's splats (4 B/px instead of a 16 B stride); stale after any writer of dn
	float* prior_d = nullptr;     // DepthData::depthMapPrior
	float4* coarse_d = nullptr;   // restore tree: nresize_normalMap / nresize_depthMap packed like dn
	float4* dnPrev_d = nullptr; float* confPrev_d = nullptr; bool hasPrev = false; // maps of the previous outer iteration (viewspread)
	float* fdepth_d = nullptr; float* fconf_d = nullptr; bool hasFiltered = false; // pending FilterDepthMap output
	float dMin = 0.f, dMax = 0.f;
	std::vector<uint32_t> nbIds;  // DepthData::neighbors (sorted by score)
	std::vector<float> nbScores;
	int nMatch = 0;               // first nMatch ids = DepthData::images[1..]
	std::vector<NbImage> nbImages; // per matching slot; w == 0: the neighbour's own image is used
	cudaEvent_t ready = nullptr;  // upload of this view's maps finished (recorded on the copy stream)
	cudaEvent_t lastUse = nullptr;  // last kernel that reads this view's images was queued before this event (compute stream)
	cudaEvent_t imgReady = nullptr; // upload of this view's images finished (copy stream); compute entry points wait on it
	float fusePriority = 0.f; bool hasFusePriority = false; // #scored neighbours (FuseDepthMaps connection score)
};

struct TimedSpan { int stage; cudaEvent_t a, b; };

// one slot of the asynchronous map read-back (hcmvs_download_depthmap_begin / _wait)
struct DownloadSlot {
	void* host = nullptr; size_t hostBytes = 0;   // page-locked: depth | normal | conf
	float* dev = nullptr; size_t devBytes = 0;    // unpacked depth | normal staging (the maps are stored packed)
	cudaEvent_t unpacked = nullptr, landed = nullptr;
	size_t n = 0; float dMin = 0.f, dMax = 0.f; bool pending = false;
};

struct FuseState; // fuse.cu

struct hcmvs_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t copyStream = nullptr;   // H2D uploads + layout kernels, overlapping the compute stream
	bool freshScene = false;             // hcmvs_begin_scene: until the first multi-view consumer, a view's maps are only touched by its own init / estimate calls
	void* upload_d = nullptr; size_t uploadBytes = 0; // staging for uploads (copy stream only)
	hcmvs_params P;
	std::vector<View> views;
	void* scratch_d = nullptr; size_t scratchBytes = 0;
	unsigned long long* counters_d = nullptr;
	std::vector<TimedSpan> timed; std::vector<cudaEvent_t> eventPool;
	double stageMs[ST_COUNT] = {0, 0, 0, 0, 0, 0, 0};
	uint32_t nLaunches = 0;
	uint64_t fuseRounds = 0, fuseSeeds = 0, fuseProbes = 0, filterBytes = 0;
	FuseState* fuse = nullptr;
	SpreadConst* spread_d = nullptr; // viewspread constants of the view being estimated
	void* comm = nullptr; int rank = 0, world = 1; // NCCL communicator of hcmvs_comm_init (exchange.cu)
	void* ctlComm = nullptr; cudaStream_t ctlStream = nullptr; float* ctl_d = nullptr; size_t ctlCap = 0; // pre-exchange agreement (status + depth ranges)
	DownloadSlot dl[HCMVS_DOWNLOAD_SLOTS]; cudaStream_t dlStream = nullptr; // device->host stream of the map read-back
	cudaStream_t commStream = nullptr; cudaEvent_t commDone = nullptr, commReady = nullptr; bool commPending = false; // asynchronous exchanges
};

void hcmvs_set_error(const char* fmt, ...);
int  hcmvs_scratch(hcmvs_ctx* ctx, size_t bytes, void** out);
void hcmvs_time_begin(hcmvs_ctx* ctx, int stage);
void hcmvs_time_end(hcmvs_ctx* ctx);
void hcmvs_fuse_release(hcmvs_ctx* ctx);
void hcmvs_fuse_invalidate_stream(hcmvs_ctx* ctx); // the resident fused cloud was edited in place
void hcmvs_comm_release(hcmvs_ctx* ctx);
void hcmvs_fill_cam(const View& v, CamConst& c);
int  hcmvs_mark_image_use(hcmvs_ctx* ctx, View& v);   // record v.lastUse on the compute stream
int  hcmvs_wait_image(hcmvs_ctx* ctx, const View& v); // make the compute stream wait for the view's image upload

// patchmatch.cu
cudaError_t hcmvs_launch_score_init(const RefConst& rc, bool tex, cudaStream_t st);
cudaError_t hcmvs_launch_score_hyp(const RefConst& rc, const float4* hyp, int smoothMode, float* out, bool tex, cudaStream_t st);
cudaError_t hcmvs_launch_sweep(const RefConst& rc, int colour, bool tex, cudaStream_t st, bool window = false);
cudaError_t hcmvs_launch_end(float4* dn, float* conf, size_t n, float keep, cudaStream_t st);
cudaError_t hcmvs_launch_median3(const float4* in, float4* out, int w, int h, cudaStream_t st);
cudaError_t hcmvs_launch_gramap(const uint8_t* bgr, uint8_t* gra, int w, int h, cudaStream_t st);
cudaError_t hcmvs_launch_pack(const float* depth, const float* normal, float4* dn, size_t n, cudaStream_t st);
cudaError_t hcmvs_launch_resize_area_up(const float* depth, const float* normal, int sw, int sh, float4* dst, int dw, int dh, cudaStream_t st);
cudaError_t hcmvs_launch_minmax_w(const float4* dn, size_t n, float* minmax_d, cudaStream_t st);
// init_tri.cu
cudaError_t hcmvs_launch_raster_triangles(const double* vtx_d, const uint32_t* tris_d, const int3* chunks_d, int nChunks, const double K[9], float4* dn, int* owner_d, int w, int h, cudaStream_t st);
cudaError_t hcmvs_launch_unpack(const float4* dn, float* depth, float* normal, size_t n, cudaStream_t st);
