// C ABI implementation (include/hcmvs_b200.h): context, device-resident scene, stage orchestration.
// Host-side mirror of DepthMapsData's state (libs/MVS/SceneDensify.h:49-88): `views` plays the role of
// arrDepthData (+ scene.images cameras); every compute step is a CUDA kernel — there is no CPU path.
#include "hcmvs_device.cuh"
#include "hcmvs_internal.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <string>
#include <vector>
#include <algorithm>

static thread_local std::string g_lastError;
void hcmvs_set_error(const char* fmt, ...) {
	char buf[1024]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
	g_lastError = buf;
}
extern "C" const char* hcmvs_last_error(void) { return g_lastError.c_str(); }

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { hcmvs_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return HCMVS_ERR_CUDA; } } while (0)

extern "C" void hcmvs_default_params(hcmvs_params* p) {
	// OPTDENSE defaults, libs/MVS/DepthMap.cpp:69-143 (CLI defaults where the CLI overrides them)
	std::memset(p, 0, sizeof(*p));
	p->nNumViews = 5; p->nMaxViews = 12; p->nMinViews = 2; p->nMinViewsTrustPoint = 2;
	p->nMinViewsFuse = 2; p->nMinViewsFilter = 2; p->nMinViewsFilterAdjust = 1; p->bFilterAdjust = 1;
	p->fNCCThresholdKeep = 0.55f;
	p->nEstimationIters = 3; p->nEstimationIters_external = 1; p->nRandomIters = 6;
	p->fRandomDepthRatio = 0.003f; p->fRandomAngle1Range = 16.f; p->fRandomAngle2Range = 10.f;
	p->fRandomSmoothDepth = 0.02f; p->fRandomSmoothNormal = 13.f; p->fRandomSmoothBonus = 0.93f;
	p->fDescriptorMinMagnitudeThreshold = 0.01f;
	p->fDepthDiffThreshold = 0.01f; p->fNormalDiffThreshold = 25.f; p->depthweight = 1.f; p->normalweight = 1.f;
	p->adapthalfwin = 5; p->propagatehalfwin = 1; p->propagatestep = 4; p->photo2geo = 2;
	p->photometric_flow = 0.f; p->para_prior = 0.3f; p->fsigmaPrior = 0.2f;
	p->rb_far_reach = 11; p->rb_prop_dirs = 2; p->sampler = 0; p->viewspread = 0;
}

// ------------------------------------------------------------------------------------------------ context
static int CheckParams(const hcmvs_params& p) {
	if (p.adapthalfwin < 1 || p.adapthalfwin > 7) { hcmvs_set_error("adapthalfwin must be in [1,7] (reference nTexels = 64, DepthMap.h:358)"); return HCMVS_ERR_ARG; }
	// the refinement walks scaleRanges[] (12 entries, DepthMap.cpp:384) from index <= 2 by at most nRandomIters steps
	if (p.nRandomIters > 9 || p.nEstimationIters > 60) { hcmvs_set_error("iteration counts out of range (nRandomIters <= 9: scaleRanges has 12 entries, DepthMap.cpp:384)"); return HCMVS_ERR_ARG; }
	if (p.rb_far_reach < 1) { hcmvs_set_error("rb_far_reach must be >= 1"); return HCMVS_ERR_ARG; }
	if (p.rb_prop_dirs != 2 && p.rb_prop_dirs != 4) { hcmvs_set_error("rb_prop_dirs must be 2 or 4"); return HCMVS_ERR_ARG; }
	if (!(p.fNCCThresholdKeep > 0.f)) { hcmvs_set_error("fNCCThresholdKeep must be > 0"); return HCMVS_ERR_ARG; }
	return HCMVS_OK;
}

extern "C" hcmvs_ctx* hcmvs_create(int device, const hcmvs_params* p) {
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0) { hcmvs_set_error("no CUDA device: %s (hcmvs_b200 has no CPU path)", cudaGetErrorString(e)); return nullptr; }
	if (device < 0 || device >= n) { hcmvs_set_error("device %d out of range (%d devices)", device, n); return nullptr; }
	hcmvs_params P; if (p) P = *p; else hcmvs_default_params(&P);
	if (CheckParams(P) != HCMVS_OK) return nullptr;
	if (cudaSetDevice(device) != cudaSuccess) { hcmvs_set_error("cudaSetDevice failed"); return nullptr; }
	hcmvs_ctx* ctx = new hcmvs_ctx();
	ctx->device = device; ctx->P = P;
	if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaMalloc(&ctx->counters_d, 8*sizeof(unsigned long long)) != cudaSuccess ||
	    cudaMemset(ctx->counters_d, 0, 8*sizeof(unsigned long long)) != cudaSuccess) {
		hcmvs_set_error("context allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
		delete ctx; return nullptr;
	}
	return ctx;
}

static void FreeNbImage(NbImage& o) {
	if (o.tex) cudaDestroyTextureObject(o.tex);
	if (o.arr) cudaFreeArray(o.arr);
	cudaFree(o.img_d);
	o = NbImage();
}
static void FreeView(View& v) {
	for (NbImage& o: v.nbImages) FreeNbImage(o);
	if (v.tex) cudaDestroyTextureObject(v.tex);
	if (v.arr) cudaFreeArray(v.arr);
	if (v.ready) cudaEventDestroy(v.ready);
	if (v.imgReady) cudaEventDestroy(v.imgReady);
	if (v.lastUse) cudaEventDestroy(v.lastUse);
	cudaFree(v.img_d); cudaFree(v.bgr_d); cudaFree(v.gra_d); cudaFree(v.dn_d); cudaFree(v.conf_d); cudaFree(v.depth_d); cudaFree(v.prior_d); cudaFree(v.coarse_d); cudaFree(v.dnPrev_d); cudaFree(v.confPrev_d); cudaFree(v.fdepth_d); cudaFree(v.fconf_d);
	v = View();
}

extern "C" void hcmvs_destroy(hcmvs_ctx* ctx) {
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	for (View& v: ctx->views) FreeView(v);
	for (auto& te: ctx->timed) { cudaEventDestroy(te.a); cudaEventDestroy(te.b); }
	for (cudaEvent_t ev: ctx->eventPool) cudaEventDestroy(ev);
	cudaFree(ctx->scratch_d); cudaFree(ctx->counters_d); cudaFree(ctx->upload_d); cudaFree(ctx->spread_d);
	if (ctx->dlStream) { cudaStreamSynchronize(ctx->dlStream); cudaStreamDestroy(ctx->dlStream); ctx->dlStream = nullptr; }
	for (DownloadSlot& d: ctx->dl) { if (d.host) cudaFreeHost(d.host); cudaFree(d.dev); if (d.unpacked) cudaEventDestroy(d.unpacked); if (d.landed) cudaEventDestroy(d.landed); }
	if (ctx->copyStream) { cudaStreamSynchronize(ctx->copyStream); cudaStreamDestroy(ctx->copyStream); }
	hcmvs_fuse_release(ctx);
	hcmvs_comm_release(ctx);
	cudaStreamDestroy(ctx->stream);
	delete ctx;
}

extern "C" int hcmvs_set_params(hcmvs_ctx* ctx, const hcmvs_params* p) {
	if (!ctx || !p) { hcmvs_set_error("null argument"); return HCMVS_ERR_ARG; }
	const int r = CheckParams(*p); if (r != HCMVS_OK) return r;
	ctx->P = *p; return HCMVS_OK;
}
extern "C" int hcmvs_pin_host_memory(void* p, uint64_t bytes) {
	if (!p || !bytes) { hcmvs_set_error("null argument"); return HCMVS_ERR_ARG; }
	CK(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault));
	return HCMVS_OK;
}
extern "C" int hcmvs_unpin_host_memory(void* p) {
	if (!p) return HCMVS_OK;
	CK(cudaHostUnregister(p));
	return HCMVS_OK;
}
extern "C" int hcmvs_sync(hcmvs_ctx* ctx) { if (!ctx) return HCMVS_ERR_ARG; cudaSetDevice(ctx->device); CK(cudaStreamSynchronize(ctx->stream)); return HCMVS_OK; }
extern "C" void* hcmvs_stream(hcmvs_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int hcmvs_scratch(hcmvs_ctx* ctx, size_t bytes, void** out) {
	if (ctx->scratchBytes < bytes) {
		CK(cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->scratch_d); ctx->scratch_d = nullptr; ctx->scratchBytes = 0;
		CK(cudaMalloc(&ctx->scratch_d, bytes));
		ctx->scratchBytes = bytes;
	}
	*out = ctx->scratch_d; return HCMVS_OK;
}

// ---- stage timing with CUDA events on the context stream
static cudaEvent_t GetEvent(hcmvs_ctx* ctx) {
	if (!ctx->eventPool.empty()) { cudaEvent_t e = ctx->eventPool.back(); ctx->eventPool.pop_back(); return e; }
	cudaEvent_t e; cudaEventCreate(&e); return e;
}
void hcmvs_time_begin(hcmvs_ctx* ctx, int stage) {
	TimedSpan ts; ts.stage = stage; ts.a = GetEvent(ctx); ts.b = GetEvent(ctx);
	cudaEventRecord(ts.a, ctx->stream);
	ctx->timed.push_back(ts);
}
void hcmvs_time_end(hcmvs_ctx* ctx) { cudaEventRecord(ctx->timed.back().b, ctx->stream); }

static void DrainTimers(hcmvs_ctx* ctx) {
	cudaStreamSynchronize(ctx->stream);
	for (TimedSpan& ts: ctx->timed) {
		float ms = 0.f; cudaEventElapsedTime(&ms, ts.a, ts.b);
		ctx->stageMs[ts.stage] += ms;
		ctx->eventPool.push_back(ts.a); ctx->eventPool.push_back(ts.b);
	}
	ctx->timed.clear();
}
extern "C" int hcmvs_get_timers(hcmvs_ctx* ctx, hcmvs_timers* t) {
	if (!ctx || !t) return HCMVS_ERR_ARG;
	cudaSetDevice(ctx->device);
	DrainTimers(ctx);
	unsigned long long c[8];
	CK(cudaMemcpy(c, ctx->counters_d, sizeof(c), cudaMemcpyDeviceToHost));
	t->ms_score = ctx->stageMs[ST_SCORE]; t->ms_sweeps = ctx->stageMs[ST_SWEEPS]; t->ms_end = ctx->stageMs[ST_END];
	t->ms_prep = ctx->stageMs[ST_PREP]; t->ms_filter = ctx->stageMs[ST_FILTER]; t->ms_fuse = ctx->stageMs[ST_FUSE]; t->ms_exchange = ctx->stageMs[ST_EXCHANGE];
	t->n_hypotheses = c[0]; t->n_view_scores = c[1]; t->n_pixel_iters = c[2]; t->n_smooth_terms = c[3]; t->n_window_walks = c[4];
	t->n_launches = ctx->nLaunches; t->n_fuse_rounds = ctx->fuseRounds;
	t->filter_bytes = ctx->filterBytes; t->fuse_seeds = ctx->fuseSeeds; t->fuse_probes = ctx->fuseProbes;
	return HCMVS_OK;
}
extern "C" int hcmvs_reset_timers(hcmvs_ctx* ctx) {
	if (!ctx) return HCMVS_ERR_ARG;
	cudaSetDevice(ctx->device);
	DrainTimers(ctx);
	for (double& m: ctx->stageMs) m = 0;
	ctx->nLaunches = 0; ctx->filterBytes = 0;
	CK(cudaMemset(ctx->counters_d, 0, 8*sizeof(unsigned long long)));
	return HCMVS_OK;
}

// ------------------------------------------------------------------------------------------------ small f64 maths
// cv::Matx semantics (plain triple loop, left-to-right accumulation), as the reference's Camera / ViewData use.
static void Mul33(const double* A, const double* B, double* C) {
	for (int i=0; i<3; ++i) for (int j=0; j<3; ++j) { double s = A[i*3]*B[j]; s += A[i*3+1]*B[3+j]; s += A[i*3+2]*B[6+j]; C[i*3+j] = s; }
}
static void Mul33Bt(const double* A, const double* B, double* C) {
	for (int i=0; i<3; ++i) for (int j=0; j<3; ++j) { double s = A[i*3]*B[j*3]; s += A[i*3+1]*B[j*3+1]; s += A[i*3+2]*B[j*3+2]; C[i*3+j] = s; }
}
static void Mul3v(const double* A, const double* v, double* o) {
	for (int i=0; i<3; ++i) { double s = A[i*3]*v[0]; s += A[i*3+1]*v[1]; s += A[i*3+2]*v[2]; o[i] = s; }
}
static void Inv33(const double* a, double* b) { // cv::Matx33 inverse: adjugate over determinant
	double d = a[0]*(a[4]*a[8]-a[5]*a[7]) - a[1]*(a[3]*a[8]-a[5]*a[6]) + a[2]*(a[3]*a[7]-a[4]*a[6]);
	d = 1./d;
	b[0] = (a[4]*a[8]-a[5]*a[7])*d; b[1] = (a[2]*a[7]-a[1]*a[8])*d; b[2] = (a[1]*a[5]-a[2]*a[4])*d;
	b[3] = (a[5]*a[6]-a[3]*a[8])*d; b[4] = (a[0]*a[8]-a[2]*a[6])*d; b[5] = (a[2]*a[3]-a[0]*a[5])*d;
	b[6] = (a[3]*a[7]-a[4]*a[6])*d; b[7] = (a[1]*a[6]-a[0]*a[7])*d; b[8] = (a[0]*a[4]-a[1]*a[3])*d;
}
static void ComposeP(const double* K, const double* R, const double* C, double* P) { // Camera.cpp:174-181
	double M[9]; Mul33(K, R, M);
	const double nC[3] = {-C[0], -C[1], -C[2]}; double t[3]; Mul3v(M, nC, t);
	for (int i=0; i<3; ++i) { P[i*4] = M[i*3]; P[i*4+1] = M[i*3+1]; P[i*4+2] = M[i*3+2]; P[i*4+3] = t[i]; }
}

// ------------------------------------------------------------------------------------------------ scene upload
static View* GetView(hcmvs_ctx* ctx, uint32_t view, bool mustExist) {
	if (!ctx) { hcmvs_set_error("null context"); return nullptr; }
	if (view >= ctx->views.size()) {
		if (mustExist) { hcmvs_set_error("view %u not set", view); return nullptr; }
		if (view > (1u<<20)) { hcmvs_set_error("view id %u too large", view); return nullptr; }
		ctx->views.resize(view+1);
	}
	View* v = &ctx->views[view];
	if (mustExist && !v->set) { hcmvs_set_error("view %u not set", view); return nullptr; }
	return v;
}

static int CreateGrayTexture(int W, int H, cudaArray_t& arr, cudaTextureObject_t& tex) {
	cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
	CK(cudaMallocArray(&arr, &desc, W, H, cudaArrayTextureGather));
	cudaResourceDesc rd; std::memset(&rd, 0, sizeof(rd)); rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
	cudaTextureDesc td; std::memset(&td, 0, sizeof(td));
	td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint;
	td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
	CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
	return HCMVS_OK;
}

extern "C" int hcmvs_set_view(hcmvs_ctx* ctx, uint32_t view, int W, int H, const double K[9], const double R[9], const double C[3],
	const float* gray, const uint8_t* bgr)
{
	if (!ctx || !K || !R || !C || !gray) { hcmvs_set_error("null argument"); return HCMVS_ERR_ARG; }
	if (W < 2*HCMVS_HW+2 || H < 2*HCMVS_HW+2 || W > 65535 || H > 65535) { hcmvs_set_error("image size %dx%d unsupported", W, H); return HCMVS_ERR_ARG; }
	if (K[1] != 0.0 || K[3] != 0.0 || K[6] != 0.0 || K[7] != 0.0) { hcmvs_set_error("K must be upper triangular with zero skew"); return HCMVS_ERR_UNSUPPORTED; }
	cudaSetDevice(ctx->device);
	View* v = GetView(ctx, view, false); if (!v) return HCMVS_ERR_ARG;
	const bool reuse = v->set && v->w == W && v->h == H && (v->bgr_d != nullptr) == (bgr != nullptr);
	if (v->set) {
		// the image may still be read by queued kernels: same-size re-uploads wait for them on the device (copy stream),
		// a change of size frees the buffers and has to wait on the host
		if (reuse) { if (v->lastUse) CK(cudaStreamWaitEvent(ctx->copyStream, v->lastUse, 0)); }
		else { CK(cudaStreamSynchronize(ctx->stream)); CK(cudaStreamSynchronize(ctx->copyStream)); FreeView(*v); }
	}
	v->w = W; v->h = H;
	std::memcpy(v->K, K, 72); std::memcpy(v->R, R, 72); std::memcpy(v->C, C, 24);
	ComposeP(K, R, C, v->P);
	const size_t n = (size_t)W*H;
	if (!reuse) {
		{ int r = CreateGrayTexture(W, H, v->arr, v->tex); if (r) return r; }
		CK(cudaMalloc(&v->img_d, n*4));
		if (bgr) CK(cudaMalloc(&v->bgr_d, n*3));
	} else {
		v->graValid = false; // a new image invalidates the derived gradient map (rebuilt by the next hcmvs_init_depthmap)
	}
	// uploads go on the COPY stream so that they overlap the kernels of other views; consumers wait on v->imgReady.
	// The caller may release its buffers on return: cudaMemcpyAsync returns once a PAGEABLE source has been staged, and for a
	// pinned source (read by the DMA engine itself) the copy stream is synchronised below.
	cudaStream_t cs = ctx->copyStream;
	CK(cudaMemcpy2DToArrayAsync(v->arr, 0, 0, gray, (size_t)W*4, (size_t)W*4, H, cudaMemcpyHostToDevice, cs));
	CK(cudaMemcpyAsync(v->img_d, gray, n*4, cudaMemcpyHostToDevice, cs));
	if (bgr) CK(cudaMemcpyAsync(v->bgr_d, bgr, n*3, cudaMemcpyHostToDevice, cs));
	if (!v->imgReady) CK(cudaEventCreateWithFlags(&v->imgReady, cudaEventDisableTiming));
	CK(cudaEventRecord(v->imgReady, cs));
	cudaPointerAttributes pa; const bool pinned = cudaPointerGetAttributes(&pa, gray) == cudaSuccess && pa.type == cudaMemoryTypeHost;
	cudaGetLastError();
	if (pinned) CK(cudaStreamSynchronize(cs)); // a pinned source is read by the DMA engine itself: wait before the caller reuses it
	v->set = true;
	return HCMVS_OK;
}

extern "C" int hcmvs_set_view_remote(hcmvs_ctx* ctx, uint32_t view, int W, int H, const double K[9], const double R[9], const double C[3], int has_bgr) {
	if (!ctx || !K || !R || !C) { hcmvs_set_error("null argument"); return HCMVS_ERR_ARG; }
	if (W < 2*HCMVS_HW+2 || H < 2*HCMVS_HW+2 || W > 65535 || H > 65535) { hcmvs_set_error("image size %dx%d unsupported", W, H); return HCMVS_ERR_ARG; }
	if (K[1] != 0.0 || K[3] != 0.0 || K[6] != 0.0 || K[7] != 0.0) { hcmvs_set_error("K must be upper triangular with zero skew"); return HCMVS_ERR_UNSUPPORTED; }
	cudaSetDevice(ctx->device);
	View* v = GetView(ctx, view, false); if (!v) return HCMVS_ERR_ARG;
	const bool reuse = v->set && v->w == W && v->h == H && (v->bgr_d != nullptr) == (has_bgr != 0);
	if (v->set && !reuse) { CK(cudaStreamSynchronize(ctx->stream)); CK(cudaStreamSynchronize(ctx->copyStream)); FreeView(*v); }
	v->w = W; v->h = H;
	std::memcpy(v->K, K, 72); std::memcpy(v->R, R, 72); std::memcpy(v->C, C, 24);
	ComposeP(K, R, C, v->P);
	const size_t n = (size_t)W*H;
	if (!reuse) {
		{ int r = CreateGrayTexture(W, H, v->arr, v->tex); if (r) return r; }
		CK(cudaMalloc(&v->img_d, n*4));
		if (has_bgr) CK(cudaMalloc(&v->bgr_d, n*3));
	} else v->graValid = false;
	v->set = true;
	return HCMVS_OK;
}

int hcmvs_mark_image_use(hcmvs_ctx* ctx, View& v) {
	if (!v.lastUse) CK(cudaEventCreateWithFlags(&v.lastUse, cudaEventDisableTiming));
	CK(cudaEventRecord(v.lastUse, ctx->stream));
	return HCMVS_OK;
}
static int MarkUse(hcmvs_ctx* ctx, View* v) { // the reference view and its matching neighbours
	int r = hcmvs_mark_image_use(ctx, *v); if (r) return r;
	for (int i=0; i<v->nMatch; ++i) { r = hcmvs_mark_image_use(ctx, ctx->views[v->nbIds[i]]); if (r) return r; }
	return HCMVS_OK;
}

int hcmvs_wait_image(hcmvs_ctx* ctx, const View& v) {
	if (v.imgReady) CK(cudaStreamWaitEvent(ctx->stream, v.imgReady, 0));
	return HCMVS_OK;
}

extern "C" int hcmvs_set_neighbors(hcmvs_ctx* ctx, uint32_t ref, const uint32_t* ids, const float* scores, int n_match, int n_all) {
	View* v = GetView(ctx, ref, true); if (!v) return HCMVS_ERR_ARG;
	if (!ids || n_all < 0 || n_match < 0 || n_match > n_all || n_match > HCMVS_MAX_MATCH_VIEWS || n_all > HCMVS_MAX_FUSE_VIEWS) {
		hcmvs_set_error("bad neighbour list (n_match %d <= %d, n_all %d <= %d)", n_match, HCMVS_MAX_MATCH_VIEWS, n_all, HCMVS_MAX_FUSE_VIEWS); return HCMVS_ERR_ARG;
	}
	for (int i=0; i<n_all; ++i) if (ids[i] == ref) { hcmvs_set_error("view %u lists itself as neighbour", ref); return HCMVS_ERR_ARG; }
	if (!v->nbImages.empty()) { // a new neighbour list drops the per-pair rescaled images
		cudaSetDevice(ctx->device);
		CK(cudaStreamSynchronize(ctx->stream));
		for (NbImage& o: v->nbImages) FreeNbImage(o);
		v->nbImages.clear();
	}
	v->nbIds.assign(ids, ids+n_all);
	v->nbScores.assign(n_all, 0.f);
	if (scores) v->nbScores.assign(scores, scores+n_all);
	v->nMatch = n_match;
	return HCMVS_OK;
}

extern "C" int hcmvs_set_neighbor_image(hcmvs_ctx* ctx, uint32_t ref, int slot, int W, int H, const double K[9], const float* gray) {
	View* v = GetView(ctx, ref, true); if (!v) return HCMVS_ERR_ARG;
	if (slot < 0 || slot >= v->nMatch) { hcmvs_set_error("view %u has no matching neighbour %d (call hcmvs_set_neighbors first)", ref, slot); return HCMVS_ERR_ARG; }
	cudaSetDevice(ctx->device);
	if (v->nbImages.size() < (size_t)v->nMatch) v->nbImages.resize(v->nMatch);
	NbImage& o = v->nbImages[slot];
	CK(cudaStreamSynchronize(ctx->stream)); // queued kernels may still sample the previous image
	FreeNbImage(o);
	if (!gray) return HCMVS_OK;
	if (!K || W < 4 || H < 4 || W > 65535 || H > 65535) { hcmvs_set_error("bad rescaled image %dx%d", W, H); return HCMVS_ERR_ARG; }
	if (K[1] != 0.0 || K[3] != 0.0 || K[6] != 0.0 || K[7] != 0.0) { hcmvs_set_error("K must be upper triangular with zero skew"); return HCMVS_ERR_UNSUPPORTED; }
	{ int r = CreateGrayTexture(W, H, o.arr, o.tex); if (r) return r; }
	CK(cudaMalloc(&o.img_d, (size_t)W*H*4));
	CK(cudaMemcpy2DToArrayAsync(o.arr, 0, 0, gray, (size_t)W*4, (size_t)W*4, H, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemcpyAsync(o.img_d, gray, (size_t)W*H*4, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	o.w = W; o.h = H; std::memcpy(o.K, K, 72);
	return HCMVS_OK;
}

static int AllocMaps(hcmvs_ctx* ctx, View* v) {
	const size_t n = (size_t)v->w*v->h;
	v->depthValid = false;
	if (!v->dn_d) { CK(cudaMalloc(&v->dn_d, n*sizeof(float4))); CK(cudaMemsetAsync(v->dn_d, 0, n*sizeof(float4), ctx->stream)); }
	if (!v->conf_d) { CK(cudaMalloc(&v->conf_d, n*4)); CK(cudaMemsetAsync(v->conf_d, 0, n*4, ctx->stream)); }
	return HCMVS_OK;
}

// Before a view's maps are overwritten on the copy stream: they may still be in use by queued compute work (re-initialisation of an
// estimated view — every scene after the first when a context is reused). Without further knowledge that is "everything queued so
// far", which makes the upload of view i+1 wait for the estimation of view i and serialises the host with the GPU (measured: one
// view enqueued every 20 ms, 44 ms of a 1.02 s C2 scene). After hcmvs_begin_scene the caller has promised that only a view's own
// init / estimate calls touch its maps until the first multi-view consumer, so the view's own last use is enough.
static int WaitMapsFree(hcmvs_ctx* ctx, View* v) {
	if (!v->dn_d) return HCMVS_OK;
	if (ctx->freshScene) { if (v->lastUse) CK(cudaStreamWaitEvent(ctx->copyStream, v->lastUse, 0)); return HCMVS_OK; }
	cudaEvent_t done; CK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming)); CK(cudaEventRecord(done, ctx->stream)); CK(cudaStreamWaitEvent(ctx->copyStream, done, 0)); CK(cudaEventDestroy(done));
	return HCMVS_OK;
}

extern "C" int hcmvs_begin_scene(hcmvs_ctx* ctx) {
	if (!ctx) { hcmvs_set_error("null argument"); return HCMVS_ERR_ARG; }
	cudaSetDevice(ctx->device);
	CK(cudaStreamSynchronize(ctx->stream)); CK(cudaStreamSynchronize(ctx->copyStream));
	if (ctx->dlStream) CK(cudaStreamSynchronize(ctx->dlStream));
	ctx->freshScene = true;
	return HCMVS_OK;
}

// Upload a view's maps on the COPY stream (H2D + packing kernel), so that it overlaps whatever the compute stream is doing
// for other views; consumers on the compute stream wait on v->ready. Returns once the host buffers may be reused.
static int UploadMaps(hcmvs_ctx* ctx, View* v, const float* depth, const float* normal, const float* conf) {
	const size_t n = (size_t)v->w*v->h;
	v->depthValid = false;
	cudaStream_t cs = ctx->copyStream;
	{ int r = WaitMapsFree(ctx, v); if (r) return r; }
	if (!v->dn_d) CK(cudaMalloc(&v->dn_d, n*sizeof(float4)));
	if (!v->conf_d) CK(cudaMalloc(&v->conf_d, n*4));
	if (ctx->uploadBytes < n*16) {
		CK(cudaStreamSynchronize(cs));
		cudaFree(ctx->upload_d); ctx->upload_d = nullptr; ctx->uploadBytes = 0;
		CK(cudaMalloc(&ctx->upload_d, n*16)); ctx->uploadBytes = n*16;
	}
	float* tmp = (float*)ctx->upload_d;
	CK(cudaMemcpyAsync(tmp, depth, n*4, cudaMemcpyHostToDevice, cs));
	if (normal) CK(cudaMemcpyAsync(tmp+n, normal, n*12, cudaMemcpyHostToDevice, cs));
	CK(hcmvs_launch_pack(tmp, normal ? tmp+n : nullptr, v->dn_d, n, cs)); ++ctx->nLaunches;
	if (conf) CK(cudaMemcpyAsync(v->conf_d, conf, n*4, cudaMemcpyHostToDevice, cs));
	else CK(cudaMemsetAsync(v->conf_d, 0, n*4, cs));
	return HCMVS_OK;
}
static int FinishUpload(hcmvs_ctx* ctx, View* v) {
	if (!v->ready) CK(cudaEventCreateWithFlags(&v->ready, cudaEventDisableTiming));
	CK(cudaEventRecord(v->ready, ctx->copyStream));
	CK(cudaStreamWaitEvent(ctx->stream, v->ready, 0)); // later compute-stream work sees the uploaded maps
	CK(cudaStreamSynchronize(ctx->copyStream));        // host buffers are free again; does not wait for compute
	return HCMVS_OK;
}

extern "C" int hcmvs_init_depthmap(hcmvs_ctx* ctx, uint32_t ref, const float* depth0, const float* normal0, float dMin, float dMax) {
	View* v = GetView(ctx, ref, true); if (!v) return HCMVS_ERR_ARG;
	if (!depth0 || !(dMin > 0.f) || !(dMin < dMax)) { hcmvs_set_error("bad depth range [%g,%g)", dMin, dMax); return HCMVS_ERR_ARG; }
	cudaSetDevice(ctx->device);
	const size_t n = (size_t)v->w*v->h;
	int r = UploadMaps(ctx, v, depth0, normal0, nullptr); if (r) return r;
	v->dMin = dMin; v->dMax = dMax; v->hasMaps = true;
	// InitGraMap, SceneDensify.cpp:815-819 (once per view; also on the copy stream)
	if (!v->gra_d) CK(cudaMalloc(&v->gra_d, n));
	if (!v->graValid) {
		if (v->bgr_d) { CK(hcmvs_launch_gramap(v->bgr_d, v->gra_d, v->w, v->h, ctx->copyStream)); ++ctx->nLaunches; }
		else CK(cudaMemsetAsync(v->gra_d, 0, n, ctx->copyStream));
		v->graValid = true;
	}
	return FinishUpload(ctx, v);
}

extern "C" int hcmvs_init_depthmap_triangles(hcmvs_ctx* ctx, uint32_t ref, const double* vertices, int n_vertices, const uint32_t* tris, int n_tris, float dMin, float dMax) {
	// InitDepthMap -> TriangulatePoints2DepthMap (SceneDensify.cpp:514-525, DepthMap.cpp:1879-1936); the caller applied dMin*0.9 / dMax*1.1
	View* v = GetView(ctx, ref, true); if (!v) return HCMVS_ERR_ARG;
	v->depthValid = false;
	if (!vertices || !tris || n_vertices < 3 || n_tris < 1 || !(dMin > 0.f) || !(dMin < dMax)) { hcmvs_set_error("bad triangulation (%d vertices, %d triangles) or depth range [%g,%g)", n_vertices, n_tris, dMin, dMax); return HCMVS_ERR_ARG; }
	cudaSetDevice(ctx->device);
	const size_t n = (size_t)v->w*v->h;
	// chunk list: the 32x8-pixel tiles of every triangle's bounding box (same fixed-point snapping as the kernel), clipped to the image
	std::vector<int3> chunks;
	for (int t=0; t<n_tris; ++t) {
		long long X[3], Y[3];
		for (int k=0; k<3; ++k) {
			const uint32_t vi = tris[(size_t)t*3+k];
			if (vi >= (uint32_t)n_vertices) { hcmvs_set_error("triangle %d refers to vertex %u of %d", t, vi, n_vertices); return HCMVS_ERR_ARG; }
			const float fx = (float)vertices[(size_t)vi*3], fy = (float)vertices[(size_t)vi*3+1];
			if (!(std::fabs(fx) < 1e7f) || !(std::fabs(fy) < 1e7f)) { hcmvs_set_error("vertex %u is not finite", vi); return HCMVS_ERR_ARG; }
			X[k] = (long long)std::floor(16.f*fx+0.5f); Y[k] = (long long)std::floor(16.f*fy+0.5f);
		}
		const int minx = std::max((int)((std::min(X[0], std::min(X[1], X[2]))+0xF)>>4), 0), maxx = std::min((int)((std::max(X[0], std::max(X[1], X[2]))+0xF)>>4), v->w);
		const int miny = std::max((int)((std::min(Y[0], std::min(Y[1], Y[2]))+0xF)>>4), 0), maxy = std::min((int)((std::max(Y[0], std::max(Y[1], Y[2]))+0xF)>>4), v->h);
		for (int y=miny; y<maxy; y+=8) for (int x=minx; x<maxx; x+=32) chunks.push_back(make_int3(t, x, y));
	}
	cudaStream_t cs = ctx->copyStream;
	{ int r = WaitMapsFree(ctx, v); if (r) return r; }
	if (!v->dn_d) CK(cudaMalloc(&v->dn_d, n*sizeof(float4)));
	if (!v->conf_d) CK(cudaMalloc(&v->conf_d, n*4));
	const size_t bV = ((size_t)n_vertices*24+255)&~(size_t)255, bT = ((size_t)n_tris*12+255)&~(size_t)255, bC = (chunks.size()*sizeof(int3)+255)&~(size_t)255, need = bV+bT+bC+n*4;
	if (ctx->uploadBytes < need) {
		CK(cudaStreamSynchronize(cs));
		cudaFree(ctx->upload_d); ctx->upload_d = nullptr; ctx->uploadBytes = 0;
		CK(cudaMalloc(&ctx->upload_d, need)); ctx->uploadBytes = need;
	}
	char* base = (char*)ctx->upload_d;
	CK(cudaMemcpyAsync(base, vertices, (size_t)n_vertices*24, cudaMemcpyHostToDevice, cs));
	CK(cudaMemcpyAsync(base+bV, tris, (size_t)n_tris*12, cudaMemcpyHostToDevice, cs));
	if (!chunks.empty()) CK(cudaMemcpyAsync(base+bV+bT, chunks.data(), chunks.size()*sizeof(int3), cudaMemcpyHostToDevice, cs));
	// pixels no triangle covers (or whose depth comes out <= 0) are uninitialised memory in the reference when bAddCorners is set
	// (DepthMap.cpp:1895-1898); here they are 0 = "unknown", which PASS A replaces by a random hypothesis
	CK(cudaMemsetAsync(v->dn_d, 0, n*sizeof(float4), cs));
	CK(cudaMemsetAsync(v->conf_d, 0, n*4, cs));
	CK(cudaMemsetAsync(base+bV+bT+bC, 0xFF, n*4, cs));
	CK(hcmvs_launch_raster_triangles((const double*)base, (const uint32_t*)(base+bV), (const int3*)(base+bV+bT), (int)chunks.size(), v->K, v->dn_d, (int*)(base+bV+bT+bC), v->w, v->h, cs)); ctx->nLaunches += 2;
	v->dMin = dMin; v->dMax = dMax; v->hasMaps = true;
	if (!v->gra_d) CK(cudaMalloc(&v->gra_d, n));
	if (!v->graValid) {
		if (v->bgr_d) { CK(hcmvs_launch_gramap(v->bgr_d, v->gra_d, v->w, v->h, cs)); ++ctx->nLaunches; }
		else CK(cudaMemsetAsync(v->gra_d, 0, n, cs));
		v->graValid = true;
	}
	return FinishUpload(ctx, v);
}

extern "C" int hcmvs_set_depthmap(hcmvs_ctx* ctx, uint32_t view, const float* depth, const float* normal, const float* conf, float dMin, float dMax) {
	View* v = GetView(ctx, view, true); if (!v) return HCMVS_ERR_ARG;
	if (!depth) { hcmvs_set_error("null depth"); return HCMVS_ERR_ARG; }
	cudaSetDevice(ctx->device);
	int r = UploadMaps(ctx, v, depth, normal, conf); if (r) return r;
	v->dMin = dMin; v->dMax = dMax; v->hasMaps = true;
	return FinishUpload(ctx, v);
}

extern "C" int hcmvs_alloc_depthmap(hcmvs_ctx* ctx, uint32_t view) {
	View* v = GetView(ctx, view, true); if (!v) return HCMVS_ERR_ARG;
	cudaSetDevice(ctx->device);
	int r = AllocMaps(ctx, v); if (r) return r;
	v->hasMaps = true;
	return HCMVS_OK;
}

extern "C" int hcmvs_get_depthmap(hcmvs_ctx* ctx, uint32_t view, float* depth, float* normal, float* conf, float* dMin, float* dMax) {
	View* v = GetView(ctx, view, true); if (!v) return HCMVS_ERR_ARG;
	if (!v->hasMaps) { hcmvs_set_error("view %u has no depth map", view); return HCMVS_ERR_STATE; }
	cudaSetDevice(ctx->device);
	const size_t n = (size_t)v->w*v->h;
	float* tmp; int r = hcmvs_scratch(ctx, n*16, (void**)&tmp); if (r) return r;
	if (depth || normal) {
		CK(hcmvs_launch_unpack(v->dn_d, tmp, tmp+n, n, ctx->stream)); ++ctx->nLaunches;
		if (depth) CK(cudaMemcpyAsync(depth, tmp, n*4, cudaMemcpyDeviceToHost, ctx->stream));
		if (normal) CK(cudaMemcpyAsync(normal, tmp+n, n*12, cudaMemcpyDeviceToHost, ctx->stream));
	}
	if (conf) CK(cudaMemcpyAsync(conf, v->conf_d, n*4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	if (dMin) *dMin = v->dMin;
	if (dMax) *dMax = v->dMax;
	return HCMVS_OK;
}

extern "C" int hcmvs_download_depthmap_begin(hcmvs_ctx* ctx, uint32_t view, int slot) {
	View* v = GetView(ctx, view, true); if (!v) return HCMVS_ERR_ARG;
	if (slot < 0 || slot >= HCMVS_DOWNLOAD_SLOTS) { hcmvs_set_error("download slot %d out of range", slot); return HCMVS_ERR_ARG; }
	if (!v->hasMaps) { hcmvs_set_error("view %u has no depth map", view); return HCMVS_ERR_STATE; }
	cudaSetDevice(ctx->device);
	DownloadSlot& d = ctx->dl[slot];
	const size_t n = (size_t)v->w*v->h;
	if (d.pending) CK(cudaEventSynchronize(d.landed)); // the caller re-used a slot it never waited for
	if (!ctx->dlStream) CK(cudaStreamCreateWithFlags(&ctx->dlStream, cudaStreamNonBlocking));
	if (!d.unpacked) { CK(cudaEventCreateWithFlags(&d.unpacked, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&d.landed, cudaEventDisableTiming)); }
	if (d.hostBytes < n*20) { if (d.host) cudaFreeHost(d.host); d.host = nullptr; d.hostBytes = 0; CK(cudaHostAlloc(&d.host, n*20, cudaHostAllocDefault)); d.hostBytes = n*20; }
	if (d.devBytes < n*20) { cudaFree(d.dev); d.dev = nullptr; d.devBytes = 0; CK(cudaMalloc(&d.dev, n*20)); d.devBytes = n*20; }
	// unpack + a private copy of the confidences on the COMPUTE stream (device to device, a few microseconds): ordered after the view's
	// estimation and before anything queued later can touch the maps; the PCIe transfer itself rides the download stream
	CK(hcmvs_launch_unpack(v->dn_d, d.dev, d.dev+n, n, ctx->stream)); ++ctx->nLaunches;
	CK(cudaMemcpyAsync(d.dev+n*4, v->conf_d, n*4, cudaMemcpyDeviceToDevice, ctx->stream));
	CK(cudaEventRecord(d.unpacked, ctx->stream));
	{ int r = hcmvs_mark_image_use(ctx, *v); if (r) return r; } // a later re-initialisation of this view waits for the unpack (hcmvs_begin_scene)
	CK(cudaStreamWaitEvent(ctx->dlStream, d.unpacked, 0));
	CK(cudaMemcpyAsync(d.host, d.dev, n*20, cudaMemcpyDeviceToHost, ctx->dlStream));
	CK(cudaEventRecord(d.landed, ctx->dlStream));
	d.n = n; d.dMin = v->dMin; d.dMax = v->dMax; d.pending = true;
	return HCMVS_OK;
}

extern "C" int hcmvs_download_depthmap_wait(hcmvs_ctx* ctx, int slot, const float** depth, const float** normal, const float** conf, float* dMin, float* dMax) {
	if (!ctx || slot < 0 || slot >= HCMVS_DOWNLOAD_SLOTS) { hcmvs_set_error("download slot %d out of range", slot); return HCMVS_ERR_ARG; }
	DownloadSlot& d = ctx->dl[slot];
	if (!d.pending) { hcmvs_set_error("download slot %d holds nothing", slot); return HCMVS_ERR_STATE; }
	// THREADING CONTRACT: this is the one entry point that may run on another thread (a .dmap writer) while the owning thread keeps
	// driving the context. It therefore touches nothing but its own slot and that slot's event — no timers, no scratch, no stream.
	cudaSetDevice(ctx->device);
	CK(cudaEventSynchronize(d.landed));
	const float* h = (const float*)d.host;
	if (depth) *depth = h;
	if (normal) *normal = h+d.n;
	if (conf) *conf = h+d.n*4;
	if (dMin) *dMin = d.dMin;
	if (dMax) *dMax = d.dMax;
	d.pending = false;
	return HCMVS_OK;
}

extern "C" int hcmvs_get_depthmap_device(hcmvs_ctx* ctx, uint32_t view, void** dn_d, void** conf_d, float* dMin, float* dMax) {
	if (ctx) ctx->freshScene = false; // a consumer of other views' maps ends the "own maps only" phase of hcmvs_begin_scene
	View* v = GetView(ctx, view, true); if (!v) return HCMVS_ERR_ARG;
	if (!v->hasMaps) { hcmvs_set_error("view %u has no depth map", view); return HCMVS_ERR_STATE; }
	if (dn_d) { *dn_d = v->dn_d; v->depthValid = false; } // the caller may write through the pointer
	if (conf_d) *conf_d = v->conf_d;
	if (dMin) *dMin = v->dMin;
	if (dMax) *dMax = v->dMax;
	return HCMVS_OK;
}
extern "C" int hcmvs_export_maps_d(hcmvs_ctx* ctx, uint32_t view, void* dn_d, void* conf_d) {
	if (ctx) ctx->freshScene = false; // a consumer of other views' maps ends the "own maps only" phase of hcmvs_begin_scene
	View* v = GetView(ctx, view, true); if (!v) return HCMVS_ERR_ARG;
	if (!v->hasMaps) { hcmvs_set_error("view %u has no depth map", view); return HCMVS_ERR_STATE; }
	cudaSetDevice(ctx->device);
	const size_t n = (size_t)v->w*v->h;
	if (dn_d) CK(cudaMemcpyAsync(dn_d, v->dn_d, n*sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
	if (conf_d) CK(cudaMemcpyAsync(conf_d, v->conf_d, n*4, cudaMemcpyDeviceToDevice, ctx->stream));
	return HCMVS_OK;
}
extern "C" int hcmvs_import_maps_d(hcmvs_ctx* ctx, uint32_t view, const void* dn_d, const void* conf_d, float dMin, float dMax) {
	if (ctx) ctx->freshScene = false; // a consumer of other views' maps ends the "own maps only" phase of hcmvs_begin_scene
	View* v = GetView(ctx, view, true); if (!v) return HCMVS_ERR_ARG;
	cudaSetDevice(ctx->device);
	int r = AllocMaps(ctx, v); if (r) return r;
	const size_t n = (size_t)v->w*v->h;
	if (dn_d) CK(cudaMemcpyAsync(v->dn_d, dn_d, n*sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
	if (conf_d) CK(cudaMemcpyAsync(v->conf_d, conf_d, n*4, cudaMemcpyDeviceToDevice, ctx->stream));
	v->dMin = dMin; v->dMax = dMax; v->hasMaps = true;
	return HCMVS_OK;
}
extern "C" int hcmvs_set_depth_range(hcmvs_ctx* ctx, uint32_t view, float dMin, float dMax) {
	View* v = GetView(ctx, view, true); if (!v) return HCMVS_ERR_ARG;
	v->dMin = dMin; v->dMax = dMax; return HCMVS_OK;
}

extern "C" int hcmvs_set_coarse_estimate(hcmvs_ctx* ctx, uint32_t view, int wc, int hc, const float* depth, const float* normal) {
	if (ctx) ctx->freshScene = false; // a consumer of other views' maps ends the "own maps only" phase of hcmvs_begin_scene
	View* v = GetView(ctx, view, true); if (!v) return HCMVS_ERR_ARG;
	cudaSetDevice(ctx->device);
	if (!depth) { CK(cudaStreamSynchronize(ctx->stream)); cudaFree(v->coarse_d); v->coarse_d = nullptr; return HCMVS_OK; }
	if (!normal || wc < 2 || hc < 2) { hcmvs_set_error("coarse estimate needs depth and normal maps of at least 2x2"); return HCMVS_ERR_ARG; }
	if (wc > v->w || hc > v->h) { hcmvs_set_error("coarse estimate %dx%d is larger than view %u (%dx%d): INTER_AREA decimation is not the hand-off the restore tree does", wc, hc, view, v->w, v->h); return HCMVS_ERR_UNSUPPORTED; }
	if (!v->hasMaps) { hcmvs_set_error("view %u has no depth map / range yet (call hcmvs_init_depthmap first)", view); return HCMVS_ERR_STATE; }
	const size_t nc = (size_t)wc*hc, n = (size_t)v->w*v->h;
	float* tmp; int r = hcmvs_scratch(ctx, nc*16+64, (void**)&tmp); if (r) return r;
	if (!v->coarse_d) CK(cudaMalloc(&v->coarse_d, n*sizeof(float4)));
	float* mm_d = tmp+nc*4;
	float mm[2] = {v->dMin, v->dMax};
	CK(cudaMemcpyAsync(tmp, depth, nc*4, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemcpyAsync(tmp+nc, normal, nc*12, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemcpyAsync(mm_d, mm, 8, cudaMemcpyHostToDevice, ctx->stream));
	CK(hcmvs_launch_resize_area_up(tmp, tmp+nc, wc, hc, v->coarse_d, v->w, v->h, ctx->stream)); ++ctx->nLaunches;
	// widen the depth range with the resized depths, restore/.../SceneDensify.cpp:526-532
	CK(hcmvs_launch_minmax_w(v->coarse_d, n, mm_d, ctx->stream)); ++ctx->nLaunches;
	CK(cudaMemcpyAsync(mm, mm_d, 8, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	v->dMin = mm[0]; v->dMax = mm[1];
	return HCMVS_OK;
}

extern "C" int hcmvs_get_coarse_estimate(hcmvs_ctx* ctx, uint32_t view, float* depth, float* normal) {
	View* v = GetView(ctx, view, true); if (!v) return HCMVS_ERR_ARG;
	if (!v->coarse_d) { hcmvs_set_error("view %u has no coarse estimate", view); return HCMVS_ERR_STATE; }
	cudaSetDevice(ctx->device);
	const size_t n = (size_t)v->w*v->h;
	float* tmp; int r = hcmvs_scratch(ctx, n*16, (void**)&tmp); if (r) return r;
	CK(hcmvs_launch_unpack(v->coarse_d, tmp, tmp+n, n, ctx->stream)); ++ctx->nLaunches;
	if (depth) CK(cudaMemcpyAsync(depth, tmp, n*4, cudaMemcpyDeviceToHost, ctx->stream));
	if (normal) CK(cudaMemcpyAsync(normal, tmp+n, n*12, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return HCMVS_OK;
}

extern "C" int hcmvs_snapshot_maps(hcmvs_ctx* ctx) {
	if (ctx) ctx->freshScene = false; // a consumer of other views' maps ends the "own maps only" phase of hcmvs_begin_scene
	if (!ctx) { hcmvs_set_error("null context"); return HCMVS_ERR_ARG; }
	cudaSetDevice(ctx->device);
	for (View& v: ctx->views) {
		if (!v.set || !v.hasMaps || !v.dn_d) continue;
		const size_t n = (size_t)v.w*v.h;
		if (!v.dnPrev_d) CK(cudaMalloc(&v.dnPrev_d, n*sizeof(float4)));
		if (!v.confPrev_d) CK(cudaMalloc(&v.confPrev_d, n*4));
		CK(cudaMemcpyAsync(v.dnPrev_d, v.dn_d, n*sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
		CK(cudaMemcpyAsync(v.confPrev_d, v.conf_d, n*4, cudaMemcpyDeviceToDevice, ctx->stream));
		v.hasPrev = true;
	}
	return HCMVS_OK;
}

extern "C" int hcmvs_restore_snapshot(hcmvs_ctx* ctx) {
	if (ctx) ctx->freshScene = false; // a consumer of other views' maps ends the "own maps only" phase of hcmvs_begin_scene
	if (!ctx) { hcmvs_set_error("null context"); return HCMVS_ERR_ARG; }
	cudaSetDevice(ctx->device);
	for (View& v: ctx->views) {
		if (!v.set || !v.hasPrev || !v.dn_d) continue;
		const size_t n = (size_t)v.w*v.h;
		CK(cudaMemcpyAsync(v.dn_d, v.dnPrev_d, n*sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
		CK(cudaMemcpyAsync(v.conf_d, v.confPrev_d, n*4, cudaMemcpyDeviceToDevice, ctx->stream));
		v.depthValid = false; v.hasMaps = true;
	}
	return HCMVS_OK;
}

extern "C" int hcmvs_get_gradient_map(hcmvs_ctx* ctx, uint32_t view, uint8_t* gra) {
	View* v = GetView(ctx, view, true); if (!v) return HCMVS_ERR_ARG;
	if (!gra || !v->gra_d) { hcmvs_set_error("view %u has no gradient map", view); return HCMVS_ERR_STATE; }
	cudaSetDevice(ctx->device);
	CK(cudaMemcpyAsync(gra, v->gra_d, (size_t)v->w*v->h, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return HCMVS_OK;
}

extern "C" int hcmvs_set_prior(hcmvs_ctx* ctx, uint32_t ref, const float* prior) {
	View* v = GetView(ctx, ref, true); if (!v) return HCMVS_ERR_ARG;
	cudaSetDevice(ctx->device);
	const size_t n = (size_t)v->w*v->h;
	if (!prior) { CK(cudaStreamSynchronize(ctx->stream)); cudaFree(v->prior_d); v->prior_d = nullptr; return HCMVS_OK; }
	if (!v->prior_d) CK(cudaMalloc(&v->prior_d, n*4));
	CK(cudaMemcpyAsync(v->prior_d, prior, n*4, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return HCMVS_OK;
}

// ------------------------------------------------------------------------------------------------ estimation
static int BuildRefConst(hcmvs_ctx* ctx, View* v, uint32_t ref, int it_external, uint64_t seed, RefConst& rc) {
	const hcmvs_params& P = ctx->P;
	if (v->nMatch < 1) { hcmvs_set_error("view %u has no matching neighbours (call hcmvs_set_neighbors)", ref); return HCMVS_ERR_STATE; }
	std::memset(&rc, 0, sizeof(rc));
	rc.w = v->w; rc.h = v->h; rc.y0 = 0; rc.y1 = v->h;
	rc.fx = v->K[0]; rc.fy = v->K[4]; rc.cx = v->K[2]; rc.cy = v->K[5];
	Inv33(v->K, rc.Hr);
	rc.img0 = v->img_d; rc.pitch0 = v->w; rc.gra = v->gra_d; rc.prior = v->prior_d;
	rc.dn = v->dn_d; rc.conf = v->conf_d; v->depthValid = false; // every user of a RefConst may write the maps
	rc.nViews = v->nMatch;
	{ int r = hcmvs_wait_image(ctx, *v); if (r) return r; }
	for (int i=0; i<v->nMatch; ++i) {
		View* nb = GetView(ctx, v->nbIds[i], true);
		if (!nb) { hcmvs_set_error("neighbour view %u of %u not set", v->nbIds[i], ref); return HCMVS_ERR_STATE; }
		{ int r = hcmvs_wait_image(ctx, *nb); if (r) return r; }
		NbViewConst& c = rc.nb[i];
		// DepthEstimator::ViewData, DepthMap.h:430-433: Hl = K1 R1 R0^T, Hm = K1 R1 (C0-C1); a neighbour rescaled for this
		// reference view (ViewData::ScaleImage, SceneDensify.cpp:370-376) brings its own image and intrinsics
		const NbImage* ov = (i < (int)v->nbImages.size() && v->nbImages[i].w) ? &v->nbImages[i] : nullptr;
		double KR[9]; Mul33(ov ? ov->K : nb->K, nb->R, KR);
		Mul33Bt(KR, v->R, c.Hl);
		const double dC[3] = {v->C[0]-nb->C[0], v->C[1]-nb->C[1], v->C[2]-nb->C[2]};
		Mul3v(KR, dC, c.Hm);
		if (ov) { c.tex = ov->tex; c.img = ov->img_d; c.pitch = ov->w; c.w = ov->w; c.h = ov->h; }
		else { c.tex = nb->tex; c.img = nb->img_d; c.pitch = nb->w; c.w = nb->w; c.h = nb->h; }
	}
	// DepthEstimator constants, DepthMap.cpp:413-433
	const float FPI = (float)3.14159265358979323846;
	auto FD2R = [FPI](float d) { return d*(FPI/180.f); };
	rc.dMin = v->dMin; rc.dMax = v->dMax; rc.dMinSqr = std::sqrt(v->dMin); rc.dMaxSqr = std::sqrt(v->dMax);
	rc.keep = P.fNCCThresholdKeep;
	rc.thConfSmall = P.fNCCThresholdKeep*0.2f; rc.thConfBig = P.fNCCThresholdKeep*0.4f;
	rc.thConfRand = P.fNCCThresholdKeep*0.9f; rc.thRobust = P.fNCCThresholdKeep*1.2f;
	rc.smoothBonusDepth = 1.f-P.fRandomSmoothBonus; rc.smoothBonusNormal = (1.f-P.fRandomSmoothBonus)*0.96f;
	rc.smoothSigmaDepth = -1.f/(2.f*(P.fRandomSmoothDepth*P.fRandomSmoothDepth));
	{ const float r = FD2R(P.fRandomSmoothNormal); rc.smoothSigmaNormal = -1.f/(2.f*(r*r)); }
	rc.angle1Range = FD2R(P.fRandomAngle1Range); rc.angle2Range = FD2R(P.fRandomAngle2Range);
	rc.depthRatio = P.fRandomDepthRatio;
	rc.nRandomIters = (int)P.nRandomIters; rc.adapthalfwin = P.adapthalfwin; rc.farReach = P.rb_far_reach; rc.propDirs = P.rb_prop_dirs;
	rc.it_external = it_external; rc.photo2geo = P.photo2geo; rc.propagatehalfwin = P.propagatehalfwin; rc.propagatestep = P.propagatestep;
	rc.photometric_flow = P.photometric_flow; rc.para_prior = P.para_prior; rc.sigmaPrior = P.fsigmaPrior;
	rc.key0 = (uint32_t)seed; rc.key1 = (uint32_t)(seed>>32)^(ref*0x9E3779B9u);
	rc.pass = 0;
	rc.counters = ctx->counters_d;
	return HCMVS_OK;
}

static int RequireMaps(View* v, uint32_t ref) {
	if (!v->hasMaps || !v->dn_d) { hcmvs_set_error("view %u has no depth map (call hcmvs_init_depthmap)", ref); return HCMVS_ERR_STATE; }
	// dMin == 0 is legal: the restore tree widens the range with the coarse map's depths, zeros included (SceneDensify.cpp:526-532)
	if (!(v->dMin >= 0.f && v->dMin < v->dMax)) { hcmvs_set_error("view %u has no valid depth range", ref); return HCMVS_ERR_STATE; }
	return HCMVS_OK;
}

extern "C" int hcmvs_score_depthmap(hcmvs_ctx* ctx, uint32_t ref, int it_external, uint64_t seed) {
	View* v = GetView(ctx, ref, true); if (!v) return HCMVS_ERR_ARG;
	int r = RequireMaps(v, ref); if (r) return r;
	cudaSetDevice(ctx->device);
	RefConst rc; r = BuildRefConst(ctx, v, ref, it_external, seed, rc); if (r) return r;
	hcmvs_time_begin(ctx, ST_SCORE);
	CK(hcmvs_launch_score_init(rc, ctx->P.sampler != 1, ctx->stream)); ++ctx->nLaunches;
	hcmvs_time_end(ctx);
	return MarkUse(ctx, v);
}

extern "C" int hcmvs_end_depthmap(hcmvs_ctx* ctx, uint32_t ref) {
	View* v = GetView(ctx, ref, true); if (!v) return HCMVS_ERR_ARG;
	int r = RequireMaps(v, ref); if (r) return r;
	v->depthValid = false;
	cudaSetDevice(ctx->device);
	hcmvs_time_begin(ctx, ST_END);
	CK(hcmvs_launch_end(v->dn_d, v->conf_d, (size_t)v->w*v->h, ctx->P.fNCCThresholdKeep, ctx->stream)); ++ctx->nLaunches;
	hcmvs_time_end(ctx);
	return hcmvs_mark_image_use(ctx, *v);
}

static int EstimateRows(hcmvs_ctx* ctx, uint32_t ref, int it_external, uint64_t seed, int rowBegin, int rowEnd);
extern "C" int hcmvs_estimate_depthmap(hcmvs_ctx* ctx, uint32_t ref, int it_external, uint64_t seed) { return EstimateRows(ctx, ref, it_external, seed, 0, -1); }
extern "C" int hcmvs_estimate_depthmap_rows(hcmvs_ctx* ctx, uint32_t ref, int it_external, uint64_t seed, int row_begin, int row_end) {
	if (row_begin < 0 || row_end <= row_begin) { hcmvs_set_error("bad row range [%d, %d)", row_begin, row_end); return HCMVS_ERR_ARG; }
	return EstimateRows(ctx, ref, it_external, seed, row_begin, row_end);
}

static int EstimateRows(hcmvs_ctx* ctx, uint32_t ref, int it_external, uint64_t seed, int rowBegin, int rowEnd) {
	View* v = GetView(ctx, ref, true); if (!v) return HCMVS_ERR_ARG;
	int r = RequireMaps(v, ref); if (r) return r;
	if (it_external < 0) { hcmvs_set_error("negative outer iteration"); return HCMVS_ERR_ARG; }
	if (it_external >= 1) {
		// the fork's "+"-shaped candidate set (DepthMap.cpp:1064-1274) must stay on the opposite checkerboard colour
		// (odd offsets 1, 1+step, ...) and fit the 8 candidate slots of the kernel
		const hcmvs_params& Q = ctx->P;
		const int phwMax = std::max(Q.propagatehalfwin, 5);
		if (Q.propagatestep < 1 || Q.propagatehalfwin < 1) { hcmvs_set_error("propagatehalfwin / propagatestep must be >= 1"); return HCMVS_ERR_ARG; }
		if ((Q.propagatestep & 1) && phwMax > 1) { hcmvs_set_error("odd propagatestep puts candidates on the pixel's own checkerboard colour: unsupported"); return HCMVS_ERR_UNSUPPORTED; }
		if (1+2*Q.propagatestep <= phwMax) { hcmvs_set_error("more than 2 candidate rings (propagatehalfwin %d, step %d): unsupported", Q.propagatehalfwin, Q.propagatestep); return HCMVS_ERR_UNSUPPORTED; }
	}
	cudaSetDevice(ctx->device);
	const hcmvs_params& P = ctx->P;
	const size_t n = (size_t)v->w*v->h;
	RefConst rc; r = BuildRefConst(ctx, v, ref, it_external, seed, rc); if (r) return r;
	const bool tex = P.sampler != 1;
	if (rowEnd >= 0) {
		// A band of a row-split view (multi-GPU load balance): a pixel of half-sweep k only depends on pixels at most `reach` rows away
		// in half-sweep k-1 (propagation sources / smoothness neighbours), so computing the band plus a halo of 2*iterations*reach rows
		// reproduces the rows [rowBegin, rowEnd) of the full-image estimate bit for bit (counter RNG per pixel) without any exchange
		// inside the estimation; halo rows hold by-products that the owner of those rows overwrites in the exchange.
		if (rowEnd > v->h) { hcmvs_set_error("row range [%d, %d) outside the %d rows of view %u", rowBegin, rowEnd, v->h, ref); return HCMVS_ERR_ARG; }
		const int reach = it_external == 0 ? std::max(P.rb_far_reach, 1) : std::max(P.propagatehalfwin, 5);
		const int halo = 2*(int)P.nEstimationIters*reach+4;
		rc.y0 = std::max(0, rowBegin-halo); rc.y1 = std::min(v->h, rowEnd+halo);
	}
	rc.coarse = v->coarse_d;
	if (P.viewspread && it_external >= 1) {
		// viewspread (DepthMap.cpp:1504-1608) reads the matching neighbours' maps of the previous outer iteration
		SpreadConst sc; std::memset(&sc, 0, sizeof(sc));
		hcmvs_fill_cam(*v, sc.camRef);
		for (int i=0; i<v->nMatch; ++i) {
			View* nb = GetView(ctx, v->nbIds[i], true);
			if (!nb || !nb->hasPrev) { hcmvs_set_error("viewspread: neighbour view %u of %u has no snapshot of the previous outer iteration (call hcmvs_snapshot_maps)", v->nbIds[i], ref); return HCMVS_ERR_STATE; }
			hcmvs_fill_cam(*nb, sc.nb[i].cam);
			sc.nb[i].dn = nb->dnPrev_d; sc.nb[i].conf = nb->confPrev_d; sc.nb[i].w = nb->w; sc.nb[i].h = nb->h;
		}
		if (!ctx->spread_d) CK(cudaMalloc(&ctx->spread_d, sizeof(SpreadConst)));
		CK(cudaMemcpyAsync(ctx->spread_d, &sc, sizeof(sc), cudaMemcpyHostToDevice, ctx->stream)); // pageable source: staged before the call returns
		rc.viewspread = 1; rc.spread = ctx->spread_d;
	}
	// cv::medianBlur(depthMap, depthMap, 3), SceneDensify.cpp:859
	hcmvs_time_begin(ctx, ST_PREP);
	float4* tmp; r = hcmvs_scratch(ctx, n*sizeof(float4), (void**)&tmp); if (r) return r;
	CK(hcmvs_launch_median3(v->dn_d, tmp, v->w, v->h, ctx->stream)); ++ctx->nLaunches;
	CK(cudaMemcpyAsync(v->dn_d, tmp, n*sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
	hcmvs_time_end(ctx);
	// PASS A, SceneDensify.cpp:915-934
	hcmvs_time_begin(ctx, ST_SCORE);
	CK(hcmvs_launch_score_init(rc, tex, ctx->stream)); ++ctx->nLaunches;
	hcmvs_time_end(ctx);
	// PASS B, SceneDensify.cpp:949-981 — each iteration = red half-sweep + black half-sweep
	hcmvs_time_begin(ctx, ST_SWEEPS);
	for (unsigned iter=0; iter<P.nEstimationIters; ++iter) {
		rc.pass = 1u+iter+(uint32_t)it_external*64u;
		rc.lastPass = it_external == (int)P.nEstimationIters_external-1 && iter == P.nEstimationIters-1;
		for (int colour=0; colour<2; ++colour) { CK(hcmvs_launch_sweep(rc, colour, tex, ctx->stream, P.sampler == 2 && iter >= 1)); ++ctx->nLaunches; }
	}
	hcmvs_time_end(ctx);
	// PASS C, SceneDensify.cpp:1035-1056
	if (it_external == (int)P.nEstimationIters_external-1) {
		hcmvs_time_begin(ctx, ST_END);
		CK(hcmvs_launch_end(v->dn_d, v->conf_d, n, P.fNCCThresholdKeep, ctx->stream)); ++ctx->nLaunches;
		hcmvs_time_end(ctx);
	}
	return MarkUse(ctx, v);
}

extern "C" int hcmvs_score_hypotheses(hcmvs_ctx* ctx, uint32_t ref, const float* depth, const float* normal, int smooth_mode, float* score_out) {
	View* v = GetView(ctx, ref, true); if (!v) return HCMVS_ERR_ARG;
	if (!depth || !normal || !score_out) { hcmvs_set_error("null argument"); return HCMVS_ERR_ARG; }
	cudaSetDevice(ctx->device);
	const size_t n = (size_t)v->w*v->h;
	// hypotheses do not need the view's own maps, only its depth range is irrelevant here
	const bool hadRange = v->dMin > 0.f && v->dMin < v->dMax;
	if (!hadRange) { v->dMin = 1e-6f; v->dMax = 3e38f; }
	RefConst rc; int r = BuildRefConst(ctx, v, ref, 0, 0, rc);
	if (!hadRange) { v->dMin = 0.f; v->dMax = 0.f; }
	if (r) return r;
	if (!v->gra_d || !v->graValid) {
		if (!v->gra_d) CK(cudaMalloc(&v->gra_d, n));
		if (v->bgr_d) { CK(hcmvs_launch_gramap(v->bgr_d, v->gra_d, v->w, v->h, ctx->stream)); ++ctx->nLaunches; }
		else CK(cudaMemsetAsync(v->gra_d, 0, n, ctx->stream));
		v->graValid = true;
		rc.gra = v->gra_d;
	}
	char* buf; r = hcmvs_scratch(ctx, n*(16+16+4), (void**)&buf); if (r) return r;
	float* stage = (float*)buf; float4* hyp = (float4*)(buf+n*16); float* out = (float*)(buf+n*32);
	CK(cudaMemcpyAsync(stage, depth, n*4, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemcpyAsync(stage+n, normal, n*12, cudaMemcpyHostToDevice, ctx->stream));
	CK(hcmvs_launch_pack(stage, stage+n, hyp, n, ctx->stream)); ++ctx->nLaunches;
	rc.counters = nullptr;
	CK(hcmvs_launch_score_hyp(rc, hyp, smooth_mode, out, ctx->P.sampler != 1, ctx->stream)); ++ctx->nLaunches;
	CK(cudaMemcpyAsync(score_out, out, n*4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return MarkUse(ctx, v);
}
