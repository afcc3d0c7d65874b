// DepthMapsData::FilterDepthMap (libs/MVS/SceneDensify.cpp:3006-3259) as two kernels.
//
//  k_filter_splat  — stage 1 (:3027-3089): every valid pixel of a neighbour depth map is lifted to world
//    (f64), projected into the reference view and splatted on the 4 surrounding pixels with a nearest-z test.
//    The CPU's raster-order "z <= current wins" is the lexicographic minimum of (z, -sourcePixelIndex), so a
//    single 64-bit atomicMin per target pixel reproduces it bit-exactly, independent of thread order.
//  k_filter_vote   — stage 2 (:3097-3248): per reference pixel, fuse/penalise (bAdjust) or strict agreement.
#include "hcmvs_internal.h"
#include "camera.cuh"
#include <climits>
#include <algorithm>

namespace hcmvs {

struct FilterConst { // passed by value (__grid_constant__): no upload, no host synchronisation per call
	int nViews;                 // N
	int wR, hR;
	CamConst camRef;
	CamConst cam[HCMVS_MAXV];
	const float* depth[HCMVS_MAXV]; // compact depth planes of the neighbours (View::depth_d)
	const float* conf[HCMVS_MAXV];
	int w[HCMVS_MAXV], h[HCMVS_MAXV];
	const float* depthRef; const float* confRef;
	unsigned nMinViews, nMinViewsAdjust;
	float thDepthDiff, thDepthDiffStrict;
	float dMin, dMax;
};

#define FILTER_EMPTY 0xFFFFFFFFFFFFFFFFull
#define FS_SRC 32  // a block splats a 32x32 tile of one neighbour map (4 pixels per thread)
#define FS_T 48    // ... through a 48x48 z-buffer in shared memory when the tile's footprint in the reference view fits

// All neighbour maps in ONE launch (blockIdx.z = neighbour). The 4 splats of a pixel and those of its neighbours land on the same
// few reference pixels, so the nearest-z reduction is done in shared memory first and only the winners of the tile go to the
// global z-buffer: ~1 global 64-bit atomicMin per touched reference pixel instead of 4 per source pixel (the round-1 kernel was
// bound by L2 atomic throughput: 61 M atomics per C2 view). min is associative, so the result is the same key.
__global__ void __launch_bounds__(256) k_filter_splat(const __grid_constant__ FilterConst fc, unsigned long long* __restrict__ projAll, size_t plane) {
	__shared__ unsigned long long tile[FS_T*FS_T];
	__shared__ int sBox[4];
	const int n = blockIdx.z;
	const int w = fc.w[n], h = fc.h[n];
	const int x0 = blockIdx.x*FS_SRC, y0 = blockIdx.y*FS_SRC;
	if (x0 >= w || y0 >= h) return; // the grid is sized for the largest neighbour
	unsigned long long* proj = projAll+(size_t)n*plane;
	const int tx = threadIdx.x&31, ty = threadIdx.x>>5;
	if (threadIdx.x == 0) { sBox[0] = INT_MAX; sBox[1] = INT_MAX; sBox[2] = INT_MIN; sBox[3] = INT_MIN; }
	int fx[4], fy[4]; unsigned flags = 0; unsigned long long key[4]; // flags: bit i valid, bit 4+i ceil(u) != floor(u), bit 8+i ceil(v) != floor(v)
	int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
	#pragma unroll
	for (int i=0; i<4; ++i) {
		const int x = x0+tx, y = y0+ty+8*i;
		fx[i] = 0; fy[i] = 0; key[i] = FILTER_EMPTY;
		if (x >= w || y >= h) continue;
		const unsigned src = (unsigned)(y*w+x);
		const float depth = fc.depth[n][src];
		if (depth == 0.f) continue;
		const D3 X = cam_I2W(fc.cam[n], (double)x, (double)y, (double)depth);
		const D3 camX = cam_W2C(fc.camRef, X);
		if (camX.z <= 0.0) continue;
		double u, v; cam_C2I(fc.camRef, camX, u, v);
		const float z = (float)camX.z;
		key[i] = ((unsigned long long)__float_as_uint(z)<<32) | (unsigned long long)(0xFFFFFFFFu-src);
		// FLOOR2INT / CEIL2INT of the f64 projection (Common/Types.h:909-936); far-off projections are clamped to a value outside any image
		const double uc = fmin(fmax(u, -1e6), 1e6), vc = fmin(fmax(v, -1e6), 1e6);
		if (!(u == u) || !(v == v)) { key[i] = FILTER_EMPTY; continue; }
		const int ax = floor2int(uc), bx = ceil2int(uc), ay = floor2int(vc), by = ceil2int(vc);
		fx[i] = ax; fy[i] = ay;
		flags |= (1u<<i) | (bx != ax ? 16u<<i : 0u) | (by != ay ? 256u<<i : 0u);
		mnx = min(mnx, ax); mny = min(mny, ay); mxx = max(mxx, bx); mxy = max(mxy, by);
	}
	mnx = __reduce_min_sync(0xffffffffu, mnx); mny = __reduce_min_sync(0xffffffffu, mny);
	mxx = __reduce_max_sync(0xffffffffu, mxx); mxy = __reduce_max_sync(0xffffffffu, mxy);
	__syncthreads();
	if (tx == 0 && mnx != INT_MAX) { atomicMin(&sBox[0], mnx); atomicMin(&sBox[1], mny); atomicMax(&sBox[2], mxx); atomicMax(&sBox[3], mxy); }
	__syncthreads();
	const int bx0 = sBox[0], by0 = sBox[1];
	if (bx0 == INT_MAX) return; // nothing valid in this tile
	const int bw = sBox[2]-bx0+1, bh = sBox[3]-by0+1;
	const int wR = fc.wR, hR = fc.hR;
	if (bw <= FS_T && bh <= FS_T) {
		for (int i=threadIdx.x; i<bw*bh; i+=256) tile[i] = FILTER_EMPTY;
		__syncthreads();
		#pragma unroll
		for (int i=0; i<4; ++i) {
			if (!(flags & (1u<<i))) continue;
			const int lx = fx[i]-bx0, ly = fy[i]-by0;
			const bool dx = flags & (16u<<i), dy = flags & (256u<<i);
			atomicMin(&tile[ly*bw+lx], key[i]);
			if (dx) atomicMin(&tile[ly*bw+lx+1], key[i]);
			if (dy) atomicMin(&tile[(ly+1)*bw+lx], key[i]);
			if (dx && dy) atomicMin(&tile[(ly+1)*bw+lx+1], key[i]);
		}
		__syncthreads();
		for (int i=threadIdx.x; i<bw*bh; i+=256) {
			const unsigned long long k = tile[i];
			if (k == FILTER_EMPTY) continue;
			const int xr = bx0+i%bw, yr = by0+i/bw;
			if (xr < 0 || yr < 0 || xr >= wR || yr >= hR) continue;
			atomicMin(&proj[(size_t)yr*wR+xr], k);
		}
	} else { // a footprint larger than the shared tile (strong zoom / rotation between the views): straight to the global z-buffer
		#pragma unroll
		for (int i=0; i<4; ++i) {
			if (!(flags & (1u<<i))) continue;
			const bool dx = flags & (16u<<i), dy = flags & (256u<<i);
			for (int a=0; a<2; ++a) for (int b=0; b<2; ++b) {
				if ((a && !dx) || (b && !dy)) continue; // same pixel twice when the projection is integral
				const int xr = fx[i]+a, yr = fy[i]+b;
				if (xr < 0 || yr < 0 || xr >= wR || yr >= hR) continue;
				atomicMin(&proj[(size_t)yr*wR+xr], key[i]);
			}
		}
	}
}

__device__ __forceinline__ float proj_depth(unsigned long long k) { return k == FILTER_EMPTY ? 0.f : __uint_as_float((unsigned)(k>>32)); }
__device__ __forceinline__ unsigned proj_src(unsigned long long k) { return 0xFFFFFFFFu-(unsigned)(k & 0xFFFFFFFFull); }

__global__ void __launch_bounds__(256) k_filter_vote(const __grid_constant__ FilterConst fcv, const unsigned long long* __restrict__ proj, int bAdjust,
	float* __restrict__ newDepth, float* __restrict__ newConf)
{
	const FilterConst* fc = &fcv;
	const float* __restrict__ confRef = fcv.confRef;
	const int wR = fc->wR, hR = fc->hR;
	const int j = blockIdx.x*blockDim.x+threadIdx.x, i = blockIdx.y*blockDim.y+threadIdx.y;
	if (j >= wR || i >= hR) return;
	const size_t o = (size_t)i*wR+j;
	const size_t plane = (size_t)wR*hR;
	const float depth = fcv.depthRef[o];
	float outD = 0.f, outC = 0.f;
	const int N = fc->nViews;
	if (depth != 0.f) {
		if (bAdjust) {
			// SceneDensify.cpp:3097-3170
			float posConf = confRef[o], negConf = 0.f;
			float avgDepth = __fmul_rn(depth, posConf);
			unsigned nPos = 0, nNeg = 0;
			bool discard = false;
			// the winners of all neighbour z-buffers and their confidences first (independent loads: the kernel was a chain of 2 x N
			// dependent DRAM trips per pixel — ncu: 23 % issue-active, long_scoreboard 26), then the reference's loop in registers
			constexpr int VB = 8; // FilterDepthMap is called with at most 8 neighbours (SceneDensify.cpp:4093-4185); more: further batches
			unsigned long long kk[VB]; float cb[VB];
			for (int n=N-1; n>=0; --n) {
				const int jb = (N-1-n)%VB; // position in the batch (j and i are the pixel)
				if (jb == 0) {
					#pragma unroll
					for (int t=0; t<VB; ++t) kk[t] = n-t >= 0 ? proj[(size_t)(n-t)*plane+o] : FILTER_EMPTY;
					#pragma unroll
					for (int t=0; t<VB; ++t) cb[t] = kk[t] != FILTER_EMPTY ? fc->conf[n-t][proj_src(kk[t])] : 0.f;
				}
				unsigned long long k = kk[0]; float cproj = cb[0];
				#pragma unroll
				for (int t=1; t<VB; ++t) if (jb == t) { k = kk[t]; cproj = cb[t]; }
				const float d = proj_depth(k);
				if (d == 0.f) {
					if (nPos+nNeg+(unsigned)n < fc->nMinViews) { discard = true; break; }
					continue;
				}
				if (depth_similar(depth, d, 0.12f)) { // hard-coded in the fork, :3127
					avgDepth = __fadd_rn(avgDepth, __fmul_rn(d, cproj));
					posConf = __fadd_rn(posConf, cproj);
					++nPos;
				} else {
					if (depth > d) negConf = __fadd_rn(negConf, cproj); // occlusion
					else {
						// free-space violation (:3141-3149)
						const D3 X = cam_I2W(fc->camRef, (double)j, (double)i, (double)depth);
						double u, v; cam_C2I(fc->cam[n], cam_W2C(fc->cam[n], X), u, v);
						const int x = round2int(u), y = round2int(v);
						float c = cproj;
						if (x >= 0 && y >= 0 && x < fc->w[n] && y < fc->h[n]) {
							const float cc = fc->conf[n][(size_t)y*fc->w[n]+x];
							if (cc > 0.f) c = cc;
						}
						negConf = __fadd_rn(negConf, c);
					}
					++nNeg;
				}
			}
			if (!discard && nPos >= fc->nMinViewsAdjust && posConf > negConf) {
				avgDepth = __fdiv_rn(avgDepth, posConf);
				if (fc->dMin <= avgDepth && avgDepth < fc->dMax) { outD = avgDepth; outC = __fsub_rn(posConf, negConf); }
			}
		} else {
			// SceneDensify.cpp:3171-3248
			bool keep = true;
			{
				unsigned nGood = 0, nViews = 0;
				for (int n=N-1; n>=0; --n) {
					const float d = proj_depth(proj[(size_t)n*plane+o]);
					if (d > 0.f) { ++nViews; if (depth_similar(depth, d, fc->thDepthDiffStrict)) ++nGood; }
				}
				if (nGood < fc->nMinViews || nGood < nViews*75u/100u) keep = false;
			}
			if (keep) {
				unsigned nGood = 0, nViews = 0;
				const int dx[4] = {-1, 1, 0, 0}, dy[4] = {0, 0, -1, 1};
				#pragma unroll
				for (int q=0; q<4; ++q) {
					const int xx = j+dx[q], yy = i+dy[q];
					const bool in = xx >= 0 && yy >= 0 && xx < wR && yy < hR; // out-of-image reads defined as 0 (oracle q8)
					for (int n=N-1; n>=0; --n) {
						const float d = in ? proj_depth(proj[(size_t)n*plane+(size_t)yy*wR+xx]) : 0.f;
						if (d > 0.f) { ++nViews; if (depth_similar(depth, d, fc->thDepthDiff)) ++nGood; }
					}
				}
				if (nGood < fc->nMinViews*2u || nGood < nViews*65u/100u) keep = false;
			}
			if (keep) { outD = depth; outC = confRef[o]; }
		}
	}
	newDepth[o] = outD; newConf[o] = outC;
}

__global__ void k_commit_filtered(float4* __restrict__ dn, float* __restrict__ conf, const float* __restrict__ fd, const float* __restrict__ fcn, size_t n) {
	const size_t i = (size_t)blockIdx.x*blockDim.x+threadIdx.x;
	if (i >= n) return;
	float4 e = dn[i]; e.w = fd[i]; dn[i] = e; conf[i] = fcn[i];
}

} // namespace hcmvs
using namespace hcmvs;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { hcmvs_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return HCMVS_ERR_CUDA; } } while (0)

void hcmvs_fill_cam(const View& v, CamConst& c) {
	memcpy(c.K, v.K, 72); memcpy(c.R, v.R, 72); memcpy(c.C, v.C, 24); memcpy(c.P, v.P, 96);
}

// the compact depth plane of a view (what the splats and the vote read instead of the 16-byte stride of dn); rebuilt when a writer of dn
// has run since (View::depthValid is cleared by every such entry point)
static int EnsureDepthPlane(hcmvs_ctx* ctx, View& v) {
	const size_t n = (size_t)v.w*v.h;
	if (!v.depth_d) { CK(cudaMalloc(&v.depth_d, n*4)); v.depthValid = false; }
	if (!v.depthValid) {
		if (v.ready) CK(cudaStreamWaitEvent(ctx->stream, v.ready, 0));
		CK(hcmvs_launch_unpack(v.dn_d, v.depth_d, nullptr, n, ctx->stream)); ++ctx->nLaunches;
		v.depthValid = true;
	}
	return HCMVS_OK;
}

extern "C" int hcmvs_filter_depthmap(hcmvs_ctx* ctx, uint32_t ref, const uint32_t* nb_idx, int n, int bAdjust, float* out_depth, float* out_conf) {
	if (ctx) ctx->freshScene = false; // a consumer of other views' maps ends the "own maps only" phase of hcmvs_begin_scene
	if (!ctx || !nb_idx) { hcmvs_set_error("null argument"); return HCMVS_ERR_ARG; }
	if (ref >= ctx->views.size() || !ctx->views[ref].set || !ctx->views[ref].hasMaps) { hcmvs_set_error("view %u has no depth map", ref); return HCMVS_ERR_STATE; }
	View& v = ctx->views[ref];
	const hcmvs_params& P = ctx->P;
	unsigned nCalib = 0; for (const View& q: ctx->views) if (q.set) ++nCalib;
	const unsigned nMinViews = std::min(P.nMinViewsFilter, nCalib-1), nMinViewsAdjust = std::min(P.nMinViewsFilterAdjust, nCalib-1);
	if (n > HCMVS_MAXV || n < (int)nMinViews || n < (int)nMinViewsAdjust) { hcmvs_set_error("depth map %u can not be filtered with %d neighbours", ref, n); return HCMVS_ERR_ARG; }
	cudaSetDevice(ctx->device);
	FilterConst fc; memset(&fc, 0, sizeof(fc));
	fc.nViews = n; fc.wR = v.w; fc.hR = v.h;
	hcmvs_fill_cam(v, fc.camRef);
	{ int r = EnsureDepthPlane(ctx, v); if (r) return r; }
	fc.depthRef = v.depth_d; fc.confRef = v.conf_d;
	int maxW = 0, maxH = 0;
	for (int i=0; i<n; ++i) {
		if (nb_idx[i] >= v.nbIds.size()) { hcmvs_set_error("neighbour index %u out of range", nb_idx[i]); return HCMVS_ERR_ARG; }
		const uint32_t id = v.nbIds[nb_idx[i]];
		if (id >= ctx->views.size() || !ctx->views[id].hasMaps) { hcmvs_set_error("neighbour view %u has no depth map", id); return HCMVS_ERR_STATE; }
		View& q = ctx->views[id];
		hcmvs_fill_cam(q, fc.cam[i]);
		{ int r = EnsureDepthPlane(ctx, q); if (r) return r; }
		fc.depth[i] = q.depth_d; fc.conf[i] = q.conf_d; fc.w[i] = q.w; fc.h[i] = q.h;
		maxW = std::max(maxW, q.w); maxH = std::max(maxH, q.h);
	}
	fc.nMinViews = nMinViews; fc.nMinViewsAdjust = nMinViewsAdjust;
	fc.thDepthDiff = P.fDepthDiffThreshold*1.2f; fc.thDepthDiffStrict = P.fDepthDiffThreshold*0.8f;
	fc.dMin = v.dMin; fc.dMax = v.dMax;
	const size_t plane = (size_t)v.w*v.h;
	const size_t bytesProj = plane*8*(size_t)n;
	char* buf; int r = hcmvs_scratch(ctx, bytesProj+256, (void**)&buf); if (r) return r;
	unsigned long long* proj = (unsigned long long*)buf;
	if (!v.fdepth_d) CK(cudaMalloc(&v.fdepth_d, plane*4));
	if (!v.fconf_d) CK(cudaMalloc(&v.fconf_d, plane*4));
	hcmvs_time_begin(ctx, ST_FILTER);
	CK(cudaMemsetAsync(proj, 0xFF, bytesProj, ctx->stream));
	if (n > 0) {
		dim3 g((maxW+FS_SRC-1)/FS_SRC, (maxH+FS_SRC-1)/FS_SRC, n);
		k_filter_splat<<<g, 256, 0, ctx->stream>>>(fc, proj, plane); ++ctx->nLaunches;
	}
	dim3 b(32, 8), g((v.w+31)/32, (v.h+7)/8);
	k_filter_vote<<<g, b, 0, ctx->stream>>>(fc, proj, bAdjust, v.fdepth_d, v.fconf_d); ++ctx->nLaunches;
	CK(cudaGetLastError());
	hcmvs_time_end(ctx);
	ctx->filterBytes += (uint64_t)(24*n+16)*(uint64_t)v.w*v.h; // n x (8 read + 8 written + 8 read back) + 8 in + 8 out per reference pixel
	v.hasFiltered = true;
	if (out_depth || out_conf) { // only a caller that wants the result on the host waits; the scratch z-buffer is reused in stream order
		if (out_depth) CK(cudaMemcpyAsync(out_depth, v.fdepth_d, plane*4, cudaMemcpyDeviceToHost, ctx->stream));
		if (out_conf) CK(cudaMemcpyAsync(out_conf, v.fconf_d, plane*4, cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
	}
	return HCMVS_OK;
}

extern "C" int hcmvs_commit_filtered(hcmvs_ctx* ctx) {
	if (ctx) ctx->freshScene = false; // a consumer of other views' maps ends the "own maps only" phase of hcmvs_begin_scene
	if (!ctx) return HCMVS_ERR_ARG;
	cudaSetDevice(ctx->device);
	hcmvs_time_begin(ctx, ST_FILTER);
	for (View& v: ctx->views) {
		if (!v.hasFiltered) continue;
		const size_t n = (size_t)v.w*v.h;
		k_commit_filtered<<<(unsigned)((n+255)/256), 256, 0, ctx->stream>>>(v.dn_d, v.conf_d, v.fdepth_d, v.fconf_d, n); ++ctx->nLaunches;
		v.hasFiltered = false; v.depthValid = false;
	}
	CK(cudaGetLastError());
	hcmvs_time_end(ctx);
	return HCMVS_OK;
}
