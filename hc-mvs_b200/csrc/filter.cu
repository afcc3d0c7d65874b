// DepthMapsData::FilterDepthMap (libs/MVS/SceneDensify.cpp:3006-3259) as two kernels.
//
//  k_filter_splat  — stage 1 (:3027-3089): every valid pixel of a neighbour depth map is lifted to world
//    (f64), projected into the reference view and splatted on the 4 surrounding pixels with a nearest-z test.
//    The CPU's raster-order "z <= current wins" is the lexicographic minimum of (z, -sourcePixelIndex), so a
//    single 64-bit atomicMin per target pixel reproduces it bit-exactly, independent of thread order.
//  k_filter_vote   — stage 2 (:3097-3248): per reference pixel, fuse/penalise (bAdjust) or strict agreement.
#include "hcmvs_internal.h"
#include "camera.cuh"

namespace hcmvs {

struct FilterConst {
	int nViews;                 // N
	int wR, hR;
	CamConst camRef;
	CamConst cam[HCMVS_MAXV];
	const float4* dn[HCMVS_MAXV];
	const float* conf[HCMVS_MAXV];
	int w[HCMVS_MAXV], h[HCMVS_MAXV];
	unsigned nMinViews, nMinViewsAdjust;
	float thDepthDiff, thDepthDiffStrict;
	float dMin, dMax;
};

#define FILTER_EMPTY 0xFFFFFFFFFFFFFFFFull

__global__ void k_filter_splat(const FilterConst* __restrict__ fc, int n, unsigned long long* __restrict__ proj) {
	const int w = fc->w[n], h = fc->h[n];
	const int x = blockIdx.x*blockDim.x+threadIdx.x, y = blockIdx.y*blockDim.y+threadIdx.y;
	if (x >= w || y >= h) return;
	const unsigned src = (unsigned)(y*w+x);
	const float depth = fc->dn[n][src].w;
	if (depth == 0.f) return;
	const D3 X = cam_I2W(fc->cam[n], (double)x, (double)y, (double)depth);
	const D3 camX = cam_W2C(fc->camRef, X);
	if (camX.z <= 0.0) return;
	double u, v; cam_C2I(fc->camRef, camX, u, v);
	const float z = (float)camX.z;
	const unsigned long long key = ((unsigned long long)__float_as_uint(z)<<32) | (unsigned long long)(0xFFFFFFFFu-src);
	const int xs[2] = {floor2int(u), ceil2int(u)}, ys[2] = {floor2int(v), ceil2int(v)};
	#pragma unroll
	for (int a=0; a<2; ++a)
		#pragma unroll
		for (int b=0; b<2; ++b) {
			const int xr = xs[a], yr = ys[b];
			if (xr < 0 || yr < 0 || xr >= fc->wR || yr >= fc->hR) continue;
			if (a == 1 && xs[1] == xs[0]) continue; // same pixel twice when the projection is integral
			if (b == 1 && ys[1] == ys[0]) continue;
			atomicMin(&proj[(size_t)yr*fc->wR+xr], key);
		}
}

__device__ __forceinline__ float proj_depth(unsigned long long k) { return k == FILTER_EMPTY ? 0.f : __uint_as_float((unsigned)(k>>32)); }
__device__ __forceinline__ unsigned proj_src(unsigned long long k) { return 0xFFFFFFFFu-(unsigned)(k & 0xFFFFFFFFull); }

__global__ void k_filter_vote(const FilterConst* __restrict__ fc, const unsigned long long* __restrict__ proj,
	const float4* __restrict__ dnRef, const float* __restrict__ confRef, int bAdjust,
	float* __restrict__ newDepth, float* __restrict__ newConf)
{
	const int wR = fc->wR, hR = fc->hR;
	const int j = blockIdx.x*blockDim.x+threadIdx.x, i = blockIdx.y*blockDim.y+threadIdx.y;
	if (j >= wR || i >= hR) return;
	const size_t o = (size_t)i*wR+j;
	const size_t plane = (size_t)wR*hR;
	const float depth = dnRef[o].w;
	float outD = 0.f, outC = 0.f;
	const int N = fc->nViews;
	if (depth != 0.f) {
		if (bAdjust) {
			// SceneDensify.cpp:3097-3170
			float posConf = confRef[o], negConf = 0.f;
			float avgDepth = __fmul_rn(depth, posConf);
			unsigned nPos = 0, nNeg = 0;
			bool discard = false;
			for (int n=N-1; n>=0; --n) {
				const unsigned long long k = proj[(size_t)n*plane+o];
				const float d = proj_depth(k);
				if (d == 0.f) {
					if (nPos+nNeg+(unsigned)n < fc->nMinViews) { discard = true; break; }
					continue;
				}
				const float cproj = fc->conf[n][proj_src(k)];
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

extern "C" int hcmvs_filter_depthmap(hcmvs_ctx* ctx, uint32_t ref, const uint32_t* nb_idx, int n, int bAdjust, float* out_depth, float* out_conf) {
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
	for (int i=0; i<n; ++i) {
		if (nb_idx[i] >= v.nbIds.size()) { hcmvs_set_error("neighbour index %u out of range", nb_idx[i]); return HCMVS_ERR_ARG; }
		const uint32_t id = v.nbIds[nb_idx[i]];
		if (id >= ctx->views.size() || !ctx->views[id].hasMaps) { hcmvs_set_error("neighbour view %u has no depth map", id); return HCMVS_ERR_STATE; }
		const View& q = ctx->views[id];
		hcmvs_fill_cam(q, fc.cam[i]);
		fc.dn[i] = q.dn_d; fc.conf[i] = q.conf_d; fc.w[i] = q.w; fc.h[i] = q.h;
	}
	fc.nMinViews = nMinViews; fc.nMinViewsAdjust = nMinViewsAdjust;
	fc.thDepthDiff = P.fDepthDiffThreshold*1.2f; fc.thDepthDiffStrict = P.fDepthDiffThreshold*0.8f;
	fc.dMin = v.dMin; fc.dMax = v.dMax;
	const size_t plane = (size_t)v.w*v.h;
	const size_t bytesProj = plane*8*(size_t)n;
	const size_t offConst = (bytesProj+255)&~(size_t)255;
	char* buf; int r = hcmvs_scratch(ctx, offConst+sizeof(FilterConst)+256, (void**)&buf); if (r) return r;
	unsigned long long* proj = (unsigned long long*)buf;
	FilterConst* fc_d = (FilterConst*)(buf+offConst);
	if (!v.fdepth_d) CK(cudaMalloc(&v.fdepth_d, plane*4));
	if (!v.fconf_d) CK(cudaMalloc(&v.fconf_d, plane*4));
	hcmvs_time_begin(ctx, ST_FILTER);
	CK(cudaMemcpyAsync(fc_d, &fc, sizeof(fc), cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemsetAsync(proj, 0xFF, bytesProj, ctx->stream));
	dim3 b(32, 8);
	for (int i=0; i<n; ++i) {
		dim3 g((fc.w[i]+31)/32, (fc.h[i]+7)/8);
		k_filter_splat<<<g, b, 0, ctx->stream>>>(fc_d, i, proj+(size_t)i*plane); ++ctx->nLaunches;
	}
	dim3 g((v.w+31)/32, (v.h+7)/8);
	k_filter_vote<<<g, b, 0, ctx->stream>>>(fc_d, proj, v.dn_d, v.conf_d, bAdjust, v.fdepth_d, v.fconf_d); ++ctx->nLaunches;
	CK(cudaGetLastError());
	hcmvs_time_end(ctx);
	ctx->filterBytes += (uint64_t)(24*n+16)*(uint64_t)v.w*v.h; // n x (8 read + 8 written + 8 read back) + 8 in + 8 out per reference pixel
	v.hasFiltered = true;
	if (out_depth) CK(cudaMemcpyAsync(out_depth, v.fdepth_d, plane*4, cudaMemcpyDeviceToHost, ctx->stream));
	if (out_conf) CK(cudaMemcpyAsync(out_conf, v.fconf_d, plane*4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream)); // FilterConst lives on the host stack / scratch is reused by the next call
	return HCMVS_OK;
}

extern "C" int hcmvs_commit_filtered(hcmvs_ctx* ctx) {
	if (!ctx) return HCMVS_ERR_ARG;
	cudaSetDevice(ctx->device);
	hcmvs_time_begin(ctx, ST_FILTER);
	for (View& v: ctx->views) {
		if (!v.hasFiltered) continue;
		const size_t n = (size_t)v.w*v.h;
		k_commit_filtered<<<(unsigned)((n+255)/256), 256, 0, ctx->stream>>>(v.dn_d, v.conf_d, v.fdepth_d, v.fconf_d, n); ++ctx->nLaunches;
		v.hasFiltered = false;
	}
	CK(cudaGetLastError());
	hcmvs_time_end(ctx);
	return HCMVS_OK;
}
