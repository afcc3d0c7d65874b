// Scene::PointCloudFilter (libs/MVS/SceneDensify.cpp:4189-4320) — visibility voting over the fused cloud.
//
// Reference: for every point X and every view v that sees it, a cone from the camera centre through X with half-angle FOV/width
// (one pixel) and height 1.02 x |X - C| is intersected with ALL cloud points (octree); every other point p inside the cone whose
// distance t along the axis is not within 1 % of X's gets  visibility[p] += #views(p)  when it lies behind X and
// visibility[p] -= #views(X)  when it lies in front; points with visibility <= thRemove are deleted. The sums are integers, so the
// order of the votes does not matter: one atomicAdd per vote reproduces the CPU result exactly, provided the f32 classification
// (TConeIntersect::Classify, libs/Common/Ray.inl:985-1002) is restated operation by operation.
//
// Here, per view: all points are binned by the pixel they project to (radix sort by bin), and every (X, v) pair tests the points of
// the bins within a radius that provably contains the cone — the cone test itself is the reference's, the bins only have to be a
// superset. The f32 cone test is coarse (cos^2 of a 1-pixel angle is 1 - 1.2e-7, one f32 ulp below 1), so rounding in t^2 and |D|^2
// moves its edge; the search radius is derived from the angle PLUS that rounding slack. Points that see the view are listed per
// view first (CSR transpose) so that the voting kernel runs with dense warps.
#include "hcmvs_internal.h"
#include <cub/cub.cuh>
#include <vector>
#include <cmath>
#include <cstring>

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { hcmvs_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return HCMVS_ERR_CUDA; } } while (0)

namespace hcmvs {

struct PcfView {
	float ox, oy, oz;          // Cast<float>(camera.C)
	float ax, ay, az;          // camera.Direction() = R row 2 (for the search radius only)
	float cosAngleSq;          // SQUARE(COS(angle)), angle = float(ComputeFOV(0) / width)
	float alphaSlack;          // half-angle of the search cone: the cone's own angle + the f32 rounding slack of the test
	float fmax;                // max(fx, fy): pixels per radian on the optical axis
	float P[12];               // projection (f32) — binning only, never a decision
	int w, h, R, GW, GH;       // image size, grid margin, grid size (w + 2R) x (h + 2R)
};

__device__ __forceinline__ bool pcf_project(const PcfView& V, float x, float y, float z, int& bx, int& by) {
	const float qz = V.P[8]*x+V.P[9]*y+V.P[10]*z+V.P[11];
	if (!(qz > 0.f)) return false;
	const float px = (V.P[0]*x+V.P[1]*y+V.P[2]*z+V.P[3])/qz, py = (V.P[4]*x+V.P[5]*y+V.P[6]*z+V.P[7])/qz;
	if (!(px >= (float)-V.R && py >= (float)-V.R && px < (float)(V.w+V.R) && py < (float)(V.h+V.R))) return false;
	bx = (int)floorf(px)+V.R; by = (int)floorf(py)+V.R;
	return true;
}

__global__ void k_pcf_bin(const PcfView V, const float* __restrict__ pts, uint32_t n, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
	const uint32_t i = blockIdx.x*blockDim.x+threadIdx.x;
	if (i >= n) return;
	int bx, by;
	keys[i] = pcf_project(V, pts[i*3], pts[i*3+1], pts[i*3+2], bx, by) ? (uint32_t)(by*V.GW+bx) : (uint32_t)(V.GW*V.GH); // outside the grid: one bin past the end
	vals[i] = i;
}
// sorted order -> binOffset[k] = number of points with key < k (k = 0 .. nBins), so that any run of consecutive bins is ONE contiguous
// segment of the sorted arrays; and a sorted copy of the positions with #views(p) in .w
__global__ void k_pcf_offsets(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ binOffset, uint32_t nBins) {
	const uint32_t k = blockIdx.x*blockDim.x+threadIdx.x;
	if (k > nBins) return;
	uint32_t lo = 0, hi = n; // lower_bound(keys, k)
	while (lo < hi) { const uint32_t mid = (lo+hi)>>1; if (keys[mid] < k) lo = mid+1; else hi = mid; }
	binOffset[k] = lo;
}
__global__ void k_pcf_sorted_pos(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t n, const float* __restrict__ pts, const uint32_t* __restrict__ offs,
	float4* __restrict__ spos, uint32_t nBins)
{
	const uint32_t j = blockIdx.x*blockDim.x+threadIdx.x;
	if (j >= n || keys[j] >= nBins) return;
	const uint32_t i = vals[j];
	spos[j] = make_float4(pts[i*3], pts[i*3+1], pts[i*3+2], __int_as_float((int)(offs[i+1]-offs[i])));
}

__device__ __forceinline__ float sq3(float a, float b, float c) { return __fadd_rn(__fmul_rn(a, a), __fadd_rn(__fmul_rn(b, b), __fmul_rn(c, c))); } // Eigen fixed-size redux: x0 + (x1 + x2)

// TConeIntersect::Classify == VISIBLE && !IsDepthSimilar -> the vote (+#views(p) behind X, -#views(X) in front), SceneDensify.cpp:4233-4246
__device__ __forceinline__ int pcf_vote_of(const PcfView& V, float px, float py, float pz, int nViewsP, float d0, float d1, float d2, float distance, float maxHeight, int weight) {
	const float E0 = __fsub_rn(px, V.ox), E1 = __fsub_rn(py, V.oy), E2 = __fsub_rn(pz, V.oz);
	const float t = __fadd_rn(__fmul_rn(d0, E0), __fadd_rn(__fmul_rn(d1, E1), __fmul_rn(d2, E2)));
	if (fabsf(t) < 0.0001f) return 0;        // ISZERO -> PLANAR
	if (t < 0.f || t > maxHeight) return 0;  // BACK / FRONT
	const float tSq = __fmul_rn(t, t), dSq = __fmul_rn(V.cosAngleSq, sq3(E0, E1, E2));
	if (!(tSq > dSq)) return 0;              // only VISIBLE votes
	if (__fdiv_rn(fabsf(__fsub_rn(distance, t)), distance) < 0.01f) return 0; // IsDepthSimilar(distance, dist, thSimilar)
	return t > distance ? nViewsP : -weight;
}

// one WARP per point that sees view V (list[] = their indices): the 32 lanes test 32 candidates at a time, read from contiguous
// segments of the sorted positions (one segment per grid row of the search disc)
__global__ void __launch_bounds__(256) k_pcf_vote(const PcfView V, const uint32_t* __restrict__ list, uint32_t nList, uint32_t n, const float* __restrict__ pts, const uint32_t* __restrict__ offs,
	const uint32_t* __restrict__ binOffset, const float4* __restrict__ spos, const uint32_t* __restrict__ vals,
	int* __restrict__ vis, unsigned long long* __restrict__ stats)
{
	const uint32_t l = (blockIdx.x*blockDim.x+threadIdx.x)>>5;
	const int lane = threadIdx.x&31;
	if (l >= nList) return;
	const uint32_t i = list[l];
	const float x = pts[i*3], y = pts[i*3+1], z = pts[i*3+2];
	// Collector::Init, SceneDensify.cpp:4221-4229
	const float D0 = __fsub_rn(x, V.ox), D1 = __fsub_rn(y, V.oy), D2 = __fsub_rn(z, V.oz);
	const float distance = __fsqrt_rn(sq3(D0, D1, D2));
	const float d0 = __fdiv_rn(D0, distance), d1 = __fdiv_rn(D1, distance), d2 = __fdiv_rn(D2, distance);
	const float maxHeight = __fmul_rn(distance, 1.02f); // MaxDepthDifference(distance, thMaxDepth)
	const int weight = (int)(offs[i+1]-offs[i]);
	int cbx, cby;
	// search radius in pixels: directions within alphaSlack of dir land within fmax*alpha/cos^2(theta+alpha) pixels
	const float cosT = d0*V.ax+d1*V.ay+d2*V.az;
	const float c = fmaxf(cosT-V.alphaSlack, 0.05f);
	const float rpx = V.fmax*V.alphaSlack/(c*c)+1.5f;
	const int r = (int)ceilf(rpx);
	unsigned long long tested = 0;
	const bool binned = pcf_project(V, x, y, z, cbx, cby) && cbx >= V.R && cby >= V.R && cbx < V.R+V.w && cby < V.R+V.h && r <= V.R;
	if (!binned) {
		// a point that does not project into the image it is listed in, or an extreme field of view: the bins cannot be trusted to
		// hold the whole cone -> test every point (rare; keeps the result exact)
		if (lane == 0) atomicAdd(&stats[0], 1ull);
		for (uint32_t k=lane; k<n; k+=32) {
			const int vote = pcf_vote_of(V, pts[k*3], pts[k*3+1], pts[k*3+2], (int)(offs[k+1]-offs[k]), d0, d1, d2, distance, maxHeight, weight);
			if (vote) atomicAdd(&vis[k], vote);
		}
		tested = lane == 0 ? n : 0;
	} else {
		const float r2 = rpx*rpx;
		for (int by=max(cby-r, 0); by<=min(cby+r, V.GH-1); ++by) {
			const int ddy = max(abs(by-cby)-1, 0);
			// bins of this row inside the disc: |bx - cbx| - 1 <= sqrt(r2 - ddy^2)
			const float rem = r2-(float)(ddy*ddy);
			if (rem < 0.f) continue;
			const int hw = (int)floorf(sqrtf(rem))+1;
			const int bx0 = max(cbx-hw, 0), bx1 = min(cbx+hw, V.GW-1);
			const uint32_t j0 = binOffset[by*V.GW+bx0], j1 = binOffset[by*V.GW+bx1+1];
			for (uint32_t j=j0+lane; j<j1; j+=32) {
				const float4 p = spos[j];
				const int vote = pcf_vote_of(V, p.x, p.y, p.z, __float_as_int(p.w), d0, d1, d2, distance, maxHeight, weight);
				if (vote) atomicAdd(&vis[vals[j]], vote);
			}
			if (lane == 0) tested += j1-j0;
		}
	}
	if (lane == 0) atomicAdd(&stats[2], tested);
}

// CSR transpose: per view the list of the points that see it
__global__ void k_pcf_count_refs(const uint32_t* __restrict__ views, uint64_t m, uint32_t nViews, uint32_t* __restrict__ counts) {
	extern __shared__ uint32_t sh[];
	for (uint32_t k=threadIdx.x; k<nViews; k+=blockDim.x) sh[k] = 0;
	__syncthreads();
	for (uint64_t j=(uint64_t)blockIdx.x*blockDim.x+threadIdx.x; j<m; j+=(uint64_t)gridDim.x*blockDim.x) { const uint32_t v = views[j]; if (v < nViews) atomicAdd(&sh[v], 1u); }
	__syncthreads();
	for (uint32_t k=threadIdx.x; k<nViews; k+=blockDim.x) if (sh[k]) atomicAdd(&counts[k], sh[k]);
}
__global__ void k_pcf_scatter_refs(const uint32_t* __restrict__ offs, const uint32_t* __restrict__ views, uint32_t n, uint32_t nViews, uint32_t* __restrict__ cursor, uint32_t* __restrict__ lists) {
	const uint32_t i = blockIdx.x*blockDim.x+threadIdx.x;
	if (i >= n) return;
	for (uint32_t j=offs[i]; j<offs[i+1]; ++j) { const uint32_t v = views[j]; if (v < nViews) lists[atomicAdd(&cursor[v], 1u)] = i; }
}

} // namespace hcmvs
using namespace hcmvs;

int hcmvs_pointcloud_filter_device(hcmvs_ctx* ctx, uint32_t n, uint64_t m, const float* pts_d, const uint32_t* offs_d, const uint32_t* views_d, int* vis_d, unsigned long long* statsOut) {
	const uint32_t V = (uint32_t)ctx->views.size();
	cudaStream_t st = ctx->stream;
	// per-view point lists
	uint32_t *counts_d = nullptr, *cursor_d = nullptr, *lists_d = nullptr;
	CK(cudaMalloc(&counts_d, (size_t)V*4)); CK(cudaMalloc(&cursor_d, (size_t)V*4)); CK(cudaMalloc(&lists_d, std::max<uint64_t>(m, 1)*4));
	CK(cudaMemsetAsync(counts_d, 0, (size_t)V*4, st));
	k_pcf_count_refs<<<592, 256, V*4, st>>>(views_d, m, V, counts_d); ++ctx->nLaunches;
	std::vector<uint32_t> counts(V), starts(V+1, 0);
	CK(cudaMemcpyAsync(counts.data(), counts_d, (size_t)V*4, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	for (uint32_t v=0; v<V; ++v) starts[v+1] = starts[v]+counts[v];
	CK(cudaMemcpyAsync(cursor_d, starts.data(), (size_t)V*4, cudaMemcpyHostToDevice, st));
	k_pcf_scatter_refs<<<(n+255)/256, 256, 0, st>>>(offs_d, views_d, n, V, cursor_d, lists_d); ++ctx->nLaunches;
	// per-view scratch
	uint32_t *keys_d = nullptr, *keys2_d = nullptr, *vals_d = nullptr, *vals2_d = nullptr, *binOffset_d = nullptr; float4* spos_d = nullptr;
	unsigned long long* stats_d = nullptr; void* tmp_d = nullptr; size_t tmpBytes = 0, binCap = 0;
	CK(cudaMalloc(&keys_d, (size_t)n*4)); CK(cudaMalloc(&keys2_d, (size_t)n*4)); CK(cudaMalloc(&vals_d, (size_t)n*4)); CK(cudaMalloc(&vals2_d, (size_t)n*4));
	CK(cudaMalloc(&spos_d, (size_t)n*16)); CK(cudaMalloc(&stats_d, 3*8)); CK(cudaMemsetAsync(stats_d, 0, 3*8, st));
	CK(cudaMemsetAsync(vis_d, 0, (size_t)n*4, st));
	cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys_d, keys2_d, vals_d, vals2_d, (int)n, 0, 32, st);
	CK(cudaMalloc(&tmp_d, tmpBytes));
	int rc = HCMVS_OK;
	for (uint32_t v=0; v<V && rc == HCMVS_OK; ++v) {
		const View& vw = ctx->views[v];
		if (!vw.set || !counts[v]) continue;
		PcfView pv; std::memset(&pv, 0, sizeof(pv));
		pv.ox = (float)vw.C[0]; pv.oy = (float)vw.C[1]; pv.oz = (float)vw.C[2];
		pv.ax = (float)vw.R[6]; pv.ay = (float)vw.R[7]; pv.az = (float)vw.R[8];
		// angle = float(image.ComputeFOV(0)/image.width), Image.cpp:215-220; cosAngle rounded from the f64 cosine (oracle q17)
		const float angle = (float)(2.0*std::atan((double)vw.w/(vw.K[0]*2.0))/(double)(unsigned)vw.w);
		const float cosAngle = (float)std::cos((double)angle);
		pv.cosAngleSq = cosAngle*cosAngle;
		// the f32 test t^2 > cosAngleSq*|D|^2 accepts true angles up to asin(sqrt(1 - cosAngleSq + slack)); slack covers the rounding of
		// t (3 ops), t^2, |D|^2 (5 ops), the product and the normalisation of dir: < 2e-6 relative in total
		pv.alphaSlack = (float)std::asin(std::sqrt(std::min(1.0, (1.0-(double)pv.cosAngleSq)+2e-6)));
		pv.fmax = (float)std::max(vw.K[0], vw.K[4]);
		for (int k=0; k<12; ++k) pv.P[k] = (float)vw.P[k];
		pv.w = vw.w; pv.h = vw.h; pv.R = 16; pv.GW = vw.w+2*pv.R; pv.GH = vw.h+2*pv.R;
		const size_t nBins = (size_t)pv.GW*pv.GH;
		if (nBins > binCap) { cudaFree(binOffset_d); CK(cudaMalloc(&binOffset_d, (nBins+1)*4)); binCap = nBins; }
		k_pcf_bin<<<(n+255)/256, 256, 0, st>>>(pv, pts_d, n, keys_d, vals_d);
		int bits = 1; while (((size_t)1<<bits) <= nBins && bits < 32) ++bits; // keys are 0 .. nBins
		if (cub::DeviceRadixSort::SortPairs(tmp_d, tmpBytes, keys_d, keys2_d, vals_d, vals2_d, (int)n, 0, bits, st) != cudaSuccess) { hcmvs_set_error("radix sort failed"); rc = HCMVS_ERR_CUDA; break; }
		k_pcf_offsets<<<(unsigned)((nBins+1+255)/256), 256, 0, st>>>(keys2_d, n, binOffset_d, (uint32_t)nBins);
		k_pcf_sorted_pos<<<(n+255)/256, 256, 0, st>>>(keys2_d, vals2_d, n, pts_d, offs_d, spos_d, (uint32_t)nBins);
		k_pcf_vote<<<(unsigned)(((size_t)counts[v]*32+255)/256), 256, 0, st>>>(pv, lists_d+starts[v], counts[v], n, pts_d, offs_d, binOffset_d, spos_d, vals2_d, vis_d, stats_d);
		ctx->nLaunches += 5;
		if (cudaGetLastError() != cudaSuccess) { hcmvs_set_error("point-cloud filter launch failed"); rc = HCMVS_ERR_CUDA; }
	}
	unsigned long long stats[3] = {0, 0, 0};
	if (rc == HCMVS_OK) { if (cudaMemcpyAsync(stats, stats_d, sizeof(stats), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { hcmvs_set_error("point-cloud filter failed: %s", cudaGetErrorString(cudaGetLastError())); rc = HCMVS_ERR_CUDA; } }
	cudaFree(counts_d); cudaFree(cursor_d); cudaFree(lists_d); cudaFree(keys_d); cudaFree(keys2_d); cudaFree(vals_d); cudaFree(vals2_d); cudaFree(spos_d); cudaFree(stats_d); cudaFree(tmp_d); cudaFree(binOffset_d);
	if (rc != HCMVS_OK) return rc;
	if (statsOut) std::memcpy(statsOut, stats, sizeof(stats));
	return HCMVS_OK;
}

extern "C" int hcmvs_pointcloud_filter(hcmvs_ctx* ctx, uint64_t n_points, const float* points, const uint32_t* view_offsets, const uint32_t* views, int32_t* visibility, uint64_t* stats3) {
	if (!ctx) { hcmvs_set_error("null context"); return HCMVS_ERR_ARG; }
	cudaSetDevice(ctx->device);
	const bool resident = points == nullptr;
	uint64_t n = n_points, m = 0;
	const float* pts_d = nullptr; const uint32_t* offs_d = nullptr; const uint32_t* ids_d = nullptr;
	float* pts_own = nullptr; uint32_t *offs_own = nullptr, *ids_own = nullptr;
	if (resident) {
		void *p = nullptr, *o = nullptr, *w = nullptr;
		if (hcmvs_get_fused_device(ctx, &n, &m, &p, nullptr, nullptr, &o, &w, nullptr) != HCMVS_OK || !n) { hcmvs_set_error("no fused cloud on the device (call hcmvs_fuse_depthmaps) and no points given"); return HCMVS_ERR_STATE; }
		pts_d = (const float*)p; offs_d = (const uint32_t*)o; ids_d = (const uint32_t*)w;
	} else {
		if (!view_offsets || !views) { hcmvs_set_error("null argument"); return HCMVS_ERR_ARG; }
		if (!n) return HCMVS_OK;
		m = view_offsets[n];
		CK(cudaMalloc(&pts_own, n*12)); CK(cudaMalloc(&offs_own, (n+1)*4)); CK(cudaMalloc(&ids_own, std::max<uint64_t>(m, 1)*4));
		CK(cudaMemcpyAsync(pts_own, points, n*12, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(offs_own, view_offsets, (n+1)*4, cudaMemcpyHostToDevice, ctx->stream));
		if (m) CK(cudaMemcpyAsync(ids_own, views, m*4, cudaMemcpyHostToDevice, ctx->stream));
		pts_d = pts_own; offs_d = offs_own; ids_d = ids_own;
	}
	if (n >= (1ull<<31)) { hcmvs_set_error("more than 2^31 points"); return HCMVS_ERR_UNSUPPORTED; }
	int* vis_d = nullptr;
	CK(cudaMalloc(&vis_d, n*4));
	unsigned long long st3[3] = {0, 0, 0};
	int rc = hcmvs_pointcloud_filter_device(ctx, (uint32_t)n, m, pts_d, offs_d, ids_d, vis_d, st3);
	if (rc == HCMVS_OK && visibility) { if (cudaMemcpy(visibility, vis_d, n*4, cudaMemcpyDeviceToHost) != cudaSuccess) { hcmvs_set_error("cannot read the visibility back"); rc = HCMVS_ERR_CUDA; } }
	if (stats3) { stats3[0] = st3[0]; stats3[1] = st3[1]; stats3[2] = st3[2]; }
	cudaFree(vis_d); cudaFree(pts_own); cudaFree(offs_own); cudaFree(ids_own);
	return rc;
}
