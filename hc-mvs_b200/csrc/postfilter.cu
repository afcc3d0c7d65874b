// Per-view post-filters between estimation and fusion: the stock speckle filter and the stock small-gap interpolation.
//
// hcmvs_remove_small_segments — DepthMapsData::RemoveSmallSegments as OpenMVS ships it, the body the fork keeps under `#if 0`
// (libs/MVS/SceneDensify.cpp:1956-2042): flood-fill segments of similar depth, segments below nSpeckleSize pixels are zeroed. The CPU
// grows each segment breadth-first from seeds in COLUMN-major order with the ASYMMETRIC test IsDepthSimilar(depth_curr, depth_nb)
// (|d0-d1|/d0 < t divides by the current pixel), so a segment is what is reachable from its seed among the pixels no earlier segment
// took — order dependent. It is reproduced exactly in two steps:
//   1. on the device, connected components over the SYMMETRIC edges only (similar in both directions): union-find with atomicMin, labels
//      = column-major pixel indices, so a component's root is its first pixel in the CPU's seed order. Such a component is atomic for the
//      CPU too: whichever segment reaches one of its pixels first takes all of it.
//   2. the few one-directional edges between different components (a few hundred per 2-Mpx map: the depth difference must fall in the
//      t^2-wide window between the two quotients) go to the host, which replays the CPU's seed order on that small directed graph and
//      sends back which components die.
// hcmvs_gap_interpolation — the small-gap branch of DepthMapsData::GapInterpolation (:2294-2352 rows, :2640-2683 columns), i.e. the stock
// OpenMVS interpolation; the fork's large-gap branches read uninitialised variables and are not built.
#include "hcmvs_internal.h"
#include "camera.cuh"
#include <vector>
#include <algorithm>
#include <unordered_map>
#include <cstring>

namespace hcmvs {

__device__ __forceinline__ unsigned uf_find(unsigned* L, unsigned i) {
	for (;;) {
		const unsigned p = *(volatile unsigned*)(L+i);
		if (p == i) return i;
		const unsigned gp = *(volatile unsigned*)(L+p);
		if (gp != p) L[i] = gp; // path halving (a benign race: parents only ever decrease)
		i = p;
	}
}
// read-only walk to the root: the flatten kernel must not halve paths — a halving store that lands after another thread has already
// written that node's final root would leave a non-root label behind
__device__ __forceinline__ unsigned uf_root(const unsigned* L, unsigned i) {
	for (;;) { const unsigned p = *(const volatile unsigned*)(L+i); if (p == i) return i; i = p; }
}
__device__ __forceinline__ void uf_union(unsigned* L, unsigned a, unsigned b) {
	for (;;) {
		a = uf_find(L, a); b = uf_find(L, b);
		if (a == b) return;
		if (a > b) { const unsigned t = a; a = b; b = t; }
		const unsigned old = atomicMin(L+b, a); // the smaller (earlier in seed order) root wins
		if (old == b) return;
		b = old;
	}
}
__device__ __forceinline__ bool sim(float d0, float d1, float th) { return __fdiv_rn(fabsf(__fsub_rn(d0, d1)), d0) < th; } // IsDepthSimilar, Util.inl:657-669

__global__ void k_ccl_init(unsigned* __restrict__ L, unsigned* __restrict__ size, size_t n) {
	const size_t i = (size_t)blockIdx.x*blockDim.x+threadIdx.x;
	if (i < n) { L[i] = (unsigned)i; size[i] = 0u; }
}
// column-major index of pixel (x, y): the CPU's seed order (for u in width: for v in height)
#define CM(x, y) ((unsigned)(x)*(unsigned)h+(unsigned)(y))
__global__ void k_ccl_union(const float4* __restrict__ dn, int w, int h, float th, unsigned* __restrict__ L) {
	const int x = blockIdx.x*blockDim.x+threadIdx.x, y = blockIdx.y*blockDim.y+threadIdx.y;
	if (x >= w || y >= h) return;
	const float d = dn[(size_t)y*w+x].w;
	if (!(d > 0.f)) return;
	if (x+1 < w) { const float e = dn[(size_t)y*w+x+1].w; if (e > 0.f && sim(d, e, th) && sim(e, d, th)) uf_union(L, CM(x, y), CM(x+1, y)); }
	if (y+1 < h) { const float e = dn[(size_t)(y+1)*w+x].w; if (e > 0.f && sim(d, e, th) && sim(e, d, th)) uf_union(L, CM(x, y), CM(x, y+1)); }
}
__global__ void k_ccl_flatten(const float4* __restrict__ dn, int w, int h, unsigned* __restrict__ L, unsigned* __restrict__ size) {
	const int x = blockIdx.x*blockDim.x+threadIdx.x, y = blockIdx.y*blockDim.y+threadIdx.y;
	if (x >= w || y >= h) return;
	const unsigned r = uf_root(L, CM(x, y));
	L[CM(x, y)] = r;
	if (dn[(size_t)y*w+x].w > 0.f) atomicAdd(size+r, 1u);
}
struct AsymEdge { unsigned from, to, sizeFrom, sizeTo; };
__global__ void k_ccl_asym(const float4* __restrict__ dn, int w, int h, float th, const unsigned* __restrict__ L, const unsigned* __restrict__ size,
	AsymEdge* __restrict__ edges, unsigned* __restrict__ nEdges, unsigned cap)
{
	const int x = blockIdx.x*blockDim.x+threadIdx.x, y = blockIdx.y*blockDim.y+threadIdx.y;
	if (x >= w || y >= h) return;
	const float d = dn[(size_t)y*w+x].w;
	if (!(d > 0.f)) return;
	#pragma unroll
	for (int k=0; k<2; ++k) {
		const int xn = x+(k == 0), yn = y+(k == 1);
		if (xn >= w || yn >= h) continue;
		const float e = dn[(size_t)yn*w+xn].w;
		if (!(e > 0.f)) continue;
		const bool ab = sim(d, e, th), ba = sim(e, d, th);
		if (ab == ba) continue;
		const unsigned ra = L[CM(x, y)], rb = L[CM(xn, yn)];
		if (ra == rb) continue;
		const unsigned slot = atomicAdd(nEdges, 1u);
		if (slot < cap) edges[slot] = ab ? AsymEdge{ra, rb, size[ra], size[rb]} : AsymEdge{rb, ra, size[rb], size[ra]};
	}
}
__global__ void k_ccl_override(const uint2* __restrict__ list, unsigned n, unsigned* __restrict__ size) { // (root, new size) decided by the host
	const unsigned i = blockIdx.x*blockDim.x+threadIdx.x;
	if (i < n) size[list[i].x] = list[i].y;
}
__global__ void k_ccl_apply(float4* __restrict__ dn, float* __restrict__ conf, int w, int h, const unsigned* __restrict__ L, const unsigned* __restrict__ size,
	unsigned speckle, unsigned* __restrict__ nRemoved)
{
	const int x = blockIdx.x*blockDim.x+threadIdx.x, y = blockIdx.y*blockDim.y+threadIdx.y;
	unsigned removed = 0;
	if (x < w && y < h) {
		const size_t o = (size_t)y*w+x;
		if (dn[o].w > 0.f && size[L[CM(x, y)]] < speckle) { dn[o] = make_float4(0.f, 0.f, 0.f, 0.f); conf[o] = 0.f; removed = 1; }
	}
	removed = __reduce_add_sync(0xffffffffu, removed);
	if (((threadIdx.y*blockDim.x+threadIdx.x)&31) == 0 && removed) atomicAdd(nRemoved, removed);
}

// ---- gap interpolation: one thread per pixel that ends a gap (valid, its predecessor along the pass direction is not)
__device__ __forceinline__ void normal2dir(const float* n, float& px, float& py) { px = atan2f(n[1], n[0]); py = acosf(n[2]); }           // Util.inl:613-618
__device__ __forceinline__ void dir2normal_(float px, float py, float* n) { float s, c, sy, cy; sincosf(px, &s, &c); sincosf(py, &sy, &cy); n[0] = c*sy; n[1] = s*sy; n[2] = cy; } // :619-626
template<bool ROWS>
__global__ void k_gap_fill(float* __restrict__ depth, float* __restrict__ normal, float* __restrict__ conf, int w, int h, unsigned gap, float th, unsigned* __restrict__ nFilled) {
	const int x = blockIdx.x*blockDim.x+threadIdx.x, y = blockIdx.y*blockDim.y+threadIdx.y;
	if (x >= w || y >= h) return;
	const int i = ROWS ? x : y;              // position along the pass
	const size_t stride = ROWS ? 1 : (size_t)w;
	const size_t base = ROWS ? (size_t)y*w : (size_t)x;
	const float d1 = depth[base+(size_t)i*stride];
	if (!(d1 > 0.f) || i == 0) return;
	unsigned count = 0;
	while (count <= gap && (int)count < i && !(depth[base+(size_t)(i-1-(int)count)*stride] > 0.f)) ++count;
	if (count == 0 || count > gap || !((unsigned)i > count)) return;
	int k = i-(int)count;
	const size_t first = base+(size_t)(k-1)*stride, last = base+(size_t)i*stride;
	const float d0 = depth[first];
	if (!sim(d0, d1, th)) return;
	const float diff = __fdiv_rn(__fsub_rn(d1, d0), (float)(count+1));
	float d = d0;
	const float c = conf ? fminf(conf[first], conf[last]) : 0.f;
	float p1x = 0.f, p1y = 0.f, dx = 0.f, dy = 0.f;
	if (normal) {
		float p2x, p2y;
		normal2dir(normal+first*3, p1x, p1y); normal2dir(normal+last*3, p2x, p2y);
		dx = __fdiv_rn(__fsub_rn(p2x, p1x), (float)(count+1)); dy = __fdiv_rn(__fsub_rn(p2y, p1y), (float)(count+1));
	}
	do {
		const size_t o = base+(size_t)k*stride;
		d = __fadd_rn(d, diff);
		depth[o] = d;
		if (normal) { p1x = __fadd_rn(p1x, dx); p1y = __fadd_rn(p1y, dy); dir2normal_(p1x, p1y, normal+o*3); }
		if (conf) conf[o] = c;
	} while (++k < i);
	atomicAdd(nFilled, count);
}

} // namespace hcmvs
using namespace hcmvs;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { hcmvs_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return HCMVS_ERR_CUDA; } } while (0)

extern "C" int hcmvs_remove_small_segments(hcmvs_ctx* ctx, uint32_t view, unsigned speckle_size, uint64_t* n_removed) {
	if (ctx) ctx->freshScene = false; // a consumer of other views' maps ends the "own maps only" phase of hcmvs_begin_scene
	if (!ctx || view >= ctx->views.size() || !ctx->views[view].set) { hcmvs_set_error("view %u not set", view); return HCMVS_ERR_ARG; }
	View& v = ctx->views[view];
	if (!v.hasMaps || !v.dn_d) { hcmvs_set_error("view %u has no depth map", view); return HCMVS_ERR_STATE; }
	cudaSetDevice(ctx->device);
	if (v.ready) CK(cudaStreamWaitEvent(ctx->stream, v.ready, 0));
	const int w = v.w, h = v.h;
	const size_t n = (size_t)w*h;
	const float th = ctx->P.fDepthDiffThreshold*0.7f; // SceneDensify.cpp:1957
	const unsigned cap = (unsigned)std::min<size_t>(2*n, (size_t)1<<24);
	auto al = [](size_t b) { return (b+255)&~(size_t)255; };
	const size_t oL = 0, oS = oL+al(n*4), oE = oS+al(n*4), oC = oE+al((size_t)cap*sizeof(AsymEdge)), total = oC+256;
	char* buf; { int r = hcmvs_scratch(ctx, total, (void**)&buf); if (r) return r; }
	unsigned* L = (unsigned*)(buf+oL); unsigned* size = (unsigned*)(buf+oS); AsymEdge* edges = (AsymEdge*)(buf+oE); unsigned* cnt = (unsigned*)(buf+oC); // cnt[0] edges, cnt[1] removed
	v.depthValid = false;
	CK(cudaMemsetAsync(cnt, 0, 8, ctx->stream));
	dim3 b(32, 8), g((w+31)/32, (h+7)/8);
	k_ccl_init<<<(unsigned)((n+255)/256), 256, 0, ctx->stream>>>(L, size, n); ++ctx->nLaunches;
	k_ccl_union<<<g, b, 0, ctx->stream>>>(v.dn_d, w, h, th, L); ++ctx->nLaunches;
	k_ccl_flatten<<<g, b, 0, ctx->stream>>>(v.dn_d, w, h, L, size); ++ctx->nLaunches;
	k_ccl_asym<<<g, b, 0, ctx->stream>>>(v.dn_d, w, h, th, L, size, edges, cnt, cap); ++ctx->nLaunches;
	unsigned nEdges = 0;
	CK(cudaMemcpyAsync(&nEdges, cnt, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	if (nEdges > cap) { hcmvs_set_error("speckle filter: %u one-directional edges exceed the buffer (%u)", nEdges, cap); return HCMVS_ERR_UNSUPPORTED; }
	if (nEdges) {
		// replay the CPU's seed order on the condensed graph: nodes = components touched by a one-directional edge, visited by increasing
		// root (= first pixel in column-major order); a segment = everything reachable that no earlier segment took
		std::vector<AsymEdge> he(nEdges);
		CK(cudaMemcpy(he.data(), edges, (size_t)nEdges*sizeof(AsymEdge), cudaMemcpyDeviceToHost));
		std::unordered_map<unsigned, unsigned> id; std::vector<unsigned> root, sz;
		auto node = [&](unsigned r, unsigned s) { auto it = id.find(r); if (it != id.end()) return it->second; const unsigned k = (unsigned)root.size(); id.emplace(r, k); root.push_back(r); sz.push_back(s); return k; };
		std::vector<std::pair<unsigned, unsigned>> arcs; arcs.reserve(nEdges);
		for (const AsymEdge& e: he) { const unsigned a = node(e.from, e.sizeFrom), c = node(e.to, e.sizeTo); arcs.emplace_back(a, c); }
		std::sort(arcs.begin(), arcs.end()); arcs.erase(std::unique(arcs.begin(), arcs.end()), arcs.end());
		const unsigned N = (unsigned)root.size();
		std::vector<unsigned> start(N+1, 0);
		for (auto& a: arcs) ++start[a.first+1];
		for (unsigned i=0; i<N; ++i) start[i+1] += start[i];
		std::vector<unsigned> order(N); for (unsigned i=0; i<N; ++i) order[i] = i;
		std::sort(order.begin(), order.end(), [&](unsigned a, unsigned c) { return root[a] < root[c]; });
		std::vector<char> done(N, 0); std::vector<uint2> over; over.reserve(N);
		std::vector<unsigned> seg;
		for (unsigned s: order) {
			if (done[s]) continue;
			seg.assign(1, s); done[s] = 1;
			unsigned long long total = 0;
			for (size_t c=0; c<seg.size(); ++c) {
				const unsigned a = seg[c]; total += sz[a];
				for (unsigned k=start[a]; k<start[a+1]; ++k) { const unsigned t = arcs[k].second; if (!done[t]) { done[t] = 1; seg.push_back(t); } }
			}
			const unsigned verdict = total < speckle_size ? 0u : speckle_size; // size the apply kernel will see: dies / survives
			for (unsigned a: seg) over.push_back(make_uint2(root[a], verdict));
		}
		uint2* over_d = (uint2*)edges; // the edge list is no longer needed
		if (over.size()*sizeof(uint2) > (size_t)cap*sizeof(AsymEdge)) { hcmvs_set_error("speckle filter: override list too long"); return HCMVS_ERR_UNSUPPORTED; }
		CK(cudaMemcpyAsync(over_d, over.data(), over.size()*sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
		k_ccl_override<<<(unsigned)((over.size()+255)/256), 256, 0, ctx->stream>>>(over_d, (unsigned)over.size(), size); ++ctx->nLaunches;
		CK(cudaStreamSynchronize(ctx->stream)); // `over` is pageable host memory
	}
	k_ccl_apply<<<g, b, 0, ctx->stream>>>(v.dn_d, v.conf_d, w, h, L, size, speckle_size, cnt+1); ++ctx->nLaunches;
	unsigned removed = 0;
	CK(cudaMemcpyAsync(&removed, cnt+1, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	if (n_removed) *n_removed = removed;
	return HCMVS_OK;
}

extern "C" int hcmvs_gap_interpolation(hcmvs_ctx* ctx, int w, int h, float* depth, float* normal, float* conf, unsigned gap_size, uint64_t* n_filled) {
	if (!ctx || !depth || w < 1 || h < 1) { hcmvs_set_error("null argument"); return HCMVS_ERR_ARG; }
	cudaSetDevice(ctx->device);
	const size_t n = (size_t)w*h;
	const float th = ctx->P.fDepthDiffThreshold*2.5f; // SceneDensify.cpp:2283
	auto al = [](size_t b) { return (b+255)&~(size_t)255; };
	const size_t oD = 0, oN = oD+al(n*4), oC = oN+al(normal ? n*12 : 0), oK = oC+al(conf ? n*4 : 0), total = oK+256;
	char* buf; { int r = hcmvs_scratch(ctx, total, (void**)&buf); if (r) return r; }
	float* d_d = (float*)(buf+oD); float* n_d = normal ? (float*)(buf+oN) : nullptr; float* c_d = conf ? (float*)(buf+oC) : nullptr; unsigned* cnt = (unsigned*)(buf+oK);
	CK(cudaMemcpyAsync(d_d, depth, n*4, cudaMemcpyHostToDevice, ctx->stream));
	if (normal) CK(cudaMemcpyAsync(n_d, normal, n*12, cudaMemcpyHostToDevice, ctx->stream));
	if (conf) CK(cudaMemcpyAsync(c_d, conf, n*4, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemsetAsync(cnt, 0, 4, ctx->stream));
	dim3 b(32, 8), g((w+31)/32, (h+7)/8);
	k_gap_fill<true><<<g, b, 0, ctx->stream>>>(d_d, n_d, c_d, w, h, gap_size, th, cnt); ++ctx->nLaunches;   // 1. row-wise
	k_gap_fill<false><<<g, b, 0, ctx->stream>>>(d_d, n_d, c_d, w, h, gap_size, th, cnt); ++ctx->nLaunches;  // 2. column-wise, on the rows' result
	unsigned filled = 0;
	CK(cudaMemcpyAsync(depth, d_d, n*4, cudaMemcpyDeviceToHost, ctx->stream));
	if (normal) CK(cudaMemcpyAsync(normal, n_d, n*12, cudaMemcpyDeviceToHost, ctx->stream));
	if (conf) CK(cudaMemcpyAsync(conf, c_d, n*4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaMemcpyAsync(&filled, cnt, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	if (n_filled) *n_filled = filled;
	return HCMVS_OK;
}
