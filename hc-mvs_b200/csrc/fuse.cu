// DepthMapsData::FuseDepthMaps (libs/MVS/SceneDensify.cpp:3265-3495) on the GPU, with the CPU's exact
// sequential semantics.
//
// The reference fuses greedily: views in connection order, pixels in raster order; a seed pixel probes one
// pixel in each neighbour view, claims the agreeing ones, and — if it survives nMinViewsFuse — zeroes the
// depths it occludes. Later seeds see those claims / zeroed depths, so the result is order dependent.
// That order is kept here by "deterministic reservations": views are still processed one after another, but
// inside a view every undecided seed reserves (atomicMin of its raster index) each neighbour pixel it would
// touch; a seed that holds all its reservations precedes every undecided seed it conflicts with, so it can
// run its CPU action immediately. Seeds with disjoint footprints commit in the same round; the lowest
// undecided seed always commits, so the loop terminates, and the outcome is bit-identical to the raster
// scan. One cooperative launch per view (grid.sync between the reserve and commit phases), then a
// raster-ordered scan/compaction writes the points in the CPU's order.
#include "hcmvs_internal.h"
#include "camera.cuh"
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include <vector>
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <cstdio>

namespace cg = cooperative_groups;

namespace hcmvs {

#define CLAIM_FREE 0xFFFFFFFFu
#define CLAIM_TAKEN 0xFFFFFFFEu

struct FuseView {
	float4* dn; const float* conf; const uint8_t* bgr; uint32_t* claim;
	int w, h; int hasMaps;
	CamConst cam;
};

struct FuseArgs {
	FuseView* views;
	int ref;
	int nNb; int nb[HCMVS_MAX_FUSE_VIEWS];
	unsigned nMinViewsFuse;
	float depthTh, normalError;
	// Seeds (valid, unclaimed pixels of the reference view) are COMPACTED in raster order before anything else: slot s <-> pixel
	// seeds[s]. Only ~1 pixel in 6 of a C2 view is a seed and they are scattered, so per-pixel kernels ran with 5 of 32 lanes active
	// (ncu: thread_inst_executed_per_inst_executed 5.3); every per-seed array below is indexed by slot.
	const uint32_t* seeds; const uint2* nSeedsPtr; // nSeedsPtr->x = number of seeds (device resident: no host round trip)
	uint8_t* state;   // per seed slot: 0 removed, 1 undecided, 2 emitted, 3 emitted but still waiting for contested probes
	uint32_t* mask;   // merged neighbours (bit k = nb[k]) of emitted seeds
	int* counters;    // [0] undecided seeds, [1] rounds, [2] seeds
	int* trace;       // optional: undecided seeds after each round (debug)
	uint32_t* probes; size_t probeStride; // [neighbour][slot] probe cache
};

struct Probe { int q; float z; };

// project the seed's 3-D point into neighbour view B (SceneDensify.cpp:3386-3394)
__device__ __forceinline__ Probe probe_view(const FuseView& B, const float3 point) {
	Probe pr; pr.q = -1; pr.z = 0.f;
	const float3 pt = cam_ProjectP3f(B.cam, point);
	if (pt.z <= 0.f) return pr;
	const int xB = round2int(__fdiv_rn(pt.x, pt.z)), yB = round2int(__fdiv_rn(pt.y, pt.z));
	if (xB < 0 || yB < 0 || xB >= B.w || yB >= B.h) return pr;
	pr.q = yB*B.w+xB; pr.z = pt.z;
	return pr;
}
__device__ __forceinline__ float3 seed_point(const FuseView& R, int x, int y, float depth) {
	const D3 P = cam_I2W(R.cam, (double)x, (double)y, (double)depth);
	return make_float3((float)P.x, (float)P.y, (float)P.z);
}
__device__ __forceinline__ float conf2weight(float conf, float depth) { // Conf2Weight, SceneDensify.cpp:154-156
	return __fdiv_rn(1.f, __fmul_rn(__fmul_rn(fmaxf(__fsub_rn(1.f, conf), 0.03f), depth), depth));
}
__device__ __forceinline__ float dot3f(const float3 a, const float3 b) { return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z)); }

// per-(seed, neighbour) probe, computed once: pixel index in the neighbour view | class << 30
#define PROBE_NONE  0u   // neither similar nor occluding: the seed never touches this pixel (no dependency through it)
#define PROBE_MERGE 1u   // depth and normal agree (SceneDensify.cpp:3400-3423): claimed if the seed survives
#define PROBE_INVAL 2u   // occluded by the seed (:3424-3427): zeroed if the seed survives
#define PROBE_DEAD  0xFFFFFFFFu

// ---- phase 0 (own kernel, full grid): find the seeds (valid depth, not yet claimed, SceneDensify.cpp:3347-3354) and classify
// their probes. The geometry of a probe is static: the seed's 3-D point, the pixel it hits in each neighbour view, and — while
// that pixel is alive (depth != 0, not claimed) — whether it would merge, be invalidated, or be left alone. Only liveness
// changes during the fusion of this view, so the f64 projections are done once and the rounds below are integer work.
// raster-ordered compaction of the seeds: per-chunk counts (k_seed_count) -> exclusive scan (k_fuse_scan) -> scatter (k_seed_scatter)
#ifndef FUSE_CHUNK
#define FUSE_CHUNK 256 // pixels (or seed slots) per block of the count / scatter / emit kernels (256 threads)
#endif
#define FUSE_ITEMS (FUSE_CHUNK/256) // consecutive pixels / slots per thread: keeps the raster order inside a block
__device__ __forceinline__ bool is_seed(const FuseView& R, int p) { return R.dn[p].w != 0.f && R.claim[p] != CLAIM_TAKEN; } // SceneDensify.cpp:3347-3354
__global__ void __launch_bounds__(256) k_seed_count(const FuseArgs a, uint2* __restrict__ blockSums) {
	__shared__ unsigned sP[8];
	const FuseView& R = a.views[a.ref];
	const int nPix = R.w*R.h;
	unsigned n = 0;
	const int base = blockIdx.x*FUSE_CHUNK;
	for (int i=threadIdx.x; i<FUSE_CHUNK; i+=256) { const int p = base+i; if (p < nPix && is_seed(R, p)) ++n; }
	for (int s=16; s>0; s>>=1) n += __shfl_xor_sync(0xffffffffu, n, s);
	if ((threadIdx.x&31) == 0) sP[threadIdx.x>>5] = n;
	__syncthreads();
	if (threadIdx.x == 0) { unsigned t = 0; for (int i=0; i<8; ++i) t += sP[i]; blockSums[blockIdx.x] = make_uint2(t, 0); }
}
__global__ void __launch_bounds__(256) k_seed_scatter(const FuseArgs a, const uint2* __restrict__ blockSums, int nBlocks, uint32_t* __restrict__ seeds) {
	// each thread owns FUSE_ITEMS consecutive pixels so that the slots keep raster order
	__shared__ unsigned sP[256];
	const FuseView& R = a.views[a.ref];
	const int nPix = R.w*R.h;
	const int p0 = blockIdx.x*FUSE_CHUNK+threadIdx.x*FUSE_ITEMS;
	bool f[FUSE_ITEMS]; unsigned n = 0;
	for (int i=0; i<FUSE_ITEMS; ++i) { f[i] = p0+i < nPix && is_seed(R, p0+i); n += f[i]; }
	sP[threadIdx.x] = n;
	__syncthreads();
	for (int off=1; off<256; off<<=1) {
		unsigned t = 0;
		if (threadIdx.x >= off) t = sP[threadIdx.x-off];
		__syncthreads();
		sP[threadIdx.x] += t;
		__syncthreads();
	}
	unsigned slot = blockSums[blockIdx.x].x+(sP[threadIdx.x]-n);
	for (int i=0; i<FUSE_ITEMS; ++i) if (f[i]) seeds[slot++] = (uint32_t)(p0+i);
	if (blockIdx.x == 0 && threadIdx.x == 0) { const int tot = (int)blockSums[nBlocks].x; a.counters[0] = tot; a.counters[2] = tot; }
}

// ---- phase 0 (own kernel, one thread per SEED): classify the probes. The geometry of a probe is static: the seed's 3-D point, the
// pixel it hits in each neighbour view, and — while that pixel is alive (depth != 0, not claimed) — whether it would merge, be
// invalidated, or be left alone. Only liveness changes during the fusion of this view, so the f64 projections are done once and the
// rounds below are integer work.
__global__ void __launch_bounds__(256) k_fuse_probe(const FuseArgs a) {
	const FuseView& R = a.views[a.ref];
	const int s = blockIdx.x*blockDim.x+threadIdx.x;
	if (s >= (int)a.nSeedsPtr->x) return;
	const int p = (int)a.seeds[s];
	const float4 e = R.dn[p];
	a.state[s] = 1;
	const int x = p%R.w, y = p/R.w;
	const float3 point = seed_point(R, x, y, e.w);
	const float3 normal = cam_NormalC2W(R.cam, make_float3(e.x, e.y, e.z));
	// neighbours in chunks of 4: the projections first, then the 4 independent gathers in flight together (the kernel is
	// bound by gather latency, not bandwidth), then the classification
	for (int k0=0; k0<a.nNb; k0+=4) {
		Probe pr[4]; float4 eB[4]; uint32_t cB[4];
		#pragma unroll
		for (int j=0; j<4; ++j) {
			pr[j].q = -1; pr[j].z = 0.f;
			if (k0+j < a.nNb) { const FuseView& B = a.views[a.nb[k0+j]]; if (B.hasMaps) pr[j] = probe_view(B, point); }
		}
		#pragma unroll
		for (int j=0; j<4; ++j) {
			eB[j] = make_float4(0.f, 0.f, 0.f, 0.f); cB[j] = CLAIM_TAKEN;
			if (pr[j].q >= 0) { const FuseView& B = a.views[a.nb[k0+j]]; eB[j] = B.dn[pr[j].q]; cB[j] = B.claim[pr[j].q]; }
		}
		#pragma unroll
		for (int j=0; j<4; ++j) {
			if (k0+j >= a.nNb) break;
			uint32_t code = PROBE_DEAD;
			if (pr[j].q >= 0 && eB[j].w != 0.f && cB[j] != CLAIM_TAKEN) {
				uint32_t cls = PROBE_NONE;
				bool merge = false;
				if (depth_similar(pr[j].z, eB[j].w, a.depthTh)) {
					const float3 normalB = cam_NormalC2W(a.views[a.nb[k0+j]].cam, make_float3(eB[j].x, eB[j].y, eB[j].z));
					merge = dot3f(normal, normalB) > a.normalError;
				}
				if (merge) cls = PROBE_MERGE; else if (pr[j].z < eB[j].w) cls = PROBE_INVAL;
				code = (uint32_t)pr[j].q | (cls<<30);
			}
			a.probes[(size_t)(k0+j)*a.probeStride+s] = code;
		}
	}
}

// ---- the reserve / resolve rounds (cooperative: grid.sync between the phases)
__global__ void __launch_bounds__(256) k_fuse_view(const FuseArgs a) {
	cg::grid_group grid = cg::this_grid();
	const FuseView& R = a.views[a.ref];
	const int nSeeds = (int)a.nSeedsPtr->x;
	const int tid = blockIdx.x*blockDim.x+threadIdx.x, nThreads = gridDim.x*blockDim.x;
	int undecided = *(volatile int*)&a.counters[0];
	int round = 0;
	while (undecided > 0) {
		// ---- phase 1: every unfinished seed reserves each live neighbour pixel it has not dealt with yet
		for (int s=tid; s<nSeeds; s+=nThreads) {
			const uint8_t st = a.state[s];
			if (st != 1 && st != 3) continue;
			const int p = (int)a.seeds[s]; // the raster index orders the reservations
			for (int k0=0; k0<a.nNb; k0+=4) { // 4 independent probe -> pixel gathers in flight
				uint32_t code[4], cl[4]; float dz[4];
				#pragma unroll
				for (int j=0; j<4; ++j) code[j] = k0+j < a.nNb ? a.probes[(size_t)(k0+j)*a.probeStride+s] : PROBE_DEAD;
				#pragma unroll
				for (int j=0; j<4; ++j) {
					dz[j] = 0.f; cl[j] = CLAIM_TAKEN;
					if (code[j] != PROBE_DEAD && (code[j]>>30) != PROBE_NONE) {
						const FuseView& B = a.views[a.nb[k0+j]];
						const uint32_t q = code[j] & 0x3FFFFFFFu;
						dz[j] = B.dn[q].w; cl[j] = B.claim[q];
					}
				}
				#pragma unroll
				for (int j=0; j<4; ++j)
					if (dz[j] != 0.f && cl[j] != CLAIM_TAKEN && cl[j] > (uint32_t)p) atomicMin(&a.views[a.nb[k0+j]].claim[code[j] & 0x3FFFFFFFu], (uint32_t)p);
			}
		}
		grid.sync();
		// ---- phase 2: resolve what the raster order already fixes.
		// A seed holding a reservation precedes every unfinished seed that touches the same pixel, so that pixel is in
		// the state the seed would find it in at its turn of the raster scan. Hence:
		//  * a seed whose own view + already merged + held agreeing pixels reach nMinViewsFuse WILL be emitted (:3429)
		//    whatever its contested probes turn into: it acts on the pixels it holds at once (claim / zero them) and only
		//    keeps waiting for the contested ones — this is what cuts the dependency chains along the rows;
		//  * a seed that cannot reach nMinViewsFuse even if every contested agreeing pixel were still alive at its turn
		//    will NOT be emitted: it releases everything;
		//  * otherwise it waits for the lower seeds it conflicts with (the lowest unfinished seed never waits).
		int nDone = 0;
		for (int s=tid; s<nSeeds; s+=nThreads) {
			const uint8_t st = a.state[s];
			if (st != 1 && st != 3) continue;
			const int p = (int)a.seeds[s];
			uint32_t merged = st == 3 ? a.mask[s] : 0u;
			uint32_t heldMerge = 0, heldInval = 0;
			unsigned nContested = 0, nContestedMerge = 0;
			for (int k0=0; k0<a.nNb; k0+=4) {
				uint32_t code[4], cl[4]; float dz[4];
				#pragma unroll
				for (int j=0; j<4; ++j) code[j] = k0+j < a.nNb ? a.probes[(size_t)(k0+j)*a.probeStride+s] : PROBE_DEAD;
				#pragma unroll
				for (int j=0; j<4; ++j) {
					dz[j] = 0.f; cl[j] = CLAIM_TAKEN;
					if (code[j] != PROBE_DEAD && (code[j]>>30) != PROBE_NONE) {
						const FuseView& B = a.views[a.nb[k0+j]];
						const uint32_t q = code[j] & 0x3FFFFFFFu;
						cl[j] = *(volatile uint32_t*)&B.claim[q]; dz[j] = *(volatile float*)&B.dn[q].w;
					}
				}
				#pragma unroll
				for (int j=0; j<4; ++j) {
					if (code[j] == PROBE_DEAD || (code[j]>>30) == PROBE_NONE) continue;
					const int k = k0+j;
					if (cl[j] == CLAIM_TAKEN || dz[j] == 0.f) { a.probes[(size_t)k*a.probeStride+s] = PROBE_DEAD; continue; }
					if (cl[j] == (uint32_t)p) { if ((code[j]>>30) == PROBE_MERGE) heldMerge |= 1u<<k; else heldInval |= 1u<<k; }
					else { ++nContested; nContestedMerge += (code[j]>>30) == PROBE_MERGE; }
				}
			}
			const unsigned nViews = 1u+__popc(merged)+__popc(heldMerge);
			if (st == 3 || nViews >= a.nMinViewsFuse) {
				for (int k=0; k<a.nNb; ++k) {
					const uint32_t bit = 1u<<k;
					if (!((heldMerge|heldInval) & bit)) continue;
					const FuseView& B = a.views[a.nb[k]];
					const size_t pi = (size_t)k*a.probeStride+s;
					const uint32_t q = a.probes[pi] & 0x3FFFFFFFu;
					if (heldMerge & bit) B.claim[q] = CLAIM_TAKEN; // merged probes keep their pixel index for k_fuse_emit
					else { B.dn[q].w = 0.f; __threadfence(); B.claim[q] = CLAIM_FREE; a.probes[pi] = PROBE_DEAD; } // :3447-3449
				}
				merged |= heldMerge;
				// a merged probe must not be looked at again: park it as PROBE_NONE (pixel index kept)
				for (int k=0; k<a.nNb; ++k) if (heldMerge & (1u<<k)) { const size_t pi = (size_t)k*a.probeStride+s; a.probes[pi] = (a.probes[pi] & 0x3FFFFFFFu) | (PROBE_NONE<<30); }
				a.mask[s] = merged;
				if (nContested == 0) { R.claim[p] = CLAIM_TAKEN; a.state[s] = 2; ++nDone; }
				else a.state[s] = 3;
			} else if (nViews+nContestedMerge < a.nMinViewsFuse) {
				for (int k=0; k<a.nNb; ++k) {
					if (!((heldMerge|heldInval) & (1u<<k))) continue;
					const FuseView& B = a.views[a.nb[k]];
					B.claim[a.probes[(size_t)k*a.probeStride+s] & 0x3FFFFFFFu] = CLAIM_FREE;
				}
				a.state[s] = 0; ++nDone;
			}
		}
		nDone = cg::reduce(cg::tiled_partition<32>(cg::this_thread_block()), nDone, cg::plus<int>());
		if ((threadIdx.x&31) == 0 && nDone) atomicSub(&a.counters[0], nDone);
		grid.sync();
		undecided = *(volatile int*)&a.counters[0];
		if (tid == 0 && a.trace && round < 256) a.trace[round] = undecided;
		++round;
	}
	if (tid == 0) a.counters[1] = round;
}

// ------------------------------------------------------------------ raster-ordered compaction of the emitted seeds (slot order == raster order)
__global__ void __launch_bounds__(256) k_fuse_count(const uint8_t* __restrict__ state, const uint32_t* __restrict__ mask, const uint2* __restrict__ nSeedsPtr, uint2* __restrict__ blockSums) {
	const int nPix = (int)nSeedsPtr->x; // number of seed slots
	__shared__ unsigned sP[8], sV[8];
	unsigned nP = 0, nV = 0;
	const int base = blockIdx.x*FUSE_CHUNK;
	for (int i=threadIdx.x; i<FUSE_CHUNK; i+=256) {
		const int p = base+i;
		if (p < nPix && state[p] == 2) { ++nP; nV += 1+__popc(mask[p]); }
	}
	for (int s=16; s>0; s>>=1) { nP += __shfl_xor_sync(0xffffffffu, nP, s); nV += __shfl_xor_sync(0xffffffffu, nV, s); }
	if ((threadIdx.x&31) == 0) { sP[threadIdx.x>>5] = nP; sV[threadIdx.x>>5] = nV; }
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned a = 0, b = 0;
		for (int i=0; i<8; ++i) { a += sP[i]; b += sV[i]; }
		blockSums[blockIdx.x] = make_uint2(a, b);
	}
}
// single-block exclusive scan of the per-chunk sums; totals go to blockSums[nBlocks]
__global__ void __launch_bounds__(1024) k_fuse_scan(uint2* __restrict__ blockSums, int nBlocks) {
	__shared__ unsigned sP[1024], sV[1024];
	__shared__ unsigned carryP, carryV;
	if (threadIdx.x == 0) { carryP = 0; carryV = 0; }
	__syncthreads();
	for (int base=0; base<nBlocks; base+=1024) {
		const int i = base+threadIdx.x;
		const uint2 v = i < nBlocks ? blockSums[i] : make_uint2(0, 0);
		sP[threadIdx.x] = v.x; sV[threadIdx.x] = v.y;
		__syncthreads();
		for (int off=1; off<1024; off<<=1) {
			unsigned tp = 0, tv = 0;
			if (threadIdx.x >= off) { tp = sP[threadIdx.x-off]; tv = sV[threadIdx.x-off]; }
			__syncthreads();
			sP[threadIdx.x] += tp; sV[threadIdx.x] += tv;
			__syncthreads();
		}
		if (i < nBlocks) blockSums[i] = make_uint2(carryP+sP[threadIdx.x]-v.x, carryV+sV[threadIdx.x]-v.y);
		__syncthreads();
		if (threadIdx.x == 1023) { carryP += sP[1023]; carryV += sV[1023]; }
		__syncthreads();
	}
	if (threadIdx.x == 0) blockSums[nBlocks] = make_uint2(carryP, carryV);
}

struct FuseOut {
	float* points; float* normals; uint8_t* colors; uint32_t* viewOffsets; uint32_t* views; float* weights;
	unsigned long long basePoint, baseView;
	int estimateColor, estimateNormal;
};

__global__ void __launch_bounds__(256) k_fuse_emit(const FuseArgs a, const uint2* __restrict__ blockSums, const FuseOut out) {
	// each thread owns FUSE_ITEMS consecutive slots of the chunk so that the output keeps raster order
	__shared__ unsigned sP[256], sV[256];
	const FuseView& R = a.views[a.ref];
	const int nSeeds = (int)a.nSeedsPtr->x;
	const int s0 = blockIdx.x*FUSE_CHUNK+threadIdx.x*FUSE_ITEMS;
	unsigned nP = 0, nV = 0;
	for (int i=0; i<FUSE_ITEMS; ++i) { const int s = s0+i; if (s < nSeeds && a.state[s] == 2) { ++nP; nV += 1+__popc(a.mask[s]); } }
	sP[threadIdx.x] = nP; sV[threadIdx.x] = nV;
	__syncthreads();
	for (int off=1; off<256; off<<=1) {
		unsigned tp = 0, tv = 0;
		if (threadIdx.x >= off) { tp = sP[threadIdx.x-off]; tv = sV[threadIdx.x-off]; }
		__syncthreads();
		sP[threadIdx.x] += tp; sV[threadIdx.x] += tv;
		__syncthreads();
	}
	const uint2 bs = blockSums[blockIdx.x];
	unsigned long long ip = out.basePoint+bs.x+(sP[threadIdx.x]-nP);
	unsigned long long iv = out.baseView+bs.y+(sV[threadIdx.x]-nV);
	for (int i=0; i<FUSE_ITEMS; ++i) {
		const int s = s0+i;
		if (s >= nSeeds || a.state[s] != 2) continue;
		const int p = (int)a.seeds[s];
		const int x = p%R.w, y = p/R.w;
		const float4 e = R.dn[p];
		const float depth = e.w;
		const float3 point = seed_point(R, x, y, depth);
		const float3 normal = cam_NormalC2W(R.cam, make_float3(e.x, e.y, e.z));
		const uint32_t merged = a.mask[s];
		// SceneDensify.cpp:3359-3379
		uint32_t vid[HCMVS_MAX_FUSE_VIEWS+1]; float vw[HCMVS_MAX_FUSE_VIEWS+1]; int nv = 1;
		vid[0] = (uint32_t)a.ref; vw[0] = conf2weight(R.conf[p], depth);
		double confidence = (double)vw[0];
		double X0 = (double)(float)dmul((double)point.x, confidence), X1 = (double)(float)dmul((double)point.y, confidence), X2 = (double)(float)dmul((double)point.z, confidence);
		float C[3] = {0.f, 0.f, 0.f};
		if (R.bgr) for (int c=0; c<3; ++c) C[c] = (float)dmul(confidence, (double)(float)R.bgr[(size_t)p*3+c]);
		float3 N = make_float3((float)dmul((double)normal.x, confidence), (float)dmul((double)normal.y, confidence), (float)dmul((double)normal.z, confidence));
		for (int k=0; k<a.nNb; ++k) {
			if (!(merged & (1u<<k))) continue;
			const FuseView& B = a.views[a.nb[k]];
			Probe pr; pr.q = (int)(a.probes[(size_t)k*a.probeStride+s] & 0x3FFFFFFFu); pr.z = 0.f;
			const int xB = pr.q%B.w, yB = pr.q/B.w;
			const float4 eB = B.dn[pr.q];
			const float depthB = eB.w;
			const float confidenceB = conf2weight(B.conf[pr.q], depthB);
			// InsertSort by view id, :3407-3409
			int pos = nv;
			while (pos > 0 && vid[pos-1] > (uint32_t)a.nb[k]) { vid[pos] = vid[pos-1]; vw[pos] = vw[pos-1]; --pos; }
			vid[pos] = (uint32_t)a.nb[k]; vw[pos] = confidenceB; ++nv;
			const D3 XB = cam_I2W(B.cam, (double)xB, (double)yB, (double)depthB);
			X0 = dadd(X0, dmul(XB.x, (double)confidenceB)); X1 = dadd(X1, dmul(XB.y, (double)confidenceB)); X2 = dadd(X2, dmul(XB.z, (double)confidenceB));
			if (out.estimateColor && B.bgr) for (int c=0; c<3; ++c) C[c] = __fadd_rn(C[c], __fmul_rn((float)B.bgr[(size_t)pr.q*3+c], confidenceB));
			if (out.estimateNormal) {
				const float3 normalB = cam_NormalC2W(B.cam, make_float3(eB.x, eB.y, eB.z));
				N.x = __fadd_rn(N.x, __fmul_rn(normalB.x, confidenceB)); N.y = __fadd_rn(N.y, __fmul_rn(normalB.y, confidenceB)); N.z = __fadd_rn(N.z, __fmul_rn(normalB.z, confidenceB));
			}
			confidence = dadd(confidence, (double)confidenceB);
		}
		const double nrm = 1.0/confidence; // :3441-3446
		out.points[ip*3+0] = (float)dmul(X0, nrm); out.points[ip*3+1] = (float)dmul(X1, nrm); out.points[ip*3+2] = (float)dmul(X2, nrm);
		const float fn = (float)nrm;
		if (out.estimateColor) for (int c=0; c<3; ++c) out.colors[ip*3+c] = (uint8_t)min(max(round2int(__fmul_rn(C[c], fn)), 0), 255);
		if (out.estimateNormal) {
			const float3 nvv = make_float3(__fmul_rn(N.x, fn), __fmul_rn(N.y, fn), __fmul_rn(N.z, fn));
			const float inv = __fdiv_rn(1.f, __fsqrt_rn(dot3f(nvv, nvv)));
			out.normals[ip*3+0] = __fmul_rn(nvv.x, inv); out.normals[ip*3+1] = __fmul_rn(nvv.y, inv); out.normals[ip*3+2] = __fmul_rn(nvv.z, inv);
		}
		out.viewOffsets[ip] = (uint32_t)iv;
		for (int j=0; j<nv; ++j) { out.views[iv+j] = vid[j]; out.weights[iv+j] = vw[j]; }
		++ip; iv += nv;
	}
}

// The fork's RemoveSmallSegments (SceneDensify.cpp:2228-2260): depthMap_fuse / normalMap_fuse = the view's estimate where the pixel
// became part of a fused point (arrDepthIdx != NO_ID), 0 elsewhere
__global__ void k_fused_support(const float4* __restrict__ dn, const uint32_t* __restrict__ claim, float* __restrict__ depth, float* __restrict__ normal, size_t n) {
	const size_t i = (size_t)blockIdx.x*blockDim.x+threadIdx.x;
	if (i >= n) return;
	const bool in = claim[i] == CLAIM_TAKEN;
	const float4 e = in ? dn[i] : make_float4(0.f, 0.f, 0.f, 0.f);
	if (depth) depth[i] = e.w;
	if (normal) { normal[i*3] = e.x; normal[i*3+1] = e.y; normal[i*3+2] = e.z; }
}

// MVS::EstimatePointColors (libs/MVS/DepthMap.cpp:2125-2161): the colour of a fused point is sampled in the view — among those that
// see it and hold an image — whose camera is nearest along its optical axis; TImage<Pixel8U>::sample (Common/Types.inl:2248-2258) with
// TPixel<uint8_t>'s operators (Common/Types.h:1930-1936): every product and every sum of the bilinear kernel is truncated to uint8.
__device__ __forceinline__ uint8_t pix_mul(uint8_t c, float v) { return (uint8_t)(int)__fmul_rn(v, (float)c); }
__global__ void __launch_bounds__(256) k_point_colors(const FuseView* __restrict__ views, int nViews, const float* __restrict__ points,
	const uint32_t* __restrict__ offs, const uint32_t* __restrict__ vids, uint8_t* __restrict__ colors, size_t n)
{
	const size_t i = (size_t)blockIdx.x*blockDim.x+threadIdx.x;
	if (i >= n) return;
	const float3 X = make_float3(points[i*3], points[i*3+1], points[i*3+2]);
	double best = (double)3.402823466e38f; int bestView = -1;
	for (uint32_t k=offs[i]; k<offs[i+1]; ++k) {
		const uint32_t id = vids[k];
		if (id >= (uint32_t)nViews || !views[id].bgr) continue; // imageData.image.empty()
		const double* P = views[id].cam.P;
		const double dist = dadd(dadd(dadd(dmul(P[8], (double)X.x), dmul(P[9], (double)X.y)), dmul(P[10], (double)X.z)), P[11]); // Camera::PointDepth
		if (best > dist) { best = dist; bestView = (int)id; }
	}
	uint8_t c0 = 255, c1 = 255, c2 = 255; // Pixel8U::WHITE
	if (bestView >= 0) {
		const FuseView& V = views[bestView];
		const float3 q = cam_ProjectP3f(V.cam, X);
		const float invZ = q.z == 0.f ? 1000000.f : __fdiv_rn(1.f, q.z); // INVERT
		const float px = __fmul_rn(q.x, invZ), py = __fmul_rn(q.y, invZ);
		if (px >= 1.f && py >= 1.f && px <= (float)(V.w-2) && py <= (float)(V.h-2)) { // isInsideWithBorder<float,1>
			const int lx = (int)px, ly = (int)py;
			const float x = __fsub_rn(px, (float)lx), x1 = __fsub_rn(1.f, x), y = __fsub_rn(py, (float)ly), y1 = __fsub_rn(1.f, y);
			const uint8_t* r0 = V.bgr+((size_t)ly*V.w+lx)*3; const uint8_t* r1 = r0+(size_t)V.w*3;
			uint8_t out[3];
			#pragma unroll
			for (int c=0; c<3; ++c) {
				const uint8_t top = (uint8_t)(pix_mul(r0[c], x1)+pix_mul(r0[3+c], x)), bot = (uint8_t)(pix_mul(r1[c], x1)+pix_mul(r1[3+c], x));
				out[c] = (uint8_t)(pix_mul(top, y1)+pix_mul(bot, y));
			}
			c0 = out[0]; c1 = out[1]; c2 = out[2];
		}
	}
	colors[i*3] = c0; colors[i*3+1] = c1; colors[i*3+2] = c2;
}

__global__ void k_fill_u32(uint32_t* p, uint32_t v, size_t n) { const size_t i = (size_t)blockIdx.x*blockDim.x+threadIdx.x; if (i < n) p[i] = v; }

} // namespace hcmvs
using namespace hcmvs;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { hcmvs_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return HCMVS_ERR_CUDA; } } while (0)

struct FuseState {
	FuseView* views_d = nullptr; size_t nViews = 0;
	uint8_t* state_d = nullptr; uint32_t* mask_d = nullptr; size_t pixCap = 0;
	uint2* blockSums_d = nullptr; size_t blockCap = 0;
	uint32_t* seeds_d = nullptr; uint2* seedSums_d = nullptr; // compacted seed pixels of the view being fused + their per-chunk offsets
	int* counters_d = nullptr; int* trace_d = nullptr;
	uint32_t* probes_d = nullptr; size_t probeCap = 0;
	// growing output
	float* points = nullptr; float* normals = nullptr; uint8_t* colors = nullptr; uint32_t* viewOffsets = nullptr; size_t capPoints = 0;
	uint32_t* oviews = nullptr; float* weights = nullptr; size_t capViews = 0;
	int coopBlocks = 0;
	void* pinned = nullptr; size_t pinnedBytes = 0; // page-locked host arena of hcmvs_download_fused_pinned (grow-only)
	// The arena is laid out for CAPACITIES (arenaPoints points, arenaViews view references), so that the ranges a view's emit kernel
	// has just written can be copied out at their final place while the next views are still being fused (download stream); a cloud
	// that outgrows the capacities falls back to one copy at the end and a larger arena for the next scene.
	size_t arenaPoints = 0, arenaViews = 0; bool arenaColor = false, arenaNormal = false;
	bool streamed = false; size_t streamedPoints = 0, streamedViews = 0; cudaEvent_t emitted = nullptr;
	size_t nPoints = 0, nViewRefs = 0; bool hasColor = false, hasNormal = false; // last fused cloud (device resident)
};

// the resident cloud was modified in place (recoloured / normals re-estimated): what was streamed to the arena is stale
void hcmvs_fuse_invalidate_stream(hcmvs_ctx* ctx) { if (ctx && ctx->fuse) ctx->fuse->streamed = false; }

void hcmvs_fuse_release(hcmvs_ctx* ctx) {
	FuseState* f = ctx->fuse; if (!f) return;
	cudaFree(f->trace_d); cudaFree(f->probes_d);
	if (ctx->dlStream) cudaStreamSynchronize(ctx->dlStream);
	if (f->pinned) cudaFreeHost(f->pinned);
	if (f->emitted) cudaEventDestroy(f->emitted);
	cudaFree(f->views_d); cudaFree(f->state_d); cudaFree(f->mask_d); cudaFree(f->blockSums_d); cudaFree(f->counters_d); cudaFree(f->seeds_d); cudaFree(f->seedSums_d);
	cudaFree(f->points); cudaFree(f->normals); cudaFree(f->colors); cudaFree(f->viewOffsets); cudaFree(f->oviews); cudaFree(f->weights);
	delete f; ctx->fuse = nullptr;
}

struct ArenaLayout { size_t oPts, oNrm, oCol, oOff, oViews, oW, total; };
static ArenaLayout LayoutFor(size_t capPoints, size_t capViews, bool hasNormal, bool hasColor) {
	auto al = [](size_t b) { return (b+255)&~(size_t)255; };
	ArenaLayout L; L.oPts = 0; L.oNrm = L.oPts+al(capPoints*12); L.oCol = L.oNrm+al(hasNormal ? capPoints*12 : 0); L.oOff = L.oCol+al(hasColor ? capPoints*3 : 0);
	L.oViews = L.oOff+al((capPoints+1)*4); L.oW = L.oViews+al(capViews*4); L.total = L.oW+al(capViews*4);
	return L;
}

template<typename T>
static int Grow(hcmvs_ctx* ctx, T*& ptr, size_t used, size_t newCap, size_t elemsPer) {
	T* np = nullptr;
	CK(cudaMalloc(&np, newCap*elemsPer*sizeof(T)));
	if (ptr && used) CK(cudaMemcpyAsync(np, ptr, used*elemsPer*sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	if (ctx->dlStream) CK(cudaStreamSynchronize(ctx->dlStream)); // streamed copies may still read the old buffer
	cudaFree(ptr); ptr = np;
	return HCMVS_OK;
}

extern "C" int hcmvs_set_fuse_priority(hcmvs_ctx* ctx, uint32_t view, float score) {
	if (!ctx || view >= ctx->views.size() || !ctx->views[view].set) { hcmvs_set_error("view %u not set", view); return HCMVS_ERR_ARG; }
	ctx->views[view].fusePriority = score; ctx->views[view].hasFusePriority = true;
	return HCMVS_OK;
}

extern "C" int hcmvs_fuse_depthmaps(hcmvs_ctx* ctx, int estimate_color, int estimate_normal, hcmvs_pointcloud* out) {
	if (!ctx) { hcmvs_set_error("null argument"); return HCMVS_ERR_ARG; }
	if (out) memset(out, 0, sizeof(*out));
	cudaSetDevice(ctx->device);
	const hcmvs_params& P = ctx->P;
	const size_t V = ctx->views.size();
	if (!ctx->fuse) ctx->fuse = new FuseState();
	FuseState* f = ctx->fuse;
	const bool debug = getenv("HCMVS_FUSE_DEBUG") != nullptr;
	// connections: valid views sorted by the size of their scored-neighbour list, SceneDensify.cpp:3286-3303 (ties by index)
	struct Conn { uint32_t idx; float score; };
	std::vector<Conn> conns;
	size_t maxPix = 0; bool anyColor = false;
	std::vector<FuseView> hv(V);
	for (size_t i=0; i<V; ++i) {
		View& v = ctx->views[i];
		FuseView& fv = hv[i]; memset(&fv, 0, sizeof(fv));
		if (!v.set || !v.hasMaps) continue;
		{ int r = hcmvs_wait_image(ctx, v); if (r) return r; } // colours
		const size_t n = (size_t)v.w*v.h;
		if (!v.claim_d) CK(cudaMalloc(&v.claim_d, n*4));
		k_fill_u32<<<(unsigned)((n+255)/256), 256, 0, ctx->stream>>>(v.claim_d, CLAIM_FREE, n); ++ctx->nLaunches;
		fv.dn = v.dn_d; fv.conf = v.conf_d; fv.bgr = v.bgr_d; fv.claim = v.claim_d; fv.w = v.w; fv.h = v.h; fv.hasMaps = 1;
		hcmvs_fill_cam(v, fv.cam);
		if (v.bgr_d) anyColor = true;
		if (!v.nbIds.empty()) { conns.push_back(Conn{(uint32_t)i, v.hasFusePriority ? v.fusePriority : (float)v.nbIds.size()}); maxPix = std::max(maxPix, n); }
	}
	if (conns.empty()) { hcmvs_set_error("no view with depth map and neighbours to fuse"); return HCMVS_ERR_STATE; }
	std::stable_sort(conns.begin(), conns.end(), [](const Conn& a, const Conn& b) { return a.score > b.score; });
	if (estimate_color && !anyColor) estimate_color = 0;
	if (f->nViews < V) { CK(cudaStreamSynchronize(ctx->stream)); cudaFree(f->views_d); CK(cudaMalloc(&f->views_d, V*sizeof(FuseView))); f->nViews = V; }
	CK(cudaMemcpyAsync(f->views_d, hv.data(), V*sizeof(FuseView), cudaMemcpyHostToDevice, ctx->stream));
	if (f->pixCap < maxPix) {
		CK(cudaStreamSynchronize(ctx->stream));
		cudaFree(f->state_d); cudaFree(f->mask_d); cudaFree(f->seeds_d);
		CK(cudaMalloc(&f->state_d, maxPix)); CK(cudaMalloc(&f->mask_d, maxPix*4)); CK(cudaMalloc(&f->seeds_d, maxPix*4)); f->pixCap = maxPix;
	}
	const size_t maxBlocks = (maxPix+FUSE_CHUNK-1)/FUSE_CHUNK;
	if (f->blockCap < maxBlocks+1) {
		CK(cudaStreamSynchronize(ctx->stream)); cudaFree(f->blockSums_d); cudaFree(f->seedSums_d);
		CK(cudaMalloc(&f->blockSums_d, (maxBlocks+1)*sizeof(uint2))); CK(cudaMalloc(&f->seedSums_d, (maxBlocks+1)*sizeof(uint2))); f->blockCap = maxBlocks+1;
	}
	if (!f->counters_d) CK(cudaMalloc(&f->counters_d, 4*sizeof(int)));
	if (!f->trace_d) CK(cudaMalloc(&f->trace_d, 256*sizeof(int)));
	size_t maxNb = 1; for (const Conn& c: conns) maxNb = std::max(maxNb, ctx->views[c.idx].nbIds.size());
	if (f->probeCap < maxNb*maxPix) { CK(cudaStreamSynchronize(ctx->stream)); cudaFree(f->probes_d); CK(cudaMalloc(&f->probes_d, maxNb*maxPix*4)); f->probeCap = maxNb*maxPix; }
	if (maxPix >= (1u<<30)) { hcmvs_set_error("depth maps above 2^30 pixels are not supported by the fusion probe cache"); return HCMVS_ERR_UNSUPPORTED; }
	if (!f->coopBlocks) {
		int dev = ctx->device, coop = 0, sms = 0, perSm = 0;
		cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
		if (!coop) { hcmvs_set_error("device lacks cooperative launch"); return HCMVS_ERR_UNSUPPORTED; }
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_fuse_view, 256, 0));
		if (perSm < 1) { hcmvs_set_error("fuse kernel does not fit"); return HCMVS_ERR_CUDA; }
		f->coopBlocks = sms*std::min(perSm, 8);
	}
	const unsigned nMinViewsFuse = std::min<unsigned>(P.nMinViewsFuse, (unsigned)std::count_if(ctx->views.begin(), ctx->views.end(), [](const View& v) { return v.set; }));
	const float FPI = (float)3.14159265358979323846;
	const float normalError = std::cos((P.fNormalDiffThreshold*P.normalweight)*(FPI/180.f));
	size_t nPoints = 0, nViewRefs = 0;
	// stream the cloud out while it is being built when the caller keeps it on the device (out == NULL, the hcmvs_download_fused_pinned
	// path) and an arena laid out for this kind of cloud exists from an earlier scene
	f->streamed = !out && f->pinned && f->arenaPoints && f->arenaColor == (estimate_color != 0) && f->arenaNormal == (estimate_normal != 0);
	f->streamedPoints = f->streamedViews = 0;
	if (f->streamed) {
		if (!ctx->dlStream) CK(cudaStreamCreateWithFlags(&ctx->dlStream, cudaStreamNonBlocking));
		if (!f->emitted) CK(cudaEventCreateWithFlags(&f->emitted, cudaEventDisableTiming));
	}
	hcmvs_time_begin(ctx, ST_FUSE);
	uint64_t totalRounds = 0, totalSeeds = 0, totalProbes = 0;
	for (const Conn& conn: conns) {
		View& v = ctx->views[conn.idx];
		FuseArgs a; memset(&a, 0, sizeof(a));
		a.views = f->views_d; a.ref = (int)conn.idx;
		a.nNb = 0;
		for (uint32_t id: v.nbIds) { if (id < V) a.nb[a.nNb++] = (int)id; }
		a.nMinViewsFuse = nMinViewsFuse;
		a.depthTh = P.fDepthDiffThreshold*P.depthweight; a.normalError = normalError;
		a.state = f->state_d; a.mask = f->mask_d; a.counters = f->counters_d; a.trace = debug ? f->trace_d : nullptr;
		a.probes = f->probes_d; a.probeStride = maxPix;
		const int nPix = v.w*v.h;
		const int nBlocks = (nPix+FUSE_CHUNK-1)/FUSE_CHUNK;
		a.seeds = f->seeds_d; a.nSeedsPtr = f->seedSums_d+nBlocks;
		CK(cudaMemsetAsync(f->counters_d, 0, 4*sizeof(int), ctx->stream));
		// seeds of this view, compacted in raster order; their number stays on the device (grids are sized for the worst case, idle
		// blocks exit at once)
		k_seed_count<<<nBlocks, 256, 0, ctx->stream>>>(a, f->seedSums_d); ++ctx->nLaunches;
		k_fuse_scan<<<1, 1024, 0, ctx->stream>>>(f->seedSums_d, nBlocks); ++ctx->nLaunches;
		k_seed_scatter<<<nBlocks, 256, 0, ctx->stream>>>(a, f->seedSums_d, nBlocks, f->seeds_d); ++ctx->nLaunches;
		k_fuse_probe<<<(nPix+255)/256, 256, 0, ctx->stream>>>(a); ++ctx->nLaunches;
		void* args[] = {(void*)&a};
		CK(cudaLaunchCooperativeKernel((void*)k_fuse_view, dim3(f->coopBlocks), dim3(256), args, 0, ctx->stream)); ++ctx->nLaunches;
		k_fuse_count<<<nBlocks, 256, 0, ctx->stream>>>(f->state_d, f->mask_d, a.nSeedsPtr, f->blockSums_d); ++ctx->nLaunches;
		k_fuse_scan<<<1, 1024, 0, ctx->stream>>>(f->blockSums_d, nBlocks); ++ctx->nLaunches;
		uint2 tot; int cnt[4];
		CK(cudaMemcpyAsync(&tot, f->blockSums_d+nBlocks, sizeof(uint2), cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaMemcpyAsync(cnt, f->counters_d, sizeof(cnt), cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
		totalRounds += (uint64_t)cnt[1]; totalSeeds += (uint64_t)cnt[2]; totalProbes += (uint64_t)cnt[2]*(uint64_t)a.nNb;
		if (debug) {
			int tr[256]; cudaMemcpy(tr, f->trace_d, sizeof(tr), cudaMemcpyDeviceToHost);
			fprintf(stderr, "[fuse] view %u: %d seeds, %d rounds, %u points, %u view refs; undecided after round:", conn.idx, cnt[2], cnt[1], tot.x, tot.y);
			for (int k=0; k<std::min(cnt[1], 12); ++k) fprintf(stderr, " %d", tr[k]);
			fprintf(stderr, "\n");
		}
		if (tot.x == 0) continue;
		if (nPoints+tot.x > f->capPoints) {
			const size_t nc = std::max<size_t>((nPoints+tot.x)*2, (size_t)1<<20);
			int r;
			if ((r = Grow(ctx, f->points, nPoints, nc, 3))) return r;
			if ((r = Grow(ctx, f->normals, nPoints, nc, 3))) return r;
			if ((r = Grow(ctx, f->colors, nPoints, nc, 3))) return r;
			if ((r = Grow(ctx, f->viewOffsets, nPoints, nc+1, 1))) return r;
			f->capPoints = nc;
		}
		if (nViewRefs+tot.y > f->capViews) {
			const size_t nc = std::max<size_t>((nViewRefs+tot.y)*2, (size_t)1<<21);
			int r;
			if ((r = Grow(ctx, f->oviews, nViewRefs, nc, 1))) return r;
			if ((r = Grow(ctx, f->weights, nViewRefs, nc, 1))) return r;
			f->capViews = nc;
		}
		FuseOut fo; fo.points = f->points; fo.normals = f->normals; fo.colors = f->colors; fo.viewOffsets = f->viewOffsets; fo.views = f->oviews; fo.weights = f->weights;
		fo.basePoint = nPoints; fo.baseView = nViewRefs; fo.estimateColor = estimate_color; fo.estimateNormal = estimate_normal;
		k_fuse_emit<<<nBlocks, 256, 0, ctx->stream>>>(a, f->blockSums_d, fo); ++ctx->nLaunches;
		CK(cudaGetLastError());
		if (f->streamed && nPoints+tot.x <= f->arenaPoints && nViewRefs+tot.y <= f->arenaViews) {
			// what this view emitted is final: copy it to its place in the arena behind the emit kernel, on the download stream
			const ArenaLayout L = LayoutFor(f->arenaPoints, f->arenaViews, f->arenaNormal, f->arenaColor);
			char* base = (char*)f->pinned;
			CK(cudaEventRecord(f->emitted, ctx->stream));
			CK(cudaStreamWaitEvent(ctx->dlStream, f->emitted, 0));
			CK(cudaMemcpyAsync(base+L.oPts+nPoints*12, f->points+nPoints*3, (size_t)tot.x*12, cudaMemcpyDeviceToHost, ctx->dlStream));
			if (estimate_normal) CK(cudaMemcpyAsync(base+L.oNrm+nPoints*12, f->normals+nPoints*3, (size_t)tot.x*12, cudaMemcpyDeviceToHost, ctx->dlStream));
			if (estimate_color) CK(cudaMemcpyAsync(base+L.oCol+nPoints*3, f->colors+nPoints*3, (size_t)tot.x*3, cudaMemcpyDeviceToHost, ctx->dlStream));
			CK(cudaMemcpyAsync(base+L.oOff+nPoints*4, f->viewOffsets+nPoints, (size_t)tot.x*4, cudaMemcpyDeviceToHost, ctx->dlStream));
			CK(cudaMemcpyAsync(base+L.oViews+nViewRefs*4, f->oviews+nViewRefs, (size_t)tot.y*4, cudaMemcpyDeviceToHost, ctx->dlStream));
			CK(cudaMemcpyAsync(base+L.oW+nViewRefs*4, f->weights+nViewRefs, (size_t)tot.y*4, cudaMemcpyDeviceToHost, ctx->dlStream));
			f->streamedPoints = nPoints+tot.x; f->streamedViews = nViewRefs+tot.y;
		} else f->streamed = false; // outgrew the arena (or not streaming): hcmvs_download_fused_pinned copies everything at the end
		nPoints += tot.x; nViewRefs += tot.y;
	}
	hcmvs_time_end(ctx);
	for (View& v: ctx->views) if (v.set && v.hasMaps) { int r = hcmvs_mark_image_use(ctx, v); if (r) return r; } // colours were read
	ctx->fuseRounds = totalRounds; ctx->fuseSeeds = totalSeeds; ctx->fuseProbes = totalProbes;
	f->nPoints = nPoints; f->nViewRefs = nViewRefs; f->hasColor = estimate_color != 0; f->hasNormal = estimate_normal != 0;
	if (nPoints) { // close the CSR offsets on the device
		const uint32_t last = (uint32_t)nViewRefs;
		CK(cudaMemcpyAsync(f->viewOffsets+nPoints, &last, 4, cudaMemcpyHostToDevice, ctx->stream));
	}
	CK(cudaStreamSynchronize(ctx->stream));
	if (!out) return HCMVS_OK; // cloud stays on the device (hcmvs_get_fused_device / hcmvs_download_fused)
	out->n_points = nPoints;
	if (nPoints) {
		out->points = (float*)malloc(nPoints*12);
		out->view_offsets = (uint32_t*)malloc((nPoints+1)*4);
		out->views = (uint32_t*)malloc(nViewRefs*4);
		out->weights = (float*)malloc(nViewRefs*4);
		if (estimate_normal) out->normals = (float*)malloc(nPoints*12);
		if (estimate_color) out->colors = (uint8_t*)malloc(nPoints*3);
		if (!out->points || !out->view_offsets || !out->views || !out->weights || (estimate_normal && !out->normals) || (estimate_color && !out->colors)) {
			hcmvs_free_pointcloud(out); hcmvs_set_error("out of host memory"); return HCMVS_ERR_ARG;
		}
		return hcmvs_download_fused(ctx, out->points, out->normals, out->colors, out->view_offsets, out->views, out->weights);
	}
	return HCMVS_OK;
}

extern "C" int hcmvs_download_fused(hcmvs_ctx* ctx, float* points, float* normals, uint8_t* colors, uint32_t* view_offsets, uint32_t* views, float* weights) {
	if (!ctx || !ctx->fuse) { hcmvs_set_error("no fused cloud"); return HCMVS_ERR_STATE; }
	FuseState* f = ctx->fuse;
	cudaSetDevice(ctx->device);
	const size_t n = f->nPoints, m = f->nViewRefs;
	if (!n) return HCMVS_OK;
	if (points) CK(cudaMemcpyAsync(points, f->points, n*12, cudaMemcpyDeviceToHost, ctx->stream));
	if (normals && f->hasNormal) CK(cudaMemcpyAsync(normals, f->normals, n*12, cudaMemcpyDeviceToHost, ctx->stream));
	if (colors && f->hasColor) CK(cudaMemcpyAsync(colors, f->colors, n*3, cudaMemcpyDeviceToHost, ctx->stream));
	if (view_offsets) CK(cudaMemcpyAsync(view_offsets, f->viewOffsets, (n+1)*4, cudaMemcpyDeviceToHost, ctx->stream));
	if (views) CK(cudaMemcpyAsync(views, f->oviews, m*4, cudaMemcpyDeviceToHost, ctx->stream));
	if (weights) CK(cudaMemcpyAsync(weights, f->weights, m*4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return HCMVS_OK;
}

extern "C" int hcmvs_download_fused_pinned(hcmvs_ctx* ctx, hcmvs_pointcloud* out) {
	if (!ctx || !ctx->fuse || !out) { hcmvs_set_error("no fused cloud"); return HCMVS_ERR_STATE; }
	FuseState* f = ctx->fuse;
	cudaSetDevice(ctx->device);
	memset(out, 0, sizeof(*out));
	const size_t n = f->nPoints, m = f->nViewRefs;
	if (!n) return HCMVS_OK;
	const bool fits = f->pinned && n <= f->arenaPoints && m <= f->arenaViews && f->arenaColor == f->hasColor && f->arenaNormal == f->hasNormal;
	if (!fits) {
		if (ctx->dlStream) CK(cudaStreamSynchronize(ctx->dlStream));
		if (f->pinned) cudaFreeHost(f->pinned);
		f->pinned = nullptr; f->pinnedBytes = 0; f->streamed = false;
		// head-room: the next scene of the same size streams into this arena while it is fused
		f->arenaPoints = n+n/8; f->arenaViews = m+m/8; f->arenaColor = f->hasColor; f->arenaNormal = f->hasNormal;
		const ArenaLayout L = LayoutFor(f->arenaPoints, f->arenaViews, f->arenaNormal, f->arenaColor);
		CK(cudaHostAlloc(&f->pinned, L.total, cudaHostAllocDefault));
		f->pinnedBytes = L.total;
	}
	const ArenaLayout L = LayoutFor(f->arenaPoints, f->arenaViews, f->arenaNormal, f->arenaColor);
	char* base = (char*)f->pinned;
	out->n_points = n;
	out->points = (float*)(base+L.oPts); out->view_offsets = (uint32_t*)(base+L.oOff); out->views = (uint32_t*)(base+L.oViews); out->weights = (float*)(base+L.oW);
	if (f->hasNormal) out->normals = (float*)(base+L.oNrm);
	if (f->hasColor) out->colors = (uint8_t*)(base+L.oCol);
	if (f->streamed && f->streamedPoints == n && f->streamedViews == m) {
		// everything but the closing CSR offset already crossed PCIe behind the emit kernels
		CK(cudaMemcpyAsync(out->view_offsets+n, f->viewOffsets+n, 4, cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaStreamSynchronize(ctx->dlStream));
		CK(cudaStreamSynchronize(ctx->stream));
		return HCMVS_OK;
	}
	if (ctx->dlStream) CK(cudaStreamSynchronize(ctx->dlStream));
	return hcmvs_download_fused(ctx, out->points, out->normals, out->colors, out->view_offsets, out->views, out->weights);
}

extern "C" int hcmvs_get_fused_support(hcmvs_ctx* ctx, uint32_t view, float* depth_fuse, float* normal_fuse) {
	if (!ctx || view >= ctx->views.size() || !ctx->views[view].set) { hcmvs_set_error("view %u not set", view); return HCMVS_ERR_ARG; }
	View& v = ctx->views[view];
	if (!v.hasMaps || !v.claim_d) { hcmvs_set_error("view %u took no part in a fusion yet (call hcmvs_fuse_depthmaps)", view); return HCMVS_ERR_STATE; }
	cudaSetDevice(ctx->device);
	const size_t n = (size_t)v.w*v.h;
	float* tmp; int r = hcmvs_scratch(ctx, n*16, (void**)&tmp); if (r) return r;
	k_fused_support<<<(unsigned)((n+255)/256), 256, 0, ctx->stream>>>(v.dn_d, v.claim_d, tmp, tmp+n, n); ++ctx->nLaunches;
	if (depth_fuse) CK(cudaMemcpyAsync(depth_fuse, tmp, n*4, cudaMemcpyDeviceToHost, ctx->stream));
	if (normal_fuse) CK(cudaMemcpyAsync(normal_fuse, tmp+n, n*12, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return HCMVS_OK;
}

extern "C" int hcmvs_estimate_point_colors(hcmvs_ctx* ctx, uint64_t n_points, const float* points, const uint32_t* view_offsets, const uint32_t* views, uint8_t* colors) {
	if (!ctx) { hcmvs_set_error("null context"); return HCMVS_ERR_ARG; }
	cudaSetDevice(ctx->device);
	const size_t V = ctx->views.size();
	std::vector<FuseView> hv(V);
	for (size_t i=0; i<V; ++i) {
		View& v = ctx->views[i];
		FuseView& fv = hv[i]; memset(&fv, 0, sizeof(fv));
		if (!v.set) continue;
		{ int r = hcmvs_wait_image(ctx, v); if (r) return r; }
		fv.bgr = v.bgr_d; fv.w = v.w; fv.h = v.h;
		hcmvs_fill_cam(v, fv.cam);
	}
	const bool resident = points == nullptr; // act on the fused cloud that lives on the device
	FuseState* f = ctx->fuse;
	if (resident && (!f || !f->nPoints)) { hcmvs_set_error("no fused cloud on the device (call hcmvs_fuse_depthmaps) and no points given"); return HCMVS_ERR_STATE; }
	if (!resident && (!view_offsets || !views || !colors)) { hcmvs_set_error("null argument"); return HCMVS_ERR_ARG; }
	const size_t n = resident ? f->nPoints : (size_t)n_points;
	if (!n) return HCMVS_OK;
	const size_t m = resident ? f->nViewRefs : (size_t)view_offsets[n];
	auto al = [](size_t b) { return (b+255)&~(size_t)255; };
	const size_t bViews = al(V*sizeof(FuseView)), bPts = resident ? 0 : al(n*12), bOff = resident ? 0 : al((n+1)*4), bIds = resident ? 0 : al(m*4), bCol = resident ? 0 : al(n*3);
	char* base; { int r = hcmvs_scratch(ctx, bViews+bPts+bOff+bIds+bCol, (void**)&base); if (r) return r; }
	CK(cudaMemcpyAsync(base, hv.data(), V*sizeof(FuseView), cudaMemcpyHostToDevice, ctx->stream));
	const float* pts_d = f ? f->points : nullptr; const uint32_t* off_d = f ? f->viewOffsets : nullptr; const uint32_t* ids_d = f ? f->oviews : nullptr; uint8_t* col_d = f ? f->colors : nullptr;
	if (!resident) {
		CK(cudaMemcpyAsync(base+bViews, points, n*12, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(base+bViews+bPts, view_offsets, (n+1)*4, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(base+bViews+bPts+bOff, views, m*4, cudaMemcpyHostToDevice, ctx->stream));
		pts_d = (const float*)(base+bViews); off_d = (const uint32_t*)(base+bViews+bPts); ids_d = (const uint32_t*)(base+bViews+bPts+bOff); col_d = (uint8_t*)(base+bViews+bPts+bOff+bIds);
	}
	k_point_colors<<<(unsigned)((n+255)/256), 256, 0, ctx->stream>>>((const FuseView*)base, (int)V, pts_d, off_d, ids_d, col_d, n); ++ctx->nLaunches;
	CK(cudaGetLastError());
	if (resident) { f->hasColor = true; f->streamed = false; } // the arena copy (if any) no longer matches
	if (colors) CK(cudaMemcpyAsync(colors, col_d, n*3, cudaMemcpyDeviceToHost, ctx->stream));
	for (View& v: ctx->views) if (v.set) { int r = hcmvs_mark_image_use(ctx, v); if (r) return r; }
	CK(cudaStreamSynchronize(ctx->stream));
	return HCMVS_OK;
}

extern "C" int hcmvs_get_fused_device(hcmvs_ctx* ctx, uint64_t* n_points, uint64_t* n_view_refs, void** points_d, void** normals_d, void** colors_d,
	void** view_offsets_d, void** views_d, void** weights_d)
{
	if (!ctx || !ctx->fuse) { hcmvs_set_error("no fused cloud"); return HCMVS_ERR_STATE; }
	FuseState* f = ctx->fuse;
	if (n_points) *n_points = f->nPoints;
	if (n_view_refs) *n_view_refs = f->nViewRefs;
	if (points_d) *points_d = f->points;
	if (normals_d) *normals_d = f->hasNormal ? f->normals : nullptr;
	if (colors_d) *colors_d = f->hasColor ? f->colors : nullptr;
	if (view_offsets_d) *view_offsets_d = f->viewOffsets;
	if (views_d) *views_d = f->oviews;
	if (weights_d) *weights_d = f->weights;
	return HCMVS_OK;
}

extern "C" void hcmvs_free_pointcloud(hcmvs_pointcloud* pc) {
	if (!pc) return;
	free(pc->points); free(pc->normals); free(pc->colors); free(pc->view_offsets); free(pc->views); free(pc->weights);
	memset(pc, 0, sizeof(*pc));
}
