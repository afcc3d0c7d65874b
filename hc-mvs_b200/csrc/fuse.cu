// DepthMapsData::FuseDepthMaps (libs/MVS/SceneDensify.cpp:3265-3495) on the GPU, with the CPU's exact
// sequential semantics.
//
// The reference fuses greedily: views in connection order, pixels in raster order; a seed pixel probes one
// pixel in each neighbour view, claims the agreeing ones, and — if it survives nMinViewsFuse — zeroes the
// depths it occludes. Later seeds see those claims / zeroed depths, so the result is order dependent.
// That order is kept here by "deterministic reservations": views are still processed one after another, but
// inside a view every undecided seed reserves (atomicMin of its raster index) each neighbour pixel it would
// touch; a seed that holds a reservation precedes every undecided seed it conflicts with, so it can run its
// CPU action on that pixel immediately. Seeds with disjoint footprints commit in the same round; the lowest
// undecided seed always commits, so the loop terminates, and the outcome is bit-identical to the raster scan.
//
// Round 2 layout (profiles/r02_notes.md): the whole scene is fused by ONE persistent cooperative kernel
// (k_fuse_scene; grid.sync between the stages of a view, no launch gaps, no host round trip per view), over
// 32-byte per-pixel read-only RECORDS {depth, world normal, weight, colour} built once per scene, so that every
// probe of a neighbour pixel is ONE aligned DRAM sector (the round-1 kernels gathered float4 + claim word +
// conf + colour from four arrays: 2 sectors per probe, 3 more per merged view in the emit); what changes during
// the fusion — liveness and reservations — lives in a 1-bit-per-pixel bitmap and three rotating planes of claim words.
#include "hcmvs_internal.h"
#include "camera.cuh"
#include <cooperative_groups.h>
#include <vector>
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <sched.h>

namespace cg = cooperative_groups;

namespace hcmvs {

#define CLAIM_FREE 0xFFFFFFFFu

// one pixel of one view as the fusion reads it: 32 bytes = one DRAM sector, READ-ONLY during the fusion (what changes — liveness
// and reservations — lives in the bitmap and the compact claim array below), derived once from the view's maps: world-space normal (Cast<float>(R^T n), SceneDensify.cpp:3358), Conf2Weight(conf, depth)
// (:154-156, :3357), the colour packed b | g<<8 | r<<16
struct __align__(32) FuseRec { float depth; uint32_t pad0; float nx, ny; float nz, weight; uint32_t bgr; uint32_t pad1; };
static_assert(sizeof(FuseRec) == 32, "one sector per pixel");

struct FuseView {
	FuseRec* rec; float4* dn; const uint8_t* bgr;
	uint32_t* alive;   // 1 bit per pixel: depth != 0, not merged into a point, not zeroed (SceneDensify.cpp:3347-3354, :3396-3398). All views
	                   // together are ~12 MB: L2 resident. It IS the liveness: every transition clears the bit before the phase's barrier.
	uint32_t* claim;   // per pixel: CLAIM_FREE or the raster index of the seed that reserved it (4 B/px: 8 pixels per sector); three planes
	                   // (FuseJob::claimPlane apart) that rotate through the rounds of a view: read / being written / being cleared
	int w, h; int hasMaps; int hasBgr;
	CamConst cam;
};

// a view in fusion order with the neighbours it probes (DepthData::neighbors, all of them — not only the matching ones)
struct FusePlanView { int view; int nNb; int nb[HCMVS_MAX_FUSE_VIEWS]; int order[HCMVS_MAX_FUSE_VIEWS+1]; }; // order: -1 (the view itself) and 0..nNb-1 sorted by view id

struct FuseOut {
	float* points; float* normals; uint8_t* colors; uint32_t* viewOffsets; uint32_t* views; float* weights;
	unsigned long long capPoints, capRefs;
	int estimateColor, estimateNormal;
};

// per-seed-slot state of the view being fused; two sets, so that the emit of view r-1 shares its stages with the seed scan of view r
struct FuseSlots {
	uint32_t* seeds;  // slot -> pixel (raster order)
	uint8_t* state;   // 0 removed, 1 undecided, 2 emitted, 3 emitted but still waiting for contested probes
	uint32_t* mask;   // merged neighbours (bit k = nb[k]) of emitted seeds
	uint32_t* probes; // [neighbour][slot] probe cache, stride probeStride
	float4* stgPoint; float4* stgNormal; // staged fused point (xyz, packed colour) and normal of a final slot
	float* stgW;      // [1 + neighbour][slot] staged weights (0: the seed's own), stride probeStride
};

struct FuseCtl {
	unsigned wlCount[3]; int overflow;        // worklist counters rotate like the claim planes: read / being filled / being zeroed
	unsigned stCount[2]; unsigned pad[2];     // deferred staging lists of the two slot sets
	unsigned long long cumPoints, cumRefs;   // cloud size after the views emitted so far
	unsigned long long rounds, seeds, probes;
};

struct FuseJob {
	FuseView* views; const FusePlanView* plan; int nPlan;
	FuseSlots sl[2]; size_t probeStride; size_t claimPlane;
	uint32_t* wl[2];                  // undecided slots of the current / next round
	uint32_t* stl[2];                 // per slot set: seeds that became final in a later round; their points are staged by the next view's count stage
	unsigned* blkSeeds; unsigned* blkEmit; // per-block counts of the two ordered compactions (blkEmit: points, view references)
	FuseCtl* ctl;
	unsigned nMinViewsFuse; float depthTh, normalError;
	FuseOut out;
	volatile unsigned long long* progress; // host-mapped: [0] views emitted, [1+2r], [2+2r] cloud size after view r (nullable)
	unsigned long long* trace;        // debug: 6 globaltimer stamps + 2 counters per view (HCMVS_FUSE_DEBUG)
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
#define PROBE_PIX   0x3FFFFFFFu

// Everything the stages of the persistent kernel exchange goes through L2 (ld.cg / st.cg): the kernel never ends between the
// stages, so L1 lines would go stale.
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define FUSE_STAMP(i) do { if (a.trace && gtid == 0 && r < a.nPlan) a.trace[8*r+(i)] = gtimer(); } while (0)

// ------------------------------------------------------------------ records: one pass over a view's maps
__global__ void __launch_bounds__(256) k_fuse_build(const FuseView* __restrict__ views, int view, const float* __restrict__ conf, unsigned long long* __restrict__ nValid) {
	const FuseView& V = views[view];
	const size_t n = (size_t)V.w*V.h;
	const size_t i = (size_t)blockIdx.x*blockDim.x+threadIdx.x;
	unsigned valid = 0;
	if (i < n) {
		const float4 e = V.dn[i];
		const float3 nw = cam_NormalC2W(V.cam, make_float3(e.x, e.y, e.z));
		uint32_t c = 0;
		if (V.bgr) { const uint8_t* s = V.bgr+i*3; c = (uint32_t)s[0] | ((uint32_t)s[1]<<8) | ((uint32_t)s[2]<<16); }
		uint4* dst = reinterpret_cast<uint4*>(V.rec+i);
		dst[0] = make_uint4(__float_as_uint(e.w), 0u, __float_as_uint(nw.x), __float_as_uint(nw.y));
		dst[1] = make_uint4(__float_as_uint(nw.z), __float_as_uint(conf2weight(conf[i], e.w)), c, 0u);
		valid = e.w != 0.f;
	}
	const unsigned word = __ballot_sync(0xffffffffu, valid);
	if ((threadIdx.x&31) == 0 && i < n) V.alive[i>>5] = word;
	// one global atomic per block (one per warp = 60 k atomics on one address per C2 view: they, not the 105 MB of traffic, set the kernel's time)
	__shared__ unsigned sValid;
	if (threadIdx.x == 0) sValid = 0;
	__syncthreads();
	if ((threadIdx.x&31) == 0 && word) atomicAdd(&sValid, (unsigned)__popc(word));
	__syncthreads();
	if (threadIdx.x == 0 && sValid) atomicAdd(nValid, (unsigned long long)sValid);
}

// ------------------------------------------------------------------ block helpers (256 threads)
#define FUSE_NT 256
// exclusive prefix of v over the block + block total; `sw` = 8 words of shared memory
__device__ __forceinline__ unsigned block_scan(unsigned v, unsigned* sw, unsigned& total) {
	const int lane = threadIdx.x&31, wid = threadIdx.x>>5;
	unsigned inc = v;
	#pragma unroll
	for (int o=1; o<32; o<<=1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
	__syncthreads(); // sw may still be read by the previous call
	if (lane == 31) sw[wid] = inc;
	__syncthreads();
	unsigned base = 0, tot = 0;
	#pragma unroll
	for (int i=0; i<FUSE_NT/32; ++i) { const unsigned t = sw[i]; if (i < wid) base += t; tot += t; }
	total = tot;
	return base+inc-v;
}
// sum over the block
__device__ __forceinline__ unsigned long long block_sum(unsigned long long v, unsigned long long* sw) {
	#pragma unroll
	for (int o=16; o>0; o>>=1) v += __shfl_xor_sync(0xffffffffu, v, o);
	__syncthreads();
	if ((threadIdx.x&31) == 0) sw[threadIdx.x>>5] = v;
	__syncthreads();
	unsigned long long t = 0;
	#pragma unroll
	for (int i=0; i<FUSE_NT/32; ++i) t += sw[i];
	return t;
}

// ------------------------------------------------------------------ stages of one view
#ifndef FUSE_FN
#define FUSE_FN __forceinline__
#endif
#ifdef FUSE_CHECK
#define FUSE_CHK(cond, id) do { if (!(cond)) { atomicCAS(&a.ctl->overflow, 0, (id)); return; } } while (0)
#define FUSE_CHKR(cond, id) do { if (!(cond)) { atomicCAS(&a.ctl->overflow, 0, (id)); return 0; } } while (0)
#else
#define FUSE_CHK(cond, id) do { } while (0)
#define FUSE_CHKR(cond, id) do { } while (0)
#endif
#ifndef FUSE_CH
#define FUSE_CH 6      // probes whose bitmap tests / record gathers are in flight together in the probe stage
#endif
#define FUSE_RCH 12    // probes per chunk of the reserve / resolve stages (DepthData::neighbors holds <= nMaxViews = 12)
__device__ __forceinline__ bool bit_alive(const FuseView& B, uint32_t q) { return (__ldcg(B.alive+(q>>5))>>(q&31)) & 1u; }
__device__ __forceinline__ void bit_clear(const FuseView& B, uint32_t q) { atomicAnd(B.alive+(q>>5), ~(1u<<(q&31))); }

// classify the probes of seed slot s (the f64 geometry, once) and place the first round's reservations. The bitmap answers "already
// part of a point / no depth" without touching the pixel's record: only probes of live pixels cost a DRAM sector.
__device__ FUSE_FN void probe_slot(const FuseJob& a, const FusePlanView& pv, const FuseView* sv, const FuseSlots& S, const int s) {
	const FuseView& R = sv[0];
	const int p = (int)__ldcg(S.seeds+s);
	FUSE_CHK(s >= 0 && (size_t)s < a.probeStride, 101); FUSE_CHK(p >= 0 && p < R.w*R.h, 102); FUSE_CHK(pv.nNb >= 0 && pv.nNb <= HCMVS_MAX_FUSE_VIEWS, 103);
	const uint4 r0 = __ldcg(reinterpret_cast<const uint4*>(R.rec+p));
	const float nz = __ldcg(&R.rec[p].nz);
	const float depth = __uint_as_float(r0.x);
	S.state[s] = 1; __stcg(S.mask+s, 0u);
	const int x = p%R.w, y = p/R.w;
	const float3 point = seed_point(R, x, y, depth);
	const float3 normal = make_float3(__uint_as_float(r0.z), __uint_as_float(r0.w), nz);
	for (int k0=0; k0<pv.nNb; k0+=FUSE_CH) {
		Probe pr[FUSE_CH]; uint32_t word[FUSE_CH]; uint4 eB[FUSE_CH]; float nzB[FUSE_CH];
		#pragma unroll
		for (int j=0; j<FUSE_CH; ++j) {
			pr[j].q = -1; pr[j].z = 0.f;
			if (k0+j < pv.nNb) { const FuseView& B = sv[1+(k0+j)]; if (B.hasMaps) pr[j] = probe_view(B, point); }
		}
		#pragma unroll
		for (int j=0; j<FUSE_CH; ++j) { word[j] = 0u; if (pr[j].q >= 0) word[j] = __ldcg(sv[1+(k0+j)].alive+(pr[j].q>>5)); }
		#pragma unroll
		for (int j=0; j<FUSE_CH; ++j) {
			eB[j] = make_uint4(0u, 0u, 0u, 0u); nzB[j] = 0.f;
			if (pr[j].q >= 0 && ((word[j]>>(pr[j].q&31)) & 1u)) { const FuseRec* rb = sv[1+(k0+j)].rec+pr[j].q; eB[j] = __ldcg(reinterpret_cast<const uint4*>(rb)); nzB[j] = __ldcg(&rb->nz); }
		}
		#pragma unroll
		for (int j=0; j<FUSE_CH; ++j) {
			if (k0+j < pv.nNb) {
				uint32_t code = PROBE_DEAD;
				const float depthB = __uint_as_float(eB[j].x); // 0 when the pixel is not alive (records of live pixels hold depth != 0)
				if (depthB != 0.f) {
					uint32_t cls = PROBE_NONE;
					bool merge = false;
					if (depth_similar(pr[j].z, depthB, a.depthTh))
						merge = dot3f(normal, make_float3(__uint_as_float(eB[j].z), __uint_as_float(eB[j].w), nzB[j])) > a.normalError;
					if (merge) cls = PROBE_MERGE; else if (pr[j].z < depthB) cls = PROBE_INVAL;
					code = (uint32_t)pr[j].q | (cls<<30);
					if (cls != PROBE_NONE) atomicMin(sv[1+(k0+j)].claim+pr[j].q, (uint32_t)p);
				}
				__stcg(S.probes+(size_t)(k0+j)*a.probeStride+s, code);
			}
		}
	}
}

// Resolve what the raster order already fixes; returns 1 while the seed stays unfinished, 2 when it became final with a point, 0 otherwise.
// A seed holding a reservation precedes every unfinished seed that touches the same pixel, so that pixel is in the state the seed
// would find it in at its turn of the raster scan. Hence:
//  * a seed whose own view + already merged + held agreeing pixels reach nMinViewsFuse WILL be emitted (:3429) whatever its
//    contested probes turn into: it acts on the pixels it holds at once (merges / zeroes them) and only keeps waiting for the
//    contested ones — this is what cuts the dependency chains along the rows;
//  * a seed that cannot reach nMinViewsFuse even if every contested agreeing pixel were still alive at its turn will NOT be
//    emitted: it releases everything;
//  * otherwise it waits for the lower seeds it conflicts with (the lowest unfinished seed never waits).
// Round k reads the claim plane (k-1)%3 (the reservations of the seeds that were unfinished after round k-1; round 1: those of the probe
// stage), and every seed that stays unfinished places its reservations for round k+1 in plane k%3 in the SAME phase — one grid-wide
// phase per round instead of reserve + resolve. Plane (k+1)%3, read a round ago and written again in the next one, is cleared on
// the way: an entry that still matters there belongs to a live pixel whose holder is unfinished, i.e. in this round's worklist (a
// holder that finished either killed the pixel or released the entry), so the worklist seeds clearing their own pixels is enough.
// Readers of the read plane tolerate the releases of this phase (a contested pixel reads as contested either way).
__device__ FUSE_FN int resolve_slot(const FuseJob& a, const FusePlanView& pv, const FuseView* sv, const FuseSlots& S, const int s, const unsigned round,
	uint32_t& pOut, uint32_t& mergedOut, uint32_t (&code0)[FUSE_RCH])
{
	// one trip for the slot's state, seed and mask (probe_slot initialises the mask)
	const uint8_t st = __ldcg(S.state+s);
	const uint32_t p = __ldcg(S.seeds+s);
	uint32_t merged = __ldcg(S.mask+s);
	if (st != 1 && st != 3) return 0;
	const size_t cR = (size_t)((round-1u)%3u)*a.claimPlane, cW = (size_t)(round%3u)*a.claimPlane, cC = (size_t)((round+1u)%3u)*a.claimPlane;
	const FuseView& R = sv[0];
	FUSE_CHKR(s >= 0 && (size_t)s < a.probeStride, 301); FUSE_CHKR((int)p < R.w*R.h, 302);
	uint32_t heldMerge = 0, heldInval = 0, want = 0; // want: live probes the seed still depends on (held or contested)
	unsigned nContested = 0, nContestedMerge = 0;
	for (int k0=0; k0<pv.nNb; k0+=FUSE_RCH) {
		uint32_t code[FUSE_RCH], word[FUSE_RCH], cl[FUSE_RCH];
		#pragma unroll
		for (int j=0; j<FUSE_RCH; ++j) code[j] = k0+j < pv.nNb ? __ldcg(S.probes+(size_t)(k0+j)*a.probeStride+s) : PROBE_DEAD;
		#pragma unroll
		for (int j=0; j<FUSE_RCH; ++j) {
			word[j] = 0u; cl[j] = CLAIM_FREE;
			if (code[j] != PROBE_DEAD && (code[j]>>30) != PROBE_NONE) {
				const FuseView& B = sv[1+(k0+j)];
				const uint32_t q = code[j]&PROBE_PIX;
				FUSE_CHKR((int)q < B.w*B.h && B.claim && B.alive, 303);
				word[j] = __ldcg(B.alive+(q>>5)); cl[j] = __ldcg(B.claim+cR+q);
				if (round >= 2u) __stcg(B.claim+cC+q, CLAIM_FREE);
			}
		}
		#pragma unroll
		for (int j=0; j<FUSE_RCH; ++j) {
			if (k0 == 0) code0[j] = code[j];
			if (code[j] == PROBE_DEAD || (code[j]>>30) == PROBE_NONE) continue;
			const int k = k0+j;
			if (!((word[j]>>(code[j]&31u)) & 1u)) { __stcg(S.probes+(size_t)k*a.probeStride+s, PROBE_DEAD); continue; }
			want |= 1u<<k;
			if (cl[j] == p) { if ((code[j]>>30) == PROBE_MERGE) heldMerge |= 1u<<k; else heldInval |= 1u<<k; }
			else { ++nContested; nContestedMerge += (code[j]>>30) == PROBE_MERGE; }
		}
	}
	pOut = p;
	const uint32_t held = heldMerge|heldInval;
	const unsigned nViews = 1u+__popc(merged)+__popc(heldMerge);
	int res;
	if (st == 3 || nViews >= a.nMinViewsFuse) {
		#pragma unroll
		for (int k=0; k<FUSE_RCH; ++k) {
			if (!(held & (1u<<k))) continue;
			const FuseView& B = sv[1+(k)];
			uint32_t* pc = S.probes+(size_t)k*a.probeStride+s;
			const uint32_t q = code0[k] & PROBE_PIX;
			if (heldMerge & (1u<<k)) __stcg(pc, q | (PROBE_NONE<<30)); // a merged probe is not looked at again; the pixel index stays for the point
			else { B.dn[q].w = 0.f; __stcg(pc, PROBE_DEAD); }          // SceneDensify.cpp:3447-3449: the occluded depth is zeroed
			bit_clear(B, q);
		}
		for (uint32_t h = held>>FUSE_RCH<<FUSE_RCH; h; h &= h-1) { // neighbours beyond the first chunk (more than 12: not the reference's default)
			const int k = __ffs(h)-1;
			const FuseView& B = sv[1+(k)];
			uint32_t* pc = S.probes+(size_t)k*a.probeStride+s;
			const uint32_t q = __ldcg(pc) & PROBE_PIX;
			if (heldMerge & (1u<<k)) __stcg(pc, q | (PROBE_NONE<<30));
			else { B.dn[q].w = 0.f; __stcg(pc, PROBE_DEAD); }
			bit_clear(B, q);
		}
		merged |= heldMerge;
		__stcg(S.mask+s, merged);
		mergedOut = merged;
		if (nContested == 0) { bit_clear(R, p); S.state[s] = 2; return 2; }
		S.state[s] = 3;
		want &= ~held; // acted on: those pixels are dead now
		res = 1;
	} else if (nViews+nContestedMerge < a.nMinViewsFuse) {
		#pragma unroll
		for (int k=0; k<FUSE_RCH; ++k) if (held & (1u<<k)) __stcg(sv[1+(k)].claim+cR+(code0[k]&PROBE_PIX), CLAIM_FREE);
		for (uint32_t h = held>>FUSE_RCH<<FUSE_RCH; h; h &= h-1) {
			const int k = __ffs(h)-1;
			__stcg(sv[1+(k)].claim+cR+(__ldcg(S.probes+(size_t)k*a.probeStride+s)&PROBE_PIX), CLAIM_FREE);
		}
		S.state[s] = 0;
		return 0;
	} else res = 1;
	// unfinished: the reservations of the next round (the raster index orders them)
	#pragma unroll
	for (int k=0; k<FUSE_RCH; ++k) if (want & (1u<<k)) atomicMin(sv[1+(k)].claim+cW+(code0[k]&PROBE_PIX), p);
	for (uint32_t h = want>>FUSE_RCH<<FUSE_RCH; h; h &= h-1) {
		const int k = __ffs(h)-1;
		atomicMin(sv[1+(k)].claim+cW+(__ldcg(S.probes+(size_t)k*a.probeStride+s)&PROBE_PIX), p);
	}
	return res;
}

// append the unfinished slots of this thread's warp to the next round's worklist: the counter's atomic is issued first and its
// answer used last (wl_commit), so that the staging in between does not wait behind the round trip. Full warps only.
__device__ __forceinline__ unsigned wl_begin(unsigned* count, bool keep, unsigned& m) {
	m = __ballot_sync(0xffffffffu, keep);
	unsigned base = 0;
	if (m && (int)(threadIdx.x&31) == __ffs(m)-1) base = atomicAdd(count, (unsigned)__popc(m));
	return base;
}
__device__ __forceinline__ void wl_commit(uint32_t* wl, unsigned base, unsigned m, bool keep, int s) {
	if (!m) return;
	base = __shfl_sync(0xffffffffu, base, __ffs(m)-1);
	if (keep) __stcg(wl+base+__popc(m & ((1u<<(threadIdx.x&31))-1u)), (uint32_t)s);
}

// The fused point of seed slot s (SceneDensify.cpp:3355-3446), computed the moment the seed becomes final — in the same phase that
// just read the heads of its merged pixels, so their records are still in L2 — and parked in per-slot staging; the ordered
// compaction of the next stage copies it to its place in the cloud. The sums run over the merged neighbours in the order the seed
// probed them (the reference's order: rounding is part of the result), four records in flight at a time.
__device__ FUSE_FN void stage_slot(const FuseJob& a, const FusePlanView& pv, const FuseView* sv, const FuseSlots& S, const int s, const int p, const uint32_t merged,
	const uint32_t (&code0)[FUSE_RCH])
{
	const FuseView& R = sv[0];
	const FuseOut& out = a.out;
	const uint4 r0 = __ldcg(reinterpret_cast<const uint4*>(R.rec+p)), r1 = __ldcg(reinterpret_cast<const uint4*>(R.rec+p)+1);
	const int x = p%R.w, y = p/R.w;
	const float depth = __uint_as_float(r0.x);
	const float3 point = seed_point(R, x, y, depth);
	const float3 normal = make_float3(__uint_as_float(r0.z), __uint_as_float(r0.w), __uint_as_float(r1.x));
	// SceneDensify.cpp:3359-3379
	const float w0 = __uint_as_float(r1.y);
	__stcg(S.stgW+s, w0);
	double confidence = (double)w0;
	double X0 = (double)(float)dmul((double)point.x, confidence), X1 = (double)(float)dmul((double)point.y, confidence), X2 = (double)(float)dmul((double)point.z, confidence);
	float C[3] = {0.f, 0.f, 0.f};
	if (R.hasBgr) for (int c=0; c<3; ++c) C[c] = (float)dmul(confidence, (double)(float)((r1.z>>(8*c))&255u));
	float3 N = make_float3((float)dmul((double)normal.x, confidence), (float)dmul((double)normal.y, confidence), (float)dmul((double)normal.z, confidence));
	for (uint32_t m = merged; m; ) {
		int kk[4]; uint32_t qq[4]; uint4 b0[4], b1[4];
		#pragma unroll
		for (int j=0; j<4; ++j) { kk[j] = m ? __ffs(m)-1 : -1; m &= m-1; }
		#pragma unroll
		for (int j=0; j<4; ++j) { // the pixel of a merged probe: the resolve stage's registers for the first 12 neighbours, the probe cache beyond
			qq[j] = 0u;
			if (kk[j] >= FUSE_RCH) qq[j] = __ldcg(S.probes+(size_t)kk[j]*a.probeStride+s) & PROBE_PIX;
			else {
				#pragma unroll
				for (int k=0; k<FUSE_RCH; ++k) if (kk[j] == k) qq[j] = code0[k] & PROBE_PIX;
			}
		}
		#pragma unroll
		for (int j=0; j<4; ++j) if (kk[j] >= 0) { const FuseRec* rb = sv[1+kk[j]].rec+qq[j]; b0[j] = __ldcg(reinterpret_cast<const uint4*>(rb)); b1[j] = __ldcg(reinterpret_cast<const uint4*>(rb)+1); }
		#pragma unroll
		for (int j=0; j<4; ++j) {
			if (kk[j] < 0) continue;
			const FuseView& B = sv[1+kk[j]];
			const int q = (int)qq[j];
			const int xB = q%B.w, yB = q/B.w;
			const float depthB = __uint_as_float(b0[j].x);
			const float confidenceB = __uint_as_float(b1[j].y);
			__stcg(S.stgW+(size_t)(1+kk[j])*a.probeStride+s, confidenceB);
			const D3 XB = cam_I2W(B.cam, (double)xB, (double)yB, (double)depthB);
			X0 = dadd(X0, dmul(XB.x, (double)confidenceB)); X1 = dadd(X1, dmul(XB.y, (double)confidenceB)); X2 = dadd(X2, dmul(XB.z, (double)confidenceB));
			if (out.estimateColor && B.hasBgr) for (int c=0; c<3; ++c) C[c] = __fadd_rn(C[c], __fmul_rn((float)((b1[j].z>>(8*c))&255u), confidenceB));
			if (out.estimateNormal) {
				const float3 normalB = make_float3(__uint_as_float(b0[j].z), __uint_as_float(b0[j].w), __uint_as_float(b1[j].x));
				N.x = __fadd_rn(N.x, __fmul_rn(normalB.x, confidenceB)); N.y = __fadd_rn(N.y, __fmul_rn(normalB.y, confidenceB)); N.z = __fadd_rn(N.z, __fmul_rn(normalB.z, confidenceB));
			}
			confidence = dadd(confidence, (double)confidenceB);
		}
	}
	const double nrm = 1.0/confidence; // :3441-3446
	const float fn = (float)nrm;
	uint32_t col = 0;
	if (out.estimateColor) for (int c=0; c<3; ++c) col |= (uint32_t)min(max(round2int(__fmul_rn(C[c], fn)), 0), 255)<<(8*c);
	__stcg(S.stgPoint+s, make_float4((float)dmul(X0, nrm), (float)dmul(X1, nrm), (float)dmul(X2, nrm), __uint_as_float(col)));
	if (out.estimateNormal) {
		const float3 nvv = make_float3(__fmul_rn(N.x, fn), __fmul_rn(N.y, fn), __fmul_rn(N.z, fn));
		const float inv = __fdiv_rn(1.f, __fsqrt_rn(dot3f(nvv, nvv)));
		__stcg(S.stgNormal+s, make_float4(__fmul_rn(nvv.x, inv), __fmul_rn(nvv.y, inv), __fmul_rn(nvv.z, inv), 0.f));
	}
}

// staging of a seed that became final in a later round, deferred to the count stage of the next view: those rounds are chains of
// dependent loads over a few hundred slots (one or two blocks busy, the rest of the grid waiting at the barrier), and the staging
// — the seed's record, its merged neighbours' records, the f64 sums — was more than half of every such chain
__device__ FUSE_FN void stage_deferred(const FuseJob& a, const FusePlanView& pv, const FuseView* sv, const FuseSlots& S, const int s) {
	const uint32_t p = __ldcg(S.seeds+s), merged = __ldcg(S.mask+s);
	uint32_t code0[FUSE_RCH];
	#pragma unroll
	for (int k=0; k<FUSE_RCH; ++k) code0[k] = (k < pv.nNb && (merged & (1u<<k))) ? __ldcg(S.probes+(size_t)k*a.probeStride+s) : PROBE_DEAD;
	stage_slot(a, pv, sv, S, s, (int)p, merged, code0);
}

// copy the staged point of final slot s to its place in the cloud; the view list is written sorted by view id (InsertSort,
// SceneDensify.cpp:3407-3409) through the plan's id order
__device__ __forceinline__ void emit_slot(const FuseJob& a, const FusePlanView& pv, const FuseSlots& S, const int s, const unsigned long long ip, const unsigned long long iv) {
	const FuseOut& out = a.out;
	const float4 P = __ldcg(S.stgPoint+s);
	const uint32_t merged = __ldcg(S.mask+s);
	float4 Nn = make_float4(0.f, 0.f, 0.f, 0.f); if (out.estimateNormal) Nn = __ldcg(S.stgNormal+s);
	out.points[ip*3+0] = P.x; out.points[ip*3+1] = P.y; out.points[ip*3+2] = P.z;
	if (out.estimateColor) { const uint32_t c = __float_as_uint(P.w); out.colors[ip*3+0] = (uint8_t)(c&255u); out.colors[ip*3+1] = (uint8_t)((c>>8)&255u); out.colors[ip*3+2] = (uint8_t)((c>>16)&255u); }
	if (out.estimateNormal) { out.normals[ip*3+0] = Nn.x; out.normals[ip*3+1] = Nn.y; out.normals[ip*3+2] = Nn.z; }
	out.viewOffsets[ip] = (uint32_t)iv;
	unsigned long long o = iv;
	for (int j0=0; j0<=pv.nNb; j0+=8) { // ids ascending: the reference view or neighbour k = order[j]; 8 weights in flight, then their stores
		float wv[8]; int kx[8];
		#pragma unroll
		for (int j=0; j<8; ++j) {
			kx[j] = -2;
			if (j0+j <= pv.nNb) { const int k = pv.order[j0+j]; if (k < 0 || (merged & (1u<<k))) { kx[j] = k; wv[j] = __ldcg(S.stgW+(size_t)(1+k)*a.probeStride+s); } }
		}
		#pragma unroll
		for (int j=0; j<8; ++j) if (kx[j] != -2) { out.views[o] = (uint32_t)(kx[j] < 0 ? pv.view : pv.nb[kx[j]]); out.weights[o] = wv[j]; ++o; }
	}
}

// ------------------------------------------------------------------ the whole scene: one persistent cooperative kernel
// Per view r (fusion order), separated by grid barriers:
//  [C] count the seeds of r per block (+ the final slots of r-1, + the deferred staging of r-1's later-round finals)
//  [D] ordered scatter of the seeds of r into slots (+ ordered copy of r-1's staged points into the cloud, progress to the host)
//  [P] classify every seed's probes (the f64 geometry, once) and place the first reservations (claim plane 0)
//  [R] round 1 over every slot: resolve, reserve for round 2 (plane 1), stage the points of the seeds that became final — their
//      merged records were fetched by [P] and are still in L2
//  [R'] rounds k >= 2 over the worklist of unfinished slots, ONE phase each: resolve against plane (k-1)%3, reserve in plane k%3,
//      clear plane (k+1)%3; seeds that become final are listed for the next view's [C]
// Blocks own contiguous bitmap-word / slot ranges in the two ordered compactions, so the slots and the cloud keep raster order with
// one block-count prefix (<= gridDim.x words) per compaction. (Processing the slots in L2-sized tiles was measured and rejected: a
// tile step cannot be shorter than its dependent chain — profiles/r02_notes.md.)
#ifndef FUSE_MINB
#define FUSE_MINB 2 // 128 registers: 3 CTAs/SM (80 registers) spill ~800 B per thread into the dependent chains (measured: 4.74 vs 4.24 ms per 16 C2 views)
#endif
__global__ void __launch_bounds__(FUSE_NT, FUSE_MINB) k_fuse_scene(const FuseJob a) {
	cg::grid_group grid = cg::this_grid();
	__shared__ unsigned long long sw64[FUSE_NT/32];
	__shared__ unsigned sw[FUSE_NT/32];
	// the plan entry of the view being fused and the views it touches ([0] the view itself, [1+k] neighbour k), per block in shared
	// memory; set r&1 belongs to view r, so the other set still describes view r-1 for its emit / deferred staging. Every stage is a
	// chain of dependent loads: fetching descriptors and cameras through pointers in global memory added two trips to each link.
	__shared__ FusePlanView sPlan[2];
	__shared__ FuseView sView[2][HCMVS_MAX_FUSE_VIEWS+1];
	const int G = (int)gridDim.x, b = (int)blockIdx.x;
	const int gtid = b*FUSE_NT+threadIdx.x, nThreads = G*FUSE_NT;
	unsigned nSeedsPrev = 0;
	for (int r=0; r<=a.nPlan; ++r) {
		const FuseSlots& S = a.sl[r&1];
		const FuseSlots& Sp = a.sl[(r&1)^1];
		const FusePlanView* pv = r < a.nPlan ? a.plan+r : nullptr;
		const FusePlanView* pvp = r > 0 ? a.plan+(r-1) : nullptr;
		if (pv) {
			static_assert(sizeof(FusePlanView)%4 == 0 && sizeof(FuseView)%8 == 0, "word copies");
			const uint32_t* src = reinterpret_cast<const uint32_t*>(pv); uint32_t* dst = reinterpret_cast<uint32_t*>(&sPlan[r&1]);
			for (int i=threadIdx.x; i<(int)(sizeof(FusePlanView)/4); i+=FUSE_NT) dst[i] = src[i];
			const int nV = 1+pv->nNb, per = (int)(sizeof(FuseView)/8);
			for (int i=threadIdx.x; i<nV*per; i+=FUSE_NT) {
				const int v = i/per, o = i-v*per;
				const int id = v == 0 ? pv->view : pv->nb[v-1];
				reinterpret_cast<unsigned long long*>(&sView[r&1][v])[o] = reinterpret_cast<const unsigned long long*>(a.views+id)[o];
			}
		}
		__syncthreads();
		int wLo = 0, wHi = 0; // bitmap words of view r this block compacts
		const uint32_t* aliveR = nullptr;
		if (pv) {
			const FuseView& R = a.views[pv->view];
			aliveR = R.alive;
			const int nWords = (R.w*R.h+31)>>5, L = (nWords+G-1)/G;
			wLo = min(b*L, nWords); wHi = min(wLo+L, nWords);
		}
		int sLo = 0, sHi = 0;
		if (pvp) { const int L = (int)((nSeedsPrev+G-1)/G); sLo = min(b*L, (int)nSeedsPrev); sHi = min(sLo+L, (int)nSeedsPrev); }
		FUSE_STAMP(0);
		// ---- [C] counts (+ the deferred staging of view r-1)
		{
			if (pvp) { const unsigned nSt = __ldcg(&a.ctl->stCount[(r&1)^1]); for (int i=gtid; i<(int)nSt; i+=nThreads) stage_deferred(a, sPlan[(r&1)^1], sView[(r&1)^1], Sp, (int)__ldcg(a.stl[(r&1)^1]+i)); }
			unsigned long long nS = 0, nP = 0, nV = 0;
			for (int i=wLo+threadIdx.x; i<wHi; i+=FUSE_NT) nS += (unsigned)__popc(__ldcg(aliveR+i));
			for (int s=sLo+threadIdx.x; s<sHi; s+=FUSE_NT)
				{ const uint8_t st = __ldcg(Sp.state+s); const uint32_t mk = __ldcg(Sp.mask+s); if (st == 2) { ++nP; nV += 1u+(unsigned)__popc(mk); } }
			nS = block_sum(nS, sw64); nP = block_sum(nP, sw64); nV = block_sum(nV, sw64);
			if (threadIdx.x == 0) { a.blkSeeds[b] = (unsigned)nS; a.blkEmit[2*b] = (unsigned)nP; a.blkEmit[2*b+1] = (unsigned)nV; }
		}
		grid.sync();
		FUSE_STAMP(1);
		// ---- [D] block prefixes, ordered scatter of the seeds of r, ordered emit of r-1
		unsigned nSeeds = 0, totP = 0, totV = 0;
		bool fits = true;
		{
			unsigned preS = 0, preP = 0, preV = 0, totS = 0;
			for (int i=threadIdx.x; i<G; i+=FUSE_NT) {
				const unsigned cs = __ldcg(a.blkSeeds+i), cp = __ldcg(a.blkEmit+2*i), cv = __ldcg(a.blkEmit+2*i+1);
				totS += cs; totP += cp; totV += cv;
				if (i < b) { preS += cs; preP += cp; preV += cv; }
			}
			// six block sums through three packed 64-bit reductions (each component < 2^32)
			const unsigned long long s1 = block_sum(((unsigned long long)preS<<32) | totS, sw64);
			const unsigned long long s2 = block_sum(((unsigned long long)preP<<32) | totP, sw64);
			const unsigned long long s3 = block_sum(((unsigned long long)preV<<32) | totV, sw64);
			preS = (unsigned)(s1>>32); totS = (unsigned)s1; preP = (unsigned)(s2>>32); totP = (unsigned)s2; preV = (unsigned)(s3>>32); totV = (unsigned)s3;
			nSeeds = totS;
			if (pv) {
				unsigned base = preS;
				for (int i0=wLo; i0<wHi; i0+=FUSE_NT) {
					const int i = i0+threadIdx.x;
					uint32_t word = i < wHi ? __ldcg(aliveR+i) : 0u;
					unsigned tot1;
					unsigned o = base+block_scan((unsigned)__popc(word), sw, tot1);
					for (; word; word &= word-1) __stcg(S.seeds+(o++), (uint32_t)(i*32+__ffs(word)-1));
					base += tot1;
				}
			}
			if (pvp) {
				const unsigned long long cumP = __ldcg(&a.ctl->cumPoints), cumV = __ldcg(&a.ctl->cumRefs);
				fits = cumP+totP <= a.out.capPoints && cumV+totV <= a.out.capRefs;
				unsigned long long ip = cumP+preP, iv = cumV+preV;
				for (int s0=sLo; s0<sHi; s0+=FUSE_NT) {
					const int s = s0+threadIdx.x;
					uint8_t st = 0; uint32_t mk = 0;
					if (s < sHi) { st = __ldcg(Sp.state+s); mk = __ldcg(Sp.mask+s); } // one trip
					const bool f = st == 2;
					const unsigned nv = f ? 1u+(unsigned)__popc(mk) : 0u;
					unsigned tot1;
					const unsigned ex = block_scan((f ? 1u<<16 : 0u) | nv, sw, tot1); // points in the high half, view references in the low half
					if (f && fits) emit_slot(a, sPlan[(r&1)^1], Sp, s, ip+(ex>>16), iv+(ex&0xFFFFu));
					ip += tot1>>16; iv += tot1&0xFFFFu;
				}
			}
			if (gtid == 0) { a.ctl->wlCount[0] = 0; a.ctl->wlCount[1] = 0; a.ctl->wlCount[2] = 0; a.ctl->stCount[r&1] = 0; if (pv) { a.ctl->seeds += nSeeds; a.ctl->probes += (unsigned long long)nSeeds*(unsigned)pv->nNb; } }
		}
		grid.sync();
		if (pvp && gtid == 0) { // every block has read the old totals: advance them; the points of view r-1 are final and visible
			if (!fits) a.ctl->overflow = 1;
			else { a.ctl->cumPoints += totP; a.ctl->cumRefs += totV; }
			if (a.progress) { // tell the host, which streams them out
				a.progress[1+2*(r-1)] = a.ctl->cumPoints; a.progress[2+2*(r-1)] = a.ctl->cumRefs;
				__threadfence_system();
				a.progress[0] = (unsigned long long)r;
			}
		}
		if (!pv) break;
		FUSE_STAMP(2);
		// ---- [P] probes + first reservations
		for (int s=gtid; s<(int)nSeeds; s+=nThreads) probe_slot(a, sPlan[r&1], sView[r&1], S, s);
		grid.sync();
		FUSE_STAMP(3);
		// ---- [R] round 1 over every slot; a seed that became final stages its point at once (its merged records were fetched by [P]);
		// round k pushes its unfinished slots to worklist k%2 / counter k%3 and zeroes counter (k+1)%3 for the round after
		for (int s0=b*FUSE_NT; s0<(int)nSeeds; s0+=nThreads) {
			const int s = s0+threadIdx.x;
			uint32_t p = 0, merged = 0, code0[FUSE_RCH];
			const int res = s < (int)nSeeds ? resolve_slot(a, sPlan[r&1], sView[r&1], S, s, 1u, p, merged, code0) : 0;
			unsigned m; const unsigned base = wl_begin(&a.ctl->wlCount[1], res == 1, m);
			if (res == 2) stage_slot(a, sPlan[r&1], sView[r&1], S, s, (int)p, merged, code0);
			wl_commit(a.wl[1], base, m, res == 1, s);
		}
		grid.sync();
		FUSE_STAMP(4);
		// ---- later rounds over the worklist of unfinished slots: ONE phase per round (resolve + next reservations, see resolve_slot)
		unsigned rounds = 1; unsigned long long wlSum = 0, syncNs = 0; // syncNs (debug): what thread 0 waits in the rounds' barriers
		for (unsigned k=2; ; ++k) {
			const unsigned nw = __ldcg(&a.ctl->wlCount[(k-1u)%3u]);
			if (!nw) break;
			wlSum += nw;
			const uint32_t* wl = a.wl[(k-1u)&1u];
			if (gtid == 0) a.ctl->wlCount[(k+1u)%3u] = 0;
			for (int i0=b*FUSE_NT; i0<(int)nw; i0+=nThreads) {
				const int i = i0+threadIdx.x;
				const int s = i < (int)nw ? (int)__ldcg(wl+i) : 0;
				uint32_t p = 0, merged = 0, code0[FUSE_RCH];
				const int res = i < (int)nw ? resolve_slot(a, sPlan[r&1], sView[r&1], S, s, k, p, merged, code0) : 0;
				unsigned m, m2; const unsigned base = wl_begin(&a.ctl->wlCount[k%3u], res == 1, m), base2 = wl_begin(&a.ctl->stCount[r&1], res == 2, m2);
				wl_commit(a.wl[k&1u], base, m, res == 1, s);
				wl_commit(a.stl[r&1], base2, m2, res == 2, s);
			}
			const unsigned long long t0 = a.trace && gtid == 0 ? gtimer() : 0ull;
			grid.sync();
			if (a.trace && gtid == 0) syncNs += gtimer()-t0;
			++rounds;
		}
		FUSE_STAMP(5);
		if (a.trace && gtid == 0) { a.trace[8*r+6] = nSeeds | ((unsigned long long)rounds<<32); a.trace[8*r+7] = (wlSum & 0xFFFFFFFFull) | (syncNs<<32); }
		if (gtid == 0) a.ctl->rounds += rounds;
		nSeedsPrev = nSeeds;
	}
}

// The fork's RemoveSmallSegments (SceneDensify.cpp:2228-2260): depthMap_fuse / normalMap_fuse = the view's estimate where the pixel
// became part of a fused point (arrDepthIdx != NO_ID), 0 elsewhere
__global__ void k_fused_support(const float4* __restrict__ dn, const uint32_t* __restrict__ alive, float* __restrict__ depth, float* __restrict__ normal, size_t n) {
	const size_t i = (size_t)blockIdx.x*blockDim.x+threadIdx.x;
	if (i >= n) return;
	const bool in = !((alive[i>>5]>>(i&31)) & 1u) && dn[i].w != 0.f; // no longer alive, yet not zeroed: merged into a point
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
	FuseRec* rec_d = nullptr; size_t recCap = 0; std::vector<size_t> recOff; uint32_t* alive_d = nullptr; size_t aliveCap = 0; std::vector<size_t> aliveOff; uint32_t* claim_d = nullptr; size_t claimCap = 0; // per-view records (kept after the fusion: claims, zeroed depths)
	FusePlanView* plan_d = nullptr; size_t planCap = 0;
	// per-seed-slot arrays (two sets) + worklists, sized for the largest view
	uint8_t* state_d[2] = {nullptr, nullptr}; uint32_t* mask_d[2] = {nullptr, nullptr}; uint32_t* seeds_d[2] = {nullptr, nullptr}; uint32_t* wl_d[2] = {nullptr, nullptr}; uint32_t* stl_d[2] = {nullptr, nullptr}; size_t pixCap = 0;
	uint32_t* probes_d[2] = {nullptr, nullptr}; float* stgW_d[2] = {nullptr, nullptr}; size_t probeCap = 0;
	float4* stgPoint_d[2] = {nullptr, nullptr}; float4* stgNormal_d[2] = {nullptr, nullptr};
	unsigned* blk_d = nullptr; int coopBlocks = 0; size_t smemEmit = 0;
	FuseCtl* ctl_d = nullptr; unsigned long long* nValid_d = nullptr;
	unsigned long long* progress = nullptr; unsigned long long* progress_dev = nullptr; size_t progressCap = 0; // page-locked, mapped
	// output (sized for the upper bound: every point consumes >= nMinViewsFuse valid pixels)
	float* points = nullptr; float* normals = nullptr; uint8_t* colors = nullptr; uint32_t* viewOffsets = nullptr; size_t capPoints = 0;
	uint32_t* oviews = nullptr; float* weights = nullptr; size_t capViews = 0;
	void* pinned = nullptr; size_t pinnedBytes = 0; // page-locked host arena of hcmvs_download_fused_pinned (grow-only)
	// The arena is laid out for CAPACITIES (arenaPoints points, arenaViews view references), so that the ranges a view has just
	// emitted can be copied out at their final place while the next views are still being fused (download stream); a cloud
	// that outgrows the capacities falls back to one copy at the end and a larger arena for the next scene.
	size_t arenaPoints = 0, arenaViews = 0; bool arenaColor = false, arenaNormal = false;
	bool streamed = false; size_t streamedPoints = 0, streamedViews = 0;
	size_t nPoints = 0, nViewRefs = 0; bool hasColor = false, hasNormal = false; // last fused cloud (device resident)
};

// the resident cloud was modified in place (recoloured / normals re-estimated): what was streamed to the arena is stale
void hcmvs_fuse_invalidate_stream(hcmvs_ctx* ctx) { if (ctx && ctx->fuse) ctx->fuse->streamed = false; }

void hcmvs_fuse_release(hcmvs_ctx* ctx) {
	FuseState* f = ctx->fuse; if (!f) return;
	if (ctx->dlStream) cudaStreamSynchronize(ctx->dlStream);
	if (f->pinned) cudaFreeHost(f->pinned);
	if (f->progress) cudaFreeHost(f->progress);
	cudaFree(f->views_d); cudaFree(f->rec_d); cudaFree(f->alive_d); cudaFree(f->claim_d); cudaFree(f->plan_d); cudaFree(f->blk_d); cudaFree(f->ctl_d); cudaFree(f->nValid_d);
	for (int i=0; i<2; ++i) { cudaFree(f->state_d[i]); cudaFree(f->mask_d[i]); cudaFree(f->seeds_d[i]); cudaFree(f->wl_d[i]); cudaFree(f->stl_d[i]); cudaFree(f->probes_d[i]); cudaFree(f->stgW_d[i]); cudaFree(f->stgPoint_d[i]); cudaFree(f->stgNormal_d[i]); }
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
static int Realloc(hcmvs_ctx* ctx, T*& ptr, size_t elems) {
	CK(cudaStreamSynchronize(ctx->stream));
	if (ctx->dlStream) CK(cudaStreamSynchronize(ctx->dlStream)); // streamed copies may still read the old buffer
	cudaFree(ptr); ptr = nullptr;
	CK(cudaMalloc(&ptr, elems*sizeof(T)));
	return HCMVS_OK;
}

extern "C" int hcmvs_set_fuse_priority(hcmvs_ctx* ctx, uint32_t view, float score) {
	if (!ctx || view >= ctx->views.size() || !ctx->views[view].set) { hcmvs_set_error("view %u not set", view); return HCMVS_ERR_ARG; }
	ctx->views[view].fusePriority = score; ctx->views[view].hasFusePriority = true;
	return HCMVS_OK;
}

extern "C" int hcmvs_fuse_depthmaps(hcmvs_ctx* ctx, int estimate_color, int estimate_normal, hcmvs_pointcloud* out) {
	if (ctx) ctx->freshScene = false; // a consumer of other views' maps ends the "own maps only" phase of hcmvs_begin_scene
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
	size_t maxPix = 0, totalPix = 0, totalWords = 0; bool anyColor = false;
	std::vector<size_t>& aliveOff = f->aliveOff; aliveOff.assign(V, 0);
	std::vector<FuseView> hv(V);
	f->recOff.assign(V, (size_t)-1);
	for (size_t i=0; i<V; ++i) {
		View& v = ctx->views[i];
		FuseView& fv = hv[i]; memset(&fv, 0, sizeof(fv));
		if (!v.set || !v.hasMaps) continue;
		v.depthValid = false; // occluded depths are zeroed in place
		{ int r = hcmvs_wait_image(ctx, v); if (r) return r; } // colours
		const size_t n = (size_t)v.w*v.h;
		if (n >= ((size_t)1<<30)) { hcmvs_set_error("depth maps above 2^30 pixels are not supported by the fusion probe cache"); return HCMVS_ERR_UNSUPPORTED; }
		f->recOff[i] = totalPix; totalPix += n; aliveOff[i] = totalWords; totalWords += (n+31)/32;
		fv.dn = v.dn_d; fv.bgr = v.bgr_d; fv.hasBgr = v.bgr_d != nullptr; fv.w = v.w; fv.h = v.h; fv.hasMaps = 1;
		hcmvs_fill_cam(v, fv.cam);
		if (v.bgr_d) anyColor = true;
		if (!v.nbIds.empty()) { conns.push_back(Conn{(uint32_t)i, v.hasFusePriority ? v.fusePriority : (float)v.nbIds.size()}); maxPix = std::max(maxPix, n); }
	}
	if (conns.empty()) { hcmvs_set_error("no view with depth map and neighbours to fuse"); return HCMVS_ERR_STATE; }
	std::stable_sort(conns.begin(), conns.end(), [](const Conn& a, const Conn& b) { return a.score > b.score; });
	if (estimate_color && !anyColor) estimate_color = 0;
	if (f->recCap < totalPix) { int r = Realloc(ctx, f->rec_d, totalPix); if (r) return r; f->recCap = totalPix; }
	if (f->aliveCap < totalWords) { int r = Realloc(ctx, f->alive_d, totalWords); if (r) return r; f->aliveCap = totalWords; }
	if (f->claimCap < 3*totalPix) { int r = Realloc(ctx, f->claim_d, 3*totalPix); if (r) return r; f->claimCap = 3*totalPix; }
	for (size_t i=0; i<V; ++i) if (f->recOff[i] != (size_t)-1) { hv[i].rec = f->rec_d+f->recOff[i]; hv[i].alive = f->alive_d+aliveOff[i]; hv[i].claim = f->claim_d+f->recOff[i]; }
	if (f->nViews < V) { int r = Realloc(ctx, f->views_d, V); if (r) return r; f->nViews = V; }
	CK(cudaMemcpyAsync(f->views_d, hv.data(), V*sizeof(FuseView), cudaMemcpyHostToDevice, ctx->stream));
	// the plan: views in fusion order with the neighbours each probes
	std::vector<FusePlanView> plan(conns.size());
	size_t maxNb = 1;
	for (size_t r=0; r<conns.size(); ++r) {
		FusePlanView& pv = plan[r]; memset(&pv, 0, sizeof(pv));
		pv.view = (int)conns[r].idx;
		for (uint32_t id: ctx->views[conns[r].idx].nbIds) if (id < V && pv.nNb < HCMVS_MAX_FUSE_VIEWS) pv.nb[pv.nNb++] = (int)id;
		maxNb = std::max(maxNb, (size_t)pv.nNb);
		// ascending view ids (the reference keeps a point's view list sorted: InsertSort, SceneDensify.cpp:3407-3409); stable like it
		for (int j=0; j<=pv.nNb; ++j) pv.order[j] = j-1;
		std::stable_sort(pv.order, pv.order+pv.nNb+1, [&](int x, int y) { return (x < 0 ? pv.view : pv.nb[x]) < (y < 0 ? pv.view : pv.nb[y]); });
	}
	if (f->planCap < plan.size()) { int r = Realloc(ctx, f->plan_d, plan.size()); if (r) return r; f->planCap = plan.size(); }
	CK(cudaMemcpyAsync(f->plan_d, plan.data(), plan.size()*sizeof(FusePlanView), cudaMemcpyHostToDevice, ctx->stream));
	if (f->pixCap < maxPix) {
		for (int i=0; i<2; ++i) {
			int r;
			if ((r = Realloc(ctx, f->state_d[i], maxPix)) || (r = Realloc(ctx, f->mask_d[i], maxPix)) || (r = Realloc(ctx, f->seeds_d[i], maxPix)) || (r = Realloc(ctx, f->wl_d[i], maxPix)) || (r = Realloc(ctx, f->stl_d[i], maxPix)) ||
			    (r = Realloc(ctx, f->stgPoint_d[i], maxPix)) || (r = Realloc(ctx, f->stgNormal_d[i], maxPix))) return r;
		}
		f->pixCap = maxPix;
	}
	if (f->probeCap < maxNb*maxPix) {
		for (int i=0; i<2; ++i) { int r; if ((r = Realloc(ctx, f->probes_d[i], maxNb*maxPix)) || (r = Realloc(ctx, f->stgW_d[i], (maxNb+1)*maxPix))) return r; }
		f->probeCap = maxNb*maxPix;
	}
	const size_t smemEmit = 0;
	if (!f->coopBlocks || f->smemEmit != smemEmit) {
		cudaFree(f->blk_d); cudaFree(f->ctl_d); cudaFree(f->nValid_d); f->blk_d = nullptr; f->ctl_d = nullptr; f->nValid_d = nullptr;
		f->smemEmit = smemEmit;
		int dev = ctx->device, coop = 0, sms = 0, perSm = 0;
		cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
		if (!coop) { hcmvs_set_error("device lacks cooperative launch"); return HCMVS_ERR_UNSUPPORTED; }
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (smemEmit > 48*1024) CK(cudaFuncSetAttribute(k_fuse_scene, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemEmit));
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_fuse_scene, FUSE_NT, smemEmit));
		if (perSm < 1) { hcmvs_set_error("fuse kernel does not fit"); return HCMVS_ERR_CUDA; }
		if (const char* e = getenv("HCMVS_FUSE_BLOCKS_PER_SM")) perSm = std::max(1, std::min(perSm, atoi(e)));
		f->coopBlocks = sms*perSm;
		CK(cudaMalloc(&f->blk_d, (size_t)f->coopBlocks*3*sizeof(unsigned)));
		CK(cudaMalloc(&f->ctl_d, sizeof(FuseCtl)));
		CK(cudaMalloc(&f->nValid_d, sizeof(unsigned long long)));
	}
	if (f->progressCap < 1+2*plan.size()) {
		if (f->progress) { CK(cudaStreamSynchronize(ctx->stream)); cudaFreeHost(f->progress); f->progress = nullptr; }
		f->progressCap = 1+2*plan.size();
		CK(cudaHostAlloc(&f->progress, f->progressCap*sizeof(unsigned long long), cudaHostAllocMapped));
		CK(cudaHostGetDevicePointer(&f->progress_dev, f->progress, 0));
	}
	const unsigned nSet = (unsigned)std::count_if(ctx->views.begin(), ctx->views.end(), [](const View& v) { return v.set; });
	const unsigned nMinViewsFuse = std::min<unsigned>(P.nMinViewsFuse, nSet);
	const float FPI = (float)3.14159265358979323846;
	const float normalError = std::cos((P.fNormalDiffThreshold*P.normalweight)*(FPI/180.f));
	hcmvs_time_begin(ctx, ST_FUSE);
	// ---- records of every view with maps + the number of valid depths (bounds the cloud)
	CK(cudaMemsetAsync(f->nValid_d, 0, sizeof(unsigned long long), ctx->stream));
	CK(cudaMemsetAsync(f->ctl_d, 0, sizeof(FuseCtl), ctx->stream));
	CK(cudaMemsetAsync(f->claim_d, 0xFF, 3*totalPix*sizeof(uint32_t), ctx->stream)); // CLAIM_FREE, the three rotating planes
	for (size_t i=0; i<V; ++i) {
		if (f->recOff[i] == (size_t)-1) continue;
		const size_t n = (size_t)hv[i].w*hv[i].h;
		k_fuse_build<<<(unsigned)((n+255)/256), 256, 0, ctx->stream>>>(f->views_d, (int)i, ctx->views[i].conf_d, f->nValid_d); ++ctx->nLaunches;
	}
	unsigned long long nValid = 0;
	CK(cudaMemcpyAsync(&nValid, f->nValid_d, sizeof(nValid), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	// a fused point consumes its seed pixel and every pixel it merges, and needs >= nMinViewsFuse of them
	const size_t needPoints = (size_t)(nValid/std::max(nMinViewsFuse, 1u))+1, needRefs = (size_t)nValid+1;
	if (f->capPoints < needPoints) {
		int r;
		if ((r = Realloc(ctx, f->points, needPoints*3)) || (r = Realloc(ctx, f->normals, needPoints*3)) || (r = Realloc(ctx, f->colors, needPoints*3)) || (r = Realloc(ctx, f->viewOffsets, needPoints+1))) return r;
		f->capPoints = needPoints;
	}
	if (f->capViews < needRefs) {
		int r;
		if ((r = Realloc(ctx, f->oviews, needRefs)) || (r = Realloc(ctx, f->weights, needRefs))) return r;
		f->capViews = needRefs;
	}
	// stream the cloud out while it is being built when the caller keeps it on the device (out == NULL, the hcmvs_download_fused_pinned
	// path) and an arena laid out for this kind of cloud exists from an earlier scene
	f->streamed = !out && f->pinned && f->arenaPoints && f->arenaColor == (estimate_color != 0) && f->arenaNormal == (estimate_normal != 0) && !getenv("HCMVS_FUSE_NO_STREAM");
	f->streamedPoints = f->streamedViews = 0;
	if (f->streamed && !ctx->dlStream) CK(cudaStreamCreateWithFlags(&ctx->dlStream, cudaStreamNonBlocking));
	memset(f->progress, 0, f->progressCap*sizeof(unsigned long long));
	FuseJob a; memset(&a, 0, sizeof(a));
	a.views = f->views_d; a.plan = f->plan_d; a.nPlan = (int)plan.size();
	for (int i=0; i<2; ++i) { a.sl[i].seeds = f->seeds_d[i]; a.sl[i].state = f->state_d[i]; a.sl[i].mask = f->mask_d[i]; a.sl[i].probes = f->probes_d[i]; a.sl[i].stgPoint = f->stgPoint_d[i]; a.sl[i].stgNormal = f->stgNormal_d[i]; a.sl[i].stgW = f->stgW_d[i]; a.wl[i] = f->wl_d[i]; a.stl[i] = f->stl_d[i]; }
	a.probeStride = maxPix; a.claimPlane = totalPix;
	a.blkSeeds = f->blk_d; a.blkEmit = f->blk_d+f->coopBlocks;
	a.ctl = f->ctl_d;
	a.nMinViewsFuse = nMinViewsFuse; a.depthTh = P.fDepthDiffThreshold*P.depthweight; a.normalError = normalError;
	a.out.points = f->points; a.out.normals = f->normals; a.out.colors = f->colors; a.out.viewOffsets = f->viewOffsets; a.out.views = f->oviews; a.out.weights = f->weights;
	a.out.capPoints = f->capPoints; a.out.capRefs = f->capViews; a.out.estimateColor = estimate_color; a.out.estimateNormal = estimate_normal;
	a.progress = f->progress_dev;
	unsigned long long* trace_d = nullptr;
	if (debug) { CK(cudaMalloc(&trace_d, plan.size()*8*sizeof(unsigned long long))); CK(cudaMemsetAsync(trace_d, 0, plan.size()*8*sizeof(unsigned long long), ctx->stream)); a.trace = trace_d; }
	void* args[] = {(void*)&a};
	CK(cudaLaunchCooperativeKernel((void*)k_fuse_scene, dim3(f->coopBlocks), dim3(FUSE_NT), args, smemEmit, ctx->stream)); ++ctx->nLaunches;
	hcmvs_time_end(ctx);
	if (f->streamed) {
		// follow the kernel: whatever a view emitted is final, so it crosses PCIe at its final place in the arena while the next views
		// are still being fused
		const ArenaLayout L = LayoutFor(f->arenaPoints, f->arenaViews, f->arenaNormal, f->arenaColor);
		char* base = (char*)f->pinned;
		volatile unsigned long long* pg = f->progress;
		size_t done = 0, sentP = 0, sentV = 0, spins = 0;
		while (done < plan.size()) {
			const size_t now = (size_t)pg[0];
			if (now == done) {
				if ((++spins & 63) == 0) { const cudaError_t q = cudaStreamQuery(ctx->stream); if (q != cudaErrorNotReady) { if ((size_t)pg[0] == done) break; } }
				sched_yield();
				continue;
			}
			done = now;
			const size_t nP = (size_t)pg[1+2*(done-1)], nV = (size_t)pg[2+2*(done-1)];
			if (nP > f->arenaPoints || nV > f->arenaViews) { f->streamed = false; break; } // outgrew the arena: one copy at the end
			if (nP > sentP) {
				CK(cudaMemcpyAsync(base+L.oPts+sentP*12, f->points+sentP*3, (nP-sentP)*12, cudaMemcpyDeviceToHost, ctx->dlStream));
				if (estimate_normal) CK(cudaMemcpyAsync(base+L.oNrm+sentP*12, f->normals+sentP*3, (nP-sentP)*12, cudaMemcpyDeviceToHost, ctx->dlStream));
				if (estimate_color) CK(cudaMemcpyAsync(base+L.oCol+sentP*3, f->colors+sentP*3, (nP-sentP)*3, cudaMemcpyDeviceToHost, ctx->dlStream));
				CK(cudaMemcpyAsync(base+L.oOff+sentP*4, f->viewOffsets+sentP, (nP-sentP)*4, cudaMemcpyDeviceToHost, ctx->dlStream));
				CK(cudaMemcpyAsync(base+L.oViews+sentV*4, f->oviews+sentV, (nV-sentV)*4, cudaMemcpyDeviceToHost, ctx->dlStream));
				CK(cudaMemcpyAsync(base+L.oW+sentV*4, f->weights+sentV, (nV-sentV)*4, cudaMemcpyDeviceToHost, ctx->dlStream));
				sentP = nP; sentV = nV;
			}
		}
		if (done < plan.size()) f->streamed = false;
		f->streamedPoints = sentP; f->streamedViews = sentV;
	}
	FuseCtl ctl;
	CK(cudaMemcpyAsync(&ctl, f->ctl_d, sizeof(ctl), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	if (ctl.overflow) { hcmvs_set_error("fused cloud exceeded its bound (%zu points / %zu view references) [code %d]", f->capPoints, f->capViews, ctl.overflow); return HCMVS_ERR_STATE; }
	const size_t nPoints = (size_t)ctl.cumPoints, nViewRefs = (size_t)ctl.cumRefs;
	if (debug && trace_d) {
		std::vector<unsigned long long> tr(plan.size()*8);
		cudaMemcpy(tr.data(), trace_d, tr.size()*8, cudaMemcpyDeviceToHost); cudaFree(trace_d);
		double acc[5] = {0, 0, 0, 0, 0};
		for (size_t r=0; r<plan.size(); ++r) for (int i=0; i<5; ++i) acc[i] += (double)(tr[8*r+i+1]-tr[8*r+i])*1e-3;
		for (size_t r=0; r<plan.size(); r += (r < 8 ? 1 : 8))
			fprintf(stderr, "[fuse] view #%zu (%d): count %.1f scatter+emit %.1f probe %.1f resolve-1 %.1f rounds %.1f us; %llu seeds, %llu rounds, worklist entries summed over the rounds %llu\n", r, plan[r].view,
				(tr[8*r+1]-tr[8*r])*1e-3, (tr[8*r+2]-tr[8*r+1])*1e-3, (tr[8*r+3]-tr[8*r+2])*1e-3, (tr[8*r+4]-tr[8*r+3])*1e-3, (tr[8*r+5]-tr[8*r+4])*1e-3, tr[8*r+6]&0xFFFFFFFFull, tr[8*r+6]>>32, tr[8*r+7]&0xFFFFFFFFull);
		for (size_t r=0; r<plan.size(); r += (r < 8 ? 1 : 8)) fprintf(stderr, "[fuse] view #%zu: thread 0 spent %.1f us in the barriers of the later rounds\n", r, (double)(tr[8*r+7]>>32)*1e-3);
		fprintf(stderr, "[fuse] us per scene: count %.0f, scatter+emit %.0f, probe %.0f, resolve-1 %.0f, rounds %.0f; total %.0f\n", acc[0], acc[1], acc[2], acc[3], acc[4], (double)(tr[8*(plan.size()-1)+5]-tr[0])*1e-3);
	}
	if (debug) fprintf(stderr, "[fuse] %zu views, %llu seeds, %llu rounds, %zu points, %zu view refs, %d blocks\n", plan.size(), ctl.seeds, ctl.rounds, nPoints, nViewRefs, f->coopBlocks);
	for (View& v: ctx->views) if (v.set && v.hasMaps) { int r = hcmvs_mark_image_use(ctx, v); if (r) return r; } // colours were read
	ctx->fuseRounds = ctl.rounds; ctx->fuseSeeds = ctl.seeds; ctx->fuseProbes = ctl.probes;
	f->nPoints = nPoints; f->nViewRefs = nViewRefs; f->hasColor = estimate_color != 0; f->hasNormal = estimate_normal != 0;
	if (f->streamed && (f->streamedPoints != nPoints || f->streamedViews != nViewRefs)) f->streamed = false;
	if (nPoints) { // close the CSR offsets on the device
		const uint32_t last = (uint32_t)nViewRefs;
		CK(cudaMemcpyAsync(f->viewOffsets+nPoints, &last, 4, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
	}
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
	if (ctx) ctx->freshScene = false; // a consumer of other views' maps ends the "own maps only" phase of hcmvs_begin_scene
	if (!ctx || view >= ctx->views.size() || !ctx->views[view].set) { hcmvs_set_error("view %u not set", view); return HCMVS_ERR_ARG; }
	View& v = ctx->views[view];
	FuseState* f = ctx->fuse;
	if (!v.hasMaps || !f || view >= f->recOff.size() || f->recOff[view] == (size_t)-1 || !f->rec_d) { hcmvs_set_error("view %u took no part in a fusion yet (call hcmvs_fuse_depthmaps)", view); return HCMVS_ERR_STATE; }
	cudaSetDevice(ctx->device);
	const size_t n = (size_t)v.w*v.h;
	float* tmp; int r = hcmvs_scratch(ctx, n*16, (void**)&tmp); if (r) return r;
	k_fused_support<<<(unsigned)((n+255)/256), 256, 0, ctx->stream>>>(v.dn_d, f->alive_d+f->aliveOff[view], tmp, tmp+n, n); ++ctx->nLaunches;
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
