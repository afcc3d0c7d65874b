// Shared device-side types for the hcmvs_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "hcmvs_b200.h"

#define HCMVS_HW 7                      // DepthEstimator::nSizeHalfWindow (libs/MVS/DepthMap.h:354)
#define HCMVS_MAXV HCMVS_MAX_MATCH_VIEWS
#define HCMVS_NT 128                    // threads per CTA of the scoring kernels
#ifndef HCMVS_MINB
#define HCMVS_MINB 3
#endif
#ifndef HCMVS_RB6
#define HCMVS_RB6 2                     // rows of a 6x6 patch whose texture gathers are in flight together (k_sweep batches)
#endif
#ifndef HCMVS_RB6_PLAIN
#define HCMVS_RB6_PLAIN 3               // ... in the plain sweep (it_external 0, no extra hypotheses): 18 gathers per thread, 168 registers, no spills
#endif
// shared-memory windows of the neighbour images (sampler 2): per CTA and matching view a HCMVS_WIN x HCMVS_WIN texel window
// around the footprint of the tile's current estimates; hypotheses whose whole patch falls inside are sampled with LDS
#define HCMVS_WIN 40
#define HCMVS_WINP 44                   // row pitch in floats (12 mod 32: the 32 pixels of a warp's 8x8 checkerboard block hit 32 distinct banks under an identity-like mapping)
#define HCMVS_WINV 5                    // views that get a window (the reference's runs match against 5)
#ifndef HCMVS_FLOOR_FADD
#define HCMVS_FLOOR_FADD 1
#endif
#ifndef HCMVS_HULL_TEST
#define HCMVS_HULL_TEST 1               // patches whose corners lie inside the neighbour image with slack walk without per-texel border tests
#endif
#ifndef HCMVS_EARLY_REJECT
#define HCMVS_EARLY_REJECT 0
#endif

// per matching-neighbour constants (DepthEstimator::ViewData, DepthMap.h:412-444)
struct NbViewConst {
	double Hl[9];                // K1 R1 R0^T
	double Hm[3];                // K1 R1 (C0-C1)
	cudaTextureObject_t tex;     // point-sampled f32 gather texture of the neighbour gray image
	const float* img;            // the same image, linear (global-load sampler)
	int pitch;                   // elements per row of img
	int w, h;
};

// everything one reference view's scoring kernels need; passed by value as a __grid_constant__
struct RefConst {
	int w, h;
	int y0, y1;                  // rows [y0, y1) this launch estimates (the whole image, or one rank's band + halo of a row-split view)
	double fx, fy, cx, cy;       // K0
	double Hr[9];                // K0^-1
	const float* img0; int pitch0;
	const uint8_t* gra;          // gradient map (u8) or nullptr
	const float* prior;          // depthMapPrior or nullptr
	float4* dn;                  // (nx,ny,nz,depth) per pixel
	float* conf;
	int nViews;
	NbViewConst nb[HCMVS_MAXV];
	float dMin, dMax, dMinSqr, dMaxSqr;
	float keep, thRobust, thConfSmall, thConfBig, thConfRand;
	float smoothBonusDepth, smoothBonusNormal, smoothSigmaDepth, smoothSigmaNormal;
	float angle1Range, angle2Range, depthRatio;
	int nRandomIters, adapthalfwin, farReach, propDirs, it_external, photo2geo, propagatehalfwin, propagatestep;
	float photometric_flow, para_prior, sigmaPrior;
	uint32_t key0, key1, pass;
	unsigned long long* counters; // [0] hypotheses, [1] view scores, [2] pixel-iterations
	// extra hypotheses after the refinement (k_sweep<.., XTRA = true> only)
	const float4* coarse;         // restore tree: the previous level's estimate resized to this view (nresize maps), or nullptr
	int lastPass;                 // last PatchMatch iteration of the last outer iteration (restore/.../DepthMap.cpp:1527)
	int viewspread;               // cross-view propagation (DepthMap.cpp:1504-1608)
	const struct SpreadConst* spread; // device memory; valid when viewspread && it_external >= 1
};

struct CamConst { // f64 camera for the filter / fuse kernels (libs/MVS/Camera.h)
	double K[9], R[9], C[3], P[12];
};

// what viewspread reads from the matching neighbours: their camera and their maps of the previous outer iteration
struct SpreadConst {
	CamConst camRef;
	struct Nb { CamConst cam; const float4* dn; const float* conf; int w, h; } nb[HCMVS_MAXV];
};
