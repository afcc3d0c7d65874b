// MVS::EstimatePointNormals (libs/MVS/DepthMap.cpp:2221-2269; `--estimate-normals 1`, SceneDensify.cpp:3570-3571): the normal of every
// cloud point is the direction of least variance of its K nearest neighbours plus itself (CGAL::pca_estimate_normals — third party,
// absent from /root/reference: K-nearest-neighbour search + linear_least_squares_fitting_3 = the eigenvector of the smallest
// eigenvalue of the neighbourhood's covariance), then flipped to face the FIRST view that sees the point (:2262-2265).
//
// Here: points are bucketed in a uniform 3-D grid (cub radix sort by cell key, run-length encode -> sorted unique cells), one thread
// per point grows a cube of cells shell by shell until its K+1 best candidates are provably the nearest (the K+1-th distance is
// within the searched cube), and the 3x3 covariance is diagonalised with cyclic Jacobi rotations in f64. HBM-bound on the sorted
// positions; the result is a function of the neighbour SET, so the order in which candidates are visited does not matter (ties at the
// K+1-th distance are the one ambiguity, as in CGAL).
#include "hcmvs_internal.h"
#include <cub/cub.cuh>
#include <vector>
#include <cmath>
#include <cstring>
#include <cfloat>

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { hcmvs_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return HCMVS_ERR_CUDA; } } while (0)
#define PN_MAXK 32 // neighbours + the point itself
#define PN_MAXR 24 // per grid level the search cube stops growing at (2*24+1)^3 cells; unfinished points go on at the next, 8x coarser level

namespace hcmvs {

struct PnGrid { float minx, miny, minz, h, invh; int nx, ny, nz; };

__device__ __forceinline__ unsigned long long pn_key(int cx, int cy, int cz) { return (unsigned long long)cx | ((unsigned long long)cy<<21) | ((unsigned long long)cz<<42); }
__device__ __forceinline__ int pn_cell(float v, float lo, float invh, int n) { return min(max((int)floorf((v-lo)*invh), 0), n-1); }

__global__ void k_pn_keys(const PnGrid G, const float* __restrict__ pts, uint32_t n, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
	const uint32_t i = blockIdx.x*blockDim.x+threadIdx.x;
	if (i >= n) return;
	keys[i] = pn_key(pn_cell(pts[i*3], G.minx, G.invh, G.nx), pn_cell(pts[i*3+1], G.miny, G.invh, G.ny), pn_cell(pts[i*3+2], G.minz, G.invh, G.nz));
	vals[i] = i;
}
__global__ void k_pn_sorted_pos(const uint32_t* __restrict__ vals, uint32_t n, const float* __restrict__ pts, float4* __restrict__ spos) {
	const uint32_t j = blockIdx.x*blockDim.x+threadIdx.x;
	if (j >= n) return;
	const uint32_t i = vals[j];
	spos[j] = make_float4(pts[i*3], pts[i*3+1], pts[i*3+2], __int_as_float((int)i));
}

// smallest-eigenvalue eigenvector of a symmetric 3x3 matrix: cyclic Jacobi, f64
__device__ __forceinline__ void pn_smallest_eigvec(double a[3][3], double out[3]) {
	double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
	for (int sweep=0; sweep<12; ++sweep) {
		const double off = fabs(a[0][1])+fabs(a[0][2])+fabs(a[1][2]);
		if (off < 1e-300 || off <= 1e-18*(fabs(a[0][0])+fabs(a[1][1])+fabs(a[2][2]))) break;
		#pragma unroll
		for (int p=0; p<2; ++p)
			#pragma unroll
			for (int q=p+1; q<3; ++q) {
				if (a[p][q] == 0.0) continue;
				const double theta = (a[q][q]-a[p][p])/(2.0*a[p][q]);
				const double t = (theta >= 0 ? 1.0 : -1.0)/(fabs(theta)+sqrt(theta*theta+1.0));
				const double c = 1.0/sqrt(t*t+1.0), s = t*c;
				#pragma unroll
				for (int k=0; k<3; ++k) { const double akp = a[k][p], akq = a[k][q]; a[k][p] = c*akp-s*akq; a[k][q] = s*akp+c*akq; }
				#pragma unroll
				for (int k=0; k<3; ++k) { const double apk = a[p][k], aqk = a[q][k]; a[p][k] = c*apk-s*aqk; a[q][k] = s*apk+c*aqk; }
				#pragma unroll
				for (int k=0; k<3; ++k) { const double vkp = v[k][p], vkq = v[k][q]; v[k][p] = c*vkp-s*vkq; v[k][q] = s*vkp+c*vkq; }
			}
	}
	int m = 0;
	if (a[1][1] < a[m][m]) m = 1;
	if (a[2][2] < a[m][m]) m = 2;
	out[0] = v[0][m]; out[1] = v[1][m]; out[2] = v[2][m];
}

struct PnCam { float cx, cy, cz; };

// one thread per point, in sorted (cell) order
// query == nullptr: every point, in sorted (cell) order; else the listed points (a coarser level of the grid hierarchy re-visits the points
// whose fine-level search hit the PN_MAXR cap). pending / nPending collect the points whose K1 nearest are still not proven.
__global__ void __launch_bounds__(128) k_pn_normals(const PnGrid G, const float4* __restrict__ spos, uint32_t n, const unsigned long long* __restrict__ cellKeys,
	const uint32_t* __restrict__ cellStart, uint32_t nCells, int K1, const uint32_t* __restrict__ offs, const uint32_t* __restrict__ views,
	const PnCam* __restrict__ cams, uint32_t nCams, float* __restrict__ normals,
	const uint32_t* __restrict__ query, uint32_t nQuery, const float* __restrict__ pts, uint32_t* __restrict__ pending, uint32_t* __restrict__ nPending)
{
	const uint32_t j = blockIdx.x*blockDim.x+threadIdx.x;
	if (j >= (query ? nQuery : n)) return;
	float4 P;
	if (query) { const uint32_t i = query[j]; P = make_float4(pts[(size_t)i*3], pts[(size_t)i*3+1], pts[(size_t)i*3+2], __int_as_float((int)i)); }
	else P = spos[j];
	const uint32_t self = (uint32_t)__float_as_int(P.w);
	const int cx = pn_cell(P.x, G.minx, G.invh, G.nx), cy = pn_cell(P.y, G.miny, G.invh, G.ny), cz = pn_cell(P.z, G.minz, G.invh, G.nz);
	float bd[PN_MAXK]; uint32_t bj[PN_MAXK]; // the K1 best so far, ascending by distance (sorted positions' indices)
	int nb = 0;
	// An isolated outlier would otherwise grow its cube through thousands of empty shells (O(R^3) cell look-ups: minutes on a 20 M-point
	// cloud). Beyond PN_MAXR cells (~25 mean point spacings) the search stops at THIS level: the point is handed to a grid with 8x
	// larger cells (hcmvs_estimate_point_normals loops over the levels), so the k-NN stays exact — CGAL's unbounded search — at a
	// bounded cost per level. (What was found so far is fitted anyway: the value of a point no level could finish.)
	const int maxR = min(max(G.nx, max(G.ny, G.nz)), PN_MAXR);
	bool proven = false;
	for (int R=0; R<=maxR; ++R) {
		for (int dz=-R; dz<=R; ++dz) {
			const int z = cz+dz; if (z < 0 || z >= G.nz) continue;
			for (int dy=-R; dy<=R; ++dy) {
				const int y = cy+dy; if (y < 0 || y >= G.ny) continue;
				const bool face = abs(dz) == R || abs(dy) == R;
				for (int dx=-R; dx<=R; dx += (face || R == 0) ? 1 : 2*R) { // only the shell of the cube: interior cells were visited at smaller R
					const int x = cx+dx; if (x < 0 || x >= G.nx) continue;
					const unsigned long long key = pn_key(x, y, z);
					uint32_t lo = 0, hi = nCells;
					while (lo < hi) { const uint32_t mid = (lo+hi)>>1; if (cellKeys[mid] < key) lo = mid+1; else hi = mid; }
					if (lo >= nCells || cellKeys[lo] != key) continue;
					for (uint32_t q=cellStart[lo]; q<cellStart[lo+1]; ++q) {
						const float4 Q = spos[q];
						const float ex = Q.x-P.x, ey = Q.y-P.y, ez = Q.z-P.z;
						const float d = ex*ex+ey*ey+ez*ez;
						if (nb == K1 && !(d < bd[K1-1])) continue;
						int pos = nb < K1 ? nb : K1-1;
						while (pos > 0 && bd[pos-1] > d) { bd[pos] = bd[pos-1]; bj[pos] = bj[pos-1]; --pos; }
						bd[pos] = d; bj[pos] = q;
						if (nb < K1) ++nb;
					}
				}
			}
		}
		if (nb == K1) {
			// everything closer than the nearest face of the searched cube has been seen
			const float lx = G.minx+(float)(cx-R)*G.h, hx = G.minx+(float)(cx+R+1)*G.h, ly = G.miny+(float)(cy-R)*G.h, hy = G.miny+(float)(cy+R+1)*G.h;
			const float lz = G.minz+(float)(cz-R)*G.h, hz = G.minz+(float)(cz+R+1)*G.h;
			float reach = FLT_MAX;
			if (cx-R > 0) reach = fminf(reach, P.x-lx); if (cx+R < G.nx-1) reach = fminf(reach, hx-P.x);
			if (cy-R > 0) reach = fminf(reach, P.y-ly); if (cy+R < G.ny-1) reach = fminf(reach, hy-P.y);
			if (cz-R > 0) reach = fminf(reach, P.z-lz); if (cz+R < G.nz-1) reach = fminf(reach, hz-P.z);
			reach = fmaxf(reach, 0.f)*0.9999f; // slack for the rounding of the cell assignment
			if (bd[K1-1] <= reach*reach) { proven = true; break; }
		}
	}
	// the cap stopped the search before the K1 nearest were certain (and the cube did not cover the whole grid): a coarser level goes on
	if (!proven && maxR < max(G.nx, max(G.ny, G.nz)) && pending) pending[atomicAdd(nPending, 1u)] = self;
	float nx = 0.f, ny = 0.f, nz = 0.f;
	if (nb >= 3) {
		// linear_least_squares_fitting_3 over the points: centroid, covariance, least-variance direction (f64)
		double mx = 0, my = 0, mz = 0;
		for (int k=0; k<nb; ++k) { const float4 Q = spos[bj[k]]; mx += Q.x; my += Q.y; mz += Q.z; }
		mx /= nb; my /= nb; mz /= nb;
		double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
		for (int k=0; k<nb; ++k) {
			const float4 Q = spos[bj[k]];
			const double ex = Q.x-mx, ey = Q.y-my, ez = Q.z-mz;
			C[0][0] += ex*ex; C[0][1] += ex*ey; C[0][2] += ex*ez; C[1][1] += ey*ey; C[1][2] += ey*ez; C[2][2] += ez*ez;
		}
		C[1][0] = C[0][1]; C[2][0] = C[0][2]; C[2][1] = C[1][2];
		double e[3]; pn_smallest_eigvec(C, e);
		const double nrm = sqrt(e[0]*e[0]+e[1]*e[1]+e[2]*e[2]);
		if (nrm > 0) { nx = (float)(e[0]/nrm); ny = (float)(e[1]/nrm); nz = (float)(e[2]/nrm); }
		// correct normal orientation: towards the first view of the point, DepthMap.cpp:2262-2265
		if (offs[self+1] > offs[self]) {
			const uint32_t v = views[offs[self]];
			if (v < nCams) { const PnCam c = cams[v]; if (nx*(c.cx-P.x)+ny*(c.cy-P.y)+nz*(c.cz-P.z) < 0.f) { nx = -nx; ny = -ny; nz = -nz; } }
		}
	}
	normals[(size_t)self*3] = nx; normals[(size_t)self*3+1] = ny; normals[(size_t)self*3+2] = nz;
}

__global__ void k_pn_minmax(const float* __restrict__ pts, uint32_t n, float* __restrict__ mm) { // mm[0..2] = min, mm[3..5] = max (ordered-int atomics)
	float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
	for (uint32_t i=blockIdx.x*blockDim.x+threadIdx.x; i<n; i+=gridDim.x*blockDim.x)
		for (int d=0; d<3; ++d) { const float v = pts[i*3+d]; lo[d] = fminf(lo[d], v); hi[d] = fmaxf(hi[d], v); }
	for (int d=0; d<3; ++d) {
		for (int s=16; s>0; s>>=1) { lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], s)); hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], s)); }
		if ((threadIdx.x&31) == 0) {
			// monotone float -> int mapping so that integer atomics order like the floats
			auto enc = [](float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; };
			atomicMin((int*)&mm[d], enc(lo[d])); atomicMax((int*)&mm[3+d], enc(hi[d]));
		}
	}
}

} // namespace hcmvs
using namespace hcmvs;

extern "C" int hcmvs_estimate_point_normals(hcmvs_ctx* ctx, uint64_t n_points, const float* points, const uint32_t* view_offsets, const uint32_t* views, int num_neighbors, float* normals) {
	if (!ctx) { hcmvs_set_error("null context"); return HCMVS_ERR_ARG; }
	if (num_neighbors < 2 || num_neighbors+1 > PN_MAXK) { hcmvs_set_error("numNeighbors must be in [2, %d]", PN_MAXK-1); return HCMVS_ERR_ARG; }
	cudaSetDevice(ctx->device);
	cudaStream_t st = ctx->stream;
	const bool resident = points == nullptr;
	uint64_t n = n_points, m = 0;
	const float* pts_d = nullptr; const uint32_t* offs_d = nullptr; const uint32_t* ids_d = nullptr; float* nrm_d = nullptr;
	float *pts_own = nullptr, *nrm_own = nullptr; uint32_t *offs_own = nullptr, *ids_own = nullptr;
	if (resident) {
		void *p = nullptr, *o = nullptr, *w = nullptr, *nn = nullptr;
		if (hcmvs_get_fused_device(ctx, &n, &m, &p, &nn, nullptr, &o, &w, nullptr) != HCMVS_OK || !n) { hcmvs_set_error("no fused cloud on the device (call hcmvs_fuse_depthmaps) and no points given"); return HCMVS_ERR_STATE; }
		if (!nn) { hcmvs_set_error("the resident cloud was fused without normals: fuse with estimate_normal = 1 (the buffer is reused) or pass host points"); return HCMVS_ERR_STATE; }
		pts_d = (const float*)p; offs_d = (const uint32_t*)o; ids_d = (const uint32_t*)w; nrm_d = (float*)nn;
		hcmvs_fuse_invalidate_stream(ctx); // the normals streamed to the page-locked arena during the fusion are about to be replaced
	} else {
		if (!view_offsets || !views || !normals) { hcmvs_set_error("null argument"); return HCMVS_ERR_ARG; }
		if (!n) return HCMVS_OK;
		m = view_offsets[n];
		CK(cudaMalloc(&pts_own, n*12)); CK(cudaMalloc(&nrm_own, n*12)); CK(cudaMalloc(&offs_own, (n+1)*4)); CK(cudaMalloc(&ids_own, std::max<uint64_t>(m, 1)*4));
		CK(cudaMemcpyAsync(pts_own, points, n*12, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(offs_own, view_offsets, (n+1)*4, cudaMemcpyHostToDevice, st));
		if (m) CK(cudaMemcpyAsync(ids_own, views, m*4, cudaMemcpyHostToDevice, st));
		pts_d = pts_own; offs_d = offs_own; ids_d = ids_own; nrm_d = nrm_own;
	}
	if (n >= (1ull<<31)) { hcmvs_set_error("more than 2^31 points"); return HCMVS_ERR_UNSUPPORTED; }
	const uint32_t N = (uint32_t)n;
	const int K1 = std::min<int>(num_neighbors+1, (int)N);
	// bounding box
	float* mm_d = nullptr; CK(cudaMalloc(&mm_d, 6*4));
	{ const int init[6] = {0x7F7FFFFF, 0x7F7FFFFF, 0x7F7FFFFF, (int)0x80800000, (int)0x80800000, (int)0x80800000}; // enc(+FLT_MAX) x3, enc(-FLT_MAX) x3
	  CK(cudaMemcpyAsync(mm_d, init, sizeof(init), cudaMemcpyHostToDevice, st)); }
	k_pn_minmax<<<592, 256, 0, st>>>(pts_d, N, mm_d); ++ctx->nLaunches;
	int mmi[6]; CK(cudaMemcpyAsync(mmi, mm_d, sizeof(mmi), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
	float mm[6]; for (int d=0; d<6; ++d) { int i = mmi[d]; if (i < 0) i ^= 0x7FFFFFFF; std::memcpy(&mm[d], &i, 4); }
	cudaFree(mm_d);
	// scratch
	unsigned long long *keys_d = nullptr, *keys2_d = nullptr, *cellKeys_d = nullptr; uint32_t *vals_d = nullptr, *vals2_d = nullptr, *cellCount_d = nullptr, *cellStart_d = nullptr, *nRuns_d = nullptr;
	float4* spos_d = nullptr; void* tmp_d = nullptr; size_t tmpBytes = 0; PnCam* cams_d = nullptr;
	CK(cudaMalloc(&keys_d, (size_t)N*8)); CK(cudaMalloc(&keys2_d, (size_t)N*8)); CK(cudaMalloc(&vals_d, (size_t)N*4)); CK(cudaMalloc(&vals2_d, (size_t)N*4));
	CK(cudaMalloc(&cellKeys_d, (size_t)N*8)); CK(cudaMalloc(&cellCount_d, (size_t)(N+1)*4)); CK(cudaMalloc(&cellStart_d, (size_t)(N+1)*4)); CK(cudaMalloc(&nRuns_d, 4));
	CK(cudaMalloc(&spos_d, (size_t)N*16));
	{ size_t a = 0, b = 0, c = 0;
	  cub::DeviceRadixSort::SortPairs(nullptr, a, keys_d, keys2_d, vals_d, vals2_d, (int)N, 0, 63, st);
	  cub::DeviceRunLengthEncode::Encode(nullptr, b, keys2_d, cellKeys_d, cellCount_d, nRuns_d, (int)N, st);
	  cub::DeviceScan::ExclusiveSum(nullptr, c, cellCount_d, cellStart_d, (int)N+1, st);
	  tmpBytes = std::max(a, std::max(b, c)); CK(cudaMalloc(&tmp_d, tmpBytes)); }
	// cell size: aim at ~4 points per occupied cell (the cloud is a surface: occupied cells ~ area / h^2); start from the volume
	// estimate and refine with the measured number of occupied cells
	const double ext[3] = {std::max((double)mm[3]-mm[0], 1e-12), std::max((double)mm[4]-mm[1], 1e-12), std::max((double)mm[5]-mm[2], 1e-12)};
	double h = std::cbrt(ext[0]*ext[1]*ext[2]/std::max<double>(N, 1))*2.0;
	h = std::max(h, std::max(ext[0], std::max(ext[1], ext[2]))/2000000.0); // 21-bit cell coordinates
	PnGrid G; uint32_t nCells = 0;
	int rc = HCMVS_OK;
	auto buildGrid = [&](double hh) -> bool { // cell keys -> sort -> occupied cells (keys + counts); nCells on the host
		G.minx = mm[0]; G.miny = mm[1]; G.minz = mm[2]; G.h = (float)hh; G.invh = (float)(1.0/hh);
		G.nx = (int)std::min(2097151.0, std::floor(ext[0]/hh)+1); G.ny = (int)std::min(2097151.0, std::floor(ext[1]/hh)+1); G.nz = (int)std::min(2097151.0, std::floor(ext[2]/hh)+1);
		k_pn_keys<<<(N+255)/256, 256, 0, st>>>(G, pts_d, N, keys_d, vals_d);
		if (cub::DeviceRadixSort::SortPairs(tmp_d, tmpBytes, keys_d, keys2_d, vals_d, vals2_d, (int)N, 0, 63, st) != cudaSuccess ||
		    cub::DeviceRunLengthEncode::Encode(tmp_d, tmpBytes, keys2_d, cellKeys_d, cellCount_d, nRuns_d, (int)N, st) != cudaSuccess) { hcmvs_set_error("cub failed"); rc = HCMVS_ERR_CUDA; return false; }
		ctx->nLaunches += 3;
		if (cudaMemcpyAsync(&nCells, nRuns_d, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { hcmvs_set_error("point normals: %s", cudaGetErrorString(cudaGetLastError())); rc = HCMVS_ERR_CUDA; return false; }
		return true;
	};
	for (int it=0; it<6; ++it) {
		if (!buildGrid(h)) break;
		const double occ = (double)N/std::max<uint32_t>(nCells, 1);
		if ((occ >= 2.0 && occ <= 8.0) || it == 5 || N <= 64) break;
		h *= std::sqrt(4.0/occ); // occupancy of a surface grows with h^2
	}
	uint32_t *pendA_d = nullptr, *pendB_d = nullptr, *nPend_d = nullptr;
	if (rc == HCMVS_OK) {
		std::vector<PnCam> cams(ctx->views.size());
		for (size_t v=0; v<cams.size(); ++v) { cams[v].cx = (float)ctx->views[v].C[0]; cams[v].cy = (float)ctx->views[v].C[1]; cams[v].cz = (float)ctx->views[v].C[2]; }
		CK(cudaMalloc(&cams_d, std::max<size_t>(cams.size(), 1)*sizeof(PnCam)));
		if (!cams.empty()) CK(cudaMemcpyAsync(cams_d, cams.data(), cams.size()*sizeof(PnCam), cudaMemcpyHostToDevice, st));
		CK(cudaMalloc(&pendA_d, (size_t)N*4)); CK(cudaMalloc(&pendB_d, (size_t)N*4)); CK(cudaMalloc(&nPend_d, 4));
		// level 0: every point on the fine grid; level L: the points still pending, on a grid with 8^L times larger cells
		uint32_t nQuery = 0; const uint32_t* query_d = nullptr;
		for (int level=0; level<12 && rc == HCMVS_OK; ++level) {
			if (level > 0) { h *= 8.0; if (!buildGrid(h)) break; }
			CK(cudaMemsetAsync(cellCount_d+nCells, 0, 4, st));
			cub::DeviceScan::ExclusiveSum(tmp_d, tmpBytes, cellCount_d, cellStart_d, (int)nCells+1, st);
			k_pn_sorted_pos<<<(N+255)/256, 256, 0, st>>>(vals2_d, N, pts_d, spos_d);
			CK(cudaMemsetAsync(nPend_d, 0, 4, st));
			uint32_t* out_d = (level&1) ? pendB_d : pendA_d;
			const uint32_t work = level == 0 ? N : nQuery;
			k_pn_normals<<<(work+127)/128, 128, 0, st>>>(G, spos_d, N, cellKeys_d, cellStart_d, nCells, K1, offs_d, ids_d, cams_d, (uint32_t)cams.size(), nrm_d,
				query_d, nQuery, pts_d, out_d, nPend_d);
			ctx->nLaunches += 3;
			if (cudaGetLastError() != cudaSuccess) { hcmvs_set_error("point normals launch failed"); rc = HCMVS_ERR_CUDA; break; }
			uint32_t nPend = 0;
			if (cudaMemcpyAsync(&nPend, nPend_d, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { hcmvs_set_error("point normals: %s", cudaGetErrorString(cudaGetLastError())); rc = HCMVS_ERR_CUDA; break; }
			if (!nPend) break;
			query_d = out_d; nQuery = nPend;
		}
		if (rc == HCMVS_OK && normals && cudaMemcpyAsync(normals, nrm_d, (size_t)N*12, cudaMemcpyDeviceToHost, st) != cudaSuccess) { hcmvs_set_error("cannot read the normals back"); rc = HCMVS_ERR_CUDA; }
		if (cudaStreamSynchronize(st) != cudaSuccess && rc == HCMVS_OK) { hcmvs_set_error("point normals: %s", cudaGetErrorString(cudaGetLastError())); rc = HCMVS_ERR_CUDA; }
	}
	cudaFree(pendA_d); cudaFree(pendB_d); cudaFree(nPend_d);
	cudaFree(keys_d); cudaFree(keys2_d); cudaFree(vals_d); cudaFree(vals2_d); cudaFree(cellKeys_d); cudaFree(cellCount_d); cudaFree(cellStart_d); cudaFree(nRuns_d);
	cudaFree(spos_d); cudaFree(tmp_d); cudaFree(cams_d); cudaFree(pts_own); cudaFree(nrm_own); cudaFree(offs_own); cudaFree(ids_own);
	return rc;
}
