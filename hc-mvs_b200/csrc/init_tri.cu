// Depth-map initialisation from the triangulated sparse points — the per-pixel half of MVS::TriangulatePoints2DepthMap
// (libs/MVS/DepthMap.cpp:1879-1936): every triangle of the host-side Delaunay mesh is rasterised with the reference's 28.4 fixed-point
// half-space rasteriser (TImage::RasterizeTriangle, libs/Common/Types.inl:2469-2606) and every covered pixel gets the depth of the
// viewing ray's intersection with the triangle's plane and the plane's normal (RasterDepthDataPlaneData, DepthMap.cpp:1899-1913).
//
// The rasteriser's top-left fill rule gives every pixel centre to exactly one triangle of a mesh — except around slivers that flip
// when their vertices snap to the 1/16-pixel grid, where the reference's answer is "the last face drawn wins". So that the parallel
// version is the sequential loop in face order whatever happens, pass 0 records per pixel the HIGHEST face index that covers it with a
// positive depth (atomicMax) and pass 1 lets exactly that face write. One CTA per (triangle, 32x8-pixel chunk of its bounding box).
// HBM-bound in principle (16 B written per pixel, 72 B read per triangle); in practice launch-latency sized: < 0.1 ms per view.
#include "hcmvs_internal.h"
#include "camera.cuh"

namespace hcmvs {

struct TriSetup {
	long long C1, C2, C3, DX12, DX23, DX31, DY12, DY23, DY31;
	int minx, maxx, miny, maxy; // bounding rectangle in pixels (max exclusive)
	float3 normal, normalPlane;
};

__device__ __forceinline__ long long round16(float v) { return (long long)floorf(__fadd_rn(__fmul_rn(16.f, v), 0.5f)); } // ROUND2INT(T(16)*v)

__device__ __forceinline__ void tri_setup(const double* __restrict__ vtx, const uint32_t* __restrict__ tri, const double fx, const double fy, const double cx, const double cy, TriSetup& s) {
	// Point3f i0..i2(face.vertex(k)->point()), DepthMap.cpp:1916-1918
	float3 I[3], c[3];
	#pragma unroll
	for (int k=0; k<3; ++k) {
		const double* p = vtx+(size_t)tri[k]*3;
		I[k] = make_float3((float)p[0], (float)p[1], (float)p[2]);
		// camera.TransformPointI2C(Point3f), Camera.h:307-312: f64 arithmetic on the widened f32 coordinates, f32 result
		c[k] = make_float3((float)(dmul(dadd((double)I[k].x, -cx), (double)I[k].z)/fx), (float)(dmul(dadd((double)I[k].y, -cy), (double)I[k].z)/fy), I[k].z);
	}
	const float3 e1 = make_float3(__fsub_rn(c[1].x, c[0].x), __fsub_rn(c[1].y, c[0].y), __fsub_rn(c[1].z, c[0].z));
	const float3 e2 = make_float3(__fsub_rn(c[2].x, c[0].x), __fsub_rn(c[2].y, c[0].y), __fsub_rn(c[2].z, c[0].z));
	// edge2.cross(edge1) (cv::Point3_::cross, un-fused f32), normalized() = cv::normalize(Vec3f): f64 norm, v * (1/norm) rounded to f32
	const float3 n = make_float3(__fsub_rn(__fmul_rn(e2.y, e1.z), __fmul_rn(e2.z, e1.y)), __fsub_rn(__fmul_rn(e2.z, e1.x), __fmul_rn(e2.x, e1.z)), __fsub_rn(__fmul_rn(e2.x, e1.y), __fmul_rn(e2.y, e1.x)));
	const double nv = sqrt(dadd(dadd(dmul((double)n.x, (double)n.x), dmul((double)n.y, (double)n.y)), dmul((double)n.z, (double)n.z)));
	const double inv = nv != 0.0 ? 1.0/nv : 0.0;
	s.normal = make_float3((float)dmul((double)n.x, inv), (float)dmul((double)n.y, inv), (float)dmul((double)n.z, inv));
	// normalPlane = normal * INVERT(normal.dot(c0)), DepthMap.cpp:1926
	const float d0 = __fadd_rn(__fadd_rn(__fmul_rn(s.normal.x, c[0].x), __fmul_rn(s.normal.y, c[0].y)), __fmul_rn(s.normal.z, c[0].z));
	const float id0 = d0 == 0.f ? 1000000.f : __fdiv_rn(1.f, d0);
	s.normalPlane = make_float3(__fmul_rn(s.normal.x, id0), __fmul_rn(s.normal.y, id0), __fmul_rn(s.normal.z, id0));
	// RasterizeTriangle(i2, i1, i0): v1 = i2, v2 = i1, v3 = i0
	const long long X1 = round16(I[2].x), Y1 = round16(I[2].y), X2 = round16(I[1].x), Y2 = round16(I[1].y), X3 = round16(I[0].x), Y3 = round16(I[0].y);
	s.DX12 = X1-X2; s.DX23 = X2-X3; s.DX31 = X3-X1;
	s.DY12 = Y1-Y2; s.DY23 = Y2-Y3; s.DY31 = Y3-Y1;
	s.minx = (int)((min(X1, min(X2, X3))+0xF)>>4); s.maxx = (int)((max(X1, max(X2, X3))+0xF)>>4);
	s.miny = (int)((min(Y1, min(Y2, Y3))+0xF)>>4); s.maxy = (int)((max(Y1, max(Y2, Y3))+0xF)>>4);
	s.C1 = s.DY12*X1-s.DX12*Y1; s.C2 = s.DY23*X2-s.DX23*Y2; s.C3 = s.DY31*X3-s.DX31*Y3;
	if (s.DY12 < 0 || (s.DY12 == 0 && s.DX12 > 0)) ++s.C1; // fill convention
	if (s.DY23 < 0 || (s.DY23 == 0 && s.DX23 > 0)) ++s.C2;
	if (s.DY31 < 0 || (s.DY31 == 0 && s.DX31 > 0)) ++s.C3;
}

// chunks[i] = (triangle, x0, y0): a 32x8 pixel tile of the triangle's bounding box; blockDim = (32, 8)
template<int PASS>
__global__ void __launch_bounds__(256) k_raster_triangles(const double* __restrict__ vtx, const uint32_t* __restrict__ tris, const int3* __restrict__ chunks,
	const double fx, const double fy, const double cx, const double cy, float4* __restrict__ dn, int* __restrict__ owner, int w, int h)
{
	__shared__ TriSetup s;
	const int3 ch = chunks[blockIdx.x];
	if (threadIdx.x == 0 && threadIdx.y == 0) tri_setup(vtx, tris+(size_t)ch.x*3, fx, fy, cx, cy, s);
	__syncthreads();
	const int x = ch.y+threadIdx.x, y = ch.z+threadIdx.y;
	// the reference walks 8x8 blocks from the bounding box corner rounded down to a multiple of 8 up to maxx / maxy (exclusive, block
	// granularity); a pixel outside the exact bounding box fails a half-space test anyway, so testing the box is equivalent
	if (x < 0 || y < 0 || x >= w || y >= h) return; // depthMap.isInside(pt)
	const long long fxp = (long long)x<<4, fyp = (long long)y<<4;
	if (!(s.C1+s.DX12*fyp-s.DY12*fxp > 0 && s.C2+s.DX23*fyp-s.DY23*fxp > 0 && s.C3+s.DX31*fyp-s.DY31*fxp > 0)) return;
	// z = INVERT(normalPlane.dot(P.TransformPointI2C(Point2f(pt)))), DepthMap.cpp:1905-1909
	const float X = (float)(dadd((double)(float)x, -cx)/fx), Y = (float)(dadd((double)(float)y, -cy)/fy);
	const float d = __fadd_rn(__fadd_rn(__fmul_rn(s.normalPlane.x, X), __fmul_rn(s.normalPlane.y, Y)), s.normalPlane.z);
	const float z = d == 0.f ? 1000000.f : __fdiv_rn(1.f, d);
	if (z <= 0.f) return; // "due to numerical instability"
	const size_t o = (size_t)y*w+x;
	if (PASS == 0) atomicMax(&owner[o], ch.x);
	else if (owner[o] == ch.x) dn[o] = make_float4(s.normal.x, s.normal.y, s.normal.z, z);
}

} // namespace hcmvs

// owner_d: w*h ints, pre-set to -1 by the caller
cudaError_t hcmvs_launch_raster_triangles(const double* vtx_d, const uint32_t* tris_d, const int3* chunks_d, int nChunks, const double K[9], float4* dn, int* owner_d, int w, int h, cudaStream_t st) {
	if (nChunks > 0) {
		hcmvs::k_raster_triangles<0><<<nChunks, dim3(32, 8), 0, st>>>(vtx_d, tris_d, chunks_d, K[0], K[4], K[2], K[5], dn, owner_d, w, h);
		hcmvs::k_raster_triangles<1><<<nChunks, dim3(32, 8), 0, st>>>(vtx_d, tris_d, chunks_d, K[0], K[4], K[2], K[5], dn, owner_d, w, h);
	}
	return cudaGetLastError();
}
