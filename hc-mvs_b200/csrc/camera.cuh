// f64 pin-hole camera maths on device, restating libs/MVS/Camera.h:273-367 with the reference's operation
// order and NO fused multiply-add (explicit _rn intrinsics), so that the integer pixel decisions made from
// these projections (FLOOR2INT/CEIL2INT/ROUND2INT) are bit-identical to the CPU path.
#pragma once
#include "hcmvs_device.cuh"

namespace hcmvs {

struct D3 { double x, y, z; };

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }

// cv::Matx row * vector, left-to-right accumulation
__device__ __forceinline__ double dot3rn(double a0, double a1, double a2, double v0, double v1, double v2) {
	return dadd(dadd(dmul(a0, v0), dmul(a1, v1)), dmul(a2, v2));
}
// TransformPointI2C(Point3), Camera.h:307-312
__device__ __forceinline__ D3 cam_I2C(const CamConst& c, double x, double y, double z) {
	return D3{dmul(dadd(x, -c.K[2]), z)/c.K[0], dmul(dadd(y, -c.K[5]), z)/c.K[4], z};
}
// TransformPointC2W: R^T X + C, Camera.h:314-316
__device__ __forceinline__ D3 cam_C2W(const CamConst& c, const D3 X) {
	return D3{dadd(dot3rn(c.R[0], c.R[3], c.R[6], X.x, X.y, X.z), c.C[0]),
	          dadd(dot3rn(c.R[1], c.R[4], c.R[7], X.x, X.y, X.z), c.C[1]),
	          dadd(dot3rn(c.R[2], c.R[5], c.R[8], X.x, X.y, X.z), c.C[2])};
}
// TransformPointW2C: R (X - C), Camera.h:354-356
__device__ __forceinline__ D3 cam_W2C(const CamConst& c, const D3 X) {
	const double v0 = dadd(X.x, -c.C[0]), v1 = dadd(X.y, -c.C[1]), v2 = dadd(X.z, -c.C[2]);
	return D3{dot3rn(c.R[0], c.R[1], c.R[2], v0, v1, v2), dot3rn(c.R[3], c.R[4], c.R[5], v0, v1, v2), dot3rn(c.R[6], c.R[7], c.R[8], v0, v1, v2)};
}
__device__ __forceinline__ D3 cam_I2W(const CamConst& c, double x, double y, double z) { return cam_C2W(c, cam_I2C(c, x, y, z)); }
// TransformPointC2I(Point3), Camera.h:350-352 via :339-343
__device__ __forceinline__ void cam_C2I(const CamConst& c, const D3 X, double& u, double& v) {
	u = dadd(c.K[2], dmul(c.K[0], X.x/X.z));
	v = dadd(c.K[5], dmul(c.K[4], X.y/X.z));
}
// ProjectPointP3<float>, Camera.h:273-279: f64 accumulate, cast to f32
__device__ __forceinline__ float3 cam_ProjectP3f(const CamConst& c, const float3 X) {
	const double x = (double)X.x, y = (double)X.y, z = (double)X.z;
	return make_float3(
		(float)dadd(dadd(dadd(dmul(c.P[0], x), dmul(c.P[1], y)), dmul(c.P[2], z)), c.P[3]),
		(float)dadd(dadd(dadd(dmul(c.P[4], x), dmul(c.P[5], y)), dmul(c.P[6], z)), c.P[7]),
		(float)dadd(dadd(dadd(dmul(c.P[8], x), dmul(c.P[9], y)), dmul(c.P[10], z)), c.P[11]));
}
// Cast<float>(R^T * Cast<REAL>(n)): camera-space normal to world
__device__ __forceinline__ float3 cam_NormalC2W(const CamConst& c, const float3 n) {
	const double x = (double)n.x, y = (double)n.y, z = (double)n.z;
	return make_float3((float)dot3rn(c.R[0], c.R[3], c.R[6], x, y, z), (float)dot3rn(c.R[1], c.R[4], c.R[7], x, y, z), (float)dot3rn(c.R[2], c.R[5], c.R[8], x, y, z));
}
__device__ __forceinline__ int floor2int(double x) { return (int)floor(x); }   // FLOOR2INT, Common/Types.h:909-922
__device__ __forceinline__ int ceil2int(double x) { return (int)ceil(x); }
__device__ __forceinline__ int round2int(double x) { return (int)floor(dadd(x, 0.5)); } // ROUND2INT(double), Types.h:944-950
__device__ __forceinline__ int round2int(float x) { return (int)floorf(__fadd_rn(x, 0.5f)); } // ROUND2INT(float), Types.h:937-943
__device__ __forceinline__ bool depth_similar(float d0, float d1, float th) { return __fdiv_rn(fabsf(__fsub_rn(d0, d1)), d0) < th; } // IsDepthSimilar, Util.inl:657-669

} // namespace hcmvs
