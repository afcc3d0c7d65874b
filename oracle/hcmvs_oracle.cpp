// HC-MVS hot-path ORACLE — TEST INFRASTRUCTURE ONLY (see hcmvs_oracle.hpp for the contract).
// Restates, function by function, /root/reference/frame_main/libs/MVS/{DepthMap,SceneDensify,Scene}.cpp.
// Build with: g++ -O3 -fno-fast-math -ffp-contract=off  (no FMA contraction: the float walk of
// ScorePixelImage is part of the reference's numerical "truth").
#include "hcmvs_oracle.hpp"
#include <cmath>
#include <cstring>
#include <cfloat>
#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <limits>

namespace orc {

static const float FPI_ = (float)3.14159265358979323846;
static inline float FD2R(float d) { return d*(FPI_/180.f); }              // Common/Types.h:566
template<typename T> static inline T SQUARE(T a) { return a*a; }
template<typename T> static inline bool ISINSIDE(T v, T l0, T l1) { return l0 <= v && v < l1; } // Types.h:1180
template<typename T> static inline T CLAMPT(T v, T a, T b) { return std::min(std::max(v, a), b); }
// INVERT / INVZERO, Types.h:1213-1219: the float and double overloads of INVZERO win over the template -> 1/0 is FINV_ZERO = 1e6f (Types.h:573)
// for float and INV_ZERO = 1e14 (:555) for double, not numeric_limits::max()
static inline float INVERT(float x) { return x == 0.f ? 1000000.f : 1.f/x; }
static inline double INVERT(double x) { return x == 0.0 ? 1e+14 : 1.0/x; }
static inline int FLOOR2INT(double x) { return int(std::floor(x)); }      // Types.h:909-922 (non-fast variant is the live one)
static inline int CEIL2INT(double x) { return int(std::ceil(x)); }
static inline int ROUND2INT(float x) { return int(std::floor(x+.5f)); }   // Types.h:937-943
static inline int ROUND2INT(double x) { return int(std::floor(x+.5)); }
static inline float DepthSimilarity(float d0, float d1) { return std::abs(d0-d1)/d0; }     // Util.inl:657-665
static inline bool IsDepthSimilar(float d0, float d1, float th) { return DepthSimilarity(d0, d1) < th; }

static double NowSec() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ------------------------------------------------------------------------------------------------
// small dense maths (cv::Matx semantics: plain triple loop, accumulate left to right)
static void Mul33(const double* A, const double* B, double* Cc) {
	for (int i=0; i<3; ++i) for (int j=0; j<3; ++j) {
		double s = A[i*3+0]*B[0*3+j];
		s += A[i*3+1]*B[1*3+j];
		s += A[i*3+2]*B[2*3+j];
		Cc[i*3+j] = s;
	}
}
static void Mul33T(const double* A, const double* B, double* Cc) { // A * B^T
	for (int i=0; i<3; ++i) for (int j=0; j<3; ++j) {
		double s = A[i*3+0]*B[j*3+0];
		s += A[i*3+1]*B[j*3+1];
		s += A[i*3+2]*B[j*3+2];
		Cc[i*3+j] = s;
	}
}
static void Mul3v(const double* A, const double* v, double* o) {
	for (int i=0; i<3; ++i) {
		double s = A[i*3+0]*v[0];
		s += A[i*3+1]*v[1];
		s += A[i*3+2]*v[2];
		o[i] = s;
	}
}
static void MulT3v(const double* A, const double* v, double* o) { // A^T v
	for (int i=0; i<3; ++i) {
		double s = A[0*3+i]*v[0];
		s += A[1*3+i]*v[1];
		s += A[2*3+i]*v[2];
		o[i] = s;
	}
}
// cv::Matx 3x3 inverse (adjugate / determinant)
static void Inv33(const double* a, double* b) {
	double d = a[0]*(a[4]*a[8]-a[5]*a[7]) - a[1]*(a[3]*a[8]-a[5]*a[6]) + a[2]*(a[3]*a[7]-a[4]*a[6]);
	d = 1./d;
	b[0] = (a[4]*a[8]-a[5]*a[7])*d; b[1] = (a[2]*a[7]-a[1]*a[8])*d; b[2] = (a[1]*a[5]-a[2]*a[4])*d;
	b[3] = (a[5]*a[6]-a[3]*a[8])*d; b[4] = (a[0]*a[8]-a[2]*a[6])*d; b[5] = (a[2]*a[3]-a[0]*a[5])*d;
	b[6] = (a[3]*a[7]-a[4]*a[6])*d; b[7] = (a[1]*a[6]-a[0]*a[7])*d; b[8] = (a[0]*a[4]-a[1]*a[3])*d;
}

// ------------------------------------------------------------------------------------------------
// Camera (libs/MVS/Camera.h:290-367, Camera.cpp:112-115,174-181)
void Camera::ComposeP() {
	double M[9]; Mul33(K, R, M);
	const double nC[3] = {-C[0], -C[1], -C[2]};
	double t[3]; Mul3v(M, nC, t);
	for (int i=0; i<3; ++i) { P[i*4+0]=M[i*3+0]; P[i*4+1]=M[i*3+1]; P[i*4+2]=M[i*3+2]; P[i*4+3]=t[i]; }
}
Vec3d Camera::TransformPointI2C(double x, double y) const { return Vec3d{(x-K[2])/K[0], (y-K[5])/K[4], 1.0}; }
Vec3d Camera::TransformPointI2C(double x, double y, double z) const { return Vec3d{(x-K[2])*z/K[0], (y-K[5])*z/K[4], z}; }
Vec3d Camera::TransformPointC2W(const Vec3d& X) const {
	const double v[3] = {X.x, X.y, X.z}; double o[3]; MulT3v(R, v, o);
	return Vec3d{o[0]+C[0], o[1]+C[1], o[2]+C[2]};
}
Vec3d Camera::TransformPointW2C(const Vec3d& X) const {
	const double v[3] = {X.x-C[0], X.y-C[1], X.z-C[2]}; double o[3]; Mul3v(R, v, o);
	return Vec3d{o[0], o[1], o[2]};
}
Vec3d Camera::TransformPointI2W(double x, double y, double z) const { return TransformPointC2W(TransformPointI2C(x, y, z)); }
void Camera::TransformPointC2I(const Vec3d& X, double& u, double& v) const {
	u = K[2]+K[0]*(X.x/X.z); v = K[5]+K[4]*(X.y/X.z);
}
Vec3f Camera::ProjectPointP3f(const Vec3f& X) const {
	const double* p = P;
	return Vec3f{
		(float)(p[0]*X.x + p[1]*X.y + p[2]*X.z + p[3]),
		(float)(p[4]*X.x + p[5]*X.y + p[6]*X.z + p[7]),
		(float)(p[8]*X.x + p[9]*X.y + p[10]*X.z + p[11])};
}
void Camera::ProjectPointPf(const Vec3f& X, float& u, float& v) const {
	const Vec3f q = ProjectPointP3f(X);
	const float invZ = INVERT(q.z);
	u = q.x*invZ; v = q.y*invZ;
}
double Camera::PointDepth(const Vec3d& X) const { return P[8]*X.x + P[9]*X.y + P[10]*X.z + P[11]; }

// ------------------------------------------------------------------------------------------------
// image helpers
void ToGray(const uint8_t* bgr, int w, int h, Image32F& out) {
	// TImage::toGray(COLOR_BGR2GRAY, bNormalize=true): Common/Types.inl:2352-2402, NormRGB_t :1588-1592
	out.w = w; out.h = h; out.d.resize((size_t)w*h);
	const float cb = 0.114f, cg = 0.587f, cr = 0.299f;
	const float s = 1.f/255.f;
	for (size_t i=0, n=(size_t)w*h; i<n; ++i) {
		const float b = float(bgr[i*3+0])*s, g = float(bgr[i*3+1])*s, r = float(bgr[i*3+2])*s;
		out.d[i] = cb*b + cg*g + cr*r;
	}
}

static inline int Reflect101(int p, int len) { // cv::BORDER_REFLECT_101
	if (len == 1) return 0;
	while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2*len-2-p; }
	return p;
}
void InitGraMap(const uint8_t* bgr, int w, int h, Image8U& gra) {
	// SceneDensify.cpp:581-595: cvtColor(BGR2GRAY) u8 -> Sobel 3x3 CV_16S (x and y) -> convertScaleAbs -> addWeighted(.5,.5)
	// third-party arithmetic (OpenCV 4.x, not vendored): gray = (B*3735 + G*19235 + R*9798 + 16384) >> 15 (the 15-bit BT.601 constants of OpenCV >= 3.4, checked against cv2 4.13);
	// Sobel border = BORDER_REFLECT_101; convertScaleAbs = saturate_u8(|v|); addWeighted = saturate_u8(round-half-even(a*.5+b*.5)).
	std::vector<uint8_t> g((size_t)w*h);
	for (size_t i=0, n=(size_t)w*h; i<n; ++i)
		g[i] = (uint8_t)((bgr[i*3+0]*3735 + bgr[i*3+1]*19235 + bgr[i*3+2]*9798 + 16384) >> 15);
	gra.w = w; gra.h = h; gra.d.resize((size_t)w*h);
	for (int y=0; y<h; ++y) {
		const int ym = Reflect101(y-1, h), yp = Reflect101(y+1, h);
		for (int x=0; x<w; ++x) {
			const int xm = Reflect101(x-1, w), xp = Reflect101(x+1, w);
			#define G(xx,yy) int(g[(size_t)(yy)*w+(xx)])
			const int gx = (G(xp,ym)+2*G(xp,y)+G(xp,yp)) - (G(xm,ym)+2*G(xm,y)+G(xm,yp));
			const int gy = (G(xm,yp)+2*G(x,yp)+G(xp,yp)) - (G(xm,ym)+2*G(x,ym)+G(xp,ym));
			#undef G
			const int ax = std::min(std::abs(gx), 255), ay = std::min(std::abs(gy), 255);
			const double v = ax*0.5 + ay*0.5;
			int r = (int)std::nearbyint(v); // cvRound: round half to even
			gra.d[(size_t)y*w+x] = (uint8_t)std::min(std::max(r, 0), 255);
		}
	}
}

void MedianBlur3(Image32F& img) {
	// cv::medianBlur(src, dst, 3) on CV_32F: 3x3 median, BORDER_REPLICATE (third-party, restated)
	const int w = img.w, h = img.h;
	std::vector<float> out((size_t)w*h);
	for (int y=0; y<h; ++y) for (int x=0; x<w; ++x) {
		float v[9]; int n = 0;
		for (int dy=-1; dy<=1; ++dy) for (int dx=-1; dx<=1; ++dx) {
			const int xx = std::min(std::max(x+dx, 0), w-1), yy = std::min(std::max(y+dy, 0), h-1);
			v[n++] = img.d[(size_t)yy*w+xx];
		}
		std::nth_element(v, v+4, v+9);
		out[(size_t)y*w+x] = v[4];
	}
	img.d.swap(out);
}

void MapMatrix2ZigzagIdx(int w, int h, std::vector<uint16_t>& coords, int rawStride) {
	// DepthMap.cpp:354-381 (no mask)
	const int w1 = w-1;
	coords.clear(); coords.reserve((size_t)w*h*2);
	for (int dy=0, hh=rawStride; dy<h; dy+=hh) {
		if (hh*2 > h-dy) hh = h-dy;
		int lastX = 0;
		int xx = 0, xy = 0;
		for (int i=0, ei=w*hh; i<ei; ++i) {
			coords.push_back((uint16_t)xx); coords.push_back((uint16_t)(xy+dy));
			if (xx-- == 0 || ++xy == hh) {
				if (++lastX < w) { xx = lastX; xy = 0; }
				else { xx = w1; xy = lastX-w1; }
			}
		}
	}
}

float SampleBilinear(const Image32F& img, float px, float py) {
	// TImage::sample, Common/Types.inl:2248-2258
	const int lx = (int)px, ly = (int)py;
	const float x = px-lx, x1 = 1.f-x;
	const float y = py-ly, y1 = 1.f-y;
	const float* r0 = &img.d[(size_t)ly*img.w+lx];
	const float* r1 = r0+img.w;
	return (r0[0]*x1 + r0[1]*x)*y1 + (r1[0]*x1 + r1[1]*x)*y;
}

void Dir2Normal(float px, float py, Vec3f& d) { // Util.inl:619-626
	const float siny = std::sin(py);
	d.x = std::cos(px)*siny;
	d.y = std::sin(px)*siny;
	d.z = std::cos(py);
}
void Normal2Dir(const Vec3f& d, float& px, float& py) { // Util.inl:613-618
	px = std::atan2(d.y, d.x);
	py = std::acos(d.z);
}
float ComputeAngleF(const float* V1, const float* V2) { // Util.inl:416-420
	return CLAMPT((V1[0]*V2[0]+V1[1]*V2[1]+V1[2]*V2[2])/std::sqrt((V1[0]*V1[0]+V1[1]*V1[1]+V1[2]*V1[2])*(V2[0]*V2[0]+V2[1]*V2[1]+V2[2]*V2[2])), -1.f, 1.f);
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11). Used by the red-black restatement.
void Philox4x32_10(const uint32_t c_in[4], const uint32_t k_in[2], uint32_t out[4]) {
	uint32_t c0=c_in[0], c1=c_in[1], c2=c_in[2], c3=c_in[3], k0=k_in[0], k1=k_in[1];
	for (int r=0; r<10; ++r) {
		const uint64_t p0 = (uint64_t)0xD2511F53u*c0, p1 = (uint64_t)0xCD9E8D57u*c2;
		const uint32_t n0 = (uint32_t)(p1>>32)^c1^k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0>>32)^c3^k1, n3 = (uint32_t)p0;
		c0=n0; c1=n1; c2=n2; c3=n3;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	out[0]=c0; out[1]=c1; out[2]=c2; out[3]=c3;
}

// ------------------------------------------------------------------------------------------------
// view selection
static float Footprint(const Camera& cam, const Vec3f& X) { // Scene.cpp:531-539
	return (float)(cam.FocalLength()/cam.PointDepth(Vec3d{X.x, X.y, X.z}));
}

bool SelectNeighborViews(Scene& scene, uint32_t ID, std::vector<uint32_t>& points,
	unsigned nMinViews, unsigned nMinPointViews, float fOptimAngle)
{
	// Scene.cpp:545-662
	ImageData& imageData = scene.images[ID];
	std::vector<ViewScore>& neighbors = imageData.neighbors;
	neighbors.clear();
	struct Score { float score, avgScale, avgAngle; uint32_t points; };
	std::vector<Score> scores(scene.images.size(), Score{0,0,0,0});
	if (nMinPointViews > scene.nCalibratedImages()) nMinPointViews = scene.nCalibratedImages();
	unsigned nPoints = 0;
	imageData.avgDepth = 0;
	const SparseCloud& pc = scene.sparse;
	for (size_t idx=0; idx<pc.points.size(); ++idx) {
		const std::vector<uint32_t>& views = pc.views[idx];
		if (!std::binary_search(views.begin(), views.end(), ID)) continue;
		const Vec3f& point = pc.points[idx];
		if (views.size() >= nMinPointViews) points.push_back((uint32_t)idx);
		imageData.avgDepth += (float)imageData.cam.PointDepth(Vec3d{point.x, point.y, point.z});
		++nPoints;
		const float V1[3] = {(float)(imageData.cam.C[0]-(double)point.x), (float)(imageData.cam.C[1]-(double)point.y), (float)(imageData.cam.C[2]-(double)point.z)};
		const float footprint1 = Footprint(imageData.cam, point);
		for (uint32_t view: views) {
			if (view == ID) continue;
			const ImageData& imageData2 = scene.images[view];
			const float V2[3] = {(float)(imageData2.cam.C[0]-(double)point.x), (float)(imageData2.cam.C[1]-(double)point.y), (float)(imageData2.cam.C[2]-(double)point.z)};
			const float fAngle = std::acos(ComputeAngleF(V1, V2));
			const float wAngle = std::min(std::pow(fAngle/fOptimAngle, 1.5f), 1.f);
			const float footprint2 = Footprint(imageData2.cam, point);
			const float fScaleRatio = footprint1/footprint2;
			float wScale;
			if (fScaleRatio > 1.6f) wScale = SQUARE(1.6f/fScaleRatio);
			else if (fScaleRatio >= 1.f) wScale = 1.f;
			else wScale = SQUARE(fScaleRatio);
			Score& score = scores[view];
			score.score += wAngle*wScale;
			score.avgScale += fScaleRatio;
			score.avgAngle += fAngle;
			++score.points;
		}
	}
	imageData.avgDepth /= nPoints;
	// select best neighborViews
	std::vector<Vec2f> projs;
	for (uint32_t IDB=0; IDB<scene.images.size(); ++IDB) {
		const ImageData& imageDataB = scene.images[IDB];
		const Score& score = scores[IDB];
		if (score.points < 3) continue;
		const float boundsA[2] = {(float)imageData.w, (float)imageData.h};
		const float boundsB[2] = {(float)imageDataB.w, (float)imageDataB.h};
		projs.clear();
		for (uint32_t idx: points) {
			const std::vector<uint32_t>& views = pc.views[idx];
			if (!std::binary_search(views.begin(), views.end(), IDB)) continue;
			const Vec3f& point = pc.points[idx];
			Vec2f ptA, ptB;
			imageData.cam.ProjectPointPf(point, ptA.x, ptA.y);
			imageDataB.cam.ProjectPointPf(point, ptB.x, ptB.y);
			const bool inA = ptA.x>=0 && ptA.y>=0 && ptA.x<boundsA[0] && ptA.y<boundsA[1]; // Camera.h:370-372
			const bool inB = ptB.x>=0 && ptB.y>=0 && ptB.x<boundsB[0] && ptB.y<boundsB[1];
			if (inA && inB) projs.push_back(ptA);
		}
		if (projs.empty()) continue;
		// ComputeCoveredArea<float,2,16,false> (Common/Util.inl:711-730)
		unsigned surface[16][16]; std::memset(surface, 0, sizeof(surface));
		for (const Vec2f& p: projs) {
			const float px = (p.x/boundsA[0]+0.f)*16.f, py = (p.y/boundsA[1]+0.f)*16.f;
			surface[FLOOR2INT((double)px)][FLOOR2INT((double)py)] = 1;
		}
		unsigned sum = 0; for (int i=0; i<16; ++i) for (int j=0; j<16; ++j) sum += surface[i][j];
		const float area = float(sum)/(16*16);
		ViewScore nb;
		nb.ID = IDB; nb.points = score.points;
		nb.scale = score.avgScale/score.points;
		nb.angle = score.avgAngle/score.points;
		nb.area = area;
		nb.score = score.score*area;
		neighbors.push_back(nb);
	}
	std::stable_sort(neighbors.begin(), neighbors.end(), [](const ViewScore& a, const ViewScore& b) { return a.score > b.score; }); // q6
	if (points.size() <= 3 || neighbors.size() < std::min(nMinViews, scene.nCalibratedImages()-1)) return false;
	return true;
}

bool FilterNeighborViews(std::vector<ViewScore>& neighbors, float fMinArea, float fMinScale, float fMaxScale,
	float fMinAngle, float fMaxAngle, unsigned nMaxViews)
{
	// Scene.cpp:665-678
	for (size_t n=neighbors.size(); n-- > 0; ) {
		const ViewScore& nb = neighbors[n];
		if (nb.area < fMinArea || !ISINSIDE(nb.scale, fMinScale, fMaxScale) || !ISINSIDE(nb.angle, fMinAngle, fMaxAngle))
			neighbors.erase(neighbors.begin()+n); // RemoveAtMove keeps order
	}
	if (neighbors.size() > nMaxViews) neighbors.resize(nMaxViews);
	return !neighbors.empty();
}

bool SelectViews(Scene& scene, uint32_t idxImage) {
	// SceneDensify.cpp:307-327
	const Params& P = scene.P;
	if (scene.arrDepthData.size() != scene.images.size()) scene.arrDepthData.resize(scene.images.size());
	DepthData& dd = scene.arrDepthData[idxImage];
	dd.idxImage = idxImage; dd.points.clear(); dd.neighbors.clear(); dd.valid = false;
	if (!SelectNeighborViews(scene, idxImage, dd.points, P.nMinViews, P.nMinViewsTrustPoint>1?P.nMinViewsTrustPoint:2, FD2R(P.fOptimAngle)))
		return false;
	dd.neighbors = scene.images[idxImage].neighbors;
	if (!FilterNeighborViews(dd.neighbors, P.fMinArea, 0.2f, 3.2f, FD2R(P.fMinAngle), FD2R(P.fMaxAngle), P.nMaxViews))
		return false;
	dd.valid = true;
	return true;
}

bool InitViews(Scene& scene, uint32_t idxImage, unsigned numNeighbors) {
	// SceneDensify.cpp:336-397, idxNeighbor==NO_ID branch. Neighbour rescaling (|scale-1|>=0.15,
	// DepthMap.h:232-238) is not restated: the synthetic scenes keep scale within that band.
	const Params& P = scene.P;
	DepthData& dd = scene.arrDepthData[idxImage];
	dd.images.clear();
	dd.images.push_back(idxImage);
	const float fMinScore = std::max(dd.neighbors.front().score*(P.fViewMinScoreRatio*0.1f), P.fViewMinScore);
	for (const ViewScore& nb: dd.neighbors) {
		if ((numNeighbors && dd.images.size() > numNeighbors) || nb.score < fMinScore) break;
		dd.images.push_back(nb.ID);
	}
	if (dd.images.size() < 2) { dd.images.clear(); return false; }
	return true;
}

// ------------------------------------------------------------------------------------------------
// DepthEstimator
static const float scaleRanges[12] = {1.f, 0.5f, 0.25f, 0.125f, 0.0625f, 0.03125f, 0.015625f, 0.0078125f, 0.00390625f, 0.001953125f, 0.0009765625f, 0.00048828125f}; // DepthMap.cpp:384

DepthEstimator::DepthEstimator(unsigned nIter, int nIterExternal, Scene& _scene, DepthData& _dd, uint64_t seed)
	: scene(_scene), dd(_dd), P(_scene.P), image0(_scene.images[_dd.images[0]]),
	nIteration(nIter), nIteration_external(nIterExternal), mt((uint32_t)seed)
{
	// DepthMap.cpp:386-439 + ViewData ctor DepthMap.h:412-444
	w = image0.w; h = image0.h;
	dMin = dd.dMin; dMax = dd.dMax; dMinSqr = std::sqrt(dd.dMin); dMaxSqr = std::sqrt(dd.dMax);
	dir = (nIter%2) ? 1 : 0;
	idxScore = dd.images.size() <= 2 ? 0u : 1u;
	smoothBonusDepth = 1.f-P.fRandomSmoothBonus; smoothBonusNormal = (1.f-P.fRandomSmoothBonus)*0.96f;
	smoothSigmaDepth = -1.f/(2.f*SQUARE(P.fRandomSmoothDepth));
	smoothSigmaNormal = -1.f/(2.f*SQUARE(FD2R(P.fRandomSmoothNormal)));
	angle1Range = FD2R(P.fRandomAngle1Range); angle2Range = FD2R(P.fRandomAngle2Range);
	thConfSmall = P.fNCCThresholdKeep*0.2f; thConfBig = P.fNCCThresholdKeep*0.4f;
	thConfRand = P.fNCCThresholdKeep*0.9f; thRobust = P.fNCCThresholdKeep*1.2f;
	for (size_t i=1; i<dd.images.size(); ++i) {
		const ImageData& image1 = (i-1 < dd.scaledImages.size() && dd.scaledImages[i-1].w) ? dd.scaledImages[i-1] : scene.images[dd.images[i]];
		EstimatorView v; v.view = &image1; v.id = dd.images[i];
		double KR[9]; Mul33(image1.cam.K, image1.cam.R, KR);
		Mul33T(KR, image0.cam.R, v.Hl);                                    // K1 R1 R0^T
		const double dC[3] = {image0.cam.C[0]-image1.cam.C[0], image0.cam.C[1]-image1.cam.C[1], image0.cam.C[2]-image1.cam.C[2]};
		Mul3v(KR, dC, v.Hm);                                               // K1 R1 (C0-C1)
		Inv33(image0.cam.K, v.Hr);                                         // K0^-1
		images.push_back(v);
	}
	scores.resize(images.size());
}

float DepthEstimator::Random() { return (float)mt()/(float)std::mt19937::max(); }
float DepthEstimator::RandomRange(float a, float b) { return a + (b-a)*Random(); }
float DepthEstimator::RandomMeanRange(float m, float d) { return m + d*(2.f*Random()-1.f); }

bool DepthEstimator::PreparePixelPatch(int x, int y) {
	x0x = x; x0y = y;
	const int hw = 7; // nSizeHalfWindow, DepthMap.h:354
	return (x-hw >= 0 && y-hw >= 0 && x-hw < w && y-hw < h) && (x+hw >= 0 && y+hw >= 0 && x+hw < w && y+hw < h);
}

bool DepthEstimator::FillPixelPatch() {
	// DepthMap.cpp:450-519 (+GetWeight DepthMap.h:537-548)
	const float tx = dd.graMap.d.empty() ? 0.f : (float)dd.graMap.at(x0x, x0y);
	adapthalfwin = (tx > 100) ? 5 : P.adapthalfwin;
	const Image32F& img = image0.gray;
	sumWeights = 0; float nsq = 0; int n = 0;
	const float colCenter = img.at(x0x, x0y);
	const float sigmaColor = -1.f/(2.f*SQUARE(0.2f));
	const float sigmaSpatial = -1.f/(2.f*SQUARE((int)adapthalfwin));
	for (int i=-adapthalfwin; i<=adapthalfwin; i+=2) {
		for (int j=-adapthalfwin; j<=adapthalfwin; j+=2) {
			const float I = img.at(x0x+j, x0y+i);
			const float wColor = SQUARE(I-colCenter)*sigmaColor;
			const float wSpatial = float(SQUARE(j)+SQUARE(i))*sigmaSpatial;
			const float wgt = (float)std::exp((double)(wColor+wSpatial)); // q10: correctly rounded DENSE_EXP
			tempWeights[n] = I; weights[n] = wgt;
			nsq += I*wgt;
			sumWeights += wgt;
			++n;
		}
	}
	const int pointnum = n;
	const float tm = nsq/sumWeights;
	nsq = 0;
	for (n=0; n<pointnum; ++n) {
		const float t = tempWeights[n]-tm;
		nsq += (tempWeights[n] = weights[n]*t)*t;
	}
	normSq0 = nsq;
	X0 = image0.cam.TransformPointI2C((double)x0x, (double)x0y);
	return true;
}

void DepthEstimator::ComputeHomography(const EstimatorView& v, Depth depth, const Vec3f& normal, float H[9]) const {
	// DepthMap.h:565-574: (Hl + Hm * (n^T * INVERT(n.X0*depth))) * Hr, all f64, then cast to f32
	const double n[3] = {(double)normal.x, (double)normal.y, (double)normal.z};
	const double ndotX = n[0]*X0.x + n[1]*X0.y + n[2]*X0.z;
	const double inv = INVERT(ndotX*(double)depth);
	const double nt[3] = {n[0]*inv, n[1]*inv, n[2]*inv};
	double A[9];
	for (int i=0; i<3; ++i) for (int j=0; j<3; ++j) A[i*3+j] = v.Hl[i*3+j] + v.Hm[i]*nt[j];
	double Hd[9]; Mul33(A, v.Hr, Hd);
	for (int i=0; i<9; ++i) H[i] = (float)Hd[i];
}

float DepthEstimator::ScorePixelImage(const EstimatorView& v, Depth depth, const Vec3f& normal) {
	// DepthMap.cpp:522-616 (NCC core + smoothness) and :890-955 (live score composition, H6-iv)
	float H[9]; ComputeHomography(v, depth, normal, H);
	const float px = float(x0x-adapthalfwin), py = float(x0y-adapthalfwin);
	float Xx = H[0]*px + H[1]*py + H[2];    // ProjectVertex_3x3_2_3, Util.inl:254-259
	float Xy = H[3]*px + H[4]*py + H[5];
	float Xz = H[6]*px + H[7]*py + H[8];
	float bx = Xx, by = Xy, bz = Xz;
	for (int i=0; i<9; ++i) H[i] *= 2.f;     // H *= float(nSizeStep)
	const Image32F& img1 = v.view->gray;
	const float maxx = float(img1.w-2), maxy = float(img1.h-2);
	int n = 0; float sum = 0, sumSq = 0, num = 0;
	for (int i=-adapthalfwin; i<=adapthalfwin; i+=2) {
		for (int j=-adapthalfwin; j<=adapthalfwin; j+=2) {
			const float ptx = Xx/Xz, pty = Xy/Xz;
			if (!(ptx >= 1.f && pty >= 1.f && ptx <= maxx && pty <= maxy)) // isInsideWithBorder<float,1>, Types.h:1632-1635
				return thRobust;
			const float val = SampleBilinear(img1, ptx, pty);
			const float vw = val*weights[n];
			sum += vw;
			sumSq += val*vw;
			num += val*tempWeights[n];
			++n;
			Xx += H[0]; Xy += H[3]; Xz += H[6];
		}
		bx += H[1]; by += H[4]; bz += H[7];
		Xx = bx; Xy = by; Xz = bz;
	}
	const float normSq1 = sumSq-SQUARE(sum)/sumWeights;
	const float nrmSq = normSq0*normSq1;
	if (nrmSq <= 0.f) return thRobust;
	const float ncc = CLAMPT(num/std::sqrt(nrmSq), -1.f, 1.f);
	float score_ncc = 1.f-ncc;
	// encourage smoothness (DENSE_SMOOTHNESS_PLANE), DepthMap.cpp:605-616
	const float nrm[3] = {normal.x, normal.y, normal.z};
	for (const NeighborEstimate& nb: neighborsClose) {
		const float dist = (planeN[0]*nb.X.x + planeN[1]*nb.X.y + planeN[2]*nb.X.z) + planeD; // Planef::Distance
		const float factorDepth = std::exp(SQUARE(dist/depth)*smoothSigmaDepth);
		const float nbn[3] = {nb.normal.x, nb.normal.y, nb.normal.z};
		const float factorNormal = std::exp(SQUARE(std::acos(ComputeAngleF(nrm, nbn)))*smoothSigmaNormal);
		score_ncc *= (1.f-smoothBonusDepth*factorDepth)*(1.f-smoothBonusNormal*factorNormal);
	}
	// live composition (DepthMap.cpp:890-893, 931): photometric_flow weighting of the (undefined) flow score;
	// benchmark/oracle setting photometric_flow=0 => score == score_ncc (q3)
	float score = (1.f-P.photometric_flow)*score_ncc + P.photometric_flow*0.f;
	if (nIteration_external >= P.photo2geo) {
		// plane-prior term, DepthMap.cpp:941-955
		if (!dd.depthMapPrior.d.empty()) {
			const float prior = dd.depthMapPrior.at(x0x, x0y);
			if (prior != 0) {
				const float depthDifference = DepthSimilarity(prior, depth);
				const float weightPrior = std::exp(-SQUARE(depthDifference)/(2*SQUARE(P.fsigmaPrior)));
				score = score*(1.f-P.para_prior) + 2*(1.f-weightPrior)*P.para_prior;
			}
		}
	}
	return score;
}

float DepthEstimator::ScorePixel(Depth depth, const Vec3f& normal) {
	// DepthMap.cpp:987-1046, DENSE_AGGNCC_MINMEAN
	++nScored;
	for (size_t i=0; i<images.size(); ++i)
		scores[i] = ScorePixelImage(images[i], depth, normal);
	if (idxScore == 0)
		return *std::min_element(scores.begin(), scores.end());
	// nth_element(idxScore=1): scores[0] <= scores[1] <= rest
	size_t i0 = 0;
	for (size_t i=1; i<scores.size(); ++i) if (scores[i] < scores[i0]) i0 = i;
	size_t i1 = (i0 == 0) ? 1 : 0;
	for (size_t i=0; i<scores.size(); ++i) if (i != i0 && scores[i] < scores[i1]) i1 = i;
	float score = scores[i0]; int n = 1;
	const float s = scores[i1];
	if (!(s >= thRobust)) { score += s; ++n; }
	return score/n;
}

void DepthEstimator::InitPlane(Depth depth, const Vec3f& normal) {
	// DepthMap.cpp:1730-1738
	planeN[0] = normal.x; planeN[1] = normal.y; planeN[2] = normal.z;
	const float X0f[3] = {(float)X0.x, (float)X0.y, (float)X0.z};
	planeD = -depth*(normal.x*X0f[0] + normal.y*X0f[1] + normal.z*X0f[2]);
}

Depth DepthEstimator::InterpolatePixel(int nx, int ny, Depth depth, const Vec3f& normal) const {
	// DepthMap.cpp:1671-1726 (ray-plane intersection branch)
	const double pn[3] = {(double)normal.x, (double)normal.y, (double)normal.z};
	const Vec3d Xn = image0.cam.TransformPointI2C((double)nx, (double)ny, (double)depth);
	const double planeDd = pn[0]*Xn.x + pn[1]*Xn.y + pn[2]*Xn.z;
	const Depth depthNew = (Depth)(planeDd/(pn[0]*X0.x + pn[1]*X0.y + pn[2]*X0.z));
	return ISINSIDE(depthNew, dMin, dMax) ? depthNew : depth;
}

void DepthEstimator::CorrectNormal(Vec3f& normal) const {
	// DepthMap.h:629-634 + TRMatrixBase::Set(axis, angle) Common/Rotation.inl:707-735
	const float vd[3] = {(float)X0.x, (float)X0.y, (float)X0.z};
	const float cosAngLen = normal.x*vd[0] + normal.y*vd[1] + normal.z*vd[2];
	if (cosAngLen >= 0) {
		const float wa[3] = {normal.y*vd[2]-normal.z*vd[1], normal.z*vd[0]-normal.x*vd[2], normal.x*vd[1]-normal.y*vd[0]};
		const float nvd = std::sqrt(vd[0]*vd[0]+vd[1]*vd[1]+vd[2]*vd[2]);
		const float phi = std::min((std::acos(cosAngLen/nvd)-FD2R(90.f))*1.01f, -0.001f);
		const float wnorm = std::sqrt(wa[0]*wa[0]+wa[1]*wa[1]+wa[2]*wa[2]);
		const float iw = 1.f/wnorm;
		const float wv[3] = {wa[0]*iw, wa[1]*iw, wa[2]*iw};
		const float O[9] = {0.f, -wv[2], wv[1],  wv[2], 0.f, -wv[0],  -wv[1], wv[0], 0.f};
		float OO[9];
		for (int i=0; i<3; ++i) for (int j=0; j<3; ++j) {
			float s = O[i*3+0]*O[0*3+j]; s += O[i*3+1]*O[1*3+j]; s += O[i*3+2]*O[2*3+j]; OO[i*3+j] = s;
		}
		const float sp = std::sin(phi), cp1 = 1.f-std::cos(phi);
		float Rm[9];
		for (int i=0; i<9; ++i) Rm[i] = ((i%4==0) ? 1.f : 0.f) + O[i]*sp + OO[i]*cp1;
		const float nn[3] = {normal.x, normal.y, normal.z};
		float o[3];
		for (int i=0; i<3; ++i) { float s = Rm[i*3+0]*nn[0]; s += Rm[i*3+1]*nn[1]; s += Rm[i*3+2]*nn[2]; o[i] = s; }
		normal.x = o[0]; normal.y = o[1]; normal.z = o[2];
	}
}

Depth DepthEstimator::RandomDepth() { return SQUARE(RandomRange(dMinSqr, dMaxSqr)); } // DepthMap.h:618-621
Vec3f DepthEstimator::RandomNormal(const Vec3f& viewRay) {
	// DepthMap.h:622-626; argument evaluation order is defined here as first draw -> azimuth
	const float a = RandomRange(FD2R(0.f), FD2R(180.f));
	const float b = RandomRange(FD2R(90.f), FD2R(180.f));
	Vec3f nrm; Dir2Normal(a, b, nrm);
	if (nrm.x*viewRay.x + nrm.y*viewRay.y + nrm.z*viewRay.z > 0) { nrm.x = -nrm.x; nrm.y = -nrm.y; nrm.z = -nrm.z; }
	return nrm;
}

static inline float Dot3(const Vec3f& a, const Vec3f& b) { return a.x*b.x + a.y*b.y + a.z*b.z; }

void DepthEstimator::ProcessPixel(int px, int py) {
	// DepthMap.cpp:1050-1501 (viewspread :1504-1608 is off in every shipped config)
	if (!PreparePixelPatch(px, py) || !FillPixelPatch()) return;
	const int hw = 7;
	Image32F& depthMap0 = dd.depthMap; Image32F& confMap0 = dd.confMap; std::vector<Vec3f>& normalMap0 = dd.normalMap;
	struct Ref { int x, y; };
	Ref neighbors[32]; int nNeighbors = 0;
	neighborsClose.clear();
	auto addClose = [&](int nx, int ny, bool asNeighbor) {
		const Depth ndepth = depthMap0.at(nx, ny);
		if (ndepth > 0) {
			if (asNeighbor) neighbors[nNeighbors++] = Ref{nx, ny};
			const Vec3d Xd = image0.cam.TransformPointI2C((double)nx, (double)ny, (double)ndepth);
			neighborsClose.push_back(NeighborEstimate{ndepth, normalMap0[(size_t)ny*w+nx], Vec3f{(float)Xd.x, (float)Xd.y, (float)Xd.z}});
		}
	};
	if (nIteration_external >= 1) {
		// DepthMap.cpp:1064-1274 ("+"-shaped candidate set)
		const float tx = dd.graMap.d.empty() ? 0.f : (float)dd.graMap.at(x0x, x0y);
		const int step = P.propagatestep;
		const int phw = (tx > 150) ? 5 : P.propagatehalfwin;
		Ref cand[64]; int nc = 0;
		if (x0x > phw && x0y > phw && x0x < w-phw && x0y < h-phw) {
			for (int i=1; i<=phw && nc+4<=64; i+=step) {
				cand[nc++] = Ref{x0x, x0y-i}; cand[nc++] = Ref{x0x, x0y+i};
				cand[nc++] = Ref{x0x-i, x0y}; cand[nc++] = Ref{x0x+i, x0y};
			}
		} else if (x0x > hw && x0y > hw && x0x < w-hw && x0y < h-hw) {
			cand[nc++] = Ref{x0x, x0y-1}; cand[nc++] = Ref{x0x, x0y+1};
			cand[nc++] = Ref{x0x-1, x0y}; cand[nc++] = Ref{x0x+1, x0y};
		}
		for (int n=0; n<nc && nNeighbors<32; ++n) addClose(cand[n].x, cand[n].y, true);
	} else if (dir == 0) {
		// LT2RB, DepthMap.cpp:1277-1331
		if (x0x > hw) addClose(x0x-1, x0y, true);
		if (x0y > hw) addClose(x0x, x0y-1, true);
		if (x0x < w-hw) addClose(x0x+1, x0y, false);
		if (x0y < h-hw) addClose(x0x, x0y+1, false);
	} else {
		// RB2LT, DepthMap.cpp:1332-1389
		if (x0x < w-hw) addClose(x0x+1, x0y, true);
		if (x0y < h-hw) addClose(x0x, x0y+1, true);
		if (x0x > hw) addClose(x0x-1, x0y, false);
		if (x0y > hw) addClose(x0x, x0y-1, false);
	}
	float& conf = confMap0.d[(size_t)x0y*w+x0x];
	Depth& depth = depthMap0.d[(size_t)x0y*w+x0x];
	Vec3f& normal = normalMap0[(size_t)x0y*w+x0x];
	const Vec3f viewDir{(float)X0.x, (float)X0.y, (float)X0.z};
	// propagation, DepthMap.cpp:1406-1440
	for (int n=0; n<nNeighbors; ++n) {
		const Ref nx = neighbors[n];
		if (confMap0.at(nx.x, nx.y) >= P.fNCCThresholdKeep) continue;
		NeighborEstimate& neighbor = neighborsClose[n];
		neighbor.depth = InterpolatePixel(nx.x, nx.y, neighbor.depth, neighbor.normal);
		CorrectNormal(neighbor.normal);
		InitPlane(neighbor.depth, neighbor.normal);
		const float nconf = ScorePixel(neighbor.depth, neighbor.normal);
		if (conf > nconf) { conf = nconf; depth = neighbor.depth; normal = neighbor.normal; }
	}
	// random refinement, DepthMap.cpp:1442-1501
	unsigned idxScaleRange = 0;
	RefineIters:
	if (conf <= thConfSmall) idxScaleRange = 2;
	else if (conf <= thConfBig) idxScaleRange = 1;
	else if (conf >= thConfRand) {
		for (unsigned iter=0; iter<P.nRandomIters; ++iter) {
			const Depth ndepth = RandomDepth();
			const Vec3f nnormal = RandomNormal(viewDir);
			const float nconf = ScorePixel(ndepth, nnormal); // q7: plane is stale here
			if (conf > nconf) {
				conf = nconf; depth = ndepth; normal = nnormal;
				if (conf < thConfRand) goto RefineIters;
			}
		}
		return;
	}
	float scaleRange = scaleRanges[idxScaleRange];
	const float depthRange = depth*P.fRandomDepthRatio; // MaxDepthDifference, Util.inl:649-656
	float pdx, pdy; Normal2Dir(normal, pdx, pdy);
	for (unsigned iter=0; iter<P.nRandomIters; ++iter) {
		const Depth ndepth = RandomMeanRange(depth, depthRange*scaleRange);
		if (!ISINSIDE(ndepth, dMin, dMax)) continue;
		const float npx = RandomMeanRange(pdx, angle1Range*scaleRange);
		const float npy = RandomMeanRange(pdy, angle2Range*scaleRange);
		Vec3f nnormal; Dir2Normal(npx, npy, nnormal);
		if (Dot3(nnormal, viewDir) >= 0) continue;
		InitPlane(ndepth, nnormal);
		const float nconf = ScorePixel(ndepth, nnormal);
		if (conf > nconf) {
			conf = nconf; depth = ndepth; normal = nnormal;
			pdx = npx; pdy = npy;
			scaleRange = scaleRanges[++idxScaleRange];
		}
	}
	ViewSpreadAndCoarse(px, py, conf, depth, normal);
}

// Cross-view propagation (DepthMap.cpp:1504-1608) and the restore tree's coarse-estimate hypothesis
// (restore/libs/MVS/DepthMap.cpp:1527-1550). Both sit after the perturbation loop, so the early `return` of a failed
// fully-random search (:1463) skips them, as in the reference.
// Defined behaviour where the reference is undefined or order dependent:
//  q11 the neighbour's maps are those at the end of the PREVIOUS outer iteration (SnapshotMaps) — the reference reads
//      whatever the neighbour holds at that moment, which depends on the order the views are processed in;
//  q12 a projection that is not finite / outside the neighbour's map is skipped ((int) of such a float is UB, and the
//      reference indexes the neighbour's map with the REFERENCE image's bounds, :1527);
//  q13 the candidates' camera-space points use the neighbour's own intrinsics (the RB2LT branch, :1572; the LT2RB
//      branch reads images[1] of the neighbour's DepthData, :1547 — another camera);
//  the neighbour's normals are used as they are, in the NEIGHBOUR's camera frame (:1546) — kept, it is what the code does.
void DepthEstimator::ViewSpreadAndCoarse(int, int, float& conf, Depth& depth, Vec3f& normal) {
	const int hw = 7;
	if (P.viewspread && nIteration_external >= 1) {
		for (const EstimatorView& v: images) {
			const uint32_t id1 = v.id;
			const DepthData& d1 = scene.arrDepthData[id1];
			if (d1.prevDepth.d.empty()) continue;
			float H[9]; ComputeHomography(v, depth, normal, H);
			const float fx0 = (float)x0x, fy0 = (float)x0y;
			const float X1x = H[0]*fx0 + H[1]*fy0 + H[2], X1y = H[3]*fx0 + H[4]*fy0 + H[5], X1z = H[6]*fx0 + H[7]*fy0 + H[8];
			const float x1f = X1x/X1z, y1f = X1y/X1z;
			if (!(std::fabs(x1f) < 1e9f) || !(std::fabs(y1f) < 1e9f)) continue; // q12
			const int x1 = (int)x1f, y1 = (int)y1f;
			if (!(x1 > hw && y1 > hw && x1 < w-hw && y1 < h-hw)) continue;
			if (x1+1 >= d1.prevDepth.w || y1+1 >= d1.prevDepth.h) continue; // q12
			const int cx[4] = {x1, x1, x1-1, x1+1}, cy[4] = {y1-1, y1+1, y1, y1};
			neighborsClose.clear();
			int candX[4], candY[4], nc = 0;
			const Camera& cam1 = scene.images[id1].cam; // the neighbour's maps live at its own resolution
			for (int k=0; k<4; ++k) {
				const Depth nd = d1.prevDepth.at(cx[k], cy[k]);
				if (nd > 0) {
					const Vec3d Xd = cam1.TransformPointI2C((double)cx[k], (double)cy[k], (double)nd); // q13
					neighborsClose.push_back(NeighborEstimate{nd, d1.prevNormal[(size_t)cy[k]*d1.prevDepth.w+cx[k]], Vec3f{(float)Xd.x, (float)Xd.y, (float)Xd.z}});
					candX[nc] = cx[k]; candY[nc] = cy[k]; ++nc;
				}
			}
			for (int n=0; n<nc; ++n) {
				if (d1.prevConf.at(candX[n], candY[n]) >= P.fNCCThresholdKeep) continue;
				const NeighborEstimate& ne = neighborsClose[n];
				const Vec3d Xw = cam1.TransformPointI2W((double)candX[n], (double)candY[n], (double)ne.depth);
				const Vec3f Xf{(float)Xw.x, (float)Xw.y, (float)Xw.z};
				const Vec3d Xc = image0.cam.TransformPointW2C(Vec3d{(double)Xf.x, (double)Xf.y, (double)Xf.z});
				const Depth nd = (float)Xc.z;
				Vec3f nn = ne.normal;
				CorrectNormal(nn);
				InitPlane(nd, nn);
				const float nconf = ScorePixel(nd, nn);
				if (conf > nconf) { conf = nconf; depth = nd; normal = nn; }
			}
		}
	}
	if (!dd.coarseDepth.d.empty() && nIteration_external == (int)P.nEstimationIters_external-1 && nIteration == P.nEstimationIters-1) {
		Depth nd = dd.coarseDepth.at(x0x, x0y);
		Vec3f nn = dd.coarseNormal[(size_t)x0y*w+x0x];
		nd = InterpolatePixel(x0x, x0y, nd, nn);
		CorrectNormal(nn);
		InitPlane(nd, nn);
		const float nconf = ScorePixel(nd, nn);
		if (conf > nconf-0.1f) { conf = nconf; depth = nd; normal = nn; } // the coarse level wins unless clearly worse
	}
}

void ResizeAreaUp(const float* src, int sw, int sh, int cn, float* dst, int dw, int dh) {
	// OpenCV resize(), interpolation INTER_AREA with the destination larger than the source in x or y: the linear kernel
	// with area_mode coordinates (imgproc/src/resize.cpp: sx = floor(dx*scale), fx = (dx+1) - (sx+1)*inv_scale, clamped
	// at 0 and reduced to its fraction; borders clamp), f32 taps, horizontal pass then vertical pass.
	const double inv_scale_x = (double)dw/sw, inv_scale_y = (double)dh/sh;
	const double scale_x = 1./inv_scale_x, scale_y = 1./inv_scale_y;
	std::vector<int> xofs(dw), yofs(dh); std::vector<float> ax(dw), ay(dh);
	auto tab = [](int d, int ssize, double scale, double inv_scale, int& so, float& f) {
		int s = (int)std::floor(d*scale);
		float fr = (float)((d+1) - (s+1)*inv_scale);
		fr = fr <= 0 ? 0.f : fr - std::floor(fr);
		if (s < 0) { fr = 0; s = 0; }
		if (s >= ssize-1) { fr = 0; s = ssize-1; }
		so = s; f = fr;
	};
	for (int dx=0; dx<dw; ++dx) tab(dx, sw, scale_x, inv_scale_x, xofs[dx], ax[dx]);
	for (int dy=0; dy<dh; ++dy) tab(dy, sh, scale_y, inv_scale_y, yofs[dy], ay[dy]);
	for (int dy=0; dy<dh; ++dy) {
		const int y0 = yofs[dy], y1 = std::min(y0+1, sh-1);
		const float b0 = 1.f-ay[dy], b1 = ay[dy];
		for (int dx=0; dx<dw; ++dx) {
			const int x0 = xofs[dx], x1 = std::min(x0+1, sw-1);
			const float a0 = 1.f-ax[dx], a1 = ax[dx];
			for (int c=0; c<cn; ++c) {
				const float r0 = src[((size_t)y0*sw+x0)*cn+c]*a0 + src[((size_t)y0*sw+x1)*cn+c]*a1;
				const float r1 = src[((size_t)y1*sw+x0)*cn+c]*a0 + src[((size_t)y1*sw+x1)*cn+c]*a1;
				dst[((size_t)dy*dw+dx)*cn+c] = r0*b0 + r1*b1;
			}
		}
	}
}

void SetCoarseEstimate(Scene& scene, uint32_t idxImage, const float* depth, const float* normal, int wc, int hc) {
	DepthData& dd = scene.arrDepthData[idxImage];
	const ImageData& image = scene.images[idxImage];
	if (!depth) { dd.coarseDepth = Image32F(); dd.coarseNormal.clear(); return; }
	dd.coarseDepth.w = image.w; dd.coarseDepth.h = image.h; dd.coarseDepth.d.assign((size_t)image.w*image.h, 0.f);
	dd.coarseNormal.assign((size_t)image.w*image.h, Vec3f{0,0,0});
	ResizeAreaUp(depth, wc, hc, 1, dd.coarseDepth.d.data(), image.w, image.h);
	ResizeAreaUp(normal, wc, hc, 3, &dd.coarseNormal[0].x, image.w, image.h);
	for (float value: dd.coarseDepth.d) { // restore/.../SceneDensify.cpp:526-532
		dd.dMin = dd.dMin > value ? value : dd.dMin;
		dd.dMax = dd.dMax > value ? dd.dMax : value;
	}
}

void SnapshotMaps(Scene& scene) {
	for (DepthData& dd: scene.arrDepthData) {
		if (dd.IsEmpty()) continue;
		dd.prevDepth = dd.depthMap; dd.prevConf = dd.confMap; dd.prevNormal = dd.normalMap;
	}
}

// ------------------------------------------------------------------------------------------------
// driver
void InitDepthMapFromSparse(Scene& scene, uint32_t idxImage) {
	// SceneDensify.cpp:772-808 (nMinViewsTrustPoint<2 branch) + graMap :815-819
	DepthData& dd = scene.arrDepthData[idxImage];
	const ImageData& image = scene.images[idxImage];
	const int w = image.w, h = image.h;
	dd.depthMap.w = w; dd.depthMap.h = h; dd.depthMap.d.assign((size_t)w*h, 0.f);
	dd.normalMap.assign((size_t)w*h, Vec3f{0,0,0});
	dd.confMap.w = w; dd.confMap.h = h; dd.confMap.d.assign((size_t)w*h, 0.f);
	const int nPixelArea = 2;
	const Camera& camera = image.cam;
	dd.dMin = FLT_MAX; dd.dMax = 0;
	for (uint32_t ip: dd.points) {
		const Vec3f& X = scene.sparse.points[ip];
		const Vec3d camX = camera.TransformPointW2C(Vec3d{X.x, X.y, X.z});
		double u, v; camera.TransformPointC2I(camX, u, v);
		const int x = ROUND2INT(u), y = ROUND2INT(v);
		const float d = (float)camX.z;
		const int sx = std::max(x-nPixelArea, 0), sy = std::max(y-nPixelArea, 0);
		const int ex = std::min(x+nPixelArea, w-1), ey = std::min(y+nPixelArea, h-1);
		for (int yy=sy; yy<=ey; ++yy) for (int xx=sx; xx<=ex; ++xx) {
			dd.depthMap.d[(size_t)yy*w+xx] = d;
			dd.normalMap[(size_t)yy*w+xx] = Vec3f{0,0,0};
		}
		if (dd.dMin > d) dd.dMin = d;
		if (dd.dMax < d) dd.dMax = d;
	}
	dd.dMin *= 0.9f; dd.dMax *= 1.1f;
	if (!image.bgr.empty()) InitGraMap(image.bgr.data(), w, h, dd.graMap);
	else { dd.graMap.w = w; dd.graMap.h = h; dd.graMap.d.assign((size_t)w*h, 0); }
}

static void ScorePixelInit(DepthEstimator& est, int x, int y) {
	// body of ScoreDepthMapTmp, SceneDensify.cpp:654-673
	DepthData& dd = est.dd; const int w = est.w;
	const size_t o = (size_t)y*w+x;
	if (!est.PreparePixelPatch(x, y) || !est.FillPixelPatch()) {
		dd.depthMap.d[o] = 0; dd.normalMap[o] = Vec3f{0,0,0}; dd.confMap.d[o] = 2.f;
		return;
	}
	Depth& depth = dd.depthMap.d[o]; Vec3f& normal = dd.normalMap[o];
	const Vec3f viewDir{(float)est.X0.x, (float)est.X0.y, (float)est.X0.z};
	if (!ISINSIDE(depth, est.dMin, est.dMax)) {
		depth = est.RandomDepth();
		normal = est.RandomNormal(viewDir);
	} else if (Dot3(normal, viewDir) >= 0) {
		normal = est.RandomNormal(viewDir);
	}
	dd.confMap.d[o] = est.ScorePixel(depth, normal);
}

template<typename F>
static void RunThreads(unsigned nThreads, F&& fn) {
	if (nThreads <= 1) { fn(0u); return; }
	std::vector<std::thread> th;
	for (unsigned t=1; t<nThreads; ++t) th.emplace_back([&fn, t]() { fn(t); });
	fn(0u);
	for (auto& t: th) t.join();
}

static uint64_t PassSeed(uint64_t seed, unsigned pass, unsigned thread) { return seed + 7919ull*pass + thread; } // q1

void ScoreDepthMap(Scene& scene, uint32_t idxImage, int it_external, uint64_t seed, unsigned nThreads) {
	DepthData& dd = scene.arrDepthData[idxImage];
	const ImageData& image = scene.images[idxImage];
	std::vector<uint16_t> coords;
	MapMatrix2ZigzagIdx(image.w, image.h, coords, std::max(64, (int)nThreads*8));
	const size_t N = coords.size()/2;
	std::atomic<size_t> idxPixel{0};
	RunThreads(nThreads, [&](unsigned t) {
		DepthEstimator est(0, it_external, scene, dd, PassSeed(seed, 0, t));
		size_t idx;
		while ((idx = idxPixel.fetch_add(1)) < N) ScorePixelInit(est, coords[idx*2], coords[idx*2+1]);
	});
}

void EndDepthMap(Scene& scene, uint32_t idxImage) {
	// EndDepthMapTmp, SceneDensify.cpp:688-744
	DepthData& dd = scene.arrDepthData[idxImage];
	const float keep = scene.P.fNCCThresholdKeep;
	for (size_t i=0, n=dd.depthMap.d.size(); i<n; ++i) {
		float& depth = dd.depthMap.d[i]; float& conf = dd.confMap.d[i];
		if (depth <= 0 || conf >= keep) { conf = 0; dd.normalMap[i] = Vec3f{0,0,0}; depth = 0; }
		else conf = conf >= 1.f ? 0.f : 1.f-conf;
	}
}

bool EstimateDepthMap(Scene& scene, uint32_t idxImage, int it_external, uint64_t seed, unsigned nThreads, EstimateStats* stats, bool runEnd) {
	// SceneDensify.cpp:758-1072; initialisation (it_external==0 block) is done by the caller
	DepthData& dd = scene.arrDepthData[idxImage];
	const ImageData& image = scene.images[idxImage];
	const Params& P = scene.P;
	std::vector<uint16_t> coords;
	MapMatrix2ZigzagIdx(image.w, image.h, coords, std::max(64, (int)nThreads*8)); // :835
	const size_t N = coords.size()/2;
	MedianBlur3(dd.depthMap); // :859
	double t0 = NowSec();
	ScoreDepthMap(scene, idxImage, it_external, seed, nThreads); // PASS A :915-934
	double t1 = NowSec();
	std::atomic<uint64_t> nHyp{0};
	for (unsigned iter=0; iter<P.nEstimationIters; ++iter) { // PASS B :949-981
		std::atomic<size_t> idxPixel{0};
		RunThreads(nThreads, [&](unsigned t) {
			DepthEstimator est(iter, it_external, scene, dd, PassSeed(seed, 1+iter, t));
			size_t idx;
			while ((idx = idxPixel.fetch_add(1)) < N) {
				const size_t k = est.dir == 0 ? idx : N-1-idx; // DepthMap.cpp:1054
				est.ProcessPixel(coords[k*2], coords[k*2+1]);
			}
			nHyp += est.nScored;
		});
	}
	double t2 = NowSec();
	if (runEnd && it_external == (int)P.nEstimationIters_external-1) EndDepthMap(scene, idxImage); // PASS C :1035-1056
	double t3 = NowSec();
	if (stats) {
		stats->secScore += t1-t0; stats->secSweeps += t2-t1; stats->secEnd += t3-t2;
		stats->nHypotheses += nHyp.load();
		const size_t inner = (size_t)std::max(image.w-14, 0)*(size_t)std::max(image.h-14, 0);
		stats->nPixelIters += (uint64_t)inner*P.nEstimationIters;
	}
	return true;
}

void ScoreHypotheses(Scene& scene, uint32_t idxImage, const float* depth, const float* normal, int smoothMode, float* scoreOut) {
	DepthData& dd = scene.arrDepthData[idxImage];
	const ImageData& image = scene.images[idxImage];
	const int w = image.w, h = image.h, hw = 7;
	DepthEstimator est(0, 0, scene, dd, 0);
	for (int y=0; y<h; ++y) for (int x=0; x<w; ++x) {
		const size_t o = (size_t)y*w+x;
		if (!est.PreparePixelPatch(x, y) || !est.FillPixelPatch()) { scoreOut[o] = 2.f; continue; }
		const Depth d = depth[o]; const Vec3f n{normal[o*3], normal[o*3+1], normal[o*3+2]};
		est.neighborsClose.clear();
		if (smoothMode) {
			auto add = [&](int nx, int ny) {
				const size_t no = (size_t)ny*w+nx;
				if (depth[no] > 0) {
					const Vec3d Xd = image.cam.TransformPointI2C((double)nx, (double)ny, (double)depth[no]);
					est.neighborsClose.push_back(NeighborEstimate{depth[no], Vec3f{normal[no*3], normal[no*3+1], normal[no*3+2]}, Vec3f{(float)Xd.x, (float)Xd.y, (float)Xd.z}});
				}
			};
			if (x > hw) add(x-1, y);
			if (y > hw) add(x, y-1);
			if (x < w-hw) add(x+1, y);
			if (y < h-hw) add(x, y+1);
			est.InitPlane(d, n);
		}
		scoreOut[o] = est.ScorePixel(d, n);
	}
}

// ------------------------------------------------------------------------------------------------
// Red-black restatement (what the CUDA sweep computes)
namespace {
struct PhiloxDraw {
	uint32_t key[2]; uint32_t pixel, pass;
	void Block(uint32_t blk, float u[4]) const {
		const uint32_t ctr[4] = {pixel, pass, blk, 0x48434D56u};
		uint32_t r[4]; Philox4x32_10(ctr, key, r);
		for (int i=0; i<4; ++i) u[i] = (float)r[i]/4294967296.f; // == Random.h:113-116 with (float)max()==2^32
	}
};
}

static void RBKey(uint64_t seed, uint32_t view, uint32_t key[2]) { key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed>>32) ^ (view*0x9E3779B9u); }

static Vec3f RBRandomNormal(float u1, float u2, const Vec3f& viewRay) {
	const float a = FD2R(0.f) + (FD2R(180.f)-FD2R(0.f))*u1;
	const float b = FD2R(90.f) + (FD2R(180.f)-FD2R(90.f))*u2;
	Vec3f nrm; Dir2Normal(a, b, nrm);
	if (Dot3(nrm, viewRay) > 0) { nrm.x = -nrm.x; nrm.y = -nrm.y; nrm.z = -nrm.z; }
	return nrm;
}

static void RBScoreInit(DepthEstimator& est, int x, int y, const uint32_t key[2]) {
	DepthData& dd = est.dd; const int w = est.w;
	const size_t o = (size_t)y*w+x;
	if (!est.PreparePixelPatch(x, y) || !est.FillPixelPatch()) {
		dd.depthMap.d[o] = 0; dd.normalMap[o] = Vec3f{0,0,0}; dd.confMap.d[o] = 2.f;
		return;
	}
	Depth& depth = dd.depthMap.d[o]; Vec3f& normal = dd.normalMap[o];
	const Vec3f viewDir{(float)est.X0.x, (float)est.X0.y, (float)est.X0.z};
	PhiloxDraw rng{{key[0], key[1]}, (uint32_t)o, 0u};
	float u[4]; rng.Block(0, u);
	if (!ISINSIDE(depth, est.dMin, est.dMax)) {
		const float s = est.dMinSqr + (est.dMaxSqr-est.dMinSqr)*u[0];
		depth = s*s;
		normal = RBRandomNormal(u[1], u[2], viewDir);
	} else if (Dot3(normal, viewDir) >= 0) {
		normal = RBRandomNormal(u[1], u[2], viewDir);
	}
	est.neighborsClose.clear();
	dd.confMap.d[o] = est.ScorePixel(depth, normal);
}

static void RBProcessPixel(DepthEstimator& est, int px, int py, const uint32_t key[2], uint32_t pass, const RedBlackCfg& cfg) {
	// Same per-pixel logic as ProcessPixel (DepthMap.cpp:1050-1501) with:
	//  - smoothness set = the 4-neighbourhood (union of the LT2RB/RB2LT sets, DepthMap.cpp:1277-1389)
	//  - propagation sources = per axis direction, the lowest-conf pixel among odd offsets 1,3,..,farReach
	//    (all of the opposite colour, hence stable during the half-sweep)
	//  - counter-based Philox draws: block 1+t for the fully-random try t, block 1+nRandomIters+t for perturbation try t
	//  - q7: stale plane := plane of the current estimate
	if (!est.PreparePixelPatch(px, py) || !est.FillPixelPatch()) return;
	DepthData& dd = est.dd; const Params& P = est.P;
	const int w = est.w, h = est.h, hw = 7;
	const int x0 = px, y0 = py;
	Image32F& depthMap0 = dd.depthMap; Image32F& confMap0 = dd.confMap; std::vector<Vec3f>& normalMap0 = dd.normalMap;
	est.neighborsClose.clear();
	auto addClose = [&](int nx, int ny) {
		const Depth nd = depthMap0.at(nx, ny);
		if (nd > 0) {
			const Vec3d Xd = est.image0.cam.TransformPointI2C((double)nx, (double)ny, (double)nd);
			est.neighborsClose.push_back(NeighborEstimate{nd, normalMap0[(size_t)ny*w+nx], Vec3f{(float)Xd.x, (float)Xd.y, (float)Xd.z}});
		}
	};
	struct Src { int x, y; };
	std::vector<Src> sources;
	if (est.nIteration_external >= 1) {
		// DepthMap.cpp:1064-1274: the fork's "+"-shaped candidate set; every candidate with depth > 0 is both a
		// propagation source and a smoothness neighbour. Offsets 1, 1+step, .. must be odd for the checkerboard
		// (the shipped runs use step 4: offsets 1 and 5).
		const float tx = dd.graMap.d.empty() ? 0.f : (float)dd.graMap.at(x0, y0);
		const int step = P.propagatestep;
		const int phw = (tx > 150) ? 5 : P.propagatehalfwin;
		std::vector<Src> cand;
		if (x0 > phw && y0 > phw && x0 < w-phw && y0 < h-phw) {
			for (int i=1; i<=phw; i+=step) { cand.push_back({x0, y0-i}); cand.push_back({x0, y0+i}); cand.push_back({x0-i, y0}); cand.push_back({x0+i, y0}); }
		} else if (x0 > hw && y0 > hw && x0 < w-hw && y0 < h-hw) {
			cand = {{x0, y0-1}, {x0, y0+1}, {x0-1, y0}, {x0+1, y0}};
		}
		for (const Src& c: cand) if (depthMap0.at(c.x, c.y) > 0) { addClose(c.x, c.y); sources.push_back(c); }
	} else {
	if (x0 > hw) addClose(x0-1, y0);
	if (y0 > hw) addClose(x0, y0-1);
	if (x0 < w-hw) addClose(x0+1, y0);
	if (y0 < h-hw) addClose(x0, y0+1);
	// propagation: one source per direction
	static const int DX[4] = {-1, 0, 1, 0}, DY[4] = {0, -1, 0, 1};
	const int reach = cfg.useFar ? cfg.farReach : 1;
	int srcX[4], srcY[4]; float srcC[4];
	for (int dirn=0; dirn<4; ++dirn) {
		int bx = -1, by = -1; float bconf = P.fNCCThresholdKeep;
		for (int k=1; k<=reach; k+=2) {
			const int nx = x0+DX[dirn]*k, ny = y0+DY[dirn]*k;
			if (nx < hw || ny < hw || nx > w-1-hw || ny > h-1-hw) break;
			if (!(depthMap0.at(nx, ny) > 0)) continue;
			const float c = confMap0.at(nx, ny);
			if (c < bconf) { bconf = c; bx = nx; by = ny; }
		}
		srcX[dirn] = bx; srcY[dirn] = by; srcC[dirn] = bconf;
	}
	if (cfg.nDirs == 2) {
		// one source per axis: the better of the two opposite directions (ties: left / up)
		for (int a=0; a<2; ++a) {
			const int d0 = a, d1 = a+2;
			if (srcX[d1] >= 0 && (srcX[d0] < 0 || srcC[d1] < srcC[d0])) { srcX[d0] = srcX[d1]; srcY[d0] = srcY[d1]; srcC[d0] = srcC[d1]; }
			srcX[d1] = -1;
		}
	}
	for (int dirn=0; dirn<4; ++dirn) if (srcX[dirn] >= 0) sources.push_back({srcX[dirn], srcY[dirn]});
	}
	float& conf = confMap0.d[(size_t)y0*w+x0];
	Depth& depth = depthMap0.d[(size_t)y0*w+x0];
	Vec3f& normal = normalMap0[(size_t)y0*w+x0];
	const Vec3f viewDir{(float)est.X0.x, (float)est.X0.y, (float)est.X0.z};
	for (const Src& sc: sources) {
		const int bx = sc.x, by = sc.y;
		if (confMap0.at(bx, by) >= P.fNCCThresholdKeep) continue; // DepthMap.cpp:1412
		Depth nd = depthMap0.at(bx, by); Vec3f nn = normalMap0[(size_t)by*w+bx];
		nd = est.InterpolatePixel(bx, by, nd, nn);
		est.CorrectNormal(nn);
		est.InitPlane(nd, nn);
		const float nconf = est.ScorePixel(nd, nn);
		if (conf > nconf) { conf = nconf; depth = nd; normal = nn; }
	}
	PhiloxDraw rng{{key[0], key[1]}, (uint32_t)((size_t)y0*w+x0), pass};
	unsigned idxScaleRange = 0;
	unsigned randIter = 0;
	RefineIters:
	if (conf <= est.thConfSmall) idxScaleRange = 2;
	else if (conf <= est.thConfBig) idxScaleRange = 1;
	else if (conf >= est.thConfRand) {
		est.InitPlane(depth, normal); // q7 definition
		for (; randIter<P.nRandomIters; ) {
			float u[4]; rng.Block(1+randIter, u); ++randIter;
			const float s = est.dMinSqr + (est.dMaxSqr-est.dMinSqr)*u[0];
			const Depth ndepth = s*s;
			const Vec3f nnormal = RBRandomNormal(u[1], u[2], viewDir);
			const float nconf = est.ScorePixel(ndepth, nnormal);
			if (conf > nconf) {
				conf = nconf; depth = ndepth; normal = nnormal;
				if (conf < est.thConfRand) goto RefineIters;
			}
		}
		return;
	}
	float scaleRange = scaleRanges[idxScaleRange];
	const float depthRange = depth*P.fRandomDepthRatio;
	float pdx, pdy; Normal2Dir(normal, pdx, pdy);
	for (unsigned iter=0; iter<P.nRandomIters; ++iter) {
		float u[4]; rng.Block(1+P.nRandomIters+iter, u);
		const Depth ndepth = depth + (depthRange*scaleRange)*(2.f*u[0]-1.f);
		if (!ISINSIDE(ndepth, est.dMin, est.dMax)) continue;
		const float npx = pdx + (est.angle1Range*scaleRange)*(2.f*u[1]-1.f);
		const float npy = pdy + (est.angle2Range*scaleRange)*(2.f*u[2]-1.f);
		Vec3f nnormal; Dir2Normal(npx, npy, nnormal);
		if (Dot3(nnormal, viewDir) >= 0) continue;
		est.InitPlane(ndepth, nnormal);
		const float nconf = est.ScorePixel(ndepth, nnormal);
		if (conf > nconf) {
			conf = nconf; depth = ndepth; normal = nnormal;
			pdx = npx; pdy = npy;
			scaleRange = scaleRanges[++idxScaleRange];
		}
	}
	est.ViewSpreadAndCoarse(px, py, conf, depth, normal);
}

// the block-best plane offered to pixel (px,py): interpolate it to the pixel, score with the pixel's own smoothness set
static void RBTestShared(DepthEstimator& est, int px, int py, int sx, int sy, Depth sd, Vec3f sn) {
	if (!est.PreparePixelPatch(px, py) || !est.FillPixelPatch()) return;
	DepthData& dd = est.dd;
	const int w = est.w, h = est.h, hw = 7;
	est.neighborsClose.clear();
	auto addClose = [&](int nx, int ny) {
		const Depth nd = dd.depthMap.at(nx, ny);
		if (nd > 0) {
			const Vec3d Xd = est.image0.cam.TransformPointI2C((double)nx, (double)ny, (double)nd);
			est.neighborsClose.push_back(NeighborEstimate{nd, dd.normalMap[(size_t)ny*w+nx], Vec3f{(float)Xd.x, (float)Xd.y, (float)Xd.z}});
		}
	};
	if (px > hw) addClose(px-1, py);
	if (py > hw) addClose(px, py-1);
	if (px < w-hw) addClose(px+1, py);
	if (py < h-hw) addClose(px, py+1);
	Depth nd = est.InterpolatePixel(sx, sy, sd, sn);
	est.CorrectNormal(sn);
	est.InitPlane(nd, sn);
	const float nconf = est.ScorePixel(nd, sn);
	float& conf = dd.confMap.d[(size_t)py*w+px];
	if (conf > nconf) { conf = nconf; dd.depthMap.d[(size_t)py*w+px] = nd; dd.normalMap[(size_t)py*w+px] = sn; }
}

bool EstimateDepthMapRedBlack(Scene& scene, uint32_t idxImage, int it_external, uint64_t seed, unsigned nThreads,
	const RedBlackCfg& cfg, EstimateStats* stats, bool runEnd)
{
	DepthData& dd = scene.arrDepthData[idxImage];
	const ImageData& image = scene.images[idxImage];
	const Params& P = scene.P;
	const int w = image.w, h = image.h;
	uint32_t key[2]; RBKey(seed, idxImage, key);
	MedianBlur3(dd.depthMap);
	double t0 = NowSec();
	{ // PASS A: every pixel independent
		std::atomic<int> row{0};
		RunThreads(nThreads, [&](unsigned) {
			DepthEstimator est(0, it_external, scene, dd, 0);
			int y; while ((y = row.fetch_add(1)) < h) for (int x=0; x<w; ++x) RBScoreInit(est, x, y, key);
		});
	}
	double t1 = NowSec();
	std::atomic<uint64_t> nHyp{0};
	for (unsigned iter=0; iter<P.nEstimationIters; ++iter) {
		for (int colour=0; colour<2; ++colour) {
			std::atomic<int> row{0};
			if (cfg.blockShare) {
				// warp-cooperative variant: the pixels of an 8x8 block (one GPU warp) first run ProcessPixel, then the pixel with
				// the lowest score offers its refined plane to the other active pixels of the block (one more hypothesis each)
				const int nby = (h+7)/8, nbx = (w+7)/8;
				std::atomic<int> blk{0};
				RunThreads(nThreads, [&](unsigned) {
					DepthEstimator est(iter, it_external, scene, dd, 0);
					int b;
					while ((b = blk.fetch_add(1)) < nby*nbx) {
						const int by0 = (b/nbx)*8, bx0 = (b%nbx)*8;
						int bestX = -1, bestY = -1; float bestC = 3.f;
						for (int ly=0; ly<8; ++ly) for (int lx=0; lx<4; ++lx) { // lane = ly*4+lx
							const int y = by0+ly, x = bx0+lx*2+((y+colour)&1);
							if (x >= w || y >= h) continue;
							RBProcessPixel(est, x, y, key, 1+iter+it_external*64u, cfg);
							if (est.PreparePixelPatch(x, y)) { const float c = dd.confMap.at(x, y); if (c < bestC) { bestC = c; bestX = x; bestY = y; } }
						}
						if (bestX < 0 || bestC >= P.fNCCThresholdKeep) continue;
						const Depth bd = dd.depthMap.at(bestX, bestY); const Vec3f bn = dd.normalMap[(size_t)bestY*w+bestX];
						for (int ly=0; ly<8; ++ly) for (int lx=0; lx<4; ++lx) {
							const int y = by0+ly, x = bx0+lx*2+((y+colour)&1);
							if (x >= w || y >= h || (x == bestX && y == bestY)) continue;
							RBTestShared(est, x, y, bestX, bestY, bd, bn);
						}
					}
					nHyp += est.nScored;
				});
			} else
			RunThreads(nThreads, [&](unsigned) {
				DepthEstimator est(iter, it_external, scene, dd, 0);
				int y;
				while ((y = row.fetch_add(1)) < h)
					for (int x=((y+colour)&1); x<w; x+=2)
						RBProcessPixel(est, x, y, key, 1+iter+it_external*64u, cfg);
				nHyp += est.nScored;
			});
		}
	}
	double t2 = NowSec();
	if (runEnd && it_external == (int)P.nEstimationIters_external-1) EndDepthMap(scene, idxImage);
	double t3 = NowSec();
	if (stats) {
		stats->secScore += t1-t0; stats->secSweeps += t2-t1; stats->secEnd += t3-t2;
		stats->nHypotheses += nHyp.load();
		const size_t inner = (size_t)std::max(w-14, 0)*(size_t)std::max(h-14, 0);
		stats->nPixelIters += (uint64_t)inner*P.nEstimationIters;
	}
	return true;
}

// ------------------------------------------------------------------------------------------------
// FilterDepthMap, SceneDensify.cpp:3006-3259
bool FilterDepthMap(Scene& scene, uint32_t idxRef, const std::vector<uint32_t>& idxNeighbors, bool bAdjust, Image32F& newDepthMap, Image32F& newConfMap) {
	const Params& P = scene.P;
	DepthData& ref = scene.arrDepthData[idxRef];
	const unsigned N = (unsigned)idxNeighbors.size();
	const unsigned nMinViews = std::min(P.nMinViewsFilter, scene.nCalibratedImages()-1);
	const unsigned nMinViewsAdjust = std::min(P.nMinViewsFilterAdjust, scene.nCalibratedImages()-1);
	if (N < nMinViews || N < nMinViewsAdjust) return false;
	const Camera& cameraRef = scene.images[idxRef].cam;
	const int wR = ref.depthMap.w, hR = ref.depthMap.h;
	std::vector<Image32F> depthMaps(N), confMaps(N);
	for (unsigned n=0; n<N; ++n) {
		Image32F& depthMap = depthMaps[n]; depthMap.w = wR; depthMap.h = hR; depthMap.d.assign((size_t)wR*hR, 0.f);
		Image32F& confMap = confMaps[n];
		if (bAdjust) { confMap.w = wR; confMap.h = hR; confMap.d.assign((size_t)wR*hR, 0.f); }
		const uint32_t idxView = ref.neighbors[idxNeighbors[n]].ID;
		const DepthData& depthData = scene.arrDepthData[idxView];
		const Camera& camera = scene.images[idxView].cam;
		const int w = depthData.depthMap.w, h = depthData.depthMap.h;
		for (int i=0; i<h; ++i) for (int j=0; j<w; ++j) {
			const Depth depth = depthData.depthMap.at(j, i);
			if (depth == 0) continue;
			const Vec3d X = camera.TransformPointI2W((double)j, (double)i, (double)depth);
			const Vec3d camX = cameraRef.TransformPointW2C(X);
			if (camX.z <= 0) continue;
			double ux, uy; cameraRef.TransformPointC2I(camX, ux, uy);
			const int xs[4] = {FLOOR2INT(ux), FLOOR2INT(ux), CEIL2INT(ux), CEIL2INT(ux)};
			const int ys[4] = {FLOOR2INT(uy), CEIL2INT(uy), FLOOR2INT(uy), CEIL2INT(uy)};
			for (int p=0; p<4; ++p) {
				const int xr = xs[p], yr = ys[p];
				if (xr < 0 || yr < 0 || xr >= wR || yr >= hR) continue;
				Depth& depthRef = depthMap.d[(size_t)yr*wR+xr];
				if (depthRef != 0 && depthRef < (Depth)camX.z) continue;
				depthRef = (Depth)camX.z;
				if (bAdjust) confMap.d[(size_t)yr*wR+xr] = depthData.confMap.at(j, i);
			}
		}
	}
	const float thDepthDiff = P.fDepthDiffThreshold*1.2f;
	newDepthMap.w = wR; newDepthMap.h = hR; newDepthMap.d.assign((size_t)wR*hR, 0.f);
	newConfMap.w = wR; newConfMap.h = hR; newConfMap.d.assign((size_t)wR*hR, 0.f);
	if (bAdjust) {
		for (int i=0; i<hR; ++i) for (int j=0; j<wR; ++j) {
			const size_t o = (size_t)i*wR+j;
			const Depth depth = ref.depthMap.d[o];
			if (depth == 0) continue;
			float posConf = ref.confMap.d[o], negConf = 0;
			Depth avgDepth = depth*posConf;
			unsigned nPosViews = 0, nNegViews = 0;
			unsigned n = N;
			bool discard = false;
			do {
				const Depth d = depthMaps[--n].d[o];
				if (d == 0) {
					if (nPosViews + nNegViews + n < nMinViews) { discard = true; break; }
					continue;
				}
				if (IsDepthSimilar(depth, d, 0.12f)) { // hard-coded in the fork, SceneDensify.cpp:3127
					const float c = confMaps[n].d[o];
					avgDepth += d*c; posConf += c; ++nPosViews;
				} else {
					if (depth > d) {
						negConf += confMaps[n].d[o]; // occlusion
					} else {
						// free-space violation
						const uint32_t idxView = ref.neighbors[idxNeighbors[n]].ID;
						const DepthData& depthData = scene.arrDepthData[idxView];
						const Camera& camera = scene.images[idxView].cam;
						const Vec3d X = cameraRef.TransformPointI2W((double)j, (double)i, (double)depth);
						double ux, uy; camera.TransformPointC2I(camera.TransformPointW2C(X), ux, uy);
						const int x = ROUND2INT(ux), y = ROUND2INT(uy);
						if (x >= 0 && y >= 0 && x < depthData.confMap.w && y < depthData.confMap.h) {
							const float c = depthData.confMap.at(x, y);
							negConf += (c > 0 ? c : confMaps[n].d[o]);
						} else
							negConf += confMaps[n].d[o];
					}
					++nNegViews;
				}
			} while (n);
			if (!discard && nPosViews >= nMinViewsAdjust && posConf > negConf && ISINSIDE(avgDepth/=posConf, ref.dMin, ref.dMax)) {
				newDepthMap.d[o] = avgDepth;
				newConfMap.d[o] = posConf-negConf;
			}
		}
	} else {
		const float thDepthDiffStrict = P.fDepthDiffThreshold*0.8f;
		const unsigned nMinGoodViewsProc = 75, nMinGoodViewsDeltaProc = 65;
		const unsigned nDeltas = 4;
		const unsigned nMinViewsDelta = nMinViews*(nDeltas-2);
		const int xDs[4][2] = {{-1,0},{1,0},{0,-1},{0,1}};
		for (int i=0; i<hR; ++i) for (int j=0; j<wR; ++j) {
			const size_t o = (size_t)i*wR+j;
			const Depth depth = ref.depthMap.d[o];
			if (depth == 0) continue;
			{
				unsigned nGoodViews = 0, nViews = 0, n = N;
				do {
					const Depth d = depthMaps[--n].d[o];
					if (d > 0) { ++nViews; if (IsDepthSimilar(depth, d, thDepthDiffStrict)) ++nGoodViews; }
				} while (n);
				if (nGoodViews < nMinViews || nGoodViews < nViews*nMinGoodViewsProc/100) continue;
			}
			{
				unsigned nGoodViews = 0, nViews = 0;
				for (unsigned dd=0; dd<nDeltas; ++dd) {
					const int xx = j+xDs[dd][0], yy = i+xDs[dd][1];
					unsigned n = N;
					do {
						--n;
						const Depth d = (xx < 0 || yy < 0 || xx >= wR || yy >= hR) ? 0.f : depthMaps[n].d[(size_t)yy*wR+xx]; // q8
						if (d > 0) { ++nViews; if (IsDepthSimilar(depth, d, thDepthDiff)) ++nGoodViews; }
					} while (n);
				}
				if (nGoodViews < nMinViewsDelta || nGoodViews < nViews*nMinGoodViewsDeltaProc/100) continue;
			}
			newDepthMap.d[o] = depth;
			newConfMap.d[o] = ref.confMap.d[o];
		}
	}
	return true;
}

// ------------------------------------------------------------------------------------------------
// FuseDepthMaps, SceneDensify.cpp:3265-3495 (+Conf2Weight :154-156)
static inline float Conf2Weight(float conf, Depth depth) { return 1.f/(std::max(1.f-conf, 0.03f)*depth*depth); }

void FuseDepthMaps(Scene& scene, PointCloud& pc, bool bEstimateColor, bool bEstimateNormal) {
	const Params& P = scene.P;
	const uint32_t NO_ID = 0xFFFFFFFFu;
	struct Conn { uint32_t idx; float score; };
	std::vector<Conn> connections;
	bool bNormalMap = true;
	for (uint32_t i=0; i<scene.images.size(); ++i) {
		DepthData& dd = scene.arrDepthData[i];
		if (!dd.valid || dd.IsEmpty()) continue;
		connections.push_back(Conn{i, (float)scene.images[i].neighbors.size()});
		if (dd.normalMap.empty()) bNormalMap = false;
	}
	std::stable_sort(connections.begin(), connections.end(), [](const Conn& a, const Conn& b) { return a.score > b.score; }); // q6
	const unsigned nMinViewsFuse = std::min(P.nMinViewsFuse, (unsigned)scene.images.size());
	const float normalError = std::cos(FD2R(P.fNormalDiffThreshold*P.normalweight));
	std::vector<std::vector<uint32_t>> arrDepthIdx(scene.images.size());
	if (bEstimateNormal && !bNormalMap) bEstimateNormal = false;
	struct Proj { uint16_t x, y; };
	std::vector<Depth*> invalidDepths;
	std::vector<uint32_t> views; std::vector<float> weights; std::vector<Proj> pointProjs;
	auto worldNormal = [&](const Camera& cam, const Vec3f& n) { // Cast<float>(R^T * Cast<REAL>(n))
		const double v[3] = {(double)n.x, (double)n.y, (double)n.z}; double o[3]; MulT3v(cam.R, v, o);
		return Vec3f{(float)o[0], (float)o[1], (float)o[2]};
	};
	for (const Conn& conn: connections) {
		const uint32_t idxImage = conn.idx;
		DepthData& depthData = scene.arrDepthData[idxImage];
		for (const ViewScore& nb: depthData.neighbors) {
			std::vector<uint32_t>& di = arrDepthIdx[nb.ID];
			if (!di.empty()) continue;
			const DepthData& ddB = scene.arrDepthData[nb.ID];
			if (ddB.IsEmpty()) continue;
			di.assign(ddB.depthMap.d.size(), NO_ID);
		}
		const ImageData& imageData = scene.images[idxImage];
		const int w = depthData.depthMap.w, h = depthData.depthMap.h;
		std::vector<uint32_t>& depthIdxs = arrDepthIdx[idxImage];
		if (depthIdxs.empty()) depthIdxs.assign((size_t)imageData.w*imageData.h, NO_ID);
		for (int i=0; i<h; ++i) for (int j=0; j<w; ++j) {
			const size_t o = (size_t)i*w+j;
			const Depth depth = depthData.depthMap.d[o];
			if (depth == 0) continue;
			uint32_t& idxPoint = depthIdxs[o];
			if (idxPoint != NO_ID) continue;
			idxPoint = (uint32_t)pc.points.size();
			const Vec3d Pw = imageData.cam.TransformPointI2W((double)(float)j, (double)(float)i, (double)depth);
			const Vec3f point{(float)Pw.x, (float)Pw.y, (float)Pw.z};
			views.clear(); weights.clear(); pointProjs.clear();
			views.push_back(idxImage);
			weights.push_back(Conf2Weight(depthData.confMap.d[o], depth));
			double confidence = (double)weights.back();
			pointProjs.push_back(Proj{(uint16_t)j, (uint16_t)i});
			const Vec3f normal = bNormalMap ? worldNormal(imageData.cam, depthData.normalMap[o]) : Vec3f{0,0,-1};
			// Point3f*double goes through cv::operator*(Point3_<float>,double) -> float-rounded products (:3377-3379)
			Vec3d X{(double)(float)((double)point.x*confidence), (double)(float)((double)point.y*confidence), (double)(float)((double)point.z*confidence)};
			float Cc[3] = {0,0,0};
			if (!imageData.bgr.empty()) for (int c=0; c<3; ++c) Cc[c] = (float)(confidence*(double)float(imageData.bgr[o*3+c]));
			Vec3f Nn{(float)((double)normal.x*confidence), (float)((double)normal.y*confidence), (float)((double)normal.z*confidence)};
			invalidDepths.clear();
			for (const ViewScore& nb: depthData.neighbors) {
				const uint32_t idxImageB = nb.ID;
				DepthData& ddB = scene.arrDepthData[idxImageB];
				if (ddB.IsEmpty()) continue;
				const ImageData& imageDataB = scene.images[idxImageB];
				const Vec3f pt = imageDataB.cam.ProjectPointP3f(point);
				if (pt.z <= 0) continue;
				const int xB = ROUND2INT(pt.x/pt.z), yB = ROUND2INT(pt.y/pt.z);
				const int wB = ddB.depthMap.w, hB = ddB.depthMap.h;
				if (xB < 0 || yB < 0 || xB >= wB || yB >= hB) continue;
				const size_t oB = (size_t)yB*wB+xB;
				Depth& depthB = ddB.depthMap.d[oB];
				if (depthB == 0) continue;
				uint32_t& idxPointB = arrDepthIdx[idxImageB][oB];
				if (idxPointB != NO_ID) continue;
				if (IsDepthSimilar(pt.z, depthB, P.fDepthDiffThreshold*P.depthweight)) {
					const Vec3f normalB = bNormalMap ? worldNormal(imageDataB.cam, ddB.normalMap[oB]) : Vec3f{0,0,-1};
					if (Dot3(normal, normalB) > normalError) {
						const float confidenceB = Conf2Weight(ddB.confMap.d[oB], depthB);
						const size_t idx = std::lower_bound(views.begin(), views.end(), idxImageB)-views.begin(); // InsertSort
						views.insert(views.begin()+idx, idxImageB);
						weights.insert(weights.begin()+idx, confidenceB);
						pointProjs.insert(pointProjs.begin()+idx, Proj{(uint16_t)xB, (uint16_t)yB});
						idxPointB = idxPoint;
						const Vec3d XB = imageDataB.cam.TransformPointI2W((double)(float)xB, (double)(float)yB, (double)depthB);
						X.x += XB.x*(double)confidenceB; X.y += XB.y*(double)confidenceB; X.z += XB.z*(double)confidenceB;
						if (bEstimateColor && !imageDataB.bgr.empty())
							for (int c=0; c<3; ++c) Cc[c] += float(imageDataB.bgr[oB*3+c])*confidenceB;
						if (bEstimateNormal) { Nn.x += normalB.x*confidenceB; Nn.y += normalB.y*confidenceB; Nn.z += normalB.z*confidenceB; }
						confidence += confidenceB;
						continue;
					}
				}
				if (pt.z < depthB) invalidDepths.push_back(&depthB);
			}
			if (views.size() < nMinViewsFuse) {
				for (size_t v=0; v<views.size(); ++v) {
					const DepthData& ddV = scene.arrDepthData[views[v]];
					arrDepthIdx[views[v]][(size_t)pointProjs[v].y*ddV.depthMap.w+pointProjs[v].x] = NO_ID;
				}
			} else {
				const double nrm = 1.0/confidence;
				pc.points.push_back(Vec3f{(float)(X.x*nrm), (float)(X.y*nrm), (float)(X.z*nrm)});
				pc.pointViews.push_back(views);
				pc.pointWeights.push_back(weights);
				if (bEstimateColor) for (int c=0; c<3; ++c)
					pc.colors.push_back((uint8_t)CLAMPT(ROUND2INT(Cc[c]*(float)nrm), 0, 255));
				if (bEstimateNormal) {
					const Vec3f nv{Nn.x*(float)nrm, Nn.y*(float)nrm, Nn.z*(float)nrm};
					const float inv = 1.f/std::sqrt(nv.x*nv.x+nv.y*nv.y+nv.z*nv.z);
					pc.normals.push_back(Vec3f{nv.x*inv, nv.y*inv, nv.z*inv});
				}
				for (Depth* pDepth: invalidDepths) *pDepth = 0;
			}
		}
	}
}

} // namespace orc
