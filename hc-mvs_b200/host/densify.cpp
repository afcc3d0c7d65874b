// Host-side mirror of DepthMapsData / Scene::DenseReconstruction over the C ABI — see densify.h.
#include "densify.h"
#include <algorithm>
#include <memory>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

namespace hcmvs_host {

static double Now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static const float kPi = (float)3.14159265358979323846;
static inline float Deg2Rad(float d) { return d*(kPi/180.f); } // FD2R, Common/Types.h:566

// ------------------------------------------------------------------------------------------------ camera
// cv::Matx products: plain loops, left-to-right accumulation (libs/MVS/Camera.cpp:174-181)
void Camera::ComposeP() {
	double M[9];
	for (int r=0; r<3; ++r) for (int c=0; c<3; ++c) {
		double acc = K[r*3]*R[c];
		acc += K[r*3+1]*R[3+c];
		acc += K[r*3+2]*R[6+c];
		M[r*3+c] = acc;
	}
	for (int r=0; r<3; ++r) {
		double acc = M[r*3]*(-C[0]);
		acc += M[r*3+1]*(-C[1]);
		acc += M[r*3+2]*(-C[2]);
		P[r*4] = M[r*3]; P[r*4+1] = M[r*3+1]; P[r*4+2] = M[r*3+2]; P[r*4+3] = acc;
	}
}
// Camera::PointDepth, Camera.cpp:112-115
static inline double PointDepth(const Camera& cam, const float* X) { return cam.P[8]*X[0] + cam.P[9]*X[1] + cam.P[10]*X[2] + cam.P[11]; }
// Camera::ProjectPointP<float>, Camera.h:273-285
static inline void ProjectPointP(const Camera& cam, const float* X, float& u, float& v) {
	const double* p = cam.P;
	const float qx = (float)(p[0]*X[0] + p[1]*X[1] + p[2]*X[2] + p[3]);
	const float qy = (float)(p[4]*X[0] + p[5]*X[1] + p[6]*X[2] + p[7]);
	const float qz = (float)(p[8]*X[0] + p[9]*X[1] + p[10]*X[2] + p[11]);
	const float invZ = qz == 0.f ? 1000000.f : 1.f/qz; // INVERT -> INVZERO(float) = FINV_ZERO, Common/Types.h:573, 1213-1219
	u = qx*invZ; v = qy*invZ;
}

void ToGray(const uint8_t* bgr, int w, int h, float* gray) {
	// Image::toGray(COLOR_BGR2GRAY, bNormalize), Common/Types.inl:2352-2402 with NormRGB_t (:1588-1592)
	const float inv255 = 1.f/255.f;
	for (size_t i=0, n=(size_t)w*h; i<n; ++i)
		gray[i] = 0.114f*(float(bgr[i*3])*inv255) + 0.587f*(float(bgr[i*3+1])*inv255) + 0.299f*(float(bgr[i*3+2])*inv255);
}

// ------------------------------------------------------------------------------------------------ image rescaling
namespace {
struct DecimateAlpha { int si, di; float alpha; };
// computeResizeAreaTab, OpenCV imgproc/src/resize.cpp
void AreaTab(int ssize, int dsize, double scale, std::vector<DecimateAlpha>& tab) {
	tab.clear();
	for (int dx=0; dx<dsize; ++dx) {
		const double fsx1 = dx*scale, fsx2 = fsx1+scale;
		const double cellWidth = std::min(scale, ssize-fsx1);
		int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
		sx2 = std::min(sx2, ssize-1);
		sx1 = std::min(sx1, sx2);
		if (sx1-fsx1 > 1e-3) tab.push_back({sx1-1, dx, (float)((sx1-fsx1)/cellWidth)});
		for (int sx=sx1; sx<sx2; ++sx) tab.push_back({sx, dx, float(1.0/cellWidth)});
		if (fsx2-sx2 > 1e-3) tab.push_back({sx2, dx, (float)(std::min(std::min(fsx2-sx2, 1.), cellWidth)/cellWidth)});
	}
}
// interpolateCubic, A = -0.75
void CubicCoeffs(float x, float c[4]) {
	const float A = -0.75f;
	c[0] = ((A*(x+1)-5*A)*(x+1)+8*A)*(x+1)-4*A;
	c[1] = ((A+2)*x-(A+3))*x*x+1;
	c[2] = ((A+2)*(1-x)-(A+3))*(1-x)*(1-x)+1;
	c[3] = 1.f-c[0]-c[1]-c[2];
}
inline int RoundHalfEven(double v) { return (int)std::nearbyint(v); } // cvRound / saturate_cast<int>(double)
}

bool ScaleImage(const std::vector<float>& src, int sw, int sh, float scale, std::vector<float>& dst, int& dw, int& dh) {
	if (std::abs(scale-1.f) < 0.15f) return false;
	const double inv_scale = (double)scale, sc = 1./inv_scale; // resize(): scale_x = 1/inv_scale_x, dsize from inv_scale
	dw = RoundHalfEven(sw*inv_scale); dh = RoundHalfEven(sh*inv_scale);
	if (dw < 1 || dh < 1) return false;
	dst.assign((size_t)dw*dh, 0.f);
	if (scale < 1.f) {
		// INTER_AREA, shrinking
		const int isc = RoundHalfEven(sc);
		if (std::abs(sc-isc) < 2.220446049250313e-16) {
			// ResizeAreaFast: block mean, f32 accumulation in raster order of the block, x (1/area)
			const float inv = 1.f/(float)(isc*isc);
			for (int dy=0; dy<dh; ++dy) for (int dx=0; dx<dw; ++dx) {
				const int sy0 = dy*isc, sx0 = dx*isc;
				if (sy0+isc <= sh && sx0+isc <= sw) {
					float sum = 0;
					for (int y=0; y<isc; ++y) for (int x=0; x<isc; ++x) sum += src[(size_t)(sy0+y)*sw+sx0+x];
					dst[(size_t)dy*dw+dx] = sum*inv;
				} else { // partial block at the border: mean of what exists
					float sum = 0; int cnt = 0;
					for (int y=0; y<isc && sy0+y<sh; ++y) for (int x=0; x<isc && sx0+x<sw; ++x) { sum += src[(size_t)(sy0+y)*sw+sx0+x]; ++cnt; }
					dst[(size_t)dy*dw+dx] = cnt ? (float)((double)sum/cnt) : 0.f;
				}
			}
			return true;
		}
		std::vector<DecimateAlpha> xtab, ytab;
		AreaTab(sw, dw, sc, xtab); AreaTab(sh, dh, sc, ytab);
		std::vector<float> buf(dw), sum(dw, 0.f);
		int prev_dy = ytab.empty() ? 0 : ytab[0].di;
		for (size_t j=0; j<ytab.size(); ++j) { // ResizeArea_Invoker
			const float beta = ytab[j].alpha; const int dy = ytab[j].di, sy = ytab[j].si;
			const float* S = &src[(size_t)sy*sw];
			std::fill(buf.begin(), buf.end(), 0.f);
			for (const DecimateAlpha& t: xtab) buf[t.di] += S[t.si]*t.alpha;
			if (dy != prev_dy) {
				float* D = &dst[(size_t)prev_dy*dw];
				for (int dx=0; dx<dw; ++dx) { D[dx] = sum[dx]; sum[dx] = beta*buf[dx]; }
				prev_dy = dy;
			} else for (int dx=0; dx<dw; ++dx) sum[dx] += beta*buf[dx];
		}
		if (!ytab.empty()) { float* D = &dst[(size_t)prev_dy*dw]; for (int dx=0; dx<dw; ++dx) D[dx] = sum[dx]; }
		return true;
	}
	// INTER_CUBIC, enlarging: taps sx-1..sx+2 / sy-1..sy+2 with replicated borders, horizontal pass then vertical pass, f32
	std::vector<int> xofs(dw), yofs(dh); std::vector<float> ax((size_t)dw*4), ay((size_t)dh*4);
	for (int dx=0; dx<dw; ++dx) { float fx = (float)((dx+0.5)*sc-0.5); const int sx = (int)std::floor(fx); fx -= sx; xofs[dx] = sx; CubicCoeffs(fx, &ax[(size_t)dx*4]); }
	for (int dy=0; dy<dh; ++dy) { float fy = (float)((dy+0.5)*sc-0.5); const int sy = (int)std::floor(fy); fy -= sy; yofs[dy] = sy; CubicCoeffs(fy, &ay[(size_t)dy*4]); }
	auto clampi = [](int v, int n) { return v < 0 ? 0 : (v >= n ? n-1 : v); };
	std::vector<float> rows[4]; for (auto& r: rows) r.resize(dw);
	int rowOf[4] = {-1000000, -1000000, -1000000, -1000000};
	for (int dy=0; dy<dh; ++dy) {
		const float* b = &ay[(size_t)dy*4];
		const float* R[4];
		for (int k=0; k<4; ++k) {
			const int sy = clampi(yofs[dy]-1+k, sh);
			int slot = -1;
			for (int q=0; q<4; ++q) if (rowOf[q] == sy) slot = q;
			if (slot < 0) { // horizontal pass of source row sy into a free slot
				for (int q=0; q<4 && slot<0; ++q) { bool used = false; for (int kk=0; kk<4; ++kk) if (rowOf[q] == clampi(yofs[dy]-1+kk, sh)) used = true; if (!used) slot = q; }
				const float* S = &src[(size_t)sy*sw];
				float* D = rows[slot].data();
				for (int dx=0; dx<dw; ++dx) {
					const float* a = &ax[(size_t)dx*4]; const int sx = xofs[dx];
					D[dx] = S[clampi(sx-1, sw)]*a[0] + S[clampi(sx, sw)]*a[1] + S[clampi(sx+1, sw)]*a[2] + S[clampi(sx+2, sw)]*a[3];
				}
				rowOf[slot] = sy;
			}
			R[k] = rows[slot].data();
		}
		float* D = &dst[(size_t)dy*dw];
		for (int dx=0; dx<dw; ++dx) D[dx] = R[0][dx]*b[0] + R[1][dx]*b[1] + R[2][dx]*b[2] + R[3][dx]*b[3];
	}
	return true;
}

// cv::resize(src, dst, Size(dw, dh), 0, 0, INTER_AREA) on an 8-bit 3-channel image, shrinking (Image::ResizeImage, Image.cpp:139-160).
// OpenCV imgproc/src/resize.cpp: integer scale factors take resizeAreaFast_ (int block sums; the 2x2 case rounds (sum+2)>>2 in
// ResizeAreaFastVec, any other factor saturate_cast<uchar>(sum*(1.f/area)) = round-half-even), everything else resizeArea_ with the
// DecimateAlpha tables, f32 accumulation and a round-half-even store. Cross-checked bit for bit with cv2 in the CPU tests.
bool ResizeAreaBGR(const uint8_t* src, int sw, int sh, int dw, int dh, uint8_t* dst) {
	if (!src || !dst || sw < 1 || sh < 1 || dw < 1 || dh < 1 || dw > sw || dh > sh) return false;
	const int cn = 3;
	const double scale_x = (double)sw/dw, scale_y = (double)sh/dh;
	const int iscale_x = RoundHalfEven(scale_x), iscale_y = RoundHalfEven(scale_y);
	auto sat = [](float v) { const int r = (int)std::nearbyintf(v); return (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r); };
	if (std::abs(scale_x-iscale_x) < 2.220446049250313e-16 && std::abs(scale_y-iscale_y) < 2.220446049250313e-16) {
		const int area = iscale_x*iscale_y;
		const float scale = 1.f/(float)area;
		const int dwidth1 = sw/iscale_x; // destination pixels whose whole block lies inside the source
		const bool fast2 = iscale_x == 2 && iscale_y == 2;
		for (int dy=0; dy<dh; ++dy) {
			uint8_t* D = dst+(size_t)dy*dw*cn;
			const int sy0 = dy*iscale_y;
			const int w = sy0+iscale_y <= sh ? dwidth1 : 0;
			if (sy0 >= sh) { memset(D, 0, (size_t)dw*cn); continue; }
			for (int dx=0; dx<dw; ++dx) for (int c=0; c<cn; ++c) {
				const int sx0 = dx*iscale_x;
				if (dx < w) {
					int sum = 0;
					for (int y=0; y<iscale_y; ++y) for (int x=0; x<iscale_x; ++x) sum += src[((size_t)(sy0+y)*sw+sx0+x)*cn+c];
					D[dx*cn+c] = fast2 ? (uint8_t)((sum+2)>>2) : sat((float)sum*scale);
				} else { // partial block at the border: mean of the pixels that exist
					int sum = 0, count = 0;
					for (int y=0; y<iscale_y && sy0+y<sh; ++y) for (int x=0; x<iscale_x && sx0+x<sw; ++x) { sum += src[((size_t)(sy0+y)*sw+sx0+x)*cn+c]; ++count; }
					D[dx*cn+c] = count ? sat((float)sum/count) : 0;
				}
			}
		}
		return true;
	}
	std::vector<DecimateAlpha> xtab, ytab;
	AreaTab(sw, dw, scale_x, xtab); AreaTab(sh, dh, scale_y, ytab);
	const int dwn = dw*cn;
	std::vector<float> buf(dwn), sum(dwn, 0.f);
	int prev_dy = ytab.empty() ? 0 : ytab[0].di;
	for (size_t j=0; j<ytab.size(); ++j) { // ResizeArea_Invoker<uchar, float>
		const float beta = ytab[j].alpha; const int dy = ytab[j].di, sy = ytab[j].si;
		const uint8_t* S = src+(size_t)sy*sw*cn;
		std::fill(buf.begin(), buf.end(), 0.f);
		for (const DecimateAlpha& t: xtab) {
			const float a = t.alpha; const uint8_t* sp = S+(size_t)t.si*cn; float* bp = &buf[(size_t)t.di*cn];
			bp[0] += sp[0]*a; bp[1] += sp[1]*a; bp[2] += sp[2]*a;
		}
		if (dy != prev_dy) {
			uint8_t* D = dst+(size_t)prev_dy*dwn;
			for (int dx=0; dx<dwn; ++dx) { D[dx] = sat(sum[dx]); sum[dx] = beta*buf[dx]; }
			prev_dy = dy;
		} else for (int dx=0; dx<dwn; ++dx) sum[dx] += beta*buf[dx];
	}
	if (!ytab.empty()) { uint8_t* D = dst+(size_t)prev_dy*dwn; for (int dx=0; dx<dwn; ++dx) D[dx] = sat(sum[dx]); }
	return true;
}

// TImage::computeMaxResolution, Common/Types.inl:2442-2460
unsigned ComputeMaxResolution(unsigned width, unsigned height, unsigned& level, unsigned minImageSize, unsigned maxImageSize) {
	const unsigned imageSize = std::max(width, height);
	if (level == 0) return std::min(imageSize, maxImageSize);
	unsigned size = imageSize>>level;
	if (size < minImageSize) {
		level = 0;
		while ((imageSize>>(level+1)) >= minImageSize) ++level;
		size = imageSize>>level;
	}
	return std::min(size, maxImageSize);
}

bool Scene::ReloadImages(unsigned nResolutionLevel, unsigned nMinResolution, unsigned nMaxResolution, std::string* err) {
	// Scene::ComputeDepthMaps, SceneDensify.cpp:3617-3631: RecomputeMaxResolution + ReloadImage + UpdateCamera per valid image
	for (Image& im: images) {
		if (!im.calibrated) continue;
		unsigned level = nResolutionLevel; // the reference passes OPTDENSE::nResolutionLevel by reference: a lowered level sticks
		const unsigned nMax = ComputeMaxResolution((unsigned)im.width, (unsigned)im.height, level, nMinResolution, nMaxResolution);
		nResolutionLevel = level;
		// Image::ResizeImage, Image.cpp:139-160 (integer arithmetic on the unsigned sizes)
		unsigned w = (unsigned)im.width, h = (unsigned)im.height;
		if (nMax == 0 || std::max(w, h) <= nMax) continue;
		if (w > h) { h = h*nMax/w; w = nMax; } else { w = w*nMax/h; h = nMax; }
		if (w < 1 || h < 1) { if (err) *err = "image '"+im.name+"' vanishes at this resolution level"; return false; }
		if (!im.bgr.empty()) {
			std::vector<uint8_t> scaled((size_t)w*h*3);
			if (!ResizeAreaBGR(im.bgr.data(), im.width, im.height, (int)w, (int)h, scaled.data())) { if (err) *err = "cannot resize '"+im.name+"'"; return false; }
			im.bgr.swap(scaled);
			im.gray.resize((size_t)w*h);
			ToGray(im.bgr.data(), (int)w, (int)h, im.gray.data());
		} else if (!im.gray.empty()) { if (err) *err = "image '"+im.name+"' has gray pixels only: reload needs the colour image"; return false; }
		// Image::UpdateCamera: K of the new resolution from the normalised intrinsics (Image.cpp:194-213, Camera.h:167-180)
		if (im.hasKnorm) {
			const float fScale = (float)std::max(w, h);
			im.camera.K[0] = im.Knorm[0]*fScale; im.camera.K[4] = im.Knorm[1]*fScale;
			if (im.Knorm[2] == 0 && im.Knorm[3] == 0) { im.camera.K[2] = 0.5*(w-1); im.camera.K[5] = 0.5*(h-1); }
			else { im.camera.K[2] = im.Knorm[2]*fScale; im.camera.K[5] = im.Knorm[3]*fScale; }
		} else {
			double Kn[9]; ScaleK(im.camera.K, im.width, im.height, (int)w, (int)h, Kn);
			memcpy(im.camera.K, Kn, sizeof(Kn));
		}
		im.width = (int)w; im.height = (int)h;
		im.camera.ComposeP();
		im.neighbors.clear();
	}
	return true;
}

void ScaleK(const double K[9], int w, int h, int newW, int newH, double Kout[9]) {
	const double s = (double)(float)std::max(newW, newH)/(double)(float)std::max(w, h);
	for (int i=0; i<9; ++i) Kout[i] = K[i];
	Kout[0] = K[0]*s; Kout[4] = K[4]*s; Kout[2] = K[2]*s; Kout[5] = K[5]*s;
}

// ------------------------------------------------------------------------------------------------ view selection
static inline float CosAngle(const float* a, const float* b) { // ComputeAngle<float,float>, Common/Util.inl:416-420
	const float c = (a[0]*b[0]+a[1]*b[1]+a[2]*b[2])/std::sqrt((a[0]*a[0]+a[1]*a[1]+a[2]*a[2])*(b[0]*b[0]+b[1]*b[1]+b[2]*b[2]));
	return std::min(std::max(c, -1.f), 1.f);
}
static inline bool Contains(const std::vector<uint32_t>& sortedViews, uint32_t id) { return std::binary_search(sortedViews.begin(), sortedViews.end(), id); }

// static chunks of [0, n) on nThreads threads (the caller's included); every item writes only its own output slot, so the result does
// not depend on the thread count
template<typename F>
static void ParallelChunks(size_t n, unsigned nThreads, size_t grain, F&& body) {
	nThreads = (unsigned)std::max<size_t>(1, std::min<size_t>(nThreads, n/grain+1));
	if (nThreads == 1) { body((size_t)0, n, 0u); return; }
	std::vector<std::thread> th;
	for (unsigned t=1; t<nThreads; ++t) th.emplace_back([&, t]() { body(n*t/nThreads, n*(t+1)/nThreads, t); });
	body((size_t)0, n/nThreads, 0u);
	for (std::thread& x: th) x.join();
}

bool Scene::SelectNeighborViews(uint32_t ID, std::vector<uint32_t>& points, unsigned nMinViews, unsigned nMinPointViews, float fOptimAngle, unsigned nThreads) {
	// Scene.cpp:545-662, "Multi-View Stereo for Community Photo Collections" style view scoring.
	// The reference runs one image per OpenMP thread (:3652-3667); the view a GPU is waiting for is on the critical path, so the work
	// INSIDE one image can be spread over nThreads as well: the expensive per-(point, view) terms (acos, pow, two projections) are
	// computed in parallel into per-point slots and the f32 sums are then formed serially in the reference's point order — the
	// result is bit-identical for every thread count (test_host_view_selection_and_image_prep_bit_exact).
	Image& ref = images[ID];
	ref.neighbors.clear();
	struct Acc { float score = 0, scale = 0, angle = 0; uint32_t n = 0; };
	std::vector<Acc> acc(images.size());
	nMinPointViews = std::min(nMinPointViews, nCalibratedImages());
	ref.avgDepth = 0;
	// pass 0: the points this image sees, in cloud order
	std::vector<uint32_t> seen;
	{
		std::vector<std::vector<uint32_t>> part(std::max(1u, nThreads));
		ParallelChunks(pointcloud.size(), nThreads, 4096, [&](size_t lo, size_t hi, unsigned t) {
			for (size_t ip=lo; ip<hi; ++ip) if (Contains(pointcloud.views[ip], ID)) part[t].push_back((uint32_t)ip);
		});
		for (const auto& v: part) seen.insert(seen.end(), v.begin(), v.end());
	}
	// pass 1: per point its depth in the reference view and one term per other view that sees it
	struct Term { uint32_t other; float w, ratio, ang; };
	std::vector<size_t> off(seen.size()+1, 0);
	for (size_t k=0; k<seen.size(); ++k) off[k+1] = off[k]+pointcloud.views[seen[k]].size()-1;
	std::vector<Term> terms(off.back());
	std::vector<float> depthRef(seen.size());
	ParallelChunks(seen.size(), nThreads, 256, [&](size_t lo, size_t hi, unsigned) {
		for (size_t k=lo; k<hi; ++k) {
			const size_t ip = seen[k];
			const std::vector<uint32_t>& pv = pointcloud.views[ip];
			const float* X = &pointcloud.xyz[ip*3];
			depthRef[k] = (float)PointDepth(ref.camera, X);
			const float toRef[3] = {(float)(ref.camera.C[0]-(double)X[0]), (float)(ref.camera.C[1]-(double)X[1]), (float)(ref.camera.C[2]-(double)X[2])};
			const float fpRef = (float)(ref.camera.K[0]/PointDepth(ref.camera, X)); // Footprint, Scene.cpp:531-539
			Term* out = terms.data()+off[k];
			for (uint32_t other: pv) {
				if (other == ID) continue;
				const Image& o = images[other];
				const float toOther[3] = {(float)(o.camera.C[0]-(double)X[0]), (float)(o.camera.C[1]-(double)X[1]), (float)(o.camera.C[2]-(double)X[2])};
				const float ang = std::acos(CosAngle(toRef, toOther));
				const float wAngle = std::min(std::pow(ang/fOptimAngle, 1.5f), 1.f);
				const float fpOther = (float)(o.camera.K[0]/PointDepth(o.camera, X));
				const float ratio = fpRef/fpOther;
				float wScale;
				if (ratio > 1.6f) { const float q = 1.6f/ratio; wScale = q*q; }
				else if (ratio >= 1.f) wScale = 1.f;
				else wScale = ratio*ratio;
				*out++ = Term{other, wAngle*wScale, ratio, ang};
			}
		}
	});
	// pass 2: the sums, serially in cloud order
	for (size_t k=0; k<seen.size(); ++k) {
		if (pointcloud.views[seen[k]].size() >= nMinPointViews) points.push_back(seen[k]);
		ref.avgDepth += depthRef[k];
		for (size_t j=off[k]; j<off[k+1]; ++j) { const Term& t = terms[j]; Acc& a = acc[t.other]; a.score += t.w; a.scale += t.ratio; a.angle += t.ang; ++a.n; }
	}
	const unsigned nSeen = (unsigned)seen.size();
	ref.avgDepth /= nSeen;
	// pass 3: the area the common points cover in the reference image, per candidate view
	std::vector<ViewScore> cand(images.size()); std::vector<char> has(images.size(), 0);
	ParallelChunks(images.size(), nThreads, 1, [&](size_t lo, size_t hi, unsigned) {
		std::vector<float> projA;
		for (size_t IDB=lo; IDB<hi; ++IDB) {
			const Acc& a = acc[IDB];
			if (a.n < 3) continue;
			const Image& B = images[IDB];
			const float wA = (float)ref.width, hA = (float)ref.height, wB = (float)B.width, hB = (float)B.height;
			projA.clear();
			for (uint32_t ip: points) {
				if (!Contains(pointcloud.views[ip], (uint32_t)IDB)) continue;
				const float* X = &pointcloud.xyz[(size_t)ip*3];
				float ua, va, ub, vb;
				ProjectPointP(ref.camera, X, ua, va);
				ProjectPointP(B.camera, X, ub, vb);
				if (ua >= 0 && va >= 0 && ua < wA && va < hA && ub >= 0 && vb >= 0 && ub < wB && vb < hB) { projA.push_back(ua); projA.push_back(va); }
			}
			if (projA.empty()) continue;
			// ComputeCoveredArea<float,2,16,false>, Common/Util.inl:711-730
			bool cell[16][16] = {};
			for (size_t k=0; k<projA.size(); k+=2) {
				const float gx = (projA[k]/wA+0.f)*16.f, gy = (projA[k+1]/hA+0.f)*16.f;
				cell[(int)std::floor(gx)][(int)std::floor(gy)] = true;
			}
			unsigned covered = 0;
			for (auto& row: cell) for (bool c: row) covered += c ? 1u : 0u;
			const float area = float(covered)/256;
			ViewScore vs;
			vs.ID = (uint32_t)IDB; vs.points = a.n; vs.scale = a.scale/a.n; vs.angle = a.angle/a.n; vs.area = area; vs.score = a.score*area;
			cand[IDB] = vs; has[IDB] = 1;
		}
	});
	for (size_t IDB=0; IDB<images.size(); ++IDB) if (has[IDB]) ref.neighbors.push_back(cand[IDB]);
	std::stable_sort(ref.neighbors.begin(), ref.neighbors.end(), [](const ViewScore& l, const ViewScore& r) { return l.score > r.score; });
	return points.size() > 3 && ref.neighbors.size() >= std::min(nMinViews, nCalibratedImages()-1);
}

bool Scene::FilterNeighborViews(std::vector<ViewScore>& nb, float fMinArea, float fMinScale, float fMaxScale, float fMinAngle, float fMaxAngle, unsigned nMaxViews) {
	// Scene.cpp:665-678
	std::vector<ViewScore> kept;
	for (const ViewScore& v: nb) {
		const bool scaleOk = fMinScale <= v.scale && v.scale < fMaxScale, angleOk = fMinAngle <= v.angle && v.angle < fMaxAngle;
		if (v.area >= fMinArea && scaleOk && angleOk) kept.push_back(v);
	}
	if (kept.size() > nMaxViews) kept.resize(nMaxViews);
	nb.swap(kept);
	return !nb.empty();
}

// ------------------------------------------------------------------------------------------------ DepthMapsData
DepthMapsData::DepthMapsData(Scene& s, hcmvs_ctx* c, const hcmvs_params& p, const ViewSelectionParams& vs)
	: arrDepthData(s.images.size()), scene(s), ctx(c), P(p), VS(vs) {}

bool DepthMapsData::Fail(const char* what) {
	lastError = std::string(what)+": "+hcmvs_last_error();
	return false;
}

bool DepthMapsData::SelectViews(uint32_t idxImage, unsigned nThreads) {
	// SceneDensify.cpp:307-327
	DepthData& dd = arrDepthData[idxImage];
	dd.points.clear(); dd.neighbors.clear(); dd.valid = false;
	if (!scene.images[idxImage].calibrated) return false; // !imageData.IsValid(), SceneDensify.cpp:3655
	if (!scene.SelectNeighborViews(idxImage, dd.points, P.nMinViews, P.nMinViewsTrustPoint > 1 ? P.nMinViewsTrustPoint : 2, Deg2Rad(VS.fOptimAngle), nThreads))
		return false;
	dd.neighbors = scene.images[idxImage].neighbors;
	if (!Scene::FilterNeighborViews(dd.neighbors, VS.fMinArea, 0.2f, 3.2f, Deg2Rad(VS.fMinAngle), Deg2Rad(VS.fMaxAngle), P.nMaxViews))
		return false;
	dd.valid = true;
	return true;
}

bool DepthMapsData::UploadView(uint32_t idxImage) {
	DepthData& dd = arrDepthData[idxImage];
	if (dd.uploaded) return true;
	Image& im = scene.images[idxImage];
	if (im.gray.empty()) {
		if (im.bgr.empty()) { lastError = "image has neither gray nor colour pixels"; return false; }
		im.gray.resize((size_t)im.width*im.height);
		ToGray(im.bgr.data(), im.width, im.height, im.gray.data());
	}
	if (hcmvs_set_view(ctx, idxImage, im.width, im.height, im.camera.K, im.camera.R, im.camera.C, im.gray.data(), im.bgr.empty() ? nullptr : im.bgr.data()) != HCMVS_OK)
		return Fail("hcmvs_set_view");
	dd.uploaded = true;
	return true;
}

bool DepthMapsData::InitViews(uint32_t idxImage, uint32_t numNeighbors) {
	// SceneDensify.cpp:336-397 (idxNeighbor == NO_ID branch), including the rescaling of neighbours with |scale-1| >= 0.15
	DepthData& dd = arrDepthData[idxImage];
	if (dd.neighbors.empty()) { lastError = "InitViews before SelectViews"; return false; }
	dd.images.assign(1, idxImage);
	const float fMinScore = std::max(dd.neighbors.front().score*(VS.fViewMinScoreRatio*0.1f), VS.fViewMinScore);
	for (const ViewScore& nb: dd.neighbors) {
		if ((numNeighbors && dd.images.size() > numNeighbors) || nb.score < fMinScore) break;
		dd.images.push_back(nb.ID);
	}
	if (dd.images.size() < 2) { dd.images.clear(); return false; }
	if (dd.images.size()-1 > HCMVS_MAX_MATCH_VIEWS) dd.images.resize(HCMVS_MAX_MATCH_VIEWS+1);
	for (uint32_t id: dd.images) if (!UploadView(id)) return false;
	std::vector<uint32_t> ids; std::vector<float> scores;
	for (const ViewScore& nb: dd.neighbors) { ids.push_back(nb.ID); scores.push_back(nb.score); }
	// matching views are the first entries of the (sorted) neighbour list
	if (hcmvs_set_neighbors(ctx, idxImage, ids.data(), scores.data(), (int)dd.images.size()-1, (int)ids.size()) != HCMVS_OK) return Fail("hcmvs_set_neighbors");
	if (hcmvs_set_fuse_priority(ctx, idxImage, (float)scene.images[idxImage].neighbors.size()) != HCMVS_OK) return Fail("hcmvs_set_fuse_priority");
	// ViewData::ScaleImage (SceneDensify.cpp:370-376): a matching view whose footprint differs by |scale-1| >= 0.15 is matched against a
	// resized copy of its image, with the camera of the new resolution
	for (size_t k=1; k<dd.images.size(); ++k) {
		const ViewScore& nb = dd.neighbors[k-1];
		Image& im = scene.images[nb.ID];
		std::vector<float> scaled; int sw = 0, sh = 0;
		if (im.gray.empty() && std::abs(nb.scale-1.f) >= 0.15f) { lastError = "a rescaled matching view needs its pixels in host memory (multi-GPU: hold that image on every rank)"; return false; }
		if (!ScaleImage(im.gray, im.width, im.height, nb.scale, scaled, sw, sh)) continue;
		double Ks[9]; ScaleK(im.camera.K, im.width, im.height, sw, sh, Ks);
		if (hcmvs_set_neighbor_image(ctx, idxImage, (int)k-1, sw, sh, Ks, scaled.data()) != HCMVS_OK) return Fail("hcmvs_set_neighbor_image");
	}
	return true;
}

void SparseInitDepth(const Scene& scene, uint32_t idxImage, const std::vector<uint32_t>& points, std::vector<float>& depth, float& dMin, float& dMax) {
	// it_external == 0 block of EstimateDepthMap, nMinViewsTrustPoint < 2 branch (SceneDensify.cpp:783-808): splat the
	// sparse points into 5x5 windows and derive the depth range. (The CGAL Delaunay initialisation of
	// InitDepthMap, DepthMap.cpp:1879-1936, is out of scope — SURVEY §8a P8.)
	const Image& im = scene.images[idxImage];
	const Camera& cam = im.camera;
	const int w = im.width, h = im.height;
	depth.assign((size_t)w*h, 0.f);
	dMin = FLT_MAX; dMax = 0;
	for (uint32_t ip: points) {
		const float* X = &scene.pointcloud.xyz[(size_t)ip*3];
		const double d0 = (double)X[0]-cam.C[0], d1 = (double)X[1]-cam.C[1], d2 = (double)X[2]-cam.C[2];
		double c[3];
		for (int r=0; r<3; ++r) { double a = cam.R[r*3]*d0; a += cam.R[r*3+1]*d1; a += cam.R[r*3+2]*d2; c[r] = a; } // TransformPointW2C
		const double u = cam.K[2]+cam.K[0]*(c[0]/c[2]), v = cam.K[5]+cam.K[4]*(c[1]/c[2]); // TransformPointC2I
		const int x = (int)std::floor(u+.5), y = (int)std::floor(v+.5); // ROUND2INT
		const float d = (float)c[2];
		const int sx = std::max(x-2, 0), sy = std::max(y-2, 0), ex = std::min(x+2, w-1), ey = std::min(y+2, h-1);
		for (int yy=sy; yy<=ey; ++yy) for (int xx=sx; xx<=ex; ++xx) depth[(size_t)yy*w+xx] = d;
		dMin = std::min(dMin, d); dMax = std::max(dMax, d);
	}
	dMin *= 0.9f; dMax *= 1.1f;
}

bool DepthMapsData::InitDepthMap(uint32_t idxImage) {
	DepthData& dd = arrDepthData[idxImage];
	if (P.nMinViewsTrustPoint >= 2) {
		// DepthMapsData::InitDepthMap, SceneDensify.cpp:514-525 (initTriangulate; bAddCorners is the reference's default)
		std::vector<double> vertices; std::vector<uint32_t> tris;
		if (!TriangulateInit(scene, idxImage, dd.points, true, vertices, tris, dd.dMin, dd.dMax)) { lastError = "cannot triangulate the sparse points of the view"; return false; }
		dd.dMin *= 0.9f; dd.dMax *= 1.1f;
		if (hcmvs_init_depthmap_triangles(ctx, idxImage, vertices.data(), (int)(vertices.size()/3), tris.data(), (int)(tris.size()/3), dd.dMin, dd.dMax) != HCMVS_OK) return Fail("hcmvs_init_depthmap_triangles");
		return true;
	}
	std::vector<float> depth;
	SparseInitDepth(scene, idxImage, dd.points, depth, dd.dMin, dd.dMax);
	if (hcmvs_init_depthmap(ctx, idxImage, depth.data(), nullptr, dd.dMin, dd.dMax) != HCMVS_OK) return Fail("hcmvs_init_depthmap");
	return true;
}

bool DepthMapsData::EstimateDepthMap(int it_external, uint32_t idxImage, uint64_t seed) {
	if (it_external == 0 && !InitDepthMap(idxImage)) return false;
	if (hcmvs_estimate_depthmap(ctx, idxImage, it_external, seed) != HCMVS_OK) return Fail("hcmvs_estimate_depthmap");
	return true;
}

bool DepthMapsData::FilterDepthMap(uint32_t idxImage, const std::vector<uint32_t>& idxNeighbors, bool bAdjust) {
	if (hcmvs_filter_depthmap(ctx, idxImage, idxNeighbors.data(), (int)idxNeighbors.size(), bAdjust ? 1 : 0, nullptr, nullptr) != HCMVS_OK) return Fail("hcmvs_filter_depthmap");
	return true;
}

bool DepthMapsData::RemoveSmallSegments(uint32_t idxImage, unsigned nSpeckleSize) {
	if (hcmvs_remove_small_segments(ctx, idxImage, nSpeckleSize, nullptr) != HCMVS_OK) return Fail("hcmvs_remove_small_segments");
	return true;
}

bool DepthMapsData::GapInterpolation(uint32_t idxImage, std::vector<float>& depthFuse, std::vector<float>& normalFuse, std::vector<float>& conf, unsigned nIpolGapSize) {
	const Image& im = scene.images[idxImage];
	const size_t n = (size_t)im.width*im.height;
	if (depthFuse.size() != n || (!normalFuse.empty() && normalFuse.size() != n*3) || (!conf.empty() && conf.size() != n)) { lastError = "GapInterpolation: map sizes do not match the image"; return false; }
	if (hcmvs_gap_interpolation(ctx, im.width, im.height, depthFuse.data(), normalFuse.empty() ? nullptr : normalFuse.data(), conf.empty() ? nullptr : conf.data(), nIpolGapSize, nullptr) != HCMVS_OK)
		return Fail("hcmvs_gap_interpolation");
	return true;
}

bool DepthMapsData::FuseDepthMaps(PointCloud& pc, bool bEstimateColor, bool bEstimateNormal) {
	// fuse on the device, then ONE copy into the context's page-locked arena; the PointCloud borrows those arrays
	// (valid until the next FuseDepthMaps on this context or hcmvs_destroy — see hcmvs_download_fused_pinned)
	if (hcmvs_fuse_depthmaps(ctx, bEstimateColor, bEstimateNormal, nullptr) != HCMVS_OK) return Fail("hcmvs_fuse_depthmaps");
	uint64_t n = 0, m = 0;
	if (hcmvs_get_fused_device(ctx, &n, &m, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr) != HCMVS_OK) return Fail("hcmvs_get_fused_device");
	pc = PointCloud();
	if (!n) return true;
	hcmvs_pointcloud host;
	if (hcmvs_download_fused_pinned(ctx, &host) != HCMVS_OK) return Fail("hcmvs_download_fused_pinned");
	pc.points.borrow(host.points, n*3); pc.viewOffsets.borrow(host.view_offsets, n+1); pc.views.borrow(host.views, m); pc.weights.borrow(host.weights, m);
	if (host.normals) pc.normals.borrow(host.normals, n*3);
	if (host.colors) pc.colors.borrow(host.colors, n*3);
	return true;
}

bool DepthMapsData::SaveDepthMapRaw(uint32_t idxImage, const std::string& fileName) {
	const Image& im = scene.images[idxImage];
	const size_t n = (size_t)im.width*im.height;
	std::vector<float> d(n), nrm(n*3), c(n); float dMin, dMax;
	if (hcmvs_get_depthmap(ctx, idxImage, d.data(), nrm.data(), c.data(), &dMin, &dMax) != HCMVS_OK) return Fail("hcmvs_get_depthmap");
	return ExportDepthDataRaw(fileName, im.name, arrDepthData[idxImage].images, im.width, im.height, im.camera.K, im.camera.R, im.camera.C,
		dMin, dMax, im.width, im.height, d.data(), nrm.data(), c.data());
}

// ------------------------------------------------------------------------------------------------ file formats
#pragma pack(push, 1)
struct HeaderDepthDataRaw { // libs/MVS/Interface.h:634-652 (28 bytes)
	uint16_t name; uint8_t type, padding; uint32_t imageWidth, imageHeight, depthWidth, depthHeight; float dMin, dMax;
};
#pragma pack(pop)
static_assert(sizeof(HeaderDepthDataRaw) == 28, "HeaderDepthDataRaw layout");

bool ExportDepthDataRaw(const std::string& fileName, const std::string& imageFileName, const std::vector<uint32_t>& IDs, int imageW, int imageH,
	const double K[9], const double R[9], const double C[3], float dMin, float dMax, int w, int h,
	const float* depth, const float* normal, const float* conf)
{
	// ExportDepthDataRaw, libs/MVS/DepthMap.cpp:2781-2846
	FILE* f = fopen(fileName.c_str(), "wb");
	if (!f) return false;
	HeaderDepthDataRaw hd;
	hd.name = 0x5244; // "DR" little endian
	hd.type = 1 | (normal ? 2 : 0) | (conf ? 4 : 0); hd.padding = 0;
	hd.imageWidth = (uint32_t)imageW; hd.imageHeight = (uint32_t)imageH; hd.depthWidth = (uint32_t)w; hd.depthHeight = (uint32_t)h;
	hd.dMin = dMin; hd.dMax = dMax;
	fwrite(&hd, sizeof(hd), 1, f);
	const uint16_t nName = (uint16_t)imageFileName.size();
	fwrite(&nName, 2, 1, f); fwrite(imageFileName.data(), 1, nName, f);
	const uint32_t nIDs = (uint32_t)IDs.size();
	fwrite(&nIDs, 4, 1, f); fwrite(IDs.data(), 4, nIDs, f);
	fwrite(K, 8, 9, f); fwrite(R, 8, 9, f); fwrite(C, 8, 3, f);
	const size_t n = (size_t)w*h;
	fwrite(depth, 4, n, f);
	if (normal) fwrite(normal, 12, n, f);
	if (conf) fwrite(conf, 4, n, f);
	const bool ok = ferror(f) == 0;
	fclose(f);
	return ok;
}

bool ImportDepthDataRaw(const std::string& fileName, std::string& imageFileName, std::vector<uint32_t>& IDs, int& imageW, int& imageH,
	double K[9], double R[9], double C[3], float& dMin, float& dMax, int& w, int& h,
	std::vector<float>& depth, std::vector<float>& normal, std::vector<float>& conf)
{
	// ImportDepthDataRaw, libs/MVS/DepthMap.cpp:2848-2925
	FILE* f = fopen(fileName.c_str(), "rb");
	if (!f) return false;
	HeaderDepthDataRaw hd;
	if (fread(&hd, sizeof(hd), 1, f) != 1 || hd.name != 0x5244 || (hd.type & 1) == 0 || hd.depthWidth == 0 || hd.depthHeight == 0 ||
	    hd.imageWidth < hd.depthWidth || hd.imageHeight < hd.depthHeight) { fclose(f); return false; }
	uint16_t nName = 0; bool ok = fread(&nName, 2, 1, f) == 1;
	imageFileName.resize(nName); if (nName) ok = ok && fread(&imageFileName[0], 1, nName, f) == nName;
	uint32_t nIDs = 0; ok = ok && fread(&nIDs, 4, 1, f) == 1;
	if (!ok || nIDs > (1u<<20)) { fclose(f); return false; }
	IDs.resize(nIDs); if (nIDs) ok = ok && fread(IDs.data(), 4, nIDs, f) == nIDs;
	ok = ok && fread(K, 8, 9, f) == 9 && fread(R, 8, 9, f) == 9 && fread(C, 8, 3, f) == 3;
	dMin = hd.dMin; dMax = hd.dMax; imageW = (int)hd.imageWidth; imageH = (int)hd.imageHeight; w = (int)hd.depthWidth; h = (int)hd.depthHeight;
	const size_t n = (size_t)w*h;
	depth.resize(n); ok = ok && fread(depth.data(), 4, n, f) == n;
	normal.clear(); conf.clear();
	if (hd.type & 2) { normal.resize(n*3); ok = ok && fread(normal.data(), 12, n, f) == n; }
	if (hd.type & 4) { conf.resize(n); ok = ok && fread(conf.data(), 4, n, f) == n; }
	fclose(f);
	return ok;
}

bool PointCloud::Save(const std::string& fileName) const {
	// PointCloud::Save, libs/MVS/PointCloud.cpp:188-242 with BasicPLY::vert_props (:105-131); header text as
	// PLY::header_complete writes it (libs/IO/PLY.cpp:269-338, new-style type names)
	const size_t n = size();
	if (!n) return false;
	FILE* f = fopen(fileName.c_str(), "wb");
	if (!f) return false;
	const bool hasN = !normals.empty();
	fprintf(f, "ply\nformat binary_little_endian 1.0\nelement vertex %d\n", (int)n);
	fprintf(f, "property float32 x\nproperty float32 y\nproperty float32 z\nproperty uint8 red\nproperty uint8 green\nproperty uint8 blue\n");
	if (hasN) fprintf(f, "property float32 nx\nproperty float32 ny\nproperty float32 nz\n");
	fprintf(f, "end_header\n");
	// records are assembled a chunk at a time and written with one fwrite per chunk (a 19.6 M-point C2 cloud is 530 MB: one call per
	// 27-byte record would make the writer, not the GPU, the slowest stage of the run)
	const size_t recBytes = hasN ? 27 : 15, chunk = (size_t)1<<18;
	std::vector<uint8_t> buf(recBytes*std::min(chunk, n));
	const bool hasC = !colors.empty();
	for (size_t i0=0; i0<n; i0+=chunk) {
		const size_t m = std::min(chunk, n-i0);
		uint8_t* rec = buf.data();
		for (size_t i=i0; i<i0+m; ++i, rec+=recBytes) {
			memcpy(rec, &points[i*3], 12);
			if (hasC) { rec[12] = colors[i*3+2]; rec[13] = colors[i*3+1]; rec[14] = colors[i*3]; } // stored b,g,r
			else rec[12] = rec[13] = rec[14] = 255; // Color::WHITE
			if (hasN) memcpy(rec+15, &normals[i*3], 12);
		}
		fwrite(buf.data(), recBytes, m, f);
	}
	const bool ok = ferror(f) == 0;
	fclose(f);
	return ok;
}

// ------------------------------------------------------------------------------------------------ point-cloud filter
long Scene::PointCloudFilter(hcmvs_ctx* ctx, int thRemove, std::string* err) {
	PointCloud& pc = densecloud;
	const size_t n = pc.size();
	if (!n) return 0;
	std::vector<int32_t> vis(n);
	if (hcmvs_pointcloud_filter(ctx, n, pc.points.data(), pc.viewOffsets.data(), pc.views.data(), vis.data(), nullptr) != HCMVS_OK) { if (err) *err = std::string("hcmvs_pointcloud_filter: ")+hcmvs_last_error(); return -1; }
	return RemovePointsByVisibility(pc, vis.data(), thRemove);
}

long RemovePointsByVisibility(PointCloud& pc, const int32_t* vis, int thRemove) {
	// RFOREACH(idxPoint) if (visibility[idxPoint] <= thRemove) pointcloud.RemovePoint(idxPoint) — SceneDensify.cpp:4310-4314 with
	// PointCloud::RemovePoint (PointCloud.cpp:54-69) = cList::RemoveAt: the LAST element takes the removed one's place
	const size_t n = pc.size();
	std::vector<uint32_t> order(n);
	for (size_t i=0; i<n; ++i) order[i] = (uint32_t)i;
	size_t size = n;
	for (size_t i=n; i-- > 0;) if (vis[i] <= thRemove) { if (i+1 != size) order[i] = order[size-1]; --size; }
	const long removed = (long)(n-size);
	if (!removed) return 0;
	const bool hasN = !pc.normals.empty(), hasC = !pc.colors.empty(), hasW = !pc.weights.empty();
	std::vector<float> P(size*3), N(hasN ? size*3 : 0), W; std::vector<uint8_t> C(hasC ? size*3 : 0); std::vector<uint32_t> off(size+1), ids;
	off[0] = 0;
	for (size_t k=0; k<size; ++k) { const uint32_t i = order[k]; off[k+1] = off[k]+(pc.viewOffsets[i+1]-pc.viewOffsets[i]); }
	ids.resize(off[size]); if (hasW) W.resize(off[size]);
	for (size_t k=0; k<size; ++k) {
		const uint32_t i = order[k];
		memcpy(&P[k*3], &pc.points[i*3], 12);
		if (hasN) memcpy(&N[k*3], &pc.normals[i*3], 12);
		if (hasC) memcpy(&C[k*3], &pc.colors[i*3], 3);
		const uint32_t a = pc.viewOffsets[i], cnt = pc.viewOffsets[i+1]-a;
		if (cnt) memcpy(&ids[off[k]], &pc.views[a], (size_t)cnt*4);
		if (hasW && cnt) memcpy(&W[off[k]], &pc.weights[a], (size_t)cnt*4);
	}
	auto put = [](auto& arr, const auto& v) { arr.resize(v.size()); if (!v.empty()) memcpy(arr.data(), v.data(), v.size()*sizeof(v[0])); };
	put(pc.points, P); put(pc.viewOffsets, off); put(pc.views, ids);
	if (hasN) put(pc.normals, N); if (hasC) put(pc.colors, C); if (hasW) put(pc.weights, W);
	return removed;
}

// ------------------------------------------------------------------------------------------------ driver
namespace {
// Raw depth-data files written behind the GPU: the maps of a view are read back into one of the context's page-locked slots
// (hcmvs_download_depthmap_begin, queued right behind the view's last estimation) while the next views are estimated, and a writer
// thread turns every landed slot into depthNNNN.dmap (ExportDepthDataRaw). The reference writes each map from the thread that
// estimated it (EVT_SAVEDEPTHMAP, SceneDensify.cpp:3960-3999), which is the same overlap on a CPU.
struct DmapWriter {
	hcmvs_ctx* ctx; const Scene& scene; const std::vector<DepthData>& dd; std::string dir;
	std::thread th; std::mutex m; std::condition_variable cv;
	std::deque<std::pair<int, uint32_t>> work; std::vector<int> freeSlots;
	bool closing = false; std::string err; uint64_t bytes = 0;
	DmapWriter(hcmvs_ctx* c, const Scene& s, const std::vector<DepthData>& d, const std::string& dr) : ctx(c), scene(s), dd(d), dir(dr) {
		for (int k=0; k<HCMVS_DOWNLOAD_SLOTS; ++k) freeSlots.push_back(k);
		th = std::thread([this]() { Run(); });
	}
	~DmapWriter() { Finish(); }
	bool Submit(uint32_t view) { // main thread: take a free slot (waits for the writer when all are in flight) and queue the read-back
		int slot;
		{ std::unique_lock<std::mutex> lk(m); cv.wait(lk, [this]() { return !freeSlots.empty() || !err.empty(); }); if (!err.empty()) return false; slot = freeSlots.back(); freeSlots.pop_back(); }
		if (hcmvs_download_depthmap_begin(ctx, view, slot) != HCMVS_OK) { std::lock_guard<std::mutex> lk(m); err = std::string("hcmvs_download_depthmap_begin: ")+hcmvs_last_error(); return false; }
		{ std::lock_guard<std::mutex> lk(m); work.emplace_back(slot, view); }
		cv.notify_all();
		return true;
	}
	void Run() {
		for (;;) {
			std::pair<int, uint32_t> job;
			{ std::unique_lock<std::mutex> lk(m); cv.wait(lk, [this]() { return !work.empty() || closing; }); if (work.empty()) return; job = work.front(); work.pop_front(); }
			const float *d = nullptr, *n = nullptr, *c = nullptr; float lo = 0, hi = 0;
			std::string e;
			if (hcmvs_download_depthmap_wait(ctx, job.first, &d, &n, &c, &lo, &hi) != HCMVS_OK) e = std::string("hcmvs_download_depthmap_wait: ")+hcmvs_last_error();
			else {
				const Image& im = scene.images[job.second];
				char name[64]; snprintf(name, sizeof(name), "/depth%04u.dmap", job.second);
				if (!ExportDepthDataRaw(dir+name, im.name, dd[job.second].images, im.width, im.height, im.camera.K, im.camera.R, im.camera.C, lo, hi, im.width, im.height, d, n, c))
					e = "cannot write "+dir+name;
				else bytes += (uint64_t)im.width*im.height*20;
			}
			{ std::lock_guard<std::mutex> lk(m); freeSlots.push_back(job.first); if (!e.empty() && err.empty()) err = e; }
			cv.notify_all();
		}
	}
	bool Finish() {
		if (th.joinable()) { { std::lock_guard<std::mutex> lk(m); closing = true; } cv.notify_all(); th.join(); }
		return err.empty();
	}
};
} // namespace

bool DenseReconstruction(Scene& scene, hcmvs_ctx* ctx, const hcmvs_params& P, const ViewSelectionParams& VS, uint64_t seed, bool runFilter,
	const std::string& dmapDir, DenseReconstructionStats* stats, std::string* err)
{
	// Scene::DenseReconstruction -> ComputeDepthMaps (SceneDensify.cpp:3532-3730) -> FuseDepthMaps (:3544)
	DenseReconstructionStats st;
	DepthMapsData data(scene, ctx, P, VS);
	auto fail = [&](const std::string& m) { if (err) *err = m; return false; };
	double t0 = Now();
	const uint32_t nImages = (uint32_t)scene.images.size();
	for (Image& im: scene.images) im.camera.ComposeP();
	// a reused context: initial-map uploads of later views only wait for their own view's previous use, not for the running estimation
	if (hcmvs_begin_scene(ctx) != HCMVS_OK) return fail(std::string("hcmvs_begin_scene: ")+hcmvs_last_error());
	// Host workers select the neighbour views (the reference does it with `#pragma omp parallel for`, :3652-3667 — each call only
	// writes image i's state) and splat the sparse points into the initial depth map (:783-808), in image order, while this
	// thread uploads the images and then feeds the GPU: view i is estimated as soon as ITS selection is done, so the host work
	// of the later views hides behind the kernels of the earlier ones. All hcmvs_* calls stay on this thread.
	struct Prepared { std::vector<float> depth; std::vector<double> vertices; std::vector<uint32_t> tris; float dMin = 0, dMax = 0; bool ok = true; };
	const bool triangulate = P.nMinViewsTrustPoint >= 2; // the reference's default initialisation (SceneDensify.cpp:781-812)
	std::vector<Prepared> prep(nImages);
	std::vector<std::atomic<int>> state(nImages); // 0 pending, 1 selected + initial depth ready, -1 rejected
	for (auto& s: state) s.store(0);
	std::atomic<uint32_t> next{0}, consumed{0};
	std::atomic<double> tSelectEnd{t0};
	const uint32_t lookahead = 24; // bounds the host memory held by prepared depth maps
	// one core stays with this thread: it feeds the GPU, and a worker that spins on it would let the launch queue run dry
	const unsigned hc = std::thread::hardware_concurrency();
	const unsigned nt = std::max(1u, std::min(hc > 1 ? hc-1 : 1u, nImages));
	std::vector<std::thread> pool;
	auto prepareOne = [&](uint32_t i, unsigned threads) {
		const bool ok = data.SelectViews(i, threads);
		if (ok && triangulate) {
			prep[i].ok = TriangulateInit(scene, i, data.arrDepthData[i].points, true, prep[i].vertices, prep[i].tris, prep[i].dMin, prep[i].dMax);
			prep[i].dMin *= 0.9f; prep[i].dMax *= 1.1f;
		} else if (ok) SparseInitDepth(scene, i, data.arrDepthData[i].points, prep[i].depth, prep[i].dMin, prep[i].dMax);
		tSelectEnd.store(Now());
		state[i].store(ok ? 1 : -1, std::memory_order_release);
	};
	// the first view is what the GPU waits for: every core works INSIDE its selection (bit-identical to the one-thread form), then the
	// workers take one view each as the reference does
	if (nImages > 0) { prepareOne(0, std::max(1u, hc)); next.store(1); }
	for (unsigned t=0; t<nt; ++t) pool.emplace_back([&]() {
		uint32_t i;
		while ((i = next.fetch_add(1)) < nImages) {
			while (i >= consumed.load(std::memory_order_acquire)+lookahead) std::this_thread::sleep_for(std::chrono::microseconds(200)); // throttled: sleep, do not spin
			prepareOne(i, 1);
		}
	});
	struct Joiner { std::vector<std::thread>& p; std::atomic<uint32_t>& c; ~Joiner() { c.store(0x7fffffffu); for (std::thread& th: p) if (th.joinable()) th.join(); } } joiner{pool, consumed};
	// images are uploaded on first use by InitViews (copy stream), so the uploads of later views overlap the kernels of earlier ones
	double t1 = Now();
	std::unique_ptr<DmapWriter> writer;
	if (!dmapDir.empty()) writer.reset(new DmapWriter(ctx, scene, data.arrDepthData, dmapDir));
	const bool singleOuter = P.nEstimationIters_external <= 1; // the maps are final (and saved, :3984) after the LAST outer iteration
	std::vector<uint32_t> valid;
	const bool hostDebug = getenv("HCMVS_HOST_DEBUG") != nullptr;
	double waitedSum = 0;
	for (uint32_t i=0; i<nImages; ++i) {
		int s;
		const double tw = Now();
		while ((s = state[i].load(std::memory_order_acquire)) == 0) std::this_thread::yield();
		const double waited = Now()-tw; waitedSum += waited;
		if (s > 0) {
			valid.push_back(i);
			DepthData& dd = data.arrDepthData[i];
			if (!data.InitViews(i, P.nNumViews)) { if (!data.lastError.empty()) return fail(data.lastError); dd.valid = false; }
			else {
				dd.dMin = prep[i].dMin; dd.dMax = prep[i].dMax;
				if (triangulate) {
					if (!prep[i].ok) return fail("cannot triangulate the sparse points of a view");
					if (hcmvs_init_depthmap_triangles(ctx, i, prep[i].vertices.data(), (int)(prep[i].vertices.size()/3), prep[i].tris.data(), (int)(prep[i].tris.size()/3), dd.dMin, dd.dMax) != HCMVS_OK)
						return fail(std::string("hcmvs_init_depthmap_triangles: ")+hcmvs_last_error());
					st.h2dBytes += (uint64_t)prep[i].vertices.size()*8+(uint64_t)prep[i].tris.size()*4;
				} else {
					if (hcmvs_init_depthmap(ctx, i, prep[i].depth.data(), nullptr, dd.dMin, dd.dMax) != HCMVS_OK) return fail(std::string("hcmvs_init_depthmap: ")+hcmvs_last_error());
					st.h2dBytes += (uint64_t)scene.images[i].width*scene.images[i].height*4;
				}
				if (hcmvs_estimate_depthmap(ctx, i, 0, seed) != HCMVS_OK) return fail(std::string("hcmvs_estimate_depthmap: ")+hcmvs_last_error());
				if (hostDebug && valid.size() <= 3) fprintf(stderr, "[host] view %u enqueued at %.1f ms (waited %.1f ms for its selection)\n", i, (Now()-t0)*1e3, waited*1e3);
				if (writer && singleOuter && !writer->Submit(i)) return fail(writer->err);
			}
		}
		std::vector<float>().swap(prep[i].depth); std::vector<double>().swap(prep[i].vertices); std::vector<uint32_t>().swap(prep[i].tris);
		consumed.store(i+1, std::memory_order_release);
	}
	for (uint32_t i=0; i<nImages; ++i) {
		const Image& im = scene.images[i];
		if (data.arrDepthData[i].uploaded) st.h2dBytes += (uint64_t)im.width*im.height*(4+(im.bgr.empty() ? 0 : 3));
	}
	st.secSelect = tSelectEnd.load()-t0; // overlaps the uploads and the estimation
	st.secUpload = 0; // image uploads ride the copy stream inside the estimation loop
	if (valid.empty()) return fail("no image has enough neighbour views");
	for (unsigned it=1; it<P.nEstimationIters_external; ++it) { // :3684
		// viewspread reads the neighbours' maps of the previous outer iteration (hcmvs_snapshot_maps)
		if (P.viewspread && hcmvs_snapshot_maps(ctx) != HCMVS_OK) return fail(std::string("hcmvs_snapshot_maps: ")+hcmvs_last_error());
		for (uint32_t i: valid) {
			if (!data.arrDepthData[i].valid) continue;
			if (!data.EstimateDepthMap((int)it, i, seed)) return fail(data.lastError);
			if (writer && it+1 == P.nEstimationIters_external && !writer->Submit(i)) return fail(writer->err);
		}
	}
	const double tEnq = Now();
	if (hcmvs_sync(ctx) != HCMVS_OK) return fail(hcmvs_last_error());
	double t3 = Now(); st.secEstimate = t3-t1;
	if (hostDebug) fprintf(stderr, "[host] all views enqueued at %.1f ms (waited %.1f ms for selections in total), GPU done at %.1f ms\n", (tEnq-t0)*1e3, waitedSum*1e3, (t3-t0)*1e3);
	// (the last read-backs / .dmap files finish behind the filter and fusion kernels; the writer is joined before returning)
	if (runFilter) {
		// Scene::DenseReconstructionFilter, SceneDensify.cpp:4093-4185: neighbours = those with a depth map, at most 8
		for (uint32_t i: valid) {
			const DepthData& dd = data.arrDepthData[i];
			if (!dd.valid) continue;
			std::vector<uint32_t> idxNb;
			for (uint32_t k=0; k<dd.neighbors.size() && idxNb.size() < 8; ++k) if (data.arrDepthData[dd.neighbors[k].ID].valid) idxNb.push_back(k);
			if (idxNb.size() < std::min(P.nMinViewsFilter, scene.nCalibratedImages()-1)) continue;
			if (!data.FilterDepthMap(i, idxNb, P.bFilterAdjust != 0)) return fail(data.lastError);
		}
		if (hcmvs_commit_filtered(ctx) != HCMVS_OK) return fail(hcmvs_last_error());
		if (hcmvs_sync(ctx) != HCMVS_OK) return fail(hcmvs_last_error());
	}
	double t4 = Now(); st.secFilter = t4-t3;
	if (!data.FuseDepthMaps(scene.densecloud, true, true)) return fail(data.lastError);
	st.d2hBytes += (uint64_t)scene.densecloud.size()*(12+12+3+4)+(uint64_t)scene.densecloud.views.size()*8;
	st.secFuse = Now()-t4;
	if (writer) { if (!writer->Finish()) return fail(writer->err); st.d2hBytes += writer->bytes; }
	if (stats) *stats = st;
	return true;
}

} // namespace hcmvs_host
