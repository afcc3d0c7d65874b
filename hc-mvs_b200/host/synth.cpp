// Synthetic MVS scene generator — see include/hcmvs_synth.h.
// Cameras follow the reference convention P = K R [I|-C] (libs/MVS/Camera.h:46-54): X_cam = R (X - C),
// camera x right, y down, z forward. Surfaces are height fields z = f(x,y) seen from above.
#include "hcmvs_synth.h"
#include <cmath>
#include <cstring>
#include <vector>
#include <thread>
#include <algorithm>

namespace {

struct SplitMix64 {
	uint64_t s;
	explicit SplitMix64(uint64_t seed) : s(seed) {}
	uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30))*0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27))*0x94D049BB133111EBull; return z ^ (z >> 31); }
	double uniform() { return (double)(next() >> 11)*(1.0/9007199254740992.0); }
	double uniform(double a, double b) { return a + (b-a)*uniform(); }
};
inline uint64_t Mix64(uint64_t z) { z = (z ^ (z >> 30))*0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27))*0x94D049BB133111EBull; return z ^ (z >> 31); }

struct Bump { double x, y, s2inv, h; };
struct Cam { double K[9], R[9], C[3]; };

} // namespace

struct hcmvs_synth_scene {
	hcmvs_synth_cfg cfg;
	std::vector<Bump> bumps;
	std::vector<Cam> cams;
	std::vector<float> sparse;               // xyz
	std::vector<int32_t> sparseOff;
	std::vector<uint32_t> sparseViews;
	double zmin, zmax;

	double Height(double x, double y, double* gx = nullptr, double* gy = nullptr) const {
		double z, dx, dy;
		if (cfg.surface == 2) {
			const double k = std::tan(40.0*M_PI/180.0);
			// wedge with a tiny smoothing so the crease has a well defined normal
			const double e = 1e-3*cfg.extent_x;
			const double r = std::sqrt(x*x + e*e);
			z = -k*(r-e); dx = -k*x/r; dy = 0;
			z += cfg.plane_b*y; dy += cfg.plane_b;
		} else {
			z = cfg.plane_a*x + cfg.plane_b*y; dx = cfg.plane_a; dy = cfg.plane_b;
		}
		for (const Bump& b: bumps) {
			const double ux = x-b.x, uy = y-b.y;
			const double q = (ux*ux+uy*uy)*b.s2inv;
			if (q > 18.0) continue; // < 1.6e-8 of the bump height
			const double g = b.h*std::exp(-q);
			z += g; dx += -2.0*ux*b.s2inv*g; dy += -2.0*uy*b.s2inv*g;
		}
		if (gx) { *gx = dx; *gy = dy; }
		return z;
	}

	// value-noise texture in [0,1]
	static inline double Lattice(uint64_t seed, int64_t ix, int64_t iy, int oct) {
		const uint64_t h = Mix64(seed ^ ((uint64_t)ix*0x9E3779B97F4A7C15ull) ^ Mix64((uint64_t)iy*0xC2B2AE3D27D4EB4Full + (uint64_t)oct*0x165667B19E3779F9ull));
		return (double)(h >> 40)*(1.0/16777216.0);
	}
	double Texture(double x, double y, int channel) const {
		const uint64_t seed = cfg.seed*0x2545F4914F6CDD1Dull + 77u;
		x += 19.37*channel*cfg.tex_wavelength; y -= 7.91*channel*cfg.tex_wavelength;
		double sum = 0, norm = 0, amp = 1, lam = cfg.tex_wavelength;
		for (int o=0; o<5; ++o) {
			const double u = x/lam, v = y/lam;
			const double fu = std::floor(u), fv = std::floor(v);
			const int64_t iu = (int64_t)fu, iv = (int64_t)fv;
			double tu = u-fu, tv = v-fv;
			tu = tu*tu*tu*(tu*(tu*6-15)+10); tv = tv*tv*tv*(tv*(tv*6-15)+10);
			const double a = Lattice(seed, iu, iv, o), b = Lattice(seed, iu+1, iv, o);
			const double c = Lattice(seed, iu, iv+1, o), d = Lattice(seed, iu+1, iv+1, o);
			sum += amp*((a+(b-a)*tu)*(1-tv) + (c+(d-c)*tu)*tv);
			norm += amp; amp *= 0.62; lam *= 0.5;
		}
		const double n = 0.5 + (sum/norm-0.5)*2.4;
		return std::min(std::max(n, 0.0), 1.0);
	}

	// ray / surface intersection: returns ray parameter t with X = o + t*d (d not normalised), or <0
	double Cast(const double o[3], const double d[3], double tGuess) const {
		double t = tGuess;
		if (!(t > 0)) {
			// intersect the mean plane z = a x + b y
			const double den = d[2]-cfg.plane_a*d[0]-cfg.plane_b*d[1];
			if (std::abs(den) < 1e-12) return -1;
			t = (cfg.plane_a*o[0]+cfg.plane_b*o[1]-o[2])/den;
			if (!(t > 0)) return -1;
		}
		for (int it=0; it<40; ++it) {
			double gx, gy;
			const double x = o[0]+t*d[0], y = o[1]+t*d[1];
			const double f = Height(x, y, &gx, &gy);
			const double g = o[2]+t*d[2]-f;
			const double dg = d[2]-gx*d[0]-gy*d[1];
			if (std::abs(dg) < 1e-12) return -1;
			const double dt = g/dg;
			t -= dt;
			if (std::abs(dt) <= 1e-12*std::abs(t)) break;
		}
		return t > 0 ? t : -1;
	}
};

static void LookAt(const double C[3], const double target[3], const double up[3], double R[9]) {
	double f[3] = {target[0]-C[0], target[1]-C[1], target[2]-C[2]};
	double n = std::sqrt(f[0]*f[0]+f[1]*f[1]+f[2]*f[2]); for (double& v: f) v /= n;
	double r[3] = {f[1]*up[2]-f[2]*up[1], f[2]*up[0]-f[0]*up[2], f[0]*up[1]-f[1]*up[0]}; // f x up
	n = std::sqrt(r[0]*r[0]+r[1]*r[1]+r[2]*r[2]); for (double& v: r) v /= n;
	double dn[3] = {f[1]*r[2]-f[2]*r[1], f[2]*r[0]-f[0]*r[2], f[0]*r[1]-f[1]*r[0]}; // f x r
	for (int i=0; i<3; ++i) { R[i] = r[i]; R[3+i] = dn[i]; R[6+i] = f[i]; }
}

extern "C" int hcmvs_synth_preset(int config, double scale, int n_views_override, hcmvs_synth_cfg* c) {
	if (!c || !(scale > 0) || scale > 1.0) return -1;
	std::memset(c, 0, sizeof(*c));
	double dist = 1;
	switch (config) {
	case 1: // 10 x 640x480 slanted plane, ring of cameras
		c->n_views = 10; c->width = 640; c->height = 480; c->focal = 600; c->cx = 319.5; c->cy = 239.5;
		c->surface = 0; c->plane_a = 0.05; c->plane_b = 0.03; c->n_bumps = 0;
		c->layout = 0; c->cam_distance = 6.0; c->cam_radius = 1.5; c->cam_step_deg = 36;
		c->extent_x = 5.5; c->extent_y = 4.5; c->n_sparse = 2000; c->seed = 1001; dist = 6.0;
		break;
	case 2: // DTU-shaped: 49 x 1600x1200, 7x7 spherical cap
		c->n_views = 49; c->width = 1600; c->height = 1200; c->focal = 2892.3; c->cx = 823.2; c->cy = 619.1;
		c->surface = 1; c->plane_a = 0.04; c->plane_b = -0.03; c->n_bumps = 8;
		c->bump_sigma_min = 40; c->bump_sigma_max = 120; c->bump_height = 60;
		c->layout = 1; c->cam_distance = 600; c->cam_step_deg = 7;
		c->extent_x = 330; c->extent_y = 280; c->n_sparse = 20000; c->seed = 1002; dist = 600;
		break;
	case 3: // ETH3D-shaped: 20 x 6048x4032, arc
		c->n_views = 20; c->width = 6048; c->height = 4032; c->focal = 3410; c->cx = 3023.5; c->cy = 2015.5;
		c->surface = 2; c->plane_a = 0; c->plane_b = 0.02; c->n_bumps = 4;
		c->bump_sigma_min = 0.3; c->bump_sigma_max = 1.0; c->bump_height = 0.4;
		c->layout = 2; c->cam_distance = 7.0; c->cam_step_deg = 5;
		c->extent_x = 16; c->extent_y = 9; c->n_sparse = 20000; c->seed = 1003; dist = 7.0;
		break;
	case 4: // video: 300 x 1920x1080 dolly
	case 5: // fusion stress: 500 x 1920x1080 maps
		c->n_views = config == 4 ? 300 : 500; c->width = 1920; c->height = 1080; c->focal = 1600; c->cx = 959.5; c->cy = 539.5;
		c->surface = 1; c->plane_a = 0.01; c->plane_b = 0.02; c->n_bumps = 16;
		c->bump_sigma_min = 0.4; c->bump_sigma_max = 1.5; c->bump_height = 0.5;
		c->layout = 3; c->cam_distance = 5.0; c->cam_radius = 0.6; c->cam_step_deg = 0.105;
		c->extent_x = (config == 4 ? 20 : 32); c->extent_y = 5; c->n_sparse = 50000; c->seed = config == 4 ? 1004 : 1005; dist = 5.0;
		break;
	default: return -1;
	}
	if (n_views_override > 0) c->n_views = n_views_override;
	if (scale != 1.0) {
		c->width = std::max(32, (int)std::lround(c->width*scale));
		c->height = std::max(32, (int)std::lround(c->height*scale));
		c->focal *= scale; c->cx = (c->cx+0.5)*scale-0.5; c->cy = (c->cy+0.5)*scale-0.5;
	}
	c->tex_wavelength = 80.0*dist/c->focal;
	return 0;
}

extern "C" hcmvs_synth_scene* hcmvs_synth_create(const hcmvs_synth_cfg* cfg) {
	if (!cfg || cfg->n_views <= 0 || cfg->width <= 0 || cfg->height <= 0) return nullptr;
	hcmvs_synth_scene* s = new hcmvs_synth_scene();
	s->cfg = *cfg;
	SplitMix64 rng(cfg->seed);
	for (int i=0; i<cfg->n_bumps; ++i) {
		Bump b;
		b.x = rng.uniform(-0.7*cfg->extent_x, 0.7*cfg->extent_x);
		b.y = rng.uniform(-0.7*cfg->extent_y, 0.7*cfg->extent_y);
		const double sg = rng.uniform(cfg->bump_sigma_min, cfg->bump_sigma_max);
		b.s2inv = 1.0/(2.0*sg*sg);
		b.h = rng.uniform(-cfg->bump_height, cfg->bump_height);
		s->bumps.push_back(b);
	}
	// cameras
	const double deg = M_PI/180.0;
	const double up[3] = {0, 1, 0}, origin[3] = {0, 0, 0};
	s->cams.resize(cfg->n_views);
	const int grid = (int)std::ceil(std::sqrt((double)cfg->n_views));
	for (int i=0; i<cfg->n_views; ++i) {
		Cam& c = s->cams[i];
		const double K[9] = {cfg->focal, 0, cfg->cx, 0, cfg->focal, cfg->cy, 0, 0, 1};
		std::memcpy(c.K, K, sizeof(K));
		switch (cfg->layout) {
		case 0: { // ring
			const double th = (2.0*M_PI*i)/cfg->n_views;
			c.C[0] = cfg->cam_radius*std::cos(th); c.C[1] = cfg->cam_radius*std::sin(th); c.C[2] = cfg->cam_distance;
			LookAt(c.C, origin, up, c.R);
			break; }
		case 1: { // grid on a spherical cap
			const int gx = i%grid, gy = i/grid;
			const double ax = (gx-(grid-1)*0.5)*cfg->cam_step_deg*deg, ay = (gy-(grid-1)*0.5)*cfg->cam_step_deg*deg;
			c.C[0] = cfg->cam_distance*std::sin(ax)*std::cos(ay); c.C[1] = cfg->cam_distance*std::sin(ay); c.C[2] = cfg->cam_distance*std::cos(ax)*std::cos(ay);
			LookAt(c.C, origin, up, c.R);
			break; }
		case 2: { // arc about the y axis
			const double ax = (i-(cfg->n_views-1)*0.5)*cfg->cam_step_deg*deg;
			const double tgt[3] = {0, 0, -2.0};
			c.C[0] = cfg->cam_distance*std::sin(ax); c.C[1] = 0.15*cfg->cam_distance*std::sin(3.0*ax); c.C[2] = cfg->cam_distance*std::cos(ax)-2.0;
			LookAt(c.C, tgt, up, c.R);
			break; }
		default: { // dolly along x, looking down with a small yaw/lateral oscillation
			const double x = (i-(cfg->n_views-1)*0.5)*cfg->cam_step_deg;
			c.C[0] = x; c.C[1] = cfg->cam_radius*std::sin(2.0*M_PI*i/97.0); c.C[2] = cfg->cam_distance + 0.15*std::sin(2.0*M_PI*i/61.0);
			const double tgt[3] = {x + 0.2*std::sin(2.0*M_PI*i/53.0), c.C[1]*0.5, 0};
			LookAt(c.C, tgt, up, c.R);
			break; }
		}
	}
	s->zmin = s->zmax = 0;
	// sparse cloud with exact visibility (projects >= 8 px inside the frame and is the first surface hit)
	std::vector<float>& P = s->sparse; s->sparseOff.push_back(0);
	for (int k=0; k<cfg->n_sparse; ++k) {
		const double x = rng.uniform(-cfg->extent_x, cfg->extent_x), y = rng.uniform(-cfg->extent_y, cfg->extent_y);
		const double z = s->Height(x, y);
		const size_t nBefore = s->sparseViews.size();
		for (int v=0; v<cfg->n_views; ++v) {
			const Cam& c = s->cams[v];
			const double d[3] = {x-c.C[0], y-c.C[1], z-c.C[2]};
			const double xc = c.R[0]*d[0]+c.R[1]*d[1]+c.R[2]*d[2], yc = c.R[3]*d[0]+c.R[4]*d[1]+c.R[5]*d[2], zc = c.R[6]*d[0]+c.R[7]*d[1]+c.R[8]*d[2];
			if (zc <= 0) continue;
			const double u = c.K[2]+c.K[0]*xc/zc, w = c.K[5]+c.K[4]*yc/zc;
			if (u < 8 || w < 8 || u > cfg->width-9 || w > cfg->height-9) continue;
			bool occluded = false; // march the segment camera -> point: any crossing before the point hides it
			for (int m=0; m<16 && !occluded; ++m) {
				const double t = 0.5+0.495*(m/15.0);
				occluded = (c.C[2]+t*d[2]) < s->Height(c.C[0]+t*d[0], c.C[1]+t*d[1]);
			}
			if (occluded) continue;
			s->sparseViews.push_back((uint32_t)v);
		}
		if (s->sparseViews.size()-nBefore < 2) { s->sparseViews.resize(nBefore); continue; }
		P.push_back((float)x); P.push_back((float)y); P.push_back((float)z);
		s->sparseOff.push_back((int32_t)s->sparseViews.size());
	}
	return s;
}

extern "C" void hcmvs_synth_destroy(hcmvs_synth_scene* s) { delete s; }
extern "C" int hcmvs_synth_get_cfg(const hcmvs_synth_scene* s, hcmvs_synth_cfg* cfg) { if (!s || !cfg) return -1; *cfg = s->cfg; return 0; }

extern "C" int hcmvs_synth_camera(const hcmvs_synth_scene* s, int view, double* K, double* R, double* C) {
	if (!s || view < 0 || view >= (int)s->cams.size()) return -1;
	const Cam& c = s->cams[view];
	if (K) std::memcpy(K, c.K, sizeof(c.K));
	if (R) std::memcpy(R, c.R, sizeof(c.R));
	if (C) std::memcpy(C, c.C, sizeof(c.C));
	return 0;
}

extern "C" int hcmvs_synth_render(const hcmvs_synth_scene* s, int view, uint8_t* bgr, float* depth, float* normal, int n_threads) {
	if (!s || view < 0 || view >= (int)s->cams.size()) return -1;
	const Cam& c = s->cams[view];
	const int W = s->cfg.width, H = s->cfg.height;
	unsigned nt = n_threads > 0 ? (unsigned)n_threads : std::max(1u, std::thread::hardware_concurrency());
	nt = std::min<unsigned>(nt, (unsigned)H);
	auto work = [&](unsigned tid) {
		for (int y=(int)tid; y<H; y+=(int)nt) {
			double tPrev = -1;
			for (int x=0; x<W; ++x) {
				const double xc = (x-c.K[2])/c.K[0], yc = (y-c.K[5])/c.K[4];
				// world ray d = R^T (xc, yc, 1): camera-z of o + t d equals t
				const double d[3] = {c.R[0]*xc+c.R[3]*yc+c.R[6], c.R[1]*xc+c.R[4]*yc+c.R[7], c.R[2]*xc+c.R[5]*yc+c.R[8]};
				double t = s->Cast(c.C, d, tPrev);
				if (t < 0 && tPrev > 0) t = s->Cast(c.C, d, -1);
				tPrev = t;
				const size_t o = (size_t)y*W+x;
				if (t < 0) {
					if (depth) depth[o] = 0;
					if (normal) { normal[o*3] = normal[o*3+1] = 0; normal[o*3+2] = 0; }
					if (bgr) { bgr[o*3] = bgr[o*3+1] = bgr[o*3+2] = 0; }
					continue;
				}
				const double wx = c.C[0]+t*d[0], wy = c.C[1]+t*d[1];
				if (depth) depth[o] = (float)t;
				if (normal) {
					double gx, gy; s->Height(wx, wy, &gx, &gy);
					double n[3] = {-gx, -gy, 1.0};
					const double nn = std::sqrt(n[0]*n[0]+n[1]*n[1]+1.0); n[0] /= nn; n[1] /= nn; n[2] /= nn;
					double nc[3] = {c.R[0]*n[0]+c.R[1]*n[1]+c.R[2]*n[2], c.R[3]*n[0]+c.R[4]*n[1]+c.R[5]*n[2], c.R[6]*n[0]+c.R[7]*n[1]+c.R[8]*n[2]};
					if (nc[0]*xc+nc[1]*yc+nc[2] > 0) { nc[0] = -nc[0]; nc[1] = -nc[1]; nc[2] = -nc[2]; }
					normal[o*3] = (float)nc[0]; normal[o*3+1] = (float)nc[1]; normal[o*3+2] = (float)nc[2];
				}
				if (bgr) {
					const bool inside = std::abs(wx) <= s->cfg.extent_x*1.5 && std::abs(wy) <= s->cfg.extent_y*1.5;
					for (int ch=0; ch<3; ++ch) {
						const double tv = inside ? s->Texture(wx, wy, ch) : 0.5;
						bgr[o*3+ch] = (uint8_t)std::lround(16.0+224.0*tv);
					}
				}
			}
		}
	};
	std::vector<std::thread> th;
	for (unsigned t=1; t<nt; ++t) th.emplace_back(work, t);
	work(0);
	for (auto& t: th) t.join();
	return 0;
}

extern "C" int hcmvs_synth_sparse_size(const hcmvs_synth_scene* s, int* n_points, int* n_view_refs) {
	if (!s) return -1;
	if (n_points) *n_points = (int)(s->sparse.size()/3);
	if (n_view_refs) *n_view_refs = (int)s->sparseViews.size();
	return 0;
}
extern "C" int hcmvs_synth_sparse(const hcmvs_synth_scene* s, float* xyz, int32_t* offsets, uint32_t* view_ids) {
	if (!s) return -1;
	if (xyz) std::memcpy(xyz, s->sparse.data(), s->sparse.size()*sizeof(float));
	if (offsets) std::memcpy(offsets, s->sparseOff.data(), s->sparseOff.size()*sizeof(int32_t));
	if (view_ids) std::memcpy(view_ids, s->sparseViews.data(), s->sparseViews.size()*sizeof(uint32_t));
	return 0;
}
extern "C" double hcmvs_synth_height(const hcmvs_synth_scene* s, double x, double y) { return s ? s->Height(x, y) : 0.0; }
