/* Synthetic MVS scene generator (host-only C ABI).
 *
 * Produces the seeded synthetic scenes BASELINE.md / SURVEY.md §8(d) name (C1..C5): analytic
 * height-field surface, ray-cast BGR images with exact ground-truth depth / camera-space normal,
 * pin-hole cameras in the reference's convention (P = K R [I|-C], libs/MVS/Camera.h:46-54) and a
 * sparse point cloud with per-point visibility lists (the input Scene::SelectNeighborViews consumes,
 * libs/MVS/Scene.cpp:545). Used by bench.py, the tests and the DensifyPointCloud-style driver; it is
 * input synthesis, not part of the measured path.
 */
#ifndef HCMVS_SYNTH_H_
#define HCMVS_SYNTH_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct hcmvs_synth_cfg {
	int32_t n_views, width, height;
	double  focal, cx, cy;
	int32_t surface;        /* 0 slanted plane, 1 plane + gaussian bumps, 2 wedge (two planes at 100 deg) + bumps */
	int32_t n_bumps;
	double  plane_a, plane_b;       /* z = a*x + b*y (+ bumps) */
	double  bump_sigma_min, bump_sigma_max, bump_height; /* bump heights uniform in [-h, h] */
	int32_t layout;         /* 0 ring, 1 grid on a spherical cap, 2 arc, 3 dolly along x */
	double  cam_distance;   /* ring: height; cap/arc: radius; dolly: height */
	double  cam_radius;     /* ring radius / dolly lateral amplitude */
	double  cam_step_deg;   /* angular step between neighbouring cameras (cap/arc) or per-frame advance in world units (dolly) */
	double  extent_x, extent_y;     /* half extents of the textured/sparse-point footprint */
	double  tex_wavelength; /* base wavelength of the value-noise texture, world units */
	int32_t n_sparse;
	uint64_t seed;
} hcmvs_synth_cfg;

typedef struct hcmvs_synth_scene hcmvs_synth_scene;

/* Fill cfg with preset `config` (1..5 = C1..C5 of SURVEY §8d). `scale` in (0,1] shrinks the image size
 * and focal length together (same geometry, fewer pixels); n_views_override > 0 replaces the view count. */
int hcmvs_synth_preset(int config, double scale, int n_views_override, hcmvs_synth_cfg* cfg);

hcmvs_synth_scene* hcmvs_synth_create(const hcmvs_synth_cfg* cfg);
void hcmvs_synth_destroy(hcmvs_synth_scene* s);
int  hcmvs_synth_get_cfg(const hcmvs_synth_scene* s, hcmvs_synth_cfg* cfg);
/* row-major K[9], R[9], C[3] */
int  hcmvs_synth_camera(const hcmvs_synth_scene* s, int view, double* K, double* R, double* C);
/* Render view: any of bgr (H*W*3 u8), depth (H*W f32, camera z), normal (H*W*3 f32, camera space, facing the
 * camera) may be NULL. n_threads <= 0 uses all hardware threads. */
int  hcmvs_synth_render(const hcmvs_synth_scene* s, int view, uint8_t* bgr, float* depth, float* normal, int n_threads);
/* Sparse cloud: n points; views are returned CSR-style (offsets has n+1 entries). Call with NULL arrays to size. */
int  hcmvs_synth_sparse_size(const hcmvs_synth_scene* s, int* n_points, int* n_view_refs);
int  hcmvs_synth_sparse(const hcmvs_synth_scene* s, float* xyz, int32_t* offsets, uint32_t* view_ids);
/* height of the analytic surface and its (unnormalised) gradient at world (x,y) */
double hcmvs_synth_height(const hcmvs_synth_scene* s, double x, double y);

#ifdef __cplusplus
}
#endif
#endif
