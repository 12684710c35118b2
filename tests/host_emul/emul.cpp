/*
 * Host emulation of the per-ray device code -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiles sdfest_b200/csrc/sdfr_core.cuh (the exact functions the sm_100a kernels inline) with
 * g++ and drives them with plain loops that mimic what one CTA does per pixel: Frame + projected
 * rectangle, table-based ray set-up, rectangle culling, slab test, sphere trace, per-pixel
 * backward.  It lets the CPU test-suite check the kernel arithmetic and the conservativeness of
 * the rectangle culling against the oracle without a GPU.  It is never loaded by sdfest_b200.
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#include "../../sdfest_b200/csrc/sdfr_core.cuh"

using namespace sdfr;

static int build_frame_host(Frame& F, const float* pos, const float* quat, const float* inv_scale,
                            const Camera& cam, HullEdge* edges = nullptr, const CellBounds* bounds = nullptr,
                            int R = 0, float threshold = 0.0f) {
  frame_pose(F, pos, quat, inv_scale, bounds, R, threshold);
  bool all_ok = true;
  float cmin = INFINITY, cmax = -INFINITY, rmin = INFINITY, rmax = -INFINITY;
  float cols[8], rows[8];
  for (int k = 0; k < 8; ++k) {
    float col = 0.f, row = 0.f;
    const bool ok = project_corner(F, cam, k, col, row);
    all_ok = all_ok && ok;
    cols[k] = col; rows[k] = row;
    cmin = fminf(cmin, col); cmax = fmaxf(cmax, col);
    rmin = fminf(rmin, row); rmax = fmaxf(rmax, row);
  }
  frame_rect(F, cam, all_ok, cmin, cmax, rmin, rmax);
  return edges ? build_hull_serial(cols, rows, all_ok, edges) : 0;
}

/* bounds: NULL or 7 ints {lo[3], hi[3], tau bits} as sdfr_grid_bounds writes them (CellBounds) */
extern "C" int emul_forward_bounds(const float* sdf, int R, const float* pos, const float* quat,
                                   const float* inv_scale, int W, int H, float cx, float cy, float fx,
                                   float fy, float threshold, float* depth, int* steps_out, int* rect,
                                   int use_rect, const void* bounds) {
  const Grid G = make_grid(R);
  const Camera cam{W, H, cx, cy, fx, fy};
  Frame F;
  HullEdge edges[kMaxHullEdges];
  const int n_edges = build_frame_host(F, pos, quat, inv_scale, cam, edges,
                                       static_cast<const CellBounds*>(bounds), R, threshold);
  /* rect[4] = hull edges found, rect[5] = pixels the hull culls inside the rectangle */
  if (rect) { rect[0] = F.x0; rect[1] = F.y0; rect[2] = F.x1; rect[3] = F.y1; rect[4] = n_edges; rect[5] = 0; rect[6] = 0; }
  for (int py = 0; py < H; ++py)
    for (int px = 0; px < W; ++px) {
      float z = 0.f;
      int steps = 0;
      bool capped = false;
      bool keep = !use_rect || (px >= F.x0 && px < F.x1 && py >= F.y0 && py < F.y1);
      if (keep && use_rect == 2) { /* the kernels test whole 8x4-pixel warp tiles */
        const float bx = (float)(px & ~7), by = (float)(py & ~3);
        for (int e = 0; e < kMaxHullEdges; ++e)
          if (hull_block_outside(edges[e], bx, by, 8.0f, 4.0f)) keep = false;
        if (!keep && rect) rect[5] += 1;
      }
      if (keep) {
        const Ray r = make_ray(F, pixel_dx(px, cx, fx), pixel_dy(py, cy, fy));
        float t_min, t_max;
        if (ray_enters_cull_box(F, r) && ray_box(F, r, t_min, t_max)) {
          z = march<0>(sdf, G, F, r, t_min, t_max, threshold, steps, capped);
          if (rect) rect[6] += 1; /* rays marched */
        }
      }
      depth[(size_t)py * W + px] = z;
      if (steps_out) steps_out[(size_t)py * W + px] = steps;
    }
  return 0;
}

extern "C" int emul_forward(const float* sdf, int R, const float* pos, const float* quat,
                            const float* inv_scale, int W, int H, float cx, float cy, float fx,
                            float fy, float threshold, float* depth, int* steps_out, int* rect,
                            int use_rect) {
  return emul_forward_bounds(sdf, R, pos, quat, inv_scale, W, H, cx, cy, fx, fy, threshold, depth, steps_out,
                             rect, use_rect, nullptr);
}

extern "C" float emul_hit_tau(const float* pos, float inv_scale, float threshold) {
  return hit_tau(pos, inv_scale, threshold);
}

extern "C" int emul_backward(const float* grad_depth, const float* depth, const float* sdf, int R,
                             const float* pos, const float* quat, const float* inv_scale, int W,
                             int H, float cx, float cy, float fx, float fy, int exact,
                             double* g_sdf, double* g_pose) {
  const Grid G = make_grid(R);
  const Camera cam{W, H, cx, cy, fx, fy};
  Frame F;
  build_frame_host(F, pos, quat, inv_scale, cam);
  for (int i = 0; i < 8; ++i) g_pose[i] = 0;
  for (int py = 0; py < H; ++py)
    for (int px = 0; px < W; ++px) {
      const float z = depth[(size_t)py * W + px];
      if (z == 0.f) continue;
      const float gup = grad_depth[(size_t)py * W + px];
      if (gup == 0.f) continue;
      const Ray r = make_ray(F, pixel_dx(px, cx, fx), pixel_dy(py, cy, fy));
      PixelGrad pg;
      pixel_backward<0, true, true>(sdf, G, F, r, z, gup, exact != 0, pg);
      const int offs[8] = {0, 1, G.R, G.R + 1, G.R2, G.R2 + 1, G.R2 + G.R, G.R2 + G.R + 1};
      for (int k = 0; k < 8; ++k) g_sdf[pg.base + offs[k]] += (double)pg.w[k];
      for (int k = 0; k < 8; ++k) g_pose[k] += (double)(pg.pose[k] * gup);
    }
  return 0;
}

/* The kernels' formulation: 13 moment sums per CTA (here: per image), one 13 -> 8 map at the end. */
extern "C" int emul_backward_moments(const float* grad_depth, const float* depth, const float* sdf, int R,
                                     const float* pos, const float* quat, const float* inv_scale, int W,
                                     int H, float cx, float cy, float fx, float fy, int exact,
                                     double* g_sdf, double* g_pose) {
  const Grid G = make_grid(R);
  const Camera cam{W, H, cx, cy, fx, fy};
  Frame F;
  build_frame_host(F, pos, quat, inv_scale, cam);
  float acc[kMoments];
  for (int i = 0; i < kMoments; ++i) acc[i] = 0.f;
  for (int py = 0; py < H; ++py)
    for (int px = 0; px < W; ++px) {
      const float z = depth[(size_t)py * W + px];
      if (z == 0.f) continue;
      const float gup = grad_depth[(size_t)py * W + px];
      if (gup == 0.f) continue;
      const Ray r = make_ray(F, pixel_dx(px, cx, fx), pixel_dy(py, cy, fy));
      int base;
      float w[8];
      pixel_backward_moments<0, true, true>(sdf, G, F, r, z, gup, exact != 0, base, w, acc);
      const int offs[8] = {0, 1, G.R, G.R + 1, G.R2, G.R2 + 1, G.R2 + G.R, G.R2 + G.R + 1};
      for (int k = 0; k < 8; ++k) g_sdf[base + offs[k]] += (double)w[k];
    }
  float out[8];
  moments_to_pose(F, G, acc, out);
  for (int k = 0; k < 8; ++k) g_pose[k] = (double)out[k];
  return 0;
}
