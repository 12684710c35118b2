/*
 * sdf_oracle_impl.h -- body of the CPU oracle, instantiated twice by sdf_oracle.c
 * (REAL = float mirrors the reference CUDA kernels, REAL = double mirrors the
 * reference numpy renderer).
 *
 * TEST INFRASTRUCTURE ONLY: nothing under sdfest_b200/ may call this; it is the
 * checker for tests/, __graft_entry__.smoke() and the cpu_baseline/reference arm
 * of bench.py.
 *
 * Citations: "cu:" = sdfest/differentiable_renderer/csrc/sdf_renderer_cuda.cu,
 *            "simple:" = sdfest/differentiable_renderer/simple_renderer.py
 * of the reference tree.  This is a restatement of the algorithm written from
 * the arithmetic specification in SURVEY.md appendix A, not a copy of either file.
 *
 * Includer defines: REAL, SUF(name), ORACLE_FINITE_BOUNDS (1: cu:164-165 start
 * values -1e-10 / 1e10, 0: simple:85-86 -inf / +inf).
 */

/* rotate v by unit quaternion q=(x,y,z,w)           (cu:112-121, simple:336-350) */
static inline void SUF(quat_apply)(const REAL q[4], const REAL v[3], REAL out[3]) {
  const REAL x = q[0], y = q[1], z = q[2], w = q[3];
  out[0] = (1 - 2 * (y * y + z * z)) * v[0] + 2 * (x * y - w * z) * v[1] + 2 * (x * z + w * y) * v[2];
  out[1] = 2 * (x * y + w * z) * v[0] + (1 - 2 * (x * x + z * z)) * v[1] + 2 * (y * z - w * x) * v[2];
  out[2] = 2 * (x * z - w * y) * v[0] + 2 * (y * z + w * x) * v[1] + (1 - 2 * (x * x + y * y)) * v[2];
}

/* pinhole ray through the centre of pixel (row, col), OpenGL camera (cu:137-154) */
static inline void SUF(pixel_direction)(int row, int col, double cx, double cy, double fx,
                                        double fy, REAL d[3]) {
  d[0] = (REAL)((col + 0.5 - cx) / fx); /* double arithmetic, then narrowed (cu:146) */
  d[1] = (REAL)(-(row + 0.5 - cy) / fy);
  d[2] = (REAL)-1.0;
  const REAL rn = (REAL)1 / (REAL)sqrt((double)(d[0] * d[0] + d[1] * d[1] + 1));
  d[0] *= rn;
  d[1] *= rn;
  d[2] *= rn;
}

/* ray / oriented-box slab test, ray origin = camera centre (cu:156-194, simple:71-118) */
static inline int SUF(obb_ray)(const REAL d[3], const REAL p[3], const REAL q[4], REAL scale,
                               REAL* t_min_out, REAL* t_max_out) {
#if ORACLE_FINITE_BOUNDS
  REAL t_min = (REAL)-1e-10, t_max = (REAL)1e10;
#else
  REAL t_min = (REAL)-INFINITY, t_max = (REAL)INFINITY;
#endif
  for (int i = 0; i < 3; ++i) {
    REAL unit[3] = {0, 0, 0}, a[3];
    unit[i] = 1;
    SUF(quat_apply)(q, unit, a);
    const REAL e = a[0] * p[0] + a[1] * p[1] + a[2] * p[2];
    const REAL f = a[0] * d[0] + a[1] * d[1] + a[2] * d[2];
    if (fabs((double)f) > 1e-20) {
      REAL t1 = (e + scale) / f, t2 = (e - scale) / f;
      if (t1 > t2) {
        const REAL tmp = t1;
        t1 = t2;
        t2 = tmp;
      }
      if (t1 > t_min) t_min = t1;
      if (t2 < t_max) t_max = t2;
      if (t_min > t_max || t_max < 0) return 0;
    } else if (-e > scale || -e < -scale) {
      return 0;
    }
  }
  *t_min_out = t_min > 0 ? t_min : (REAL)0;
  *t_max_out = t_max;
  return 1;
}

/* base cell of a normalised object point (cu:196-207, simple:158-172) */
static inline int SUF(cell_of)(REAL u, int R) {
  int b = (int)floor((double)((u + (REAL)1.0) * (REAL)(R - 1) * (REAL)0.5));
  if (b > R - 2) b = R - 2;
  if (b < 0) b = 0;
  return b;
}

/* trilinear sample of grid G[R][R][R] (z fastest) at object point x*inv_scale
 * (cu:217-239, simple:185-219); offsets are NOT clamped (points just outside
 * the unit cube extrapolate). */
static inline REAL SUF(trilinear)(const REAL* G, int R, const REAL x[3], REAL inv_scale) {
  const REAL h = (REAL)(2.0 / (R - 1));
  const REAL hinv = (REAL)((R - 1) / 2.0);
  int b[3];
  REAL off[3];
  for (int a = 0; a < 3; ++a) {
    const REAL u = x[a] * inv_scale;
    b[a] = SUF(cell_of)(u, R);
    const REAL pos0 = (REAL)b[a] * h - (REAL)1.0;
    off[a] = hinv * (u - pos0);
  }
  const size_t sx = (size_t)R * R, sy = (size_t)R;
  const REAL* c = G + b[0] * sx + b[1] * sy + b[2];
  const REAL c00 = c[0] * (1 - off[0]) + c[sx] * off[0];
  const REAL c01 = c[1] * (1 - off[0]) + c[sx + 1] * off[0];
  const REAL c10 = c[sy] * (1 - off[0]) + c[sx + sy] * off[0];
  const REAL c11 = c[sy + 1] * (1 - off[0]) + c[sx + sy + 1] * off[0];
  const REAL c0 = c00 * (1 - off[1]) + c10 * off[1];
  const REAL c1 = c01 * (1 - off[1]) + c11 * off[1];
  return c0 * (1 - off[2]) + c1 * off[2];
}

/*
 * Forward: depth image of one posed, scaled SDF grid (cu:241-298, simple:253-314).
 * depth[row*W+col] = -t*d_z at the first sample with dist < threshold*t, else 0.
 * steps (nullable) receives the number of trilinear samples taken per pixel,
 * t_out (nullable) the ray parameter at the hit (0 where no hit).
 */
int SUF(oracle_render)(const REAL* sdf, int R, const REAL* pos, const REAL* quat,
                       const REAL* inv_scale_p, int W, int H, double cx, double cy, double fx,
                       double fy, double threshold_d, REAL* depth, int* steps, REAL* t_out,
                       int max_steps, int nthreads) {
  if (!sdf || !pos || !quat || !inv_scale_p || !depth || R < 2 || W < 0 || H < 0) return -1;
  const REAL inv_scale = *inv_scale_p;
  const REAL scale = (REAL)(1. / (double)inv_scale); /* cu:259 */
  const REAL threshold = (REAL)threshold_d;
  const REAL q[4] = {quat[0], quat[1], quat[2], quat[3]};
  const REAL qinv[4] = {-quat[0], -quat[1], -quat[2], quat[3]};
  const REAL p[3] = {pos[0], pos[1], pos[2]};
  REAL origin_o[3];
  {
    const REAL mp[3] = {(REAL)0 - p[0], (REAL)0 - p[1], (REAL)0 - p[2]};
    SUF(quat_apply)(qinv, mp, origin_o); /* cu:279-281 */
  }
  if (max_steps <= 0) max_steps = 1 << 20;
  if (nthreads <= 0) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
  for (int row = 0; row < H; ++row) {
    for (int col = 0; col < W; ++col) {
      REAL d[3], t_min, t_max, d_o[3];
      REAL out = 0, t_hit = 0;
      int n = 0;
      SUF(pixel_direction)(row, col, cx, cy, fx, fy, d);
      if (SUF(obb_ray)(d, p, q, scale, &t_min, &t_max)) {
        SUF(quat_apply)(qinv, d, d_o);
        REAL t = t_min;
        while (t < t_max && n < max_steps) {
          const REAL x[3] = {origin_o[0] + t * d_o[0], origin_o[1] + t * d_o[1],
                             origin_o[2] + t * d_o[2]};
          const REAL dist = SUF(trilinear)(sdf, R, x, inv_scale) * scale;
          ++n;
          if (dist < threshold * t) {
            out = -t * d[2];
            t_hit = t;
            break;
          }
          t += dist;
        }
      }
      depth[(size_t)row * W + col] = out;
      if (steps) steps[(size_t)row * W + col] = n;
      if (t_out) t_out[(size_t)row * W + col] = t_hit;
    }
  }
  return 0;
}

/*
 * Backward (cu:300-468, simple:317-458).  For every pixel with depth != 0 the hit
 * point is re-derived from the stored depth (t = -z/d_z, cu:336-338), and
 *   g_sdf[corner] += g * w_corner * scale*|d_z|
 *   g_pose[k]     += g * scale*|d_z| * d(trilinear)/d(theta_k),  k = x,y,z,qx,qy,qz,qw,s_inv
 *   g_pose[7]     -= g * v * scale^2 * |d_z|
 * sdf_grad_mode 0 = the corner weights the reference CUDA kernel uses (cu:373-388;
 * a permuted list), 1 = the true trilinear weights (simple:399-408).
 * Sums are accumulated in double.  deriv (nullable, [8][H][W]) receives the
 * per-pixel derivative images d depth / d theta_k (the reference numpy
 * renderer's `derivatives[k]`).
 */
int SUF(oracle_backward)(const REAL* grad_depth, const REAL* depth, const REAL* sdf, int R,
                         const REAL* pos, const REAL* quat, const REAL* inv_scale_p, int W,
                         int H, double cx, double cy, double fx, double fy, int sdf_grad_mode,
                         double* g_sdf, double* g_pose, REAL* deriv, int nthreads) {
  if (!depth || !sdf || !pos || !quat || !inv_scale_p || R < 2) return -1;
  const REAL inv_scale = *inv_scale_p;
  const REAL scale = (REAL)(1. / (double)inv_scale);             /* cu:324 */
  const REAL grid_size = (REAL)(2.0 / (R - 1));                  /* cu:327 */
  const REAL grid_size_inv = (REAL)(1. / (double)grid_size);     /* cu:328 */
  const REAL qx = quat[0], qy = quat[1], qz = quat[2], qw = quat[3];
  const REAL qinv[4] = {-qx, -qy, -qz, qw};
  const REAL p[3] = {pos[0], pos[1], pos[2]};
  REAL origin_o[3];
  {
    const REAL mp[3] = {(REAL)0 - p[0], (REAL)0 - p[1], (REAL)0 - p[2]};
    SUF(quat_apply)(qinv, mp, origin_o);
  }
  const size_t sx = (size_t)R * R, sy = (size_t)R;
  double pose_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (nthreads <= 0) nthreads = 1;
  if (deriv) memset(deriv, 0, sizeof(REAL) * 8 * (size_t)W * H);

#pragma omp parallel num_threads(nthreads)
  {
    double local[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma omp for schedule(dynamic, 4)
    for (int row = 0; row < H; ++row) {
      for (int col = 0; col < W; ++col) {
        const size_t pix = (size_t)row * W + col;
        const REAL z = depth[pix];
        if (z == 0) continue;
        const REAL g = grad_depth ? grad_depth[pix] : (REAL)1;
        REAL d[3], d_o[3];
        SUF(pixel_direction)(row, col, cx, cy, fx, fy, d);
        const REAL t = -z / d[2];
        SUF(quat_apply)(qinv, d, d_o);
        const REAL x[3] = {t * d[0], t * d[1], t * d[2]};
        const REAL o[3] = {origin_o[0] + t * d_o[0], origin_o[1] + t * d_o[1],
                           origin_o[2] + t * d_o[2]};
        int b[3];
        REAL c[3];
        for (int a = 0; a < 3; ++a) {
          const REAL no = o[a] * inv_scale;
          b[a] = SUF(cell_of)(no, R);
          const REAL pos0 = (REAL)b[a] * grid_size - (REAL)1.0;
          c[a] = grid_size_inv * (no - pos0);
        }
        const size_t base = b[0] * sx + b[1] * sy + b[2];
        const REAL* G = sdf + base;
        const REAL c000 = G[0], c001 = G[1], c010 = G[sy], c011 = G[sy + 1];
        const REAL c100 = G[sx], c101 = G[sx + 1], c110 = G[sx + sy], c111 = G[sx + sy + 1];
        const REAL c00 = c000 * (1 - c[0]) + c100 * c[0];
        const REAL c01 = c001 * (1 - c[0]) + c101 * c[0];
        const REAL c10 = c010 * (1 - c[0]) + c110 * c[0];
        const REAL c11 = c011 * (1 - c[0]) + c111 * c[0];
        const REAL c0 = c00 * (1 - c[1]) + c10 * c[1];
        const REAL c1 = c01 * (1 - c[1]) + c11 * c[1];
        const REAL t_diff = c0 * (1 - c[2]) + c1 * c[2];
        const REAL absdz = (REAL)fabs((double)d[2]);
        const REAL f = scale * absdz;

        if (g_sdf) {
          REAL w[8]; /* corner order 000,001,010,011,100,101,110,111 as (x,y,z) */
          if (sdf_grad_mode == 0) { /* cu:373-388 */
            w[0] = (1 - c[0]) * (1 - c[1]) * c[2];
            w[1] = (1 - c[0]) * c[1] * (1 - c[2]);
            w[2] = (1 - c[0]) * c[1] * c[2];
            w[3] = c[0] * (1 - c[1]) * (1 - c[2]);
            w[4] = c[0] * (1 - c[1]) * c[2];
            w[5] = c[0] * (1 - c[1]) * c[2];
            w[6] = c[0] * c[1] * (1 - c[2]);
            w[7] = c[0] * c[1] * c[2];
          } else { /* simple:399-408 */
            w[0] = (1 - c[0]) * (1 - c[1]) * (1 - c[2]);
            w[1] = (1 - c[0]) * (1 - c[1]) * c[2];
            w[2] = (1 - c[0]) * c[1] * (1 - c[2]);
            w[3] = (1 - c[0]) * c[1] * c[2];
            w[4] = c[0] * (1 - c[1]) * (1 - c[2]);
            w[5] = c[0] * (1 - c[1]) * c[2];
            w[6] = c[0] * c[1] * (1 - c[2]);
            w[7] = c[0] * c[1] * c[2];
          }
          const size_t offs[8] = {0, 1, sy, sy + 1, sx, sx + 1, sx + sy, sx + sy + 1};
          for (int k = 0; k < 8; ++k) {
            const double v = (double)(g * w[k] * f);
#pragma omp atomic
            g_sdf[base + offs[k]] += v;
          }
        }

        /* d c / d theta for theta = x,y,z,qx,qy,qz,qw,s_inv (cu:391-438) */
        const REAL s = inv_scale * grid_size_inv;
        const REAL ox[3] = {x[0] - p[0], x[1] - p[1], x[2] - p[2]};
        REAL dc[8][3];
        dc[0][0] = (2 * (qy * qy + qz * qz) - 1) * s;
        dc[0][1] = 2 * (qw * qz - qx * qy) * s;
        dc[0][2] = -2 * (qx * qz + qw * qy) * s;
        dc[1][0] = -2 * (qx * qy + qw * qz) * s;
        dc[1][1] = (2 * (qx * qx + qz * qz) - 1) * s;
        dc[1][2] = 2 * (qw * qx - qy * qz) * s;
        dc[2][0] = 2 * (qw * qy - qx * qz) * s;
        dc[2][1] = -2 * (qy * qz + qw * qx) * s;
        dc[2][2] = (2 * (qx * qx + qy * qy) - 1) * s;
        /* qx */
        dc[3][0] = (2 * qx * ox[0] + 2 * qy * ox[1] + 2 * qz * ox[2] - 2 * qx * o[0]) * s;
        dc[3][1] = (2 * qy * ox[0] - 2 * qx * ox[1] + 2 * qw * ox[2] - 2 * qx * o[1]) * s;
        dc[3][2] = (2 * qz * ox[0] - 2 * qw * ox[1] - 2 * qx * ox[2] - 2 * qx * o[2]) * s;
        /* qy */
        dc[4][0] = (-2 * qy * ox[0] + 2 * qx * ox[1] - 2 * qw * ox[2] - 2 * qy * o[0]) * s;
        dc[4][1] = (2 * qx * ox[0] + 2 * qy * ox[1] + 2 * qz * ox[2] - 2 * qy * o[1]) * s;
        dc[4][2] = (2 * qw * ox[0] + 2 * qz * ox[1] - 2 * qy * ox[2] - 2 * qy * o[2]) * s;
        /* qz */
        dc[5][0] = (-2 * qz * ox[0] + 2 * qw * ox[1] + 2 * qx * ox[2] - 2 * qz * o[0]) * s;
        dc[5][1] = (-2 * qw * ox[0] - 2 * qz * ox[1] + 2 * qy * ox[2] - 2 * qz * o[1]) * s;
        dc[5][2] = (2 * qx * ox[0] + 2 * qy * ox[1] + 2 * qz * ox[2] - 2 * qz * o[2]) * s;
        /* qw */
        dc[6][0] = (2 * qw * ox[0] + 2 * qz * ox[1] - 2 * qy * ox[2] - 2 * qw * o[0]) * s;
        dc[6][1] = (-2 * qz * ox[0] + 2 * qw * ox[1] + 2 * qx * ox[2] - 2 * qw * o[1]) * s;
        dc[6][2] = (2 * qy * ox[0] - 2 * qx * ox[1] + 2 * qw * ox[2] - 2 * qw * o[2]) * s;
        /* s_inv */
        dc[7][0] = o[0] * grid_size_inv;
        dc[7][1] = o[1] * grid_size_inv;
        dc[7][2] = o[2] * grid_size_inv;

        for (int k = 0; k < 8; ++k) { /* chain rule through the trilinear form (cu:444-456) */
          const REAL dc00 = -c000 * dc[k][0] + c100 * dc[k][0];
          const REAL dc01 = -c001 * dc[k][0] + c101 * dc[k][0];
          const REAL dc10 = -c010 * dc[k][0] + c110 * dc[k][0];
          const REAL dc11 = -c011 * dc[k][0] + c111 * dc[k][0];
          const REAL dc0 = dc00 * (1 - c[1]) - c00 * dc[k][1] + dc10 * c[1] + c10 * dc[k][1];
          const REAL dc1 = dc01 * (1 - c[1]) - c01 * dc[k][1] + dc11 * c[1] + c11 * dc[k][1];
          const REAL dtdiff = dc0 * (1 - c[2]) - c0 * dc[k][2] + dc1 * c[2] + c1 * dc[k][2];
          REAL dz = scale * dtdiff * absdz;
          if (k == 7) dz -= (t_diff * scale * scale) * absdz; /* cu:457 */
          if (deriv) deriv[(size_t)k * W * H + pix] = dz;
          local[k] += (double)(dz * g);
        }
      }
    }
#pragma omp critical
    for (int k = 0; k < 8; ++k) pose_acc[k] += local[k];
  }
  if (g_pose)
    for (int k = 0; k < 8; ++k) g_pose[k] = pose_acc[k];
  return 0;
}
