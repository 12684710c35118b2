/*
 * sdfr_step.cuh -- the optimiser side of one render-and-compare iteration as ONE kernel.
 * Included by sdfrender.cu inside its anonymous namespace.
 *
 * Reference (sdfest/estimation/simple_setup.py), per iteration and hypothesis:
 *   :400-406  torch.optim.Adam over four parameter groups: position lr 1e-3, orientation 1e-2,
 *             scale 1e-3, latent 1e-2 (default betas 0.9 / 0.999, eps 1e-8)
 *   :411      norm_orientation = orientation / sqrt(sum(orientation**2))
 *   :431      the renderer receives 1 / scale
 *   :447-452  loss = depth_weight * loss_depth + pc_weight * loss_pc (+ terms that are zero in the
 *             shipped configuration)
 *   :456-462  backward, optimizer.step(), orientation /= |orientation|
 * Run through torch this is ~60 launches of 1-2 us kernels per iteration (the chain rule through
 * the normalisation and the reciprocal, nan_to_num, four foreach-Adam groups of ~10 kernels, the
 * renormalisation) -- as much device time as the renderer itself once the loop is replayed from a
 * CUDA graph (profiles/r01m_loop_ops_fused_iteration.txt).  Here thread b owns hypothesis b and does,
 * in registers:
 *   coef      = n_overlap > 0 ? depth_weight / n_overlap : 0        (mean over the overlap, deferred
 *                                                                    by sdfr_compare_fused)
 *   g_pos     = coef * gr_pos + g2_pos
 *   g_unit_q  = coef * gr_quat + g2_quat
 *   g_orient  = (g_unit_q - q (q . g_unit_q)) / |orientation|       (through :411)
 *   g_scale   = -coef * gr_inv_scale / scale^2 + g2_scale           (through :431)
 *   Adam (torch/optim/adam.py _single_tensor_adam, no weight decay / amsgrad):
 *     m = m + (g - m)(1 - b1);  v = b2 v + (1 - b2) g g;
 *     p += -(lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
 *   orientation /= |orientation|;  unit_orientation = orientation / |orientation|;  inv_scale = 1/scale
 *   loss      = depth_weight * (n_overlap > 0 ? loss_sum / n_overlap : NaN) + point_weight * point_sum
 *               (NaN = the reference's mean over an empty overlap; such a hypothesis must never rank
 *               best -- its gradients are 0, not NaN, so the optimiser state stays finite)
 * and optionally clears the gradient inputs it consumed, so that the next iteration's kernels can
 * accumulate into them without a memset node.
 */
#ifndef SDFR_STEP_CUH_
#define SDFR_STEP_CUH_

constexpr int kStepMaxLatent = 64;

struct StepParams {
  float* __restrict__ position;     /* [B,3] */
  float* __restrict__ orientation;  /* [B,4] */
  float* __restrict__ scale;        /* [B]   */
  float* __restrict__ latent;       /* [B,L] or NULL */
  int latent_size; /* 0 when there is no latent gradient this call */
  int state_stride; /* 8 + the caller's latent_size */
  int batch;
  float* __restrict__ loss_sum;     /* [B] or NULL: masked-L1 sums of sdfr_compare_fused */
  float* __restrict__ n_overlap;    /* [B] or NULL (= no render gradient: coef = 0) */
  float* __restrict__ gr_position;  /* raw (unnormalised) render gradients, any may be NULL */
  float* __restrict__ gr_orientation;
  float* __restrict__ gr_inv_scale;
  float depth_weight;
  float* __restrict__ point_sum;    /* [B] or NULL */
  float point_weight;               /* pc_weight / n_points */
  float* __restrict__ g2_position;  /* second gradient set, already weighted; w.r.t. the UNIT */
  float* __restrict__ g2_orientation; /* quaternion and w.r.t. scale (not its inverse) */
  float* __restrict__ g2_scale;
  float* __restrict__ g_latent;     /* [B,L] or NULL */
  float* __restrict__ g_orientation_raw; /* [B,4] or NULL: added AFTER the chain rule, i.e. w.r.t. the
                                          * un-normalised orientation (point constraint, :164-175) */
  float* __restrict__ loss_extra;        /* [B] or NULL: added to the loss as is */
  float* __restrict__ exp_avg;      /* [B, 8+L] */
  float* __restrict__ exp_avg_sq;   /* [B, 8+L] */
  int* __restrict__ step;           /* [B] */
  float lr[4];                      /* position, orientation, scale, latent */
  float beta1, beta2, eps;
  float* __restrict__ unit_orientation; /* [B,4] out or NULL */
  float* __restrict__ inv_scale;        /* [B] out or NULL */
  float* __restrict__ loss;             /* [B] out or NULL */
  unsigned flags;
};

__device__ __forceinline__ float adam_update(float p, float g, float& m, float& v, float lr_over_bc1,
                                             float bc2_sqrt, float b1, float b2, float eps) {
  m = m + (g - m) * (1.0f - b1);
  v = v * b2 + (1.0f - b2) * g * g;
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  return p + (-lr_over_bc1) * (m / denom);
}

/* beta^t by repeated squaring (<= 2 log2 t double multiplies; pow() costs microseconds here) */
__device__ __forceinline__ double step_powi(double base, int t) {
  double r = 1.0;
  while (t > 0) {
    if (t & 1) r *= base;
    base *= base;
    t >>= 1;
  }
  return r;
}

/* Every input of the hypothesis is loaded before the first store (the pointers may alias as far as
 * the compiler knows, so interleaved loads and stores serialise on memory latency: the first version
 * took 11 us for 64 hypotheses, ~15 dependent round trips). */
__global__ void __launch_bounds__(128)
sdfr_hypothesis_step_kernel(const __grid_constant__ StepParams P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.batch) return;
  const bool clear = (P.flags & SDFR_STEP_CLEAR_INPUTS) != 0;
  const bool frozen = (P.flags & SDFR_STEP_NO_UPDATE) != 0;
  const int NS = P.state_stride;
  float* __restrict__ m = P.exp_avg + (size_t)b * NS;
  float* __restrict__ v = P.exp_avg_sq + (size_t)b * NS;

  /* ---- loads ---- */
  const float n = P.n_overlap ? P.n_overlap[b] : 0.0f;
  const float lsum = P.loss_sum ? P.loss_sum[b] : 0.0f;
  const float psum = P.point_sum ? P.point_sum[b] : 0.0f;
  float p[8], gr[8], g2[8], mi[8], vi[8];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    p[i] = P.position[3 * b + i];
    gr[i] = P.gr_position ? P.gr_position[3 * b + i] : 0.0f;
    g2[i] = P.g2_position ? P.g2_position[3 * b + i] : 0.0f;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    p[3 + i] = P.orientation[4 * b + i];
    gr[3 + i] = P.gr_orientation ? P.gr_orientation[4 * b + i] : 0.0f;
    g2[3 + i] = P.g2_orientation ? P.g2_orientation[4 * b + i] : 0.0f;
  }
  p[7] = P.scale[b];
  gr[7] = P.gr_inv_scale ? P.gr_inv_scale[b] : 0.0f;
  g2[7] = P.g2_scale ? P.g2_scale[b] : 0.0f;
  float graw[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) graw[i] = P.g_orientation_raw ? P.g_orientation_raw[4 * b + i] : 0.0f;
  const float lextra = P.loss_extra ? P.loss_extra[b] : 0.0f;
  int t = 0;
  if (!frozen) {
    t = P.step[b] + 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mi[i] = m[i];
      vi[i] = v[i];
    }
  }

  /* ---- arithmetic ---- */
  const float coef = n > 0.0f ? P.depth_weight / n : 0.0f;
  float l = 0.0f;
  /* no overlap: the reference's mean over an empty selection is NaN (simple_setup.py:131) and so is its
   * loss; the GRADIENTS stay 0 here (coef above) instead of poisoning Adam as the reference's do */
  if (P.loss_sum) l = n > 0.0f ? P.depth_weight * (lsum / n) : __int_as_float(0x7fc00000);
  if (P.point_sum) l += P.point_weight * psum;
  l += lextra;
  float bc1 = 1.0f, bc2_sqrt = 1.0f;
  if (!frozen) {
    float g[8]; /* position 0-2, orientation 3-6, scale 7 */
#pragma unroll
    for (int i = 0; i < 3; ++i) g[i] = coef * gr[i] + g2[i];
    float gq[4], q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) gq[i] = coef * gr[3 + i] + g2[3 + i];
    const float nrm = sqrtf(p[3] * p[3] + p[4] * p[4] + p[5] * p[5] + p[6] * p[6]);
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = p[3 + i] / nrm;
    const float dot = q[0] * gq[0] + q[1] * gq[1] + q[2] * gq[2] + q[3] * gq[3];
#pragma unroll
    for (int i = 0; i < 4; ++i) g[3 + i] = (gq[i] - q[i] * dot) / nrm + graw[i];
    g[7] = -(coef * gr[7]) / (p[7] * p[7]) + g2[7];
    bc1 = (float)(1.0 - step_powi((double)P.beta1, t));
    bc2_sqrt = sqrtf((float)(1.0 - step_powi((double)P.beta2, t)));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float lr = P.lr[i < 3 ? 0 : (i < 7 ? 1 : 2)];
      p[i] = adam_update(p[i], g[i], mi[i], vi[i], lr / bc1, bc2_sqrt, P.beta1, P.beta2, P.eps);
    }
    /* simple_setup.py:462 */
    const float n2 = sqrtf(p[3] * p[3] + p[4] * p[4] + p[5] * p[5] + p[6] * p[6]);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[3 + i] = p[3 + i] / n2;
  }
  /* what the next iteration's renderer reads (simple_setup.py:411, :431) */
  const float n3 = sqrtf(p[3] * p[3] + p[4] * p[4] + p[5] * p[5] + p[6] * p[6]);

  /* ---- stores ---- */
  if (P.loss) P.loss[b] = l;
  if (!frozen) {
    P.step[b] = t;
#pragma unroll
    for (int i = 0; i < 3; ++i) P.position[3 * b + i] = p[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) P.orientation[4 * b + i] = p[3 + i];
    P.scale[b] = p[7];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      m[i] = mi[i];
      v[i] = vi[i];
    }
  }
  if (P.unit_orientation) {
#pragma unroll
    for (int i = 0; i < 4; ++i) P.unit_orientation[4 * b + i] = p[3 + i] / n3;
  }
  if (P.inv_scale) P.inv_scale[b] = 1.0f / p[7];
  if (clear) {
    if (P.loss_sum) P.loss_sum[b] = 0.0f;
    if (P.n_overlap) P.n_overlap[b] = 0.0f;
    if (P.point_sum) P.point_sum[b] = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (P.gr_position) P.gr_position[3 * b + i] = 0.0f;
      if (P.g2_position) P.g2_position[3 * b + i] = 0.0f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (P.gr_orientation) P.gr_orientation[4 * b + i] = 0.0f;
      if (P.g2_orientation) P.g2_orientation[4 * b + i] = 0.0f;
    }
    if (P.gr_inv_scale) P.gr_inv_scale[b] = 0.0f;
    if (P.g2_scale) P.g2_scale[b] = 0.0f;
    if (P.loss_extra) P.loss_extra[b] = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (P.g_orientation_raw) P.g_orientation_raw[4 * b + i] = 0.0f;
  }

  /* ---- latent group: 8 values at a time, loads before stores ---- */
  if (!frozen && P.latent && P.g_latent) {
    float* __restrict__ z = P.latent + (size_t)b * P.latent_size;
    float* __restrict__ gz = P.g_latent + (size_t)b * P.latent_size;
    for (int i0 = 0; i0 < P.latent_size; i0 += 8) {
      float zz[8], gg[8], mm[8], vv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool ok = i0 + i < P.latent_size;
        zz[i] = ok ? z[i0 + i] : 0.0f;
        gg[i] = ok ? gz[i0 + i] : 0.0f;
        mm[i] = ok ? m[8 + i0 + i] : 0.0f;
        vv[i] = ok ? v[8 + i0 + i] : 0.0f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        zz[i] = adam_update(zz[i], gg[i], mm[i], vv[i], P.lr[3] / bc1, bc2_sqrt, P.beta1, P.beta2, P.eps);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i0 + i < P.latent_size) {
          z[i0 + i] = zz[i];
          m[8 + i0 + i] = mm[i];
          v[8 + i0 + i] = vv[i];
          if (clear) gz[i0 + i] = 0.0f;
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Result selection of the loop (estimation/simple_setup.py):
 *   :177-188  _compute_inlier_ratio: rel = |depth_input - depth_estimate| / depth_input;
 *             inliers = count(rel < relative_inlier_threshold); ratio = inliers / count(depth_input != 0)
 *   :190-211  _update_best_estimate: keep the estimate with the highest ratio so far
 *   :463      called every iteration, after optimizer.step(), with the depth rendered before the step
 * The reference needs ~8 torch kernels and, through `inlier_ratio > self._best_inlier_ratio`, a host
 * synchronisation per iteration.  Here: one streaming pass over the estimate the fused traversal wrote
 * anyway (sdfr_inlier_count_kernel, two counters per hypothesis) and one thread per hypothesis for the
 * bookkeeping (sdfr_track_best_kernel).  The division is IEEE (obs == 0 gives inf or nan: never an
 * inlier, exactly as the torch expression).
 * ---------------------------------------------------------------------------------------- */
struct InlierParams {
  const float* __restrict__ depth; /* [B, pixels] */
  const float* __restrict__ obs;   /* + b * obs_stride */
  long long obs_stride;
  int pixels;
  float threshold;
  float* __restrict__ n_inlier; /* [B] += */
  float* __restrict__ n_valid;  /* [B] += */
  int z_offset;
};

template <bool VEC>
__global__ void __launch_bounds__(256)
sdfr_inlier_count_kernel(const __grid_constant__ InlierParams P) {
  __shared__ float red[8][2];
  const int b = blockIdx.y + P.z_offset;
  const float* __restrict__ est = P.depth + (size_t)b * P.pixels;
  const float* __restrict__ obs = P.obs + (size_t)b * P.obs_stride;
  const float thr = P.threshold;
  float inl = 0.0f, val = 0.0f;
  if (VEC) {
    const int n4 = P.pixels >> 2;
    const float4* __restrict__ e4 = reinterpret_cast<const float4*>(est);
    const float4* __restrict__ o4 = reinterpret_cast<const float4*>(obs);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n4; i += gridDim.x * 256) {
      const float4 e = __ldg(e4 + i), o = __ldg(o4 + i);
      val += (o.x != 0.0f) + (o.y != 0.0f) + (o.z != 0.0f) + (o.w != 0.0f);
      inl += (__fdiv_rn(fabsf(o.x - e.x), o.x) < thr) + (__fdiv_rn(fabsf(o.y - e.y), o.y) < thr) +
             (__fdiv_rn(fabsf(o.z - e.z), o.z) < thr) + (__fdiv_rn(fabsf(o.w - e.w), o.w) < thr);
    }
  } else {
    for (int i = blockIdx.x * 256 + threadIdx.x; i < P.pixels; i += gridDim.x * 256) {
      const float e = __ldg(est + i), o = __ldg(obs + i);
      val += (o != 0.0f);
      inl += (__fdiv_rn(fabsf(o - e), o) < thr);
    }
  }
  inl = warp_sum(inl);
  val = warp_sum(val);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[warp][0] = inl;
    red[warp][1] = val;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.0f;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    /* counts are integers < 2^24: the float sums are exact in any order */
    if (t != 0.0f) atomicAdd((threadIdx.x == 0 ? P.n_inlier : P.n_valid) + b, t);
  }
}

struct TrackParams {
  float* __restrict__ n_inlier; /* [B] */
  float* __restrict__ n_valid;  /* [B] */
  const float* __restrict__ position;
  const float* __restrict__ orientation;
  const float* __restrict__ scale;
  const float* __restrict__ latent; /* [B,L] or NULL */
  int latent_size, batch;
  const int* __restrict__ step;     /* [B] iteration counter of sdfr_hypothesis_step, or NULL */
  float* __restrict__ ratio;        /* [B] out or NULL */
  float* __restrict__ best_ratio;   /* [B] */
  int* __restrict__ best_iteration; /* [B], < 0 = nothing kept yet */
  float* __restrict__ best_position;
  float* __restrict__ best_orientation;
  float* __restrict__ best_scale;
  float* __restrict__ best_latent;
  unsigned flags;
};

__global__ void __launch_bounds__(128)
sdfr_track_best_kernel(const __grid_constant__ TrackParams P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.batch) return;
  const float r = __fdiv_rn(P.n_inlier[b], P.n_valid[b]); /* 0/0 = nan, like the torch expression */
  const int it = P.step ? P.step[b] : 0;
  const bool first = P.best_iteration[b] < 0;
  const bool better = first || r > P.best_ratio[b]; /* simple_setup.py:205 */
  if (P.ratio) P.ratio[b] = r;
  if (P.flags & SDFR_STEP_CLEAR_INPUTS) {
    P.n_inlier[b] = 0.0f;
    if (!(P.flags & SDFR_TRACK_KEEP_VALID)) P.n_valid[b] = 0.0f;
  }
  if (!better) return;
  P.best_ratio[b] = r;
  P.best_iteration[b] = it;
#pragma unroll
  for (int i = 0; i < 3; ++i) P.best_position[3 * b + i] = P.position[3 * b + i];
#pragma unroll
  for (int i = 0; i < 4; ++i) P.best_orientation[4 * b + i] = P.orientation[4 * b + i];
  P.best_scale[b] = P.scale[b];
  if (P.latent && P.best_latent)
    for (int i = 0; i < P.latent_size; ++i)
      P.best_latent[(size_t)b * P.latent_size + i] = P.latent[(size_t)b * P.latent_size + i];
}

/* ------------------------------------------------------------------------------------------
 * Several camera views of one object (estimation/simple_setup.py:420-446).  The object pose lives in the
 * world frame; view v sees it at
 *     position_c    = R(q_w2c,v) (position - camera_position_v)
 *     orientation_c = q_w2c,v (x) unit_orientation                 q_w2c,v = conj(camera_orientation_v)
 * (quaternion_utils.py:12-66, scalar-last; camera orientations are unit quaternions).  Both maps are
 * linear in the pose, so the per-view gradients the renderer and the point loss return in the camera
 * frames are pulled back with the transposes -- R^T = rotation by camera_orientation_v, L(q)^T = L(conj q)
 * = left-multiplication by camera_orientation_v -- and summed over the views.  The reference does this
 * with ~20 autograd nodes per view and iteration; here: one thread per (view, hypothesis) forward, one
 * thread per hypothesis backward.
 * ---------------------------------------------------------------------------------------- */
struct ViewParams {
  const float* __restrict__ position;      /* [B,3] world */
  const float* __restrict__ unit_orientation; /* [B,4] world, unit */
  const float* __restrict__ inv_scale;     /* [B] */
  const float* __restrict__ scale;         /* [B] (pull-back only) */
  const float* __restrict__ cam_position;  /* [V,3] */
  const float* __restrict__ cam_orientation; /* [V,4] camera-to-world, unit */
  int n_views, batch;
  /* forward outputs, [V,B,...] */
  float* __restrict__ position_c;
  float* __restrict__ orientation_c;
  float* __restrict__ inv_scale_c;
  /* pull-back inputs, [V,B,...]; any may be NULL */
  float* __restrict__ gr_position;   /* d (weighted depth loss of view v) / d position_c */
  float* __restrict__ gr_orientation;
  float* __restrict__ gr_inv_scale;
  float* __restrict__ g2_position;   /* d (weighted point loss of view v) / d position_c, orientation_c, scale */
  float* __restrict__ g2_orientation;
  float* __restrict__ g2_scale;
  float* __restrict__ loss_sum;      /* [V,B] masked-L1 sums */
  float* __restrict__ n_overlap;     /* [V,B] */
  float* __restrict__ point_sum;     /* [V,B] already weighted */
  float depth_weight;
  /* pull-back outputs, [B,...] */
  float* __restrict__ g_position;
  float* __restrict__ g_orientation; /* w.r.t. the world-frame UNIT quaternion */
  float* __restrict__ g_scale;
  float* __restrict__ loss;          /* [B] += */
  unsigned flags;
};

__device__ __forceinline__ void quat_rotate(const float* q, const float* v, float* out) {
  /* R(q) v for a unit quaternion (x,y,z,w): v + 2 w (u x v) + 2 u x (u x v) */
  const float ux = q[0], uy = q[1], uz = q[2], w = q[3];
  const float cx = uy * v[2] - uz * v[1], cy = uz * v[0] - ux * v[2], cz = ux * v[1] - uy * v[0];
  const float dx = uy * cz - uz * cy, dy = uz * cx - ux * cz, dz = ux * cy - uy * cx;
  out[0] = v[0] + 2.0f * (w * cx + dx);
  out[1] = v[1] + 2.0f * (w * cy + dy);
  out[2] = v[2] + 2.0f * (w * cz + dz);
}

__device__ __forceinline__ void quat_mul(const float* a, const float* b, float* out) {
  /* Hamilton product a (x) b, scalar-last (quaternion_utils.py:28-34) */
  out[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  out[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
  out[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
  out[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}

__global__ void __launch_bounds__(128)
sdfr_view_poses_kernel(const __grid_constant__ ViewParams P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_views * P.batch) return;
  const int v = i / P.batch, b = i - v * P.batch;
  const float cq[4] = {P.cam_orientation[4 * v], P.cam_orientation[4 * v + 1], P.cam_orientation[4 * v + 2],
                       P.cam_orientation[4 * v + 3]};
  const float w2c[4] = {-cq[0], -cq[1], -cq[2], cq[3]};
  const float d[3] = {P.position[3 * b] - P.cam_position[3 * v], P.position[3 * b + 1] - P.cam_position[3 * v + 1],
                      P.position[3 * b + 2] - P.cam_position[3 * v + 2]};
  const float q[4] = {P.unit_orientation[4 * b], P.unit_orientation[4 * b + 1], P.unit_orientation[4 * b + 2],
                      P.unit_orientation[4 * b + 3]};
  float pc[3], qc[4];
  quat_rotate(w2c, d, pc);
  quat_mul(w2c, q, qc);
#pragma unroll
  for (int k = 0; k < 3; ++k) P.position_c[3 * i + k] = pc[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) P.orientation_c[4 * i + k] = qc[k];
  P.inv_scale_c[i] = P.inv_scale[b];
}

__global__ void __launch_bounds__(128)
sdfr_views_pull_back_kernel(const __grid_constant__ ViewParams P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.batch) return;
  const bool clear = (P.flags & SDFR_STEP_CLEAR_INPUTS) != 0;
  float gp[3] = {0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f}, gis = 0.f, gs = 0.f, l = 0.f;
  for (int v = 0; v < P.n_views; ++v) {
    const int i = v * P.batch + b;
    const float cq[4] = {P.cam_orientation[4 * v], P.cam_orientation[4 * v + 1], P.cam_orientation[4 * v + 2],
                         P.cam_orientation[4 * v + 3]};
    float a[3] = {0.f, 0.f, 0.f}, c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 3; ++k)
      a[k] = (P.gr_position ? P.gr_position[3 * i + k] : 0.f) + (P.g2_position ? P.g2_position[3 * i + k] : 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      c[k] = (P.gr_orientation ? P.gr_orientation[4 * i + k] : 0.f) + (P.g2_orientation ? P.g2_orientation[4 * i + k] : 0.f);
    float ra[3], rc[4];
    quat_rotate(cq, a, ra); /* R(q_w2c)^T = R(camera_orientation) */
    quat_mul(cq, c, rc);    /* L(q_w2c)^T = L(camera_orientation) */
#pragma unroll
    for (int k = 0; k < 3; ++k) gp[k] += ra[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) gq[k] += rc[k];
    gis += P.gr_inv_scale ? P.gr_inv_scale[i] : 0.f;
    gs += P.g2_scale ? P.g2_scale[i] : 0.f;
    if (P.loss_sum && P.n_overlap) {
      const float n = P.n_overlap[i];
      /* no overlap in a view: NaN, the reference's mean over an empty selection (:131) */
      l += n > 0.f ? P.depth_weight * (P.loss_sum[i] / n) : __int_as_float(0x7fc00000);
    }
    if (P.point_sum) l += P.point_sum[i];
    if (clear) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (P.gr_position) P.gr_position[3 * i + k] = 0.f;
        if (P.g2_position) P.g2_position[3 * i + k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (P.gr_orientation) P.gr_orientation[4 * i + k] = 0.f;
        if (P.g2_orientation) P.g2_orientation[4 * i + k] = 0.f;
      }
      if (P.gr_inv_scale) P.gr_inv_scale[i] = 0.f;
      if (P.g2_scale) P.g2_scale[i] = 0.f;
      if (P.loss_sum) P.loss_sum[i] = 0.f;
      if (P.n_overlap) P.n_overlap[i] = 0.f;
      if (P.point_sum) P.point_sum[i] = 0.f;
    }
  }
  const float s = P.scale[b];
#pragma unroll
  for (int k = 0; k < 3; ++k) P.g_position[3 * b + k] = gp[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) P.g_orientation[4 * b + k] = gq[k];
  P.g_scale[b] = gs - gis / (s * s); /* d(1/s)/ds */
  P.loss[b] += l;
}

/* ------------------------------------------------------------------------------------------
 * Point constraint (estimation/simple_setup.py:164-175, estimation/losses.py:138-153): per hypothesis
 *     loss = weight * | q (x) (source,0) (x) conj(q) - target |
 * with q the UN-NORMALISED orientation parameter (quaternion_utils.py:37-54 does not normalise: a
 * non-unit q also scales by |q|^2, and the gradient below is that function's).  For q = (u, w):
 *     y   = (w^2 - |u|^2) s + 2 (u.s) u + 2 w (u x s),    d = y - t,   g = d / |d|  (0 at d = 0, torch's
 *                                                                         subgradient of the norm)
 *     dL/du = 2 [ (u.s) g + (g.u) s - (g.s) u + w (s x g) ],   dL/dw = 2 [ w (g.s) + g.(u x s) ]
 * Both outputs are ACCUMULATED (they are sdfr_hypothesis_step's g_orientation_raw / loss_extra).
 * ---------------------------------------------------------------------------------------- */
struct ConstraintParams {
  const float* __restrict__ orientation; /* [B,4] un-normalised */
  int batch;
  float source[3], target[3], weight;
  float* __restrict__ g_orientation_raw; /* [B,4] += or NULL */
  float* __restrict__ loss;              /* [B]   += or NULL */
};

__global__ void __launch_bounds__(128)
sdfr_point_constraint_kernel(const __grid_constant__ ConstraintParams P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.batch) return;
  const float u[3] = {P.orientation[4 * b], P.orientation[4 * b + 1], P.orientation[4 * b + 2]};
  const float w = P.orientation[4 * b + 3];
  const float* s = P.source;
  const float uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
  const float us = u[0] * s[0] + u[1] * s[1] + u[2] * s[2];
  const float c[3] = {u[1] * s[2] - u[2] * s[1], u[2] * s[0] - u[0] * s[2], u[0] * s[1] - u[1] * s[0]}; /* u x s */
  float d[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) d[k] = (w * w - uu) * s[k] + 2.0f * us * u[k] + 2.0f * w * c[k] - P.target[k];
  const float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (P.loss) P.loss[b] += P.weight * nrm;
  if (!P.g_orientation_raw) return;
  const float inv = nrm > 0.0f ? P.weight / nrm : 0.0f;
  const float g[3] = {d[0] * inv, d[1] * inv, d[2] * inv};
  const float gs = g[0] * s[0] + g[1] * s[1] + g[2] * s[2];
  const float gu = g[0] * u[0] + g[1] * u[1] + g[2] * u[2];
  const float gc = g[0] * c[0] + g[1] * c[1] + g[2] * c[2];
  const float sg[3] = {s[1] * g[2] - s[2] * g[1], s[2] * g[0] - s[0] * g[2], s[0] * g[1] - s[1] * g[0]}; /* s x g */
#pragma unroll
  for (int k = 0; k < 3; ++k)
    P.g_orientation_raw[4 * b + k] += 2.0f * (us * g[k] + gu * s[k] - gs * u[k] + w * sg[k]);
  P.g_orientation_raw[4 * b + 3] += 2.0f * (w * gs + gc);
}

#endif /* SDFR_STEP_CUH_ */
