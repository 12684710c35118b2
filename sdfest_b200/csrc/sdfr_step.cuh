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
    for (int i = 0; i < 4; ++i) g[3 + i] = (gq[i] - q[i] * dot) / nrm;
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

#endif /* SDFR_STEP_CUH_ */
