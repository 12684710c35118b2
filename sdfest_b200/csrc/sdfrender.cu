/*
 * sdfrender.cu -- sm_100a kernels and C ABI of libsdfrender.so (see include/sdfrender.h).
 *
 * Replaces the reference's sdf_renderer_cpp extension (sdf_renderer.cpp:42-91,
 * sdf_renderer_cuda.cu:241-556).  Design (DESIGN.md has the long version):
 *
 *  - one launch renders a whole BATCH of hypotheses (blockIdx.z), the reference renders one
 *    object per launch plus 1 (fwd) / 4 (bwd) memset launches;
 *  - a CTA owns a 32x8 pixel tile; warp w owns the 8x4 sub-tile so that the 32 rays of a warp
 *    stay spatially coherent (their 8-corner gathers fall into 1-4 L1 lines) and the depth
 *    stores fill whole 32-byte sectors;
 *  - warp 0 builds the per-hypothesis Frame (rotation, object-frame origin, slab constants and
 *    the screen rectangle of the projected box) once per CTA; CTAs / warps outside the
 *    rectangle only zero-fill -- 64-99 % of the pixels of the reference workloads;
 *  - un-normalised ray components come from two per-CTA tables evaluated in double exactly as
 *    the reference does per thread (cu:146-147): 40 double divisions per CTA instead of 512;
 *  - the grid is read through the read-only path (LDG.E.CONSTANT) and stays L1/L2 resident
 *    (1 MiB at 64^3; 126 MB of L2 hold 64 distinct hypothesis grids);
 *  - backward: pose/scale gradients reduce warp-shuffle -> shared memory -> 8 atomics per CTA
 *    (the reference issues 8 same-address atomics per hit pixel, cu:459-466); SDF gradients
 *    scatter with fire-and-forget RED.ADD.F32; `flags` prunes what is not needed;
 *  - the fused render-and-compare pair evaluates the masked-L1 depth loss inside the render
 *    kernel and rebuilds the loss gradient inside the backward kernel, so no grad_depth image
 *    is ever materialised.
 *
 * No tensor cores: the path is a dependent gather chain, not a contraction.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/sdfrender.h"
#include "sdfr_core.cuh"

namespace {

using namespace sdfr;

constexpr int kTileW = 32;
constexpr int kTileH = 8;
constexpr int kThreads = kTileW * kTileH;  // 256 = 8 warps of 8x4 pixels
constexpr int kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;

thread_local char g_err[256] = "";

int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

struct Pose {
  const float* __restrict__ position;     // [B,3]
  const float* __restrict__ orientation;  // [B,4]
  const float* __restrict__ inv_scale;    // [B]
};

struct FwdParams {
  const float* __restrict__ sdf;
  long long sdf_stride;
  Pose pose;
  Grid grid;
  Camera cam;
  float threshold;
  float* __restrict__ depth;  // [B,H,W]
  int z_offset;               // first hypothesis of this launch (gridDim.z chunking)
  // compare
  const float* __restrict__ depth_obs;
  long long obs_stride;
  float* __restrict__ loss_sum;
  float* __restrict__ n_overlap;
  // stats
  unsigned long long* __restrict__ stats;
};

struct BwdParams {
  const float* __restrict__ grad_depth;  // PLAIN / COMPOSITE
  const float* __restrict__ depth;
  const float* __restrict__ sdf;
  long long sdf_stride;
  Pose pose;
  Grid grid;
  Camera cam;
  float* __restrict__ grad_sdf;
  long long grad_sdf_stride;
  float* __restrict__ grad_position;
  float* __restrict__ grad_orientation;
  float* __restrict__ grad_inv_scale;
  unsigned flags;
  int z_offset;
  // compare
  const float* __restrict__ depth_obs;
  long long obs_stride;
  const float* __restrict__ n_overlap;
  const float* __restrict__ upstream;
  // composite
  const int* __restrict__ winner;
  int n_objects;
};

/* Built by warp 0: pose part by every lane (registers), the 8 box corners by lanes 0-7. */
__device__ __forceinline__ void build_frame(Frame& smemF, const Pose& pose, int b,
                                            const Camera& cam, int lane) {
  Frame F;
  frame_pose(F, pose.position + 3 * b, pose.orientation + 4 * b, pose.inv_scale + b);
  float col = 0.f, row = 0.f;
  const bool ok = project_corner(F, cam, lane & 7, col, row);
  const bool all_ok = __all_sync(kFull, ok);
  float cmin = col, cmax = col, rmin = row, rmax = row;
#pragma unroll
  for (int o = 4; o >= 1; o >>= 1) {
    cmin = fminf(cmin, __shfl_xor_sync(kFull, cmin, o));
    cmax = fmaxf(cmax, __shfl_xor_sync(kFull, cmax, o));
    rmin = fminf(rmin, __shfl_xor_sync(kFull, rmin, o));
    rmax = fmaxf(rmax, __shfl_xor_sync(kFull, rmax, o));
  }
  if (lane == 0) {
    frame_rect(F, cam, all_ok, cmin, cmax, rmin, rmax);
    smemF = F;
  }
}

/* CTA prologue shared by all kernels: Frame (warp 0) and the ray tables (warps 1, 2). */
__device__ __forceinline__ void cta_prologue(Frame& F, float* colx, float* rowy, const Pose& pose,
                                             int b, const Camera& cam, int bx0, int by0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    build_frame(F, pose, b, cam, lane);
  } else if (warp == 1) {
    colx[lane] = pixel_dx(bx0 + lane, cam.cx, cam.fx);
  } else if (warp == 2 && lane < kTileH) {
    rowy[lane] = pixel_dy(by0 + lane, cam.cy, cam.fy);
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

/* ------------------------------------------------------------------------------------------
 * Forward (replaces sdf_renderer_cuda_forward_kernel, cu:241-298).
 * ---------------------------------------------------------------------------------------- */
template <bool COMPARE, bool STATS>
__global__ void __launch_bounds__(kThreads)
sdfr_forward_kernel(const __grid_constant__ FwdParams P) {
  __shared__ Frame F;
  __shared__ float colx[kTileW];
  __shared__ float rowy[kTileH];
  __shared__ float red[2][kWarps];

  const int b = blockIdx.z + P.z_offset;
  const int bx0 = blockIdx.x * kTileW, by0 = blockIdx.y * kTileH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lx = ((warp & 3) << 3) + (lane & 7);
  const int ly = ((warp >> 2) << 2) + (lane >> 3);
  const int px = bx0 + lx, py = by0 + ly;
  const bool inside = px < P.cam.W && py < P.cam.H;
  float* __restrict__ out = P.depth + ((size_t)b * P.cam.H + py) * P.cam.W + px;

  cta_prologue(F, colx, rowy, P.pose, b, P.cam, bx0, by0);
  __syncthreads();

  /* CTA outside the projected box: nothing can be hit (cu:294-296 writes 0 for these) */
  if (bx0 >= F.x1 || bx0 + kTileW <= F.x0 || by0 >= F.y1 || by0 + kTileH <= F.y0) {
    if (inside) *out = 0.0f;
    return;
  }

  float z = 0.0f;
  int steps = 0;
  bool entered = false, capped = false;
  if (inside && px >= F.x0 && px < F.x1 && py >= F.y0 && py < F.y1) {
    const Ray r = make_ray(F, colx[lx], rowy[ly]);
    float t_min, t_max;
    if (ray_box(F, r, t_min, t_max)) {
      entered = true;
      const float* __restrict__ g = P.sdf + (size_t)b * P.sdf_stride;
      z = march(g, P.grid, F, r, t_min, t_max, P.threshold, steps, capped);
    }
  }
  if (inside) *out = z;

  if (STATS) {
    const unsigned s = __reduce_add_sync(kFull, (unsigned)steps);
    const unsigned e = __popc(__ballot_sync(kFull, entered));
    const unsigned h = __popc(__ballot_sync(kFull, z != 0.0f));
    const unsigned c = __popc(__ballot_sync(kFull, capped));
    if (lane == 0) {
      if (s) atomicAdd(P.stats + 0, (unsigned long long)s);
      if (e) atomicAdd(P.stats + 1, (unsigned long long)e);
      if (h) atomicAdd(P.stats + 2, (unsigned long long)h);
      if (c) atomicAdd(P.stats + 3, (unsigned long long)c);
    }
  }

  if (COMPARE) {
    /* masked L1 against the observation (estimation/simple_setup.py:125-131) */
    float err = 0.0f, cnt = 0.0f;
    if (z > 0.0f) {
      const float obs =
          __ldg(P.depth_obs + (size_t)b * P.obs_stride + (size_t)py * P.cam.W + px);
      if (obs > 0.0f) {
        err = fabsf(z - obs);
        cnt = 1.0f;
      }
    }
    err = warp_sum(err);
    cnt = warp_sum(cnt);
    if (lane == 0) {
      red[0][warp] = err;
      red[1][warp] = cnt;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
      float s = 0.0f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += red[threadIdx.x][w];
      if (s != 0.0f) atomicAdd((threadIdx.x == 0 ? P.loss_sum : P.n_overlap) + b, s);
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Backward (replaces sdf_renderer_cuda_backward_kernel, cu:300-468).
 * MODE 0: explicit grad_depth.  MODE 1: fused compare (gradient of the masked L1 rebuilt here).
 * ---------------------------------------------------------------------------------------- */
template <int MODE, bool WANT_SDF, bool WANT_POSE>
__global__ void __launch_bounds__(kThreads)
sdfr_backward_kernel(const __grid_constant__ BwdParams P) {
  __shared__ Frame F;
  __shared__ float colx[kTileW];
  __shared__ float rowy[kTileH];
  __shared__ float red[kWarps][8];
  __shared__ float coef_s;

  const int b = blockIdx.z + P.z_offset;
  const int bx0 = blockIdx.x * kTileW, by0 = blockIdx.y * kTileH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lx = ((warp & 3) << 3) + (lane & 7);
  const int ly = ((warp >> 2) << 2) + (lane >> 3);
  const int px = bx0 + lx, py = by0 + ly;
  const bool inside = px < P.cam.W && py < P.cam.H;
  const size_t pix = ((size_t)b * P.cam.H + py) * P.cam.W + px;

  /* upstream gradient of this pixel */
  float z = 0.0f, gup = 0.0f;
  if (MODE == 1 && threadIdx.x == 0) {
    const float n = __ldg(P.n_overlap + b);
    const float u = P.upstream ? __ldg(P.upstream + b) : 1.0f;
    coef_s = n > 0.0f ? u / n : 0.0f;
  }
  if (inside) z = __ldg(P.depth + pix);
  if (MODE == 0) {
    if (z != 0.0f) gup = __ldg(P.grad_depth + pix);
  } else {
    if (z > 0.0f) {
      const float obs =
          __ldg(P.depth_obs + (size_t)b * P.obs_stride + (size_t)py * P.cam.W + px);
      if (obs > 0.0f) gup = (z > obs) ? 1.0f : ((z < obs) ? -1.0f : 0.0f);
    }
  }
  const bool active = (z != 0.0f) && (gup != 0.0f);
  if (!__syncthreads_or(active)) return; /* also orders coef_s */
  if (MODE == 1) gup *= coef_s;

  cta_prologue(F, colx, rowy, P.pose, b, P.cam, bx0, by0);
  __syncthreads();

  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0f;

  if (active && gup != 0.0f) {
    const Ray r = make_ray(F, colx[lx], rowy[ly]);
    const float* __restrict__ g = P.sdf + (size_t)b * P.sdf_stride;
    PixelGrad pg;
    pixel_backward<WANT_SDF, WANT_POSE>(g, P.grid, F, r, z, gup,
                                        (P.flags & SDFR_SDF_GRAD_EXACT) != 0, pg);
    if (WANT_SDF) {
      float* __restrict__ gs = P.grad_sdf + (size_t)b * P.grad_sdf_stride + pg.base;
      const int R = P.grid.R, R2 = P.grid.R2;
      atomicAdd(gs, pg.w[0]);
      atomicAdd(gs + 1, pg.w[1]);
      atomicAdd(gs + R, pg.w[2]);
      atomicAdd(gs + R + 1, pg.w[3]);
      atomicAdd(gs + R2, pg.w[4]);
      atomicAdd(gs + R2 + 1, pg.w[5]);
      atomicAdd(gs + R2 + R, pg.w[6]);
      atomicAdd(gs + R2 + R + 1, pg.w[7]);
    }
    if (WANT_POSE) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = pg.pose[i] * gup;
    }
  }

  if (WANT_POSE) {
    /* warp shuffle -> shared -> 8 atomics per CTA (the reference: 8 per hit pixel, cu:459-466) */
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = warp_sum(acc[i]);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
      float s = 0.0f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += red[w][threadIdx.x];
      const int i = threadIdx.x;
      if (s != 0.0f) {
        if (i < 3) {
          if (P.flags & SDFR_GRAD_POSITION) atomicAdd(P.grad_position + 3 * b + i, s);
        } else if (i < 7) {
          if (P.flags & SDFR_GRAD_ORIENTATION) atomicAdd(P.grad_orientation + 4 * b + (i - 3), s);
        } else {
          if (P.flags & SDFR_GRAD_INV_SCALE) atomicAdd(P.grad_inv_scale + b, s);
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Multi-object composite: one depth map, per-pixel minimum positive depth over K objects.
 * Frames of up to kMaxObjPerPass objects live in shared memory; a pixel only traces the
 * objects whose projected rectangle contains it.
 * ---------------------------------------------------------------------------------------- */
constexpr int kMaxObjPerPass = 32;

__global__ void __launch_bounds__(kThreads)
sdfr_forward_composite_kernel(const __grid_constant__ FwdParams P, int n_objects,
                              int* __restrict__ winner_out) {
  __shared__ Frame Fs[kMaxObjPerPass];
  __shared__ float colx[kTileW];
  __shared__ float rowy[kTileH];

  const int bx0 = blockIdx.x * kTileW, by0 = blockIdx.y * kTileH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lx = ((warp & 3) << 3) + (lane & 7);
  const int ly = ((warp >> 2) << 2) + (lane >> 3);
  const int px = bx0 + lx, py = by0 + ly;
  const bool inside = px < P.cam.W && py < P.cam.H;

  if (warp == 1) colx[lane] = pixel_dx(bx0 + lane, P.cam.cx, P.cam.fx);
  if (warp == 2 && lane < kTileH) rowy[lane] = pixel_dy(by0 + lane, P.cam.cy, P.cam.fy);

  float best = 0.0f;
  int win = -1;
  for (int k0 = 0; k0 < n_objects; k0 += kMaxObjPerPass) {
    const int nk = min(kMaxObjPerPass, n_objects - k0);
    __syncthreads(); /* previous pass done with Fs; also publishes the tables */
    for (int k = warp; k < nk; k += kWarps) build_frame(Fs[k], P.pose, k0 + k, P.cam, lane);
    __syncthreads();
    if (!inside) continue;
    for (int k = 0; k < nk; ++k) {
      const Frame& F = Fs[k];
      if (px < F.x0 || px >= F.x1 || py < F.y0 || py >= F.y1) continue;
      const Ray r = make_ray(F, colx[lx], rowy[ly]);
      float t_min, t_max;
      if (!ray_box(F, r, t_min, t_max)) continue;
      int steps;
      bool capped;
      const float z = march(P.sdf + (size_t)(k0 + k) * P.sdf_stride, P.grid, F, r, t_min,
                            t_max, P.threshold, steps, capped);
      if (z > 0.0f && (win < 0 || z < best)) {
        best = z;
        win = k0 + k;
      }
    }
  }
  if (inside) {
    const size_t pix = (size_t)py * P.cam.W + px;
    P.depth[pix] = best;
    winner_out[pix] = win;
  }
}

template <bool WANT_SDF, bool WANT_POSE>
__global__ void __launch_bounds__(kThreads)
sdfr_backward_composite_kernel(const __grid_constant__ BwdParams P) {
  __shared__ float colx[kTileW];
  __shared__ float rowy[kTileH];

  const int bx0 = blockIdx.x * kTileW, by0 = blockIdx.y * kTileH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lx = ((warp & 3) << 3) + (lane & 7);
  const int ly = ((warp >> 2) << 2) + (lane >> 3);
  const int px = bx0 + lx, py = by0 + ly;
  const bool inside = px < P.cam.W && py < P.cam.H;
  const size_t pix = (size_t)py * P.cam.W + px;

  float z = 0.0f, gup = 0.0f;
  int win = -1;
  if (inside) {
    z = __ldg(P.depth + pix);
    if (z != 0.0f) {
      win = __ldg(P.winner + pix);
      gup = __ldg(P.grad_depth + pix);
    }
  }
  const bool active = z != 0.0f && gup != 0.0f && win >= 0 && win < P.n_objects;
  if (!__syncthreads_or(active)) return;
  if (warp == 1) colx[lane] = pixel_dx(bx0 + lane, P.cam.cx, P.cam.fx);
  if (warp == 2 && lane < kTileH) rowy[lane] = pixel_dy(by0 + lane, P.cam.cy, P.cam.fy);
  __syncthreads();

  /* a warp's 8x4 pixels almost always belong to one object: loop over the distinct winners */
  unsigned todo = __ballot_sync(kFull, active);
  while (todo) {
    const int leader = __ffs(todo) - 1;
    const int k = __shfl_sync(kFull, win, leader);
    const bool mine = active && win == k;
    todo &= ~__ballot_sync(kFull, mine);

    Frame F; /* every lane builds the pose part in registers; no rectangle needed */
    frame_pose(F, P.pose.position + 3 * k, P.pose.orientation + 4 * k, P.pose.inv_scale + k);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
    if (mine) {
      const Ray r = make_ray(F, colx[lx], rowy[ly]);
      const float* __restrict__ g = P.sdf + (size_t)k * P.sdf_stride;
      PixelGrad pg;
      pixel_backward<WANT_SDF, WANT_POSE>(g, P.grid, F, r, z, gup,
                                          (P.flags & SDFR_SDF_GRAD_EXACT) != 0, pg);
      if (WANT_SDF) {
        float* __restrict__ gs = P.grad_sdf + (size_t)k * P.grad_sdf_stride + pg.base;
        const int R = P.grid.R, R2 = P.grid.R2;
        atomicAdd(gs, pg.w[0]);
        atomicAdd(gs + 1, pg.w[1]);
        atomicAdd(gs + R, pg.w[2]);
        atomicAdd(gs + R + 1, pg.w[3]);
        atomicAdd(gs + R2, pg.w[4]);
        atomicAdd(gs + R2 + 1, pg.w[5]);
        atomicAdd(gs + R2 + R, pg.w[6]);
        atomicAdd(gs + R2 + R + 1, pg.w[7]);
      }
      if (WANT_POSE) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = pg.pose[i] * gup;
      }
    }
    if (WANT_POSE) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = warp_sum(acc[i]);
      if (lane < 8) {
        float s = acc[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) s = (lane == i) ? acc[i] : s;
        if (s != 0.0f) {
          if (lane < 3) {
            if (P.flags & SDFR_GRAD_POSITION) atomicAdd(P.grad_position + 3 * k + lane, s);
          } else if (lane < 7) {
            if (P.flags & SDFR_GRAD_ORIENTATION)
              atomicAdd(P.grad_orientation + 4 * k + (lane - 3), s);
          } else {
            if (P.flags & SDFR_GRAD_INV_SCALE) atomicAdd(P.grad_inv_scale + k, s);
          }
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Host side
 * ---------------------------------------------------------------------------------------- */
int check_common(const float* sdf, int R, long long sdf_stride, const float* pos,
                 const float* quat, const float* inv_scale, int batch, int W, int H) {
  if (batch < 0 || W < 0 || H < 0) return fail(SDFR_E_SHAPE, "negative batch/width/height");
  if (R < 2 || R > 1024) return fail(SDFR_E_SHAPE, "resolution must be in [2, 1024]");
  if (sdf_stride < 0) return fail(SDFR_E_SHAPE, "negative sdf_stride");
  if (W > (1 << 20) || H > (1 << 19)) return fail(SDFR_E_SHAPE, "image too large");
  if (batch == 0 || W == 0 || H == 0) return 0;
  if (!sdf || !pos || !quat || !inv_scale) return fail(SDFR_E_NULL, "NULL input pointer");
  return 0;
}

dim3 tile_grid(int W, int H, int z) {
  return dim3((W + kTileW - 1) / kTileW, (H + kTileH - 1) / kTileH, z);
}

int zero_async(void* p, size_t bytes, cudaStream_t s) {
  if (!p || bytes == 0) return 0;
  const cudaError_t e = cudaMemsetAsync(p, 0, bytes, s);
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "cudaMemsetAsync: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

template <bool COMPARE, bool STATS>
int launch_forward(FwdParams P, int batch, cudaStream_t s) {
  for (int z0 = 0; z0 < batch; z0 += 65535) {
    P.z_offset = z0;
    const int nz = batch - z0 < 65535 ? batch - z0 : 65535;
    sdfr_forward_kernel<COMPARE, STATS><<<tile_grid(P.cam.W, P.cam.H, nz), kThreads, 0, s>>>(P);
  }
  return check_launch("sdfr_forward_kernel");
}

template <int MODE>
int launch_backward(BwdParams P, int batch, cudaStream_t s) {
  const bool want_sdf = (P.flags & SDFR_GRAD_SDF) != 0;
  const bool want_pose =
      (P.flags & (SDFR_GRAD_POSITION | SDFR_GRAD_ORIENTATION | SDFR_GRAD_INV_SCALE)) != 0;
  if (!want_sdf && !want_pose) return 0;
  for (int z0 = 0; z0 < batch; z0 += 65535) {
    P.z_offset = z0;
    const int nz = batch - z0 < 65535 ? batch - z0 : 65535;
    const dim3 grid = tile_grid(P.cam.W, P.cam.H, nz);
    if (want_sdf && want_pose)
      sdfr_backward_kernel<MODE, true, true><<<grid, kThreads, 0, s>>>(P);
    else if (want_sdf)
      sdfr_backward_kernel<MODE, true, false><<<grid, kThreads, 0, s>>>(P);
    else
      sdfr_backward_kernel<MODE, false, true><<<grid, kThreads, 0, s>>>(P);
  }
  return check_launch("sdfr_backward_kernel");
}

int check_backward_outputs(unsigned flags, const float* gs, long long gs_stride, const float* gp,
                           const float* gq, const float* gi) {
  if (flags & ~(SDFR_GRAD_ALL | SDFR_SDF_GRAD_EXACT | SDFR_ZERO_GRADS))
    return fail(SDFR_E_FLAGS, "unknown flag bits");
  if (gs_stride < 0) return fail(SDFR_E_SHAPE, "negative grad_sdf_stride");
  if (((flags & SDFR_GRAD_SDF) && !gs) || ((flags & SDFR_GRAD_POSITION) && !gp) ||
      ((flags & SDFR_GRAD_ORIENTATION) && !gq) || ((flags & SDFR_GRAD_INV_SCALE) && !gi))
    return fail(SDFR_E_NULL, "NULL gradient buffer for a requested gradient");
  return 0;
}

int zero_grads(unsigned flags, int R, int batch, float* gs, long long gs_stride, float* gp,
               float* gq, float* gi, cudaStream_t s) {
  if (!(flags & SDFR_ZERO_GRADS)) return 0;
  int rc = 0;
  const size_t grid_elems = (size_t)R * R * R;
  if (flags & SDFR_GRAD_SDF) {
    if (gs_stride == 0 || (size_t)gs_stride == grid_elems) {
      const size_t n = gs_stride == 0 ? grid_elems : grid_elems * batch;
      rc = zero_async(gs, n * sizeof(float), s);
    } else {
      for (int b = 0; b < batch && rc == 0; ++b)
        rc = zero_async(gs + (size_t)b * gs_stride, grid_elems * sizeof(float), s);
    }
  }
  if (rc == 0 && (flags & SDFR_GRAD_POSITION)) rc = zero_async(gp, sizeof(float) * 3 * batch, s);
  if (rc == 0 && (flags & SDFR_GRAD_ORIENTATION)) rc = zero_async(gq, sizeof(float) * 4 * batch, s);
  if (rc == 0 && (flags & SDFR_GRAD_INV_SCALE)) rc = zero_async(gi, sizeof(float) * batch, s);
  return rc;
}

FwdParams fwd_params(const float* sdf, int R, long long sdf_stride, const float* pos,
                     const float* quat, const float* inv_scale, int W, int H, float cx, float cy,
                     float fx, float fy, float threshold, float* depth) {
  FwdParams P;
  memset(&P, 0, sizeof(P));
  P.sdf = sdf;
  P.sdf_stride = sdf_stride;
  P.pose = Pose{pos, quat, inv_scale};
  P.grid = make_grid(R);
  P.cam = Camera{W, H, cx, cy, fx, fy};
  P.threshold = threshold;
  P.depth = depth;
  return P;
}

BwdParams bwd_params(const float* depth, const float* sdf, int R, long long sdf_stride,
                     const float* pos, const float* quat, const float* inv_scale, int W, int H,
                     float cx, float cy, float fx, float fy, float* gs, long long gs_stride,
                     float* gp, float* gq, float* gi, unsigned flags) {
  BwdParams P;
  memset(&P, 0, sizeof(P));
  P.depth = depth;
  P.sdf = sdf;
  P.sdf_stride = sdf_stride;
  P.pose = Pose{pos, quat, inv_scale};
  P.grid = make_grid(R);
  P.cam = Camera{W, H, cx, cy, fx, fy};
  P.grad_sdf = gs;
  P.grad_sdf_stride = gs_stride;
  P.grad_position = gp;
  P.grad_orientation = gq;
  P.grad_inv_scale = gi;
  P.flags = flags;
  return P;
}

}  // namespace

extern "C" {

int sdfr_abi_version(void) { return SDFR_ABI_VERSION; }
const char* sdfr_last_error(void) { return g_err; }
const char* sdfr_build_info(void) {
  return "libsdfrender sm_100a; tile 32x8, 8 warps of 8x4 px; fp32; nvcc " __DATE__;
}
int sdfr_max_steps(void) { return sdfr::kMaxSteps; }

int sdfr_forward(const float* sdf, int R, long long sdf_stride, const float* pos,
                 const float* quat, const float* inv_scale, int batch, int W, int H, float cx,
                 float cy, float fx, float fy, float threshold, float* depth, void* stream) {
  if (int rc = check_common(sdf, R, sdf_stride, pos, quat, inv_scale, batch, W, H)) return rc;
  if (batch == 0 || W == 0 || H == 0) return 0;
  if (!depth) return fail(SDFR_E_NULL, "depth is NULL");
  FwdParams P = fwd_params(sdf, R, sdf_stride, pos, quat, inv_scale, W, H, cx, cy, fx, fy,
                           threshold, depth);
  return launch_forward<false, false>(P, batch, (cudaStream_t)stream);
}

int sdfr_forward_stats(const float* sdf, int R, long long sdf_stride, const float* pos,
                       const float* quat, const float* inv_scale, int batch, int W, int H,
                       float cx, float cy, float fx, float fy, float threshold, float* depth,
                       unsigned long long* stats, void* stream) {
  if (int rc = check_common(sdf, R, sdf_stride, pos, quat, inv_scale, batch, W, H)) return rc;
  if (batch == 0 || W == 0 || H == 0) return 0;
  if (!depth || !stats) return fail(SDFR_E_NULL, "depth or stats is NULL");
  FwdParams P = fwd_params(sdf, R, sdf_stride, pos, quat, inv_scale, W, H, cx, cy, fx, fy,
                           threshold, depth);
  P.stats = stats;
  return launch_forward<false, true>(P, batch, (cudaStream_t)stream);
}

int sdfr_backward(const float* grad_depth, const float* depth, const float* sdf, int R,
                  long long sdf_stride, const float* pos, const float* quat,
                  const float* inv_scale, int batch, int W, int H, float cx, float cy, float fx,
                  float fy, float* gs, long long gs_stride, float* gp, float* gq, float* gi,
                  unsigned flags, void* stream) {
  if (int rc = check_common(sdf, R, sdf_stride, pos, quat, inv_scale, batch, W, H)) return rc;
  if (int rc = check_backward_outputs(flags, gs, gs_stride, gp, gq, gi)) return rc;
  if (batch == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = zero_grads(flags, R, batch, gs, gs_stride, gp, gq, gi, s)) return rc;
  if (W == 0 || H == 0) return 0;
  if (!grad_depth || !depth) return fail(SDFR_E_NULL, "grad_depth or depth is NULL");
  BwdParams P = bwd_params(depth, sdf, R, sdf_stride, pos, quat, inv_scale, W, H, cx, cy, fx, fy,
                           gs, gs_stride, gp, gq, gi, flags);
  P.grad_depth = grad_depth;
  return launch_backward<0>(P, batch, s);
}

int sdfr_compare_forward(const float* sdf, int R, long long sdf_stride, const float* pos,
                         const float* quat, const float* inv_scale, int batch, int W, int H,
                         float cx, float cy, float fx, float fy, float threshold,
                         const float* depth_obs, long long obs_stride, float* depth,
                         float* loss_sum, float* n_overlap, unsigned flags, void* stream) {
  if (int rc = check_common(sdf, R, sdf_stride, pos, quat, inv_scale, batch, W, H)) return rc;
  if (flags & ~SDFR_ZERO_GRADS) return fail(SDFR_E_FLAGS, "unknown flag bits");
  if (obs_stride < 0) return fail(SDFR_E_SHAPE, "negative obs_stride");
  if (batch == 0) return 0;
  if (!loss_sum || !n_overlap) return fail(SDFR_E_NULL, "loss_sum or n_overlap is NULL");
  cudaStream_t s = (cudaStream_t)stream;
  if (flags & SDFR_ZERO_GRADS) {
    if (int rc = zero_async(loss_sum, sizeof(float) * batch, s)) return rc;
    if (int rc = zero_async(n_overlap, sizeof(float) * batch, s)) return rc;
  }
  if (W == 0 || H == 0) return 0;
  if (!depth || !depth_obs) return fail(SDFR_E_NULL, "depth or depth_obs is NULL");
  FwdParams P = fwd_params(sdf, R, sdf_stride, pos, quat, inv_scale, W, H, cx, cy, fx, fy,
                           threshold, depth);
  P.depth_obs = depth_obs;
  P.obs_stride = obs_stride;
  P.loss_sum = loss_sum;
  P.n_overlap = n_overlap;
  return launch_forward<true, false>(P, batch, s);
}

int sdfr_compare_backward(const float* depth, const float* depth_obs, long long obs_stride,
                          const float* n_overlap, const float* upstream, const float* sdf, int R,
                          long long sdf_stride, const float* pos, const float* quat,
                          const float* inv_scale, int batch, int W, int H, float cx, float cy,
                          float fx, float fy, float* gs, long long gs_stride, float* gp,
                          float* gq, float* gi, unsigned flags, void* stream) {
  if (int rc = check_common(sdf, R, sdf_stride, pos, quat, inv_scale, batch, W, H)) return rc;
  if (int rc = check_backward_outputs(flags, gs, gs_stride, gp, gq, gi)) return rc;
  if (obs_stride < 0) return fail(SDFR_E_SHAPE, "negative obs_stride");
  if (batch == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = zero_grads(flags, R, batch, gs, gs_stride, gp, gq, gi, s)) return rc;
  if (W == 0 || H == 0) return 0;
  if (!depth || !depth_obs || !n_overlap)
    return fail(SDFR_E_NULL, "depth, depth_obs or n_overlap is NULL");
  BwdParams P = bwd_params(depth, sdf, R, sdf_stride, pos, quat, inv_scale, W, H, cx, cy, fx, fy,
                           gs, gs_stride, gp, gq, gi, flags);
  P.depth_obs = depth_obs;
  P.obs_stride = obs_stride;
  P.n_overlap = n_overlap;
  P.upstream = upstream;
  return launch_backward<1>(P, batch, s);
}

int sdfr_forward_composite(const float* sdf, int R, long long sdf_stride, const float* pos,
                           const float* quat, const float* inv_scale, int n_objects, int W,
                           int H, float cx, float cy, float fx, float fy, float threshold,
                           float* depth, int* winner, void* stream) {
  if (int rc = check_common(sdf, R, sdf_stride, pos, quat, inv_scale, n_objects, W, H)) return rc;
  if (W == 0 || H == 0) return 0;
  if (!depth || !winner) return fail(SDFR_E_NULL, "depth or winner is NULL");
  FwdParams P = fwd_params(sdf, R, sdf_stride, pos, quat, inv_scale, W, H, cx, cy, fx, fy,
                           threshold, depth);
  sdfr_forward_composite_kernel<<<tile_grid(W, H, 1), kThreads, 0, (cudaStream_t)stream>>>(
      P, n_objects, winner);
  return check_launch("sdfr_forward_composite_kernel");
}

int sdfr_backward_composite(const float* grad_depth, const float* depth, const int* winner,
                            const float* sdf, int R, long long sdf_stride, const float* pos,
                            const float* quat, const float* inv_scale, int n_objects, int W,
                            int H, float cx, float cy, float fx, float fy, float* gs,
                            long long gs_stride, float* gp, float* gq, float* gi, unsigned flags,
                            void* stream) {
  if (int rc = check_common(sdf, R, sdf_stride, pos, quat, inv_scale, n_objects, W, H)) return rc;
  if (int rc = check_backward_outputs(flags, gs, gs_stride, gp, gq, gi)) return rc;
  if (n_objects == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = zero_grads(flags, R, n_objects, gs, gs_stride, gp, gq, gi, s)) return rc;
  if (W == 0 || H == 0) return 0;
  if (!grad_depth || !depth || !winner)
    return fail(SDFR_E_NULL, "grad_depth, depth or winner is NULL");
  BwdParams P = bwd_params(depth, sdf, R, sdf_stride, pos, quat, inv_scale, W, H, cx, cy, fx, fy,
                           gs, gs_stride, gp, gq, gi, flags);
  P.grad_depth = grad_depth;
  P.winner = winner;
  P.n_objects = n_objects;
  const bool want_sdf = (flags & SDFR_GRAD_SDF) != 0;
  const bool want_pose =
      (flags & (SDFR_GRAD_POSITION | SDFR_GRAD_ORIENTATION | SDFR_GRAD_INV_SCALE)) != 0;
  if (!want_sdf && !want_pose) return 0;
  const dim3 grid = tile_grid(W, H, 1);
  if (want_sdf && want_pose)
    sdfr_backward_composite_kernel<true, true><<<grid, kThreads, 0, s>>>(P);
  else if (want_sdf)
    sdfr_backward_composite_kernel<true, false><<<grid, kThreads, 0, s>>>(P);
  else
    sdfr_backward_composite_kernel<false, true><<<grid, kThreads, 0, s>>>(P);
  return check_launch("sdfr_backward_composite_kernel");
}

}  // extern "C"
