// Gather-bandwidth micro-benchmark ("L2 peak is not in MEASURED_PEAKS.json -- measure with a
// gather microbenchmark"): how many 8-corner trilinear samples per second can a B200 deliver when the
// kernel does NOTHING but the gathers (no dependent march, no ALU work), for the access patterns of the
// SDF renderer?  The numbers are the denominators of bench.py's roofline (profiles/r02_gather_peaks.json).
//
//   pattern A  "coherent"  the renderer's own pattern: the 32 lanes of a warp hold an 8x4 pixel footprint
//              that covers a few neighbouring cells (FOOT_X x FOOT_Y cells; ~4 pixels per cell at
//              640x480 / 64^3) and walks through its grid one cell per step; served by L1 (ld.global.nc)
//   pattern B  the same addresses with ld.global.cg (L1 bypassed): what L2 alone sustains for pattern A
//   pattern C  "random"    every lane samples a random cell of 64 resident 1 MiB grids: L2 sector rate
//
// layouts: dense [64][64][64]; skewed (pitch_y = 67, pitch_x = 4297, sdfr_core.cuh); z-pair float2.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather_peaks gather_peaks.cu
// run:   ./gather_peaks > gpurun_out/gather_peaks.json
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

constexpr int R = 64;
constexpr int kGrids = 64;

template <int MODE>  // 0: ld.global.nc (L1), 1: ld.global.cg (L2 only)
__device__ __forceinline__ float ld(const float* p) {
  float v;
  if (MODE == 0) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
template <int MODE>
__device__ __forceinline__ float2 ld2(const float2* p) {
  float2 v;
  if (MODE == 0) asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  else asm volatile("ld.global.cg.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}

__device__ __forceinline__ unsigned hash(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// PAIR = 0: 8 x LDG.32 with pitches (py, px) floats.  PAIR = 1: 4 x LDG.64 on a float2-per-voxel array
// (pitches in float2 units).  RANDOM: every lane its own random cell, else the coherent footprint walk.
template <int MODE, int PAIR, int RANDOM>
__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ grids, long long grid_stride,
                                                     int py, int px, int iters, float* out) {
  const int lane = threadIdx.x & 31, warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const float* g = grids + (size_t)(blockIdx.x % kGrids) * grid_stride;
  // footprint: 8x4 pixels over ~2.5 x 1.5 cells (4 pixels per cell, sub-cell phase from the warp id)
  const int fx = ((lane & 7) + (warp_global & 3)) >> 2, fy = ((lane >> 3) + ((warp_global >> 2) & 3)) >> 2;
  unsigned s = hash(warp_global * 2654435761u + 12345u);
  int cx = s % (R - 4), cy = (s >> 8) % (R - 4), cz = (s >> 16) % (R - 4);
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    int ix, iy, iz;
    const float* gg = g;
    if (RANDOM) {
      const unsigned h = hash((blockIdx.x * blockDim.x + threadIdx.x) * 747796405u + i * 2891336453u);
      ix = h % (R - 1); iy = (h >> 8) % (R - 1); iz = (h >> 16) % (R - 1);
      gg = grids + (size_t)((h >> 24) % kGrids) * grid_stride;
    } else {
      ix = cx + fx; iy = cy + fy; iz = cz + ((lane * 7 + i) & 1);  // depth differs by <= 1 cell in a warp
      // one cell per step along a diagonal, wrapping inside the grid
      cx = cx + 1 < R - 4 ? cx + 1 : 0; cy = (i & 1) ? (cy + 1 < R - 4 ? cy + 1 : 0) : cy;
      cz = (i & 3) == 3 ? (cz + 1 < R - 4 ? cz + 1 : 0) : cz;
    }
    if (PAIR) {
      const float2* c = reinterpret_cast<const float2*>(gg) + (ix * px + iy * py + iz);
      const float2 a = ld2<MODE>(c), b = ld2<MODE>(c + py), d = ld2<MODE>(c + px), e = ld2<MODE>(c + px + py);
      acc += (a.x + a.y) + (b.x + b.y) + (d.x + d.y) + (e.x + e.y);
    } else {
      const float* c = gg + (ix * px + iy * py + iz);
      acc += ld<MODE>(c) + ld<MODE>(c + 1) + ld<MODE>(c + py) + ld<MODE>(c + py + 1) + ld<MODE>(c + px) +
             ld<MODE>(c + px + 1) + ld<MODE>(c + px + py) + ld<MODE>(c + px + py + 1);
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

// ---- north-star option "fp16 copy staged in shared memory for <= 48^3" (n1): the same coherent walk on a
// 48^3 grid, (a) fp32 skewed through L1 as the product does it, (b) a __half copy of the whole grid
// (221 184 B, the only size class that fits 227 KB) staged per CTA into shared memory and gathered with
// 8 x LDS.U16.  One CTA of 1024 threads per SM (the grid owns the SM's shared memory).  The staging copy
// is inside the timed region once per CTA, as it would be once per (CTA, hypothesis).
#include <cuda_fp16.h>
constexpr int R48 = 48;
template <int SMEM>
__global__ void __launch_bounds__(1024) gather48_kernel(const float* __restrict__ grids, const __half* __restrict__ hgrids,
                                                         long long grid_stride, int py, int px, int iters, float* out) {
  extern __shared__ __half sh[];
  const int lane = threadIdx.x & 31, warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int gidx = blockIdx.x % kGrids;
  if (SMEM) {
    const uint4* src = reinterpret_cast<const uint4*>(hgrids + (size_t)gidx * R48 * R48 * R48);
    uint4* dst = reinterpret_cast<uint4*>(sh);
    for (int i = threadIdx.x; i < R48 * R48 * R48 / 8; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
  }
  const float* g = grids + (size_t)gidx * grid_stride;
  const int fx = ((lane & 7) + (warp_global & 3)) >> 2, fy = ((lane >> 3) + ((warp_global >> 2) & 3)) >> 2;
  unsigned s = hash(warp_global * 2654435761u + 12345u);
  int cx = s % (R48 - 4), cy = (s >> 8) % (R48 - 4), cz = (s >> 16) % (R48 - 4);
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    const int ix = cx + fx, iy = cy + fy, iz = cz + ((lane * 7 + i) & 1);
    cx = cx + 1 < R48 - 4 ? cx + 1 : 0; cy = (i & 1) ? (cy + 1 < R48 - 4 ? cy + 1 : 0) : cy;
    cz = (i & 3) == 3 ? (cz + 1 < R48 - 4 ? cz + 1 : 0) : cz;
    if (SMEM) {
      const __half* c = sh + ((ix * R48 + iy) * R48 + iz);
      acc += __half2float(c[0]) + __half2float(c[1]) + __half2float(c[R48]) + __half2float(c[R48 + 1]) +
             __half2float(c[R48 * R48]) + __half2float(c[R48 * R48 + 1]) + __half2float(c[R48 * R48 + R48]) +
             __half2float(c[R48 * R48 + R48 + 1]);
    } else {
      const float* c = g + (ix * px + iy * py + iz);
      acc += ld<0>(c) + ld<0>(c + 1) + ld<0>(c + py) + ld<0>(c + py + 1) + ld<0>(c + px) + ld<0>(c + px + 1) +
             ld<0>(c + px + py) + ld<0>(c + px + py + 1);
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

struct Result { const char* name; double gsamples, gbytes, ms; };

template <int SMEM>
Result run48(const char* name, const float* grids, const __half* hgrids, long long stride, int py, int px, int ctas,
             int iters) {
  float* out; cudaMalloc(&out, 4);
  const size_t smem = SMEM ? (size_t)R48 * R48 * R48 * sizeof(__half) : 0;
  if (SMEM) cudaFuncSetAttribute(gather48_kernel<SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 2; ++w) gather48_kernel<SMEM><<<ctas, 1024, smem>>>(grids, hgrids, stride, py, px, iters, out);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { fprintf(stderr, "%s: %s\n", name, cudaGetErrorString(err)); return {name, 0, 0, 0}; }
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a);
    gather48_kernel<SMEM><<<ctas, 1024, smem>>>(grids, hgrids, stride, py, px, iters, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  const double samples = (double)ctas * 1024 * iters;
  cudaFree(out);
  return {name, samples / (best * 1e-3) / 1e9, samples * 32 / (best * 1e-3) / 1e9, best};
}

template <int MODE, int PAIR, int RANDOM>
Result run(const char* name, const float* grids, long long stride, int py, int px, int ctas, int iters) {
  float* out; cudaMalloc(&out, 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 2; ++w) gather_kernel<MODE, PAIR, RANDOM><<<ctas, 256>>>(grids, stride, py, px, iters, out);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a);
    gather_kernel<MODE, PAIR, RANDOM><<<ctas, 256>>>(grids, stride, py, px, iters, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  const double samples = (double)ctas * 256 * iters;
  cudaFree(out);
  return {name, samples / (best * 1e-3) / 1e9, samples * 32 / (best * 1e-3) / 1e9, best};
}

int main() {
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int PY = 67, PX = 64 * 67 + ((9 - (64 * 67) % 32) + 32) % 32;       // skewed pitches (floats)
  const int PY2 = 64 + 3, PX2 = 64 * PY2 + ((9 - (64 * PY2) % 16) + 16) % 16;  // z-pair pitches (float2; 16 per 128 B line)
  const long long dense = (long long)R * R * R, skew = (long long)R * PX, pair = 2ll * R * PX2;
  float *d_dense, *d_skew, *d_pair;
  cudaMalloc(&d_dense, dense * kGrids * 4); cudaMalloc(&d_skew, skew * kGrids * 4); cudaMalloc(&d_pair, pair * kGrids * 4);
  cudaMemset(d_dense, 0, dense * kGrids * 4); cudaMemset(d_skew, 0, skew * kGrids * 4); cudaMemset(d_pair, 0, pair * kGrids * 4);
  const int ctas = sms * 8, iters = 4096;
  std::vector<Result> rs;
  rs.push_back(run<0, 0, 0>("coherent_l1_dense_ldg32", d_dense, dense, R, R * R, ctas, iters));
  rs.push_back(run<0, 0, 0>("coherent_l1_skewed_ldg32", d_skew, skew, PY, PX, ctas, iters));
  rs.push_back(run<0, 1, 0>("coherent_l1_zpair_ldg64", d_pair, pair, PY2, PX2, ctas, iters));
  rs.push_back(run<1, 0, 0>("coherent_l2only_skewed_ldg32", d_skew, skew, PY, PX, ctas, iters / 4));
  rs.push_back(run<1, 1, 0>("coherent_l2only_zpair_ldg64", d_pair, pair, PY2, PX2, ctas, iters / 4));
  rs.push_back(run<0, 0, 1>("random_l2_skewed_ldg32", d_skew, skew, PY, PX, ctas, iters / 8));
  rs.push_back(run<0, 1, 1>("random_l2_zpair_ldg64", d_pair, pair, PY2, PX2, ctas, iters / 8));
  {
    // n1: 48^3, fp32 skewed through L1 (2 CTAs x 1024 threads per SM) against the fp16 shared-memory copy
    // (1 CTA x 1024 per SM); short walks (256 samples per thread ~ one batch of rays per staged grid) and
    // long ones (staging amortised away)
    const int PY48 = 48 + (((3 - 48 % 32) + 32) % 32), PX48 = 48 * PY48 + (((9 - (48 * PY48) % 32) + 32) % 32);
    const long long skew48 = (long long)R48 * PX48;
    float* d48; __half* h48;
    cudaMalloc(&d48, skew48 * kGrids * 4); cudaMemset(d48, 0, skew48 * kGrids * 4);
    cudaMalloc(&h48, (size_t)R48 * R48 * R48 * kGrids * 2); cudaMemset(h48, 0, (size_t)R48 * R48 * R48 * kGrids * 2);
    rs.push_back(run48<0>("r48_l1_fp32_skewed_ldg32", d48, h48, skew48, PY48, PX48, sms * 2, 4096));
    rs.push_back(run48<1>("r48_smem_fp16_lds16_long_walk", d48, h48, skew48, PY48, PX48, sms, 4096));
    rs.push_back(run48<0>("r48_l1_fp32_skewed_ldg32_256_samples", d48, h48, skew48, PY48, PX48, sms * 2, 256));
    rs.push_back(run48<1>("r48_smem_fp16_lds16_256_samples_incl_staging", d48, h48, skew48, PY48, PX48, sms, 256));
    cudaFree(d48); cudaFree(h48);
  }
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\n \"sms\": %d, \"sm_clock_khz_attr\": %d, \"ctas\": %d, \"threads\": 256, \"bytes_per_sample\": 32,\n", sms, clk, ctas);
  printf(" \"what\": \"pure 8-corner gather kernels (no march, no ALU): Gsample/s and algorithmic GB/s (32 B per sample)\",\n");
  for (size_t i = 0; i < rs.size(); ++i)
    printf(" \"%s\": {\"gsamples_per_s\": %.3f, \"gb_per_s\": %.1f, \"ms\": %.4f}%s\n", rs[i].name, rs[i].gsamples,
           rs[i].gbytes, rs[i].ms, i + 1 < rs.size() ? "," : "");
  printf("}\n");
  return 0;
}
