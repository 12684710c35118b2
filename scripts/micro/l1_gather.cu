// L1 gather micro-benchmark: how many data-stage cycles does one warp-level LDG.32 cost as a
// function of (distinct 128B lines, bank overlap)?  Decides the SDF grid layout.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l1_gather l1_gather.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__global__ void gather_kernel(const float* __restrict__ buf, const int* __restrict__ offs,
                              int iters, int step, float* out, long long* cycles) {
  const int lane = threadIdx.x & 31;
  const int off = offs[lane];  // in floats
  const float* p = buf + off;
  float acc = 0.f;
  // warm L1
  for (int k = 0; k < 8; ++k) acc += __ldg(p + k * 2048);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {  // 8 independent gathers, 8 KB apart
      float v;
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p + i * step + k * 2048));
      acc += v;
    }
  }
  const long long t1 = clock64();
  if (acc == 123.456f) out[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

int main() {
  const int n = 1 << 16;  // 256 KB buffer; each warp touches <= 8 x 32 lines
  float* buf; cudaMalloc(&buf, n * sizeof(float)); cudaMemset(buf, 0, n * sizeof(float));
  int* d_off; cudaMalloc(&d_off, 32 * sizeof(int));
  float* out; cudaMalloc(&out, 4);
  long long* cyc; cudaMalloc(&cyc, 8);
  struct Pat { const char* name; int (*f)(int); };
  Pat pats[] = {
      {"coalesced: 1 line, 32 banks", [](int l) { return l; }},
      {"broadcast: 1 word", [](int l) { return 0; }},
      {"32 lines, same bank", [](int l) { return l * 32; }},
      {"32 lines, distinct banks", [](int l) { return l * 33; }},
      {"6 lines, same bank (5-6 lanes share a word)", [](int l) { return (l % 6) * 32; }},
      {"6 lines, distinct banks", [](int l) { return (l % 6) * 33; }},
      {"8 lines x 4 words, same 4 banks", [](int l) { return (l / 4) * 32 + (l % 4); }},
      {"8 lines x 4 words, skewed banks", [](int l) { return (l / 4) * 36 + (l % 4); }},
      {"2 lines, same bank", [](int l) { return (l % 2) * 32; }},
      {"2 lines, distinct banks", [](int l) { return (l % 2) * 33; }},
      {"4 lines, same bank", [](int l) { return (l % 4) * 32; }},
      {"4 lines, distinct banks", [](int l) { return (l % 4) * 33; }},
      {"4 sectors of one line x 8 words", [](int l) { return l; }},
      {"12 lines, same bank", [](int l) { return (l % 12) * 32; }},
      {"12 lines, distinct banks", [](int l) { return (l % 12) * 33; }},
  };
  const int iters = 2000;
  for (auto& p : pats) {
    std::vector<int> h(32);
    for (int l = 0; l < 32; ++l) h[l] = p.f(l);
    cudaMemcpy(d_off, h.data(), 32 * sizeof(int), cudaMemcpyHostToDevice);
    for (int warps : {1, 8, 32}) {
      // one CTA per SM, `warps` warps each: per-SM throughput
      gather_kernel<<<148, warps * 32>>>(buf, d_off, iters, 0, out, cyc);
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double per_ldg = (double)c / (iters * 8.0 * warps);
      printf("%-48s warps/SM=%2d  cycles per warp-LDG (SM-wide) = %6.2f\n", p.name, warps, per_ldg);
    }
  }
  return 0;
}
