// Does the MUFU (and the FMA pipe) slow down while the tensor core of the same SM is busy? Warps 4..7 (one per scheduler) run the
// softmax instruction mix of the attention kernel; warp 1 optionally keeps the tensor pipe busy with back-to-back tcgen05.mma
// (128 x 128 x 16, bf16, shared-memory operands with arbitrary contents, fp32 accumulators in TMEM) for the whole duration.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I yume_b200/csrc -o tools/microbench/mufu_under_mma tools/microbench/mufu_under_mma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "yb_ptx.cuh"
using namespace yb;

__device__ __forceinline__ float ex2v(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fmav(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float addv(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ uint32_t packv(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

__device__ __forceinline__ float maxv(float a, float b) { float d; asm volatile("max.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ uint32_t madv(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
// exp2 on the FMA / ALU pipes: round-to-nearest split x = n + f (magic-number add), degree-3 polynomial for 2^f, exponent by integer
// multiply-add (8 instructions per value, max relative error 2e-4)
__device__ __forceinline__ float exp2_fma(float x) {
  x = maxv(x, -125.0f);
  const float t = addv(x, 12582912.0f);
  const float f = addv(x, -addv(t, -12582912.0f));
  float p = fmav(0.05550411f, f, 0.24022651f);
  p = fmav(p, f, 0.69314720f);
  p = fmav(p, f, 1.0f);
  return __uint_as_float(madv(__float_as_uint(t), 1u << 23, __float_as_uint(p)));
}

constexpr int U = 16;
// MODE 0: MUFU only; 1: softmax mix (2 FFMA + 2 EX2 + 2 FADD + 1 F2FP per pair); 2: FFMA only
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(int iters, int mma_on, int mma_n, long long* cyc, long long* mma_cyc, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
  if (warp == 1) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 1) {
    if (lane == 0 && mma_on) {
      const uint32_t idesc = make_idesc_bf16(128, mma_n, 0, 0);
      const uint64_t ad = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
      const uint64_t bd = make_smem_desc_sw128(smem_u32(smem + 32768), 16, 1024);
      const long long t0 = clock64();
      long long n = 0;
      uint32_t phase = 0;
      while (!stop) {                       // batches of 64 MMAs, one commit + wait per batch (keeps the queue full, bounds the run)
        for (int i = 0; i < 64; ++i) umma_ss(tmem_base + (i & 1) * 256, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, 1u);
        umma_commit(&bar);
        mbar_wait(&bar, phase);
        phase ^= 1;
        n += 64;
      }
      const long long t1 = clock64();
      mma_cyc[2 * blockIdx.x] = t1 - t0;
      mma_cyc[2 * blockIdx.x + 1] = n;
    }
  } else if (warp >= 4) {
    float x[U];
    uint32_t w[U / 2];
#pragma unroll
    for (int i = 0; i < U; ++i) x[i] = 0.25f + 0.001f * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < U / 2; ++i) w[i] = threadIdx.x + i;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < U; ++i) x[i] = ex2v(x[i]);
      } else if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < U / 2; ++i) {
          const float a = ex2v(fmav(x[2 * i], 1.0001f, -0.5f)), b = ex2v(fmav(x[2 * i + 1], 1.0001f, -0.5f));
          x[2 * i] = addv(x[2 * i], a); x[2 * i + 1] = addv(x[2 * i + 1], b);
          w[i] ^= packv(a, b);
        }
      } else if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < U; ++i) x[i] = fmav(x[i], 1.0001f, 0.5f);
      } else {   // MODE 3 / 4: the mix with one of every 4 (3) / 2 (4) pairs on the FMA-pipe polynomial instead of the MUFU
#pragma unroll
        for (int i = 0; i < U / 2; ++i) {
          const bool poly = (MODE == 3) ? (i % 4 == 3) : (i % 2 == 1);
          const float xa = fmav(x[2 * i], 1.0001f, -0.5f), xb = fmav(x[2 * i + 1], 1.0001f, -0.5f);
          const float a = poly ? exp2_fma(xa) : ex2v(xa), b = poly ? exp2_fma(xb) : ex2v(xb);
          x[2 * i] = addv(x[2 * i], a); x[2 * i + 1] = addv(x[2 * i + 1], b);
          w[i] ^= packv(a, b);
        }
      }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < U; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < U / 2; ++i) s += __uint_as_float(w[i]);
    if (threadIdx.x == 128) cyc[blockIdx.x] = t1 - t0;
    if (s == 12345.678f) sink[0] = s;
    __syncwarp();
    if (threadIdx.x == 128) stop = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int MODE>
static void run(const char* name, int ops_per_iter) {
  long long *cyc, *mc; float* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&mc, 148 * 16); cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 4096;
  for (int cfg = 0; cfg < 3; ++cfg) {
    const int mma_on = cfg > 0, mma_n = cfg == 2 ? 256 : 128;
    cudaMemset(mc, 0, 148 * 16);
    for (int rep = 0; rep < 2; ++rep) k<MODE><<<148, 256, 100 * 1024>>>(iters, mma_on, mma_n, cyc, mc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[148], hm[296];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hm, mc, sizeof(hm), cudaMemcpyDeviceToHost);
    double mean = 0, mm = 0, mn = 0;
    for (int i = 0; i < 148; ++i) { mean += h[i]; mm += hm[2 * i]; mn += hm[2 * i + 1]; }
    mean /= 148;
    printf("%-44s tensor pipe %-22s %9.0f cycles = %6.2f per warp instruction per scheduler", name,
           mma_on ? (mma_n == 256 ? "busy (128x256x16 MMAs)" : "busy (128x128x16 MMAs)") : "idle", mean, mean / (double(iters) * ops_per_iter));
    if (mma_on && mn > 0) printf("   [%.1f cycles per MMA]", mm / mn);
    printf("\n");
  }
  cudaFree(cyc); cudaFree(mc); cudaFree(sink);
}

int main() {
  run<0>("MUFU.EX2, 1 warp per scheduler", U);
  run<1>("softmax mix (7 instr / pair), 1 warp/sched", 7 * U / 2);
  run<2>("FFMA, 1 warp per scheduler", U);
  run<3>("mix, 1 of 4 pairs on the FMA polynomial", 7 * U / 2);   // reported per 'mix instruction' so that x 7 = cycles per pair
  run<4>("mix, 1 of 2 pairs on the FMA polynomial", 7 * U / 2);
  return 0;
}
