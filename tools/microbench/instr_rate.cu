// Issue-rate microbenchmark for the softmax inner loop of the attention kernel (sm_100a): how many cycles does one warp
// instruction of each kind cost per SM sub-partition, alone and in the mixes the kernel issues? Answers (profiles/r02_instr_rate.md):
// the MUFU.EX2 rate, whether the fp32->bf16x2 pack (F2FP) shares a pipe with it, what a packed-FMA polynomial costs beside it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/instr_rate tools/microbench/instr_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) {
  uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
}
__device__ __forceinline__ float fma_(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float add_(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b) { uint32_t d; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float max3(float a, float b, float c) { float d; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

constexpr int U = 16;   // independent chains per thread

template <int MODE>
__global__ void __launch_bounds__(512, 1) rate_kernel(int iters, float seed, long long* cycles, float* sink) {
  float x[U];
  uint32_t w[U / 2];
#pragma unroll
  for (int i = 0; i < U; ++i) x[i] = seed + 0.001f * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < U / 2; ++i) w[i] = threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {            // MUFU.EX2 only
#pragma unroll
      for (int i = 0; i < U; ++i) x[i] = ex2(x[i]);
    } else if (MODE == 1) {     // F2FP pack only (one per pair)
#pragma unroll
      for (int i = 0; i < U / 2; ++i) { w[i] = pack(x[2 * i], x[2 * i + 1]); x[2 * i] = __uint_as_float(w[i]); }
    } else if (MODE == 2) {     // 2 ex2 + 1 pack per pair
#pragma unroll
      for (int i = 0; i < U / 2; ++i) { x[2 * i] = ex2(x[2 * i]); x[2 * i + 1] = ex2(x[2 * i + 1]); w[i] ^= pack(x[2 * i], x[2 * i + 1]); }
    } else if (MODE == 3) {     // FFMA only
#pragma unroll
      for (int i = 0; i < U; ++i) x[i] = fma_(x[i], 1.0001f, 0.5f);
    } else if (MODE == 4) {     // the kernel's mix per pair: 2 FFMA (scale, -max), 2 ex2, 2 FADD (row sum), 1 pack
#pragma unroll
      for (int i = 0; i < U / 2; ++i) {
        const float a = ex2(fma_(x[2 * i], 1.0001f, -0.5f)), b = ex2(fma_(x[2 * i + 1], 1.0001f, -0.5f));
        x[2 * i] = add_(x[2 * i], a); x[2 * i + 1] = add_(x[2 * i + 1], b);
        w[i] ^= pack(a, b);
      }
    } else if (MODE == 5) {     // same mix, bf16 rounding on the FMA pipe (Veltkamp split: 1 FMUL + 2 FADD per value) + PRMT
#pragma unroll
      for (int i = 0; i < U / 2; ++i) {
        const float a = ex2(fma_(x[2 * i], 1.0001f, -0.5f)), b = ex2(fma_(x[2 * i + 1], 1.0001f, -0.5f));
        x[2 * i] = add_(x[2 * i], a); x[2 * i + 1] = add_(x[2 * i + 1], b);
        const float ca = a * 65537.0f, cb = b * 65537.0f;
        const float ha = add_(ca, -add_(ca, -a)), hb = add_(cb, -add_(cb, -b));
        w[i] ^= prmt(__float_as_uint(ha), __float_as_uint(hb));
      }
    } else if (MODE == 6) {     // same mix, truncating pack (PRMT only)
#pragma unroll
      for (int i = 0; i < U / 2; ++i) {
        const float a = ex2(fma_(x[2 * i], 1.0001f, -0.5f)), b = ex2(fma_(x[2 * i + 1], 1.0001f, -0.5f));
        x[2 * i] = add_(x[2 * i], a); x[2 * i + 1] = add_(x[2 * i + 1], b);
        w[i] ^= prmt(__float_as_uint(a), __float_as_uint(b));
      }
    } else if (MODE == 7) {     // packed bf16x2 exponentials only
#pragma unroll
      for (int i = 0; i < U / 2; ++i) w[i] = ex2_bf16x2(w[i]);
    } else if (MODE == 8) {     // packed FMA (FFMA2) only
      unsigned long long* p = reinterpret_cast<unsigned long long*>(x);
#pragma unroll
      for (int i = 0; i < U / 2; ++i) p[i] = fma2(p[i], 0x3f8003473f800347ull, 0x3f0000003f000000ull);
    } else if (MODE == 9) {     // 3-input max only
#pragma unroll
      for (int i = 0; i < U / 2; ++i) x[2 * i] = max3(x[2 * i], x[2 * i + 1], seed);
    } else if (MODE == 10) {    // mix of MODE 4 with the row sum on packed adds (1 FADD2 per pair)
      unsigned long long* p = reinterpret_cast<unsigned long long*>(x);
#pragma unroll
      for (int i = 0; i < U / 2; ++i) {
        const float a = ex2(fma_(x[2 * i], 1.0001f, -0.5f)), b = ex2(fma_(x[2 * i + 1], 1.0001f, -0.5f));
        unsigned long long ab = (static_cast<unsigned long long>(__float_as_uint(b)) << 32) | __float_as_uint(a);
        p[i] = fma2(ab, 0x3f8000003f800000ull, p[i]);
        w[i] ^= pack(a, b);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < U; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < U / 2; ++i) s += __uint_as_float(w[i]);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (s == 12345.678f) sink[0] = s;
}


// ---- does a warp parked on an mbarrier slow down a computing warp on the same scheduler? (attention kernel: while query tile A
// runs its exponentials, tile B's softmax warp waits for its S on the same SM sub-partition)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
template <int WAITER>   // 0 = partner warps exit, 1 = try_wait loop, 2 = try_wait with a 1 ms suspend hint, 3 = test_wait + nanosleep(256)
__global__ void __launch_bounds__(256, 1) parked_kernel(int iters, float seed, long long* cycles, float* sink) {
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(128));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x >= 128) {
    if (WAITER == 0) return;
    uint32_t ok = 0;
    while (!ok) {
      if (WAITER == 1)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      else if (WAITER == 2)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0), "r"(1000000) : "memory");
      else {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        if (!ok) __nanosleep(256);
      }
    }
    return;
  }
  float x[U];
  uint32_t w[U / 2];
#pragma unroll
  for (int i = 0; i < U; ++i) x[i] = seed + 0.001f * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < U / 2; ++i) w[i] = threadIdx.x + i;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < U / 2; ++i) {
      const float a = ex2(fma_(x[2 * i], 1.0001f, -0.5f)), b = ex2(fma_(x[2 * i + 1], 1.0001f, -0.5f));
      x[2 * i] = add_(x[2 * i], a); x[2 * i + 1] = add_(x[2 * i + 1], b);
      w[i] ^= pack(a, b);
    }
  }
  const long long t1 = clock64();
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
  float sacc = 0.f;
#pragma unroll
  for (int i = 0; i < U; ++i) sacc += x[i];
#pragma unroll
  for (int i = 0; i < U / 2; ++i) sacc += __uint_as_float(w[i]);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (sacc == 12345.678f) sink[0] = sacc;
}

template <int WAITER>
static void run_parked(const char* name) {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * sizeof(long long)); cudaMalloc(&sink, 4);
  const int iters = 4096;
  parked_kernel<WAITER><<<148, 256>>>(iters, 0.25f, cyc, sink);
  parked_kernel<WAITER><<<148, 256>>>(iters, 0.25f, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < 148; ++i) mean += h[i]; mean /= 148;
  printf("mix on 1 warp/SMSP, partner warp: %-44s %8.0f cycles, %.2f cycles per pair of exponentials\n", name, mean,
         mean / (static_cast<double>(iters) * U / 2));
  cudaFree(cyc); cudaFree(sink);
}


// ---- the attention kernel's softmax body on registers only (no TMEM, no barriers): 128 scores per thread, row max, then the two
// 64-column halves of exponentials + row sums + bf16 packs exactly as attention.cu writes them. One warp per scheduler.
// VARIANT 0: packed FFMA2/FADD2 (the product kernel); 1: scalar FFMA/FADD; 2: scalar, software-pipelined by hand — the MUFU
// stream of the whole tile is issued in program order (volatile asm) and the FADD / F2FP of pair i - DIST ride in its shadow
__device__ __forceinline__ unsigned long long f2pack(float lo, float hi) {
  return (static_cast<unsigned long long>(__float_as_uint(hi)) << 32) | __float_as_uint(lo);
}
__device__ __forceinline__ unsigned long long f2add(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ unsigned long long f2fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
template <int VARIANT>
__global__ void __launch_bounds__(128, 3) softmax_body_kernel(int iters, const float* __restrict__ in, long long* cycles, float* sink) {
  __shared__ uint4 stage[128][9];
  float s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = in[(threadIdx.x * 131 + i * 7) & 4095];
  const float sc = 0.1275f;
  float m_used = -1e30f, l = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float mxa[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) mxa[a] = -INFINITY;
#pragma unroll
    for (int i = 0; i < 128; ++i) mxa[i & 7] = fmaxf(mxa[i & 7], s[i]);
    const float mx = fmaxf(fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])), fmaxf(fmaxf(mxa[4], mxa[5]), fmaxf(mxa[6], mxa[7])));
    m_used = fmaxf(m_used, mx * sc) + 1e-3f * it;      // varies per iteration: nothing below is loop-invariant
    const float negm = -m_used;
    auto store_half = [&](const uint32_t (&pk)[32], int h) {   // stands in for one TMEM store of P: 4 x 16-byte smem stores
#pragma unroll
      for (int c = 0; c < 4; ++c)
        stage[threadIdx.x][h * 4 + c] = make_uint4(pk[8 * c] ^ pk[8 * c + 4], pk[8 * c + 1] ^ pk[8 * c + 5], pk[8 * c + 2] ^ pk[8 * c + 6], pk[8 * c + 3] ^ pk[8 * c + 7]);
    };
    if (VARIANT == 0) {
      const unsigned long long sc2 = f2pack(sc, sc), negm2 = f2pack(negm, negm);
      unsigned long long ls2[2] = {0ull, 0ull};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int c0 = h * 64 + 2 * i;
          const unsigned long long x2 = fma2(f2pack(s[c0], s[c0 + 1]), sc2, negm2);
          const float p0 = ex2(__uint_as_float(static_cast<uint32_t>(x2))), p1 = ex2(__uint_as_float(static_cast<uint32_t>(x2 >> 32)));
          ls2[i & 1] = f2add(ls2[i & 1], f2pack(p0, p1));
          pk[i] = pack(p0, p1);
        }
        store_half(pk, h);
      }
      l += (__uint_as_float(static_cast<uint32_t>(ls2[0])) + __uint_as_float(static_cast<uint32_t>(ls2[0] >> 32))) +
           (__uint_as_float(static_cast<uint32_t>(ls2[1])) + __uint_as_float(static_cast<uint32_t>(ls2[1] >> 32)));
    } else if (VARIANT == 1) {
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int c0 = h * 64 + 2 * i;
          const float p0 = ex2(fma_(s[c0], sc, negm)), p1 = ex2(fma_(s[c0 + 1], sc, negm));
          ls[(2 * i) & 3] += p0; ls[(2 * i + 1) & 3] += p1;
          pk[i] = pack(p0, p1);
        }
        store_half(pk, h);
      }
      l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
    } else {
      constexpr int DIST = 6;
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      float pv[128];
      uint32_t pk0[32], pk1[32];
#pragma unroll
      for (int i = 0; i < 64 + DIST; ++i) {
        if (i < 64) { pv[2 * i] = ex2(fma_(s[2 * i], sc, negm)); pv[2 * i + 1] = ex2(fma_(s[2 * i + 1], sc, negm)); }
        if (i >= DIST) {
          const int k = i - DIST;
          ls[(2 * k) & 3] = add_(ls[(2 * k) & 3], pv[2 * k]);
          ls[(2 * k + 1) & 3] = add_(ls[(2 * k + 1) & 3], pv[2 * k + 1]);
          if (k < 32) pk0[k] = pack(pv[2 * k], pv[2 * k + 1]); else pk1[k - 32] = pack(pv[2 * k], pv[2 * k + 1]);
          if (k == 31) store_half(pk0, 0);
          if (k == 63) store_half(pk1, 1);
        }
      }
      l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (l == 12345.678f) sink[0] = l + stage[threadIdx.x][3].x;
}

template <int VARIANT>
static void run_body(const char* name) {
  long long* cyc; float* sink; float* in;
  cudaMalloc(&cyc, 148 * sizeof(long long)); cudaMalloc(&sink, 4); cudaMalloc(&in, 4096 * 4);
  float h_in[4096];
  for (int i = 0; i < 4096; ++i) h_in[i] = static_cast<float>((i * 2654435761u) % 2001) / 100.0f - 10.0f;
  cudaMemcpy(in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
  const int iters = 2048;
  softmax_body_kernel<VARIANT><<<148, 128>>>(iters, in, cyc, sink);
  softmax_body_kernel<VARIANT><<<148, 128>>>(iters, in, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < 148; ++i) mean += h[i]; mean /= 148;
  printf("softmax body (128 scores/thread, 1 warp/SMSP), %-40s %7.0f cycles per tile (ideal MUFU time 1024)\n", name, mean / iters);
  cudaFree(cyc); cudaFree(sink); cudaFree(in);
}

template <int MODE>
static void run(const char* name, int ops_per_iter_per_thread) {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * sizeof(long long)); cudaMalloc(&sink, 4);
  const int iters = 4096;
  for (int threads : {128, 256, 512}) {
    rate_kernel<MODE><<<148, threads>>>(iters, 0.25f, cyc, sink);   // warm-up
    rate_kernel<MODE><<<148, threads>>>(iters, 0.25f, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < 148; ++i) mean += h[i]; mean /= 148;
    const double warps_per_smsp = threads / 128.0;
    const double winstr = static_cast<double>(iters) * ops_per_iter_per_thread * warps_per_smsp;
    printf("%-46s warps/SMSP %.0f: %8.0f cycles, %.2f cycles per warp instruction per SMSP (%.1f thread-ops/clk/SM)\n", name,
           warps_per_smsp, mean, mean / winstr, 4.0 * 32.0 * winstr / mean);
  }
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  run<0>("MUFU.EX2", U);
  run<1>("F2FP pack (cvt.rn.bf16x2.f32)", U / 2);
  run<2>("2 EX2 + 1 F2FP", U + U / 2);
  run<3>("FFMA", U);
  run<8>("FFMA2 (fma.rn.f32x2)", U / 2);
  run<9>("FMNMX3", U / 2);
  run<7>("MUFU.EX2 bf16x2", U / 2);
  run<4>("mix: 2 FFMA + 2 EX2 + 2 FADD + 1 F2FP", 7 * U / 2);
  run<10>("mix: 2 FFMA + 2 EX2 + 1 FADD2 + 1 F2FP", 6 * U / 2);
  run<5>("mix, bf16 rounding on FMA pipe + PRMT", 13 * U / 2);
  run<6>("mix, truncating PRMT pack", 7 * U / 2);
  run_body<0>("packed FFMA2 / FADD2 (product kernel)");
  run_body<1>("scalar FFMA / FADD");
  run_body<2>("scalar, hand-pipelined (distance 6 pairs)");
  run_parked<0>("none (exits)");
  run_parked<1>("mbarrier.try_wait loop");
  run_parked<2>("mbarrier.try_wait, 1 ms suspend hint");
  run_parked<3>("mbarrier.test_wait + nanosleep(256)");
  return 0;
}
