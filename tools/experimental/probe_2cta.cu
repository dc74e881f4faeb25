// probe_2cta.cu — EXPERIMENTAL, not part of libyume_b200.so and not on any product path.
//
// Stand-alone self-test of the `cta_group::2` (SM pair) building blocks the next revision of the GEMM / conv / attention
// kernels needs: cluster launch, paired TMEM allocation, TMA loads that signal the LEADER CTA's
// mbarrier, one `tcgen05.mma.cta_group::2` with M = 256 (128 rows per SM) x N = 256 (each SM stages HALF of B), the
// multicast commit, and a per-SM epilogue. Written when no GPU time was left in round 1: it assembles with nvcc 12.9 for
// sm_100a but has NOT run on hardware yet. PTX forms follow the public CUTLASS sm100 headers (cute/arch/copy_sm100_tma.hpp,
// cutlass/arch/barrier.h, cute/arch/mma_sm100_umma.hpp, cute/arch/tmem_allocator_sm100.hpp).
//
// Build + run (GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 --expt-relaxed-constexpr -Iinclude -Iyume_b200/csrc \
//        tools/experimental/probe_2cta.cu -o /tmp/probe_2cta && timeout 60 /tmp/probe_2cta
// Prints the max abs error of D[256,256] = A[256,128] . B[256,128]^T against a CPU reference (expect ~1e-5 relative).
#include <cstdio>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "yb_host.h"
#include "yb_ptx.cuh"

namespace yb {

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;           // this CTA's 128 rows of A: 2 K-slabs x 16 KB
  uint8_t* sB = smem + 32768;   // this CTA's HALF of B (128 of the 256 N rows): 2 K-slabs x 16 KB
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + 65536);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_mma + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc_2cta(tmem_ptr, 256);   // same warp id and same dst offset in both CTAs
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // both CTAs' barriers exist before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (threadIdx.x == 0) {
    if (rank == 0) mbar_arrive_expect_tx(bar_load, 2 * 65536);   // the bytes of BOTH CTAs land on the leader's barrier
    const int row0 = static_cast<int>(rank) * 128;
    tma_load_2d_2cta(sA, &tmA, bar_load, 0, row0);
    tma_load_2d_2cta(sA + 16384, &tmA, bar_load, 64, row0);
    tma_load_2d_2cta(sB, &tmB, bar_load, 0, row0);
    tma_load_2d_2cta(sB + 16384, &tmB, bar_load, 64, row0);
    if (rank == 0) {
      mbar_wait(bar_load, 0);
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc_bf16(256, 256, 0, 0);
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
        umma_ss_2cta(tmem_base, make_smem_desc_sw128(a0 + off, 16, 1024), make_smem_desc_sw128(b0 + off, 16, 1024), idesc,
                     kk != 0);
      }
      umma_commit_2cta(bar_mma);
    }
  }
  __syncwarp();
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  const int row = static_cast<int>(rank) * 128 + warp * 32 + lane;   // this SM's TMEM holds its own 128 rows x 256 columns
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {
    uint32_t r[32];
    tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) D[row * 256 + c * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 256);
  }
}

}  // namespace yb

static float bf16_round(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  u &= 0xFFFF0000u;
  memcpy(&x, &u, 4);
  return x;
}

int main() {
  using namespace yb;
  const int M = 256, N = 256, K = 128;
  std::vector<float> A(M * K), B(N * K);
  std::vector<uint16_t> Ah(M * K), Bh(N * K);
  srand(7);
  auto fill = [&](std::vector<float>& f, std::vector<uint16_t>& h) {
    for (size_t i = 0; i < f.size(); ++i) {
      f[i] = bf16_round((rand() % 2001 - 1000) / 1000.0f);
      uint32_t u;
      memcpy(&u, &f[i], 4);
      h[i] = static_cast<uint16_t>(u >> 16);
    }
  };
  fill(A, Ah);
  fill(B, Bh);
  void *dA, *dB;
  float* dD;
  cudaMalloc(&dA, Ah.size() * 2);
  cudaMalloc(&dB, Bh.size() * 2);
  cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xFF, M * N * 4);
  CUtensorMap tmA, tmB;
  if (make_tmap_bf16_2d(&tmA, dA, M, K, K, 128, 64) || make_tmap_bf16_2d(&tmB, dB, N, K, K, 128, 64)) {
    printf("tensor map encode failed\n");
    return 2;
  }
  const int smem = 65536 + 1024 + 64;
  cudaFuncSetAttribute(probe_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_2cta_kernel<<<2, 128, smem>>>(tmA, tmB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("kernel failed: %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> D(M * N);
  cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
  double max_err = 0, max_ref = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += static_cast<double>(A[m * K + k]) * B[n * K + k];
      max_err = std::max(max_err, std::abs(ref - D[m * N + n]));
      max_ref = std::max(max_ref, std::abs(ref));
    }
  printf("probe_2cta: max |err| %.3e (max |ref| %.3e) -> %s\n", max_err, max_ref, max_err < 1e-3 * max_ref ? "OK" : "MISMATCH");
  return max_err < 1e-3 * max_ref ? 0 : 1;
}
