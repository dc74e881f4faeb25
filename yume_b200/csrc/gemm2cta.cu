// gemm2cta.cu — EXPERIMENTAL SM-pair GEMM (`tcgen05.mma.cta_group::2`). Opt-in entry point `yb_gemm_bf16_2cta`; nothing on the
// product path calls it. Written at the end of round 1 when no GPU time was left: it assembles for sm_100a (SASS:
// UTCHMMA.2CTA, UTMALDG.2D.2CTA, UTCBAR.2CTA.MULTICAST) but has NOT run on hardware. Its GPU test is skipped unless
// YB_RUN_EXPERIMENTAL=1. Purpose: 1-CTA MMAs are bound by operand fetch (12 KB of smem per 128x256x16
// instruction); a CTA pair sharing one 256 x 256 tile fetches 8 KB per SM for the same math.
//
//   out[M,N] (bf16) = A[M,K] (bf16, row-major) x B[N,K]^T (bf16, row-major) + bias[N]
//
// Cluster of 2 CTAs, persistent over 256 x 256 output tiles; CTA r of the pair owns rows [r*128, r*128+128) of the tile.
//   warp 0      TMA producer (both CTAs): own 128x64 A tile + own HALF (128 rows of N) of the B tile per k-step into a
//               6-stage ring; completion bytes of both CTAs are credited to the LEADER's full barrier (2-SM TMA form)
//   warp 1      MMA issuer (leader CTA only): M=256, N=256, K=16 x4 per stage; commits are multicast to both CTAs'
//               empty / tmem_full barriers
//   warps 2..5  epilogue (both CTAs): own 128 rows, TMEM -> smem transpose -> coalesced bf16 stores; they arrive on the
//               LEADER's tmem_empty barrier (count 256)
#include "yb_host.h"
#include "yb_ptx.cuh"

namespace yb {

constexpr int G2_THREADS = 192;
constexpr int G2_BLOCK_N = 256;
constexpr int G2_BLOCK_K = 64;
constexpr int G2_STAGES = 6;
constexpr int G2_A_BYTES = 128 * G2_BLOCK_K * 2;            // this CTA's 128 rows of A
constexpr int G2_B_BYTES = (G2_BLOCK_N / 2) * G2_BLOCK_K * 2;  // this CTA's half of the B tile
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_BAR_OFF = G2_STAGES * G2_STAGE_BYTES;
constexpr int G2_EPI_OFF = G2_BAR_OFF + 256;
constexpr int G2_SMEM_BYTES = G2_EPI_OFF + 4 * 32 * 36 * 4 + 1024;
constexpr int G2_GROUP_N = 8;

struct Gemm2Params {
  int M, N, K;
  const float* bias;
  __nv_bfloat16* out;
  long long ldo;
  int num_m_tiles, num_n_tiles;   // tiles of 256 x 256
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Gemm2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G2_BAR_OFF);   // used in the leader CTA only
  uint64_t* empty_bar = full_bar + G2_STAGES;
  uint64_t* tmem_full = empty_bar + G2_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;                                  // used in the leader CTA only
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int tile_first = blockIdx.x >> 1, tile_step = gridDim.x >> 1;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_kb = (p.K + G2_BLOCK_K - 1) / G2_BLOCK_K;
  auto tile_coords = [&](int tile, int& m_tile, int& n_tile) {
    const int per_group = G2_GROUP_N * p.num_m_tiles;
    const int g = tile / per_group, r = tile - g * per_group;
    const int n_first = g * G2_GROUP_N;
    const int n_in_group = min(G2_GROUP_N, p.num_n_tiles - n_first);
    m_tile = r / n_in_group;
    n_tile = n_first + (r - m_tile * n_in_group);
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < G2_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 256);   // 4 epilogue warps of each CTA of the pair
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_ptr, 512);   // same warp and same smem slot in both CTAs
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // every barrier of the pair exists before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
        int m_tile, n_tile;
        tile_coords(tile, m_tile, n_tile);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * G2_STAGE_BYTES;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);
          tma_load_2d_2cta(sa, &tmA, &full_bar[stage], kb * G2_BLOCK_K, (m_tile * 2 + rank) * 128);
          tma_load_2d_2cta(sa + G2_A_BYTES, &tmB, &full_bar[stage], kb * G2_BLOCK_K,
                           n_tile * G2_BLOCK_N + rank * (G2_BLOCK_N / 2));
          if (++stage == G2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, G2_BLOCK_N, 0, 0);
      int stage = 0, local = 0;
      uint32_t phase = 0;
      for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++local) {
        const int acc = local & 1;
        mbar_wait_cluster(&tmem_empty[acc], ((local >> 1) & 1) ^ 1);   // arrivals come from both CTAs
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * G2_BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * G2_STAGE_BYTES);
          const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sa + G2_A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < G2_BLOCK_K / 16; ++k) umma_ss_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_2cta(&empty_bar[stage]);        // frees the slot in BOTH CTAs
          if (++stage == G2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2cta(&tmem_full[acc]);            // accumulator ready in BOTH CTAs
      }
    }
  } else {
    const int quad = warp & 3;
    float* tr = reinterpret_cast<float*>(smem + G2_EPI_OFF) + quad * (32 * 36);
    int local = 0;
    for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++local) {
      int m_tile, n_tile;
      tile_coords(tile, m_tile, n_tile);
      const int acc = local & 1;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * G2_BLOCK_N;
      const int row = (m_tile * 2 + rank) * 128 + quad * 32 + lane;
      const int my_row = row < p.M ? row : -1;
      mbar_wait(&tmem_full[acc], (local >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < G2_BLOCK_N / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(t_row + c * 32, r);
        tmem_ld_wait();
        const int col0 = n_tile * G2_BLOCK_N + c * 32;
        float4* st = reinterpret_cast<float4*>(tr + lane * 36);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          st[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                              __uint_as_float(r[4 * i + 3]));
        __syncwarp();
        if (col0 < p.N) {
          const int cq = (lane & 3) * 8;
          float b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (p.bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + cq));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + cq + 4));
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
          }
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int rr = it * 8 + (lane >> 2);
            const int rowi = __shfl_sync(0xffffffffu, my_row, rr);
            const float4 a0 = *reinterpret_cast<const float4*>(tr + rr * 36 + cq);
            const float4 a1 = *reinterpret_cast<const float4*>(tr + rr * 36 + cq + 4);
            if (rowi >= 0) {
              uint4 w;
              w.x = pack_bf16x2(a0.x + b[0], a0.y + b[1]);
              w.y = pack_bf16x2(a0.z + b[2], a0.w + b[3]);
              w.z = pack_bf16x2(a1.x + b[4], a1.y + b[5]);
              w.w = pack_bf16x2(a1.z + b[6], a1.w + b[7]);
              *reinterpret_cast<uint4*>(p.out + static_cast<long long>(rowi) * p.ldo + col0 + cq) = w;
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive_leader(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace yb

extern "C" int yb_gemm_bf16_2cta(const void* A, long long lda, const void* B, long long ldb, const void* bias, void* out,
                                 long long ldo, int M, int N, int K, void* stream_) {
  using namespace yb;
  if (!A || !B || !out || M <= 0 || N <= 0 || K <= 0) return YB_ERR_ARG;
  if (N % 32 != 0 || K % 8 != 0) return YB_ERR_SHAPE;
  if ((lda % 8) || (ldb % 8) || (ldo % 8) || (reinterpret_cast<uintptr_t>(out) & 0xF)) return YB_ERR_ALIGNMENT;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_2d(&tmA, A, M, K, lda, 128, G2_BLOCK_K);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, B, N, K, ldb, G2_BLOCK_N / 2, G2_BLOCK_K);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES) != cudaSuccess) {
      (void)cudaGetLastError();
      return YB_ERR_LAUNCH;
    }
    attr_set = true;
  }
  Gemm2Params p;
  p.M = M; p.N = N; p.K = K;
  p.bias = static_cast<const float*>(bias);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.num_m_tiles = (M + 255) / 256;
  p.num_n_tiles = (N + G2_BLOCK_N - 1) / G2_BLOCK_N;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int clusters = tiles < sm_count() / 2 ? tiles : sm_count() / 2;
  gemm2cta_kernel<<<2 * clusters, G2_THREADS, G2_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream_)>>>(tmA, tmB, p);
  return check_launch("gemm2cta");
}
