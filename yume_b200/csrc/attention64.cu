// attention64.cu — EXPERIMENTAL attention variant: Q resident in TMEM, 64-key tiles. Selected with the YB_ATT_Q64 flag of
// yb_attention only; the product path never sets it. Written at the end of round 1 when no GPU time was left: it assembles
// for sm_100a but has NOT run on hardware; its GPU tests are skipped unless YB_RUN_EXPERIMENTAL=1.
//
// Why: a 1-CTA `tcgen05.mma` is bound by operand fetch from shared memory (~64 B/clk/SM). In the
// product kernel S = Q.K^T takes BOTH operands from smem (8 KB per 65-cycle 128x128x16 instruction -> ~128 cycles), while
// O += P.V (P in TMEM) runs at the math rate: 8*128 + 8*65 = 1544 cycles per 128 keys and query tile against 1040 of math.
// Here Q (constant over the whole key loop) is the TMEM-resident A operand of S = Q.K^T as well, so every MMA fetches one smem
// operand only. TMEM has no room for Q next to S (2x128) and O (2x128), so the key tile shrinks to 64:
//   columns [0,64) S_0 / P_0   [64,128) S_1 / P_1   [128,192) Q_0   [192,256) Q_1   [256,384) O_0   [384,512) O_1
// Per 64 keys and query tile: 8 MMAs 128x64x16 (32 cycles of math, 2 KB of K) + 4 MMAs 128x128x16 (65 cycles, 4 KB of V)
// = 516 cycles, i.e. 1032 per 128 keys — the math rate, co-saturated with the MUFU (16 ex2/clk/SM -> 1024 cycles).
// Shared memory holds only the K/V ring (Q is read once from global memory by the thread that owns the row).
//
// Same contract as attention.cu (non-causal, keys >= Lk dropped, bf16 in, fp32 accumulate, [L, heads, 128] output,
// YB_ATT_ACCUMULATE, Ulysses peer scatter, automatic tail split with the same planner); no trace in this variant.
#include "yb_host.h"
#include "yb_ptx.cuh"

namespace yb {

constexpr int A64_THREADS = 384;
constexpr int A64_TILE_BYTES = 64 * 128 * 2;   // one 64-key x 128-dim bf16 tile = 2 swizzled slabs [64 keys x 64 dims] of 8 KB
constexpr int A64_NS = 12;                     // K/V ring slots (6 key tiles of lookahead)
constexpr int A64_BAR_OFF = A64_NS * A64_TILE_BYTES;
constexpr int A64_SMEM_BYTES = A64_BAR_OFF + 1024 + 256;

struct Att64Params {
  const __nv_bfloat16* q;
  long long ldq;
  __nv_bfloat16* out;
  long long ldo;
  int Lq, Lk, nkv;          // nkv = ceil(Lk / 64)
  int accumulate;
  __nv_bfloat16* out_peers[8];
  int sp_world, sp_rank, sp_Lp;
  float scale_log2;
  // tail split (same decomposition as attention.cu): unit u = head * nq + q_block; CTAs [0, full_units) run whole units,
  // CTA full_units + r runs KV segment r % ns of unit full_units + r / ns and leaves a partial result in the workspace
  int nq, full_units, ns;
  float* ws_o;    // [tail CTAs, 256, 128] unnormalised O
  float* ws_ml;   // [tail CTAs, 256, 2]   (row max in the log2 domain, row sum)
};

// HALVES: P is handed to the MMA in two 32-key halves (tcgen05.st x16 + two barriers per tile) so that P.V of the first half
// overlaps the second half of the exponentials. Off in the first validation pass (YB_ATT64_HALVES=1 turns it on).
template <int EMU, bool HALVES>
__global__ void __launch_bounds__(A64_THREADS, 1)
attention64_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const Att64Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* q_ready = reinterpret_cast<uint64_t*>(smem + A64_BAR_OFF);   // [X]: Q_X is in TMEM (128 arrivals)
  uint64_t* kv_full = q_ready + 2;
  uint64_t* kv_empty = kv_full + A64_NS;
  uint64_t* s_full = kv_empty + A64_NS;
  uint64_t* p_ready = s_full + 2;                                        // [X][half]: P_X (or its 32-key half) is in TMEM
  uint64_t* o_done = p_ready + 4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // every role decodes the work unit INSIDE its own branch from an opaque copy of blockIdx.x (nothing stays live across
  // the register-starved softmax loop); kv_begin is even so barrier parities can follow the global tile index
  auto decode = [&](int& unit, int& kv_begin, int& nkv) {
    int bx = blockIdx.x;
    asm volatile("" : "+r"(bx));
    unit = bx;
    kv_begin = 0;
    nkv = p.nkv;
    if (bx >= p.full_units) {
      const int r = bx - p.full_units;
      unit = p.full_units + r / p.ns;
      const int per = (((p.nkv + p.ns - 1) / p.ns) + 1) & ~1;
      kv_begin = (r % p.ns) * per;
      nkv = min(per, p.nkv - kv_begin);
    }
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < A64_NS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_ready[i], 128);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[2 * i], 128);
      mbar_init(&p_ready[2 * i + 1], 128);
      mbar_init(&o_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < 4) {
    setmaxnreg_dec<80>();
    if (warp == 0 && lane == 0) {
      // ------------------------------- TMA producer: K_0, V_0, K_1, V_1, ... -------------------------------
      int unit, kv_begin, nkv;
      decode(unit, kv_begin, nkv);
      const int head = unit / p.nq;
      for (int it = 0; it < 2 * nkv; ++it) {
        const int slot = it % A64_NS;
        const uint32_t ph = (it / A64_NS) & 1;
        mbar_wait(&kv_empty[slot], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[slot], A64_TILE_BYTES);
        const CUtensorMap* tm = (it & 1) ? &tmV : &tmK;
        uint8_t* dst = smem + slot * A64_TILE_BYTES;
        const int j = kv_begin + (it >> 1);
        tma_load_2d(dst, tm, &kv_full[slot], head * 128, j * 64);
        tma_load_2d(dst + 8192, tm, &kv_full[slot], head * 128 + 64, j * 64);
      }
    } else if (warp == 1 && lane == 0) {
      // ------------------------------- MMA issuer -------------------------------
      int unit, kv_begin, nkv;
      decode(unit, kv_begin, nkv);
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);     // S[128 q, 64 keys] = Q (TMEM) x K^T (K-major smem)
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, 0, 1);   // O[128 q, 128 d] += P (TMEM) x V (MN-major smem)
      const uint32_t sKV = smem_u32(smem);
      const uint64_t kdesc0 = make_smem_desc_sw128(0, 16, 1024);       // K tile: 64 rows (keys) x 128-byte rows per d-half slab
      const uint64_t vdesc0 = make_smem_desc_sw128(0, 8192, 1024);     // V tile: MN-major, the two 64-dim slabs are 8 KB apart
      auto issue_S = [&](int X, uint32_t kbase) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {   // 16 head-dim elements per MMA: 8 TMEM columns of packed Q, 32 bytes of each K row
          const uint32_t off = (kk >> 2) * 8192 + (kk & 3) * 32;
          umma_ts(tmem_base + X * 64, tmem_base + 128 + X * 64 + kk * 8, kdesc0 + ((kbase + off) >> 4), idesc_s,
                  kk != 0 ? 1u : 0u);
        }
      };
      auto issue_PV = [&](int X, uint32_t vbase, bool acc, int k0, int k1) {
#pragma unroll
        for (int kk = k0; kk < k1; ++kk)   // 16 keys per MMA: 8 TMEM columns of packed P, 16 rows (2048 B) of the V tile
          umma_ts(tmem_base + 256 + X * 128, tmem_base + X * 64 + kk * 8, vdesc0 + ((vbase + kk * 2048) >> 4), idesc_pv,
                  (acc || kk != 0) ? 1u : 0u);
      };
      mbar_wait(&q_ready[0], 0);
      mbar_wait(&q_ready[1], 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_S(0, sKV);
      umma_commit(&s_full[0]);
      issue_S(1, sKV);
      umma_commit(&s_full[1]);
      umma_commit(&kv_empty[0]);
      for (int j = 0; j < nkv; ++j) {
        const int iv = 2 * j + 1, ik = 2 * j + 2;
        const int slot_v = iv % A64_NS, slot_k = ik % A64_NS;
        const bool has_next = (j + 1 < nkv);
        mbar_wait(&kv_full[slot_v], (iv / A64_NS) & 1);
        if (has_next) mbar_wait(&kv_full[slot_k], (ik / A64_NS) & 1);
        tc_fence_after();
        const uint32_t vbase = sKV + slot_v * A64_TILE_BYTES;
        const uint32_t kbase = sKV + slot_k * A64_TILE_BYTES;
#pragma unroll
        for (int X = 0; X < 2; ++X) {
          mbar_wait(&p_ready[2 * X], j & 1);
          tc_fence_after();
          if (HALVES) {
            issue_PV(X, vbase, j > 0, 0, 2);          // keys 0..31 while the softmax warps finish 32..63
            mbar_wait(&p_ready[2 * X + 1], j & 1);
            tc_fence_after();
            issue_PV(X, vbase, true, 2, 4);
          } else {
            issue_PV(X, vbase, j > 0, 0, 4);
          }
          if (has_next) {
            issue_S(X, kbase);          // overwrites S_X / P_X: ordered behind PV_X(j) on the in-order tensor pipe
            umma_commit(&s_full[X]);
          } else {
            umma_commit(&o_done[X]);
          }
        }
        umma_commit(&kv_empty[slot_v]);
        if (has_next) umma_commit(&kv_empty[slot_k]);
      }
    }
  } else {
    // ------------------------------- softmax / correction / epilogue -------------------------------
    setmaxnreg_inc<208>();
    const int X = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + X * 64;
    const uint32_t tQ = tmem_base + lane_off + 128 + X * 64;
    const uint32_t tO = tmem_base + lane_off + 256 + X * 128;
    const float sc = p.scale_log2;
    int head, q_row, kv_begin, kv_end;
    {
      int unit, nkv;
      decode(unit, kv_begin, nkv);
      kv_end = kv_begin + nkv;
      head = unit / p.nq;
      q_row = (unit - head * p.nq) * 256 + X * 128 + row_in_tile;
    }

    {  // Q row -> TMEM: column c of lane m holds (Q[m][2c], Q[m][2c+1]) packed lo/hi (the layout probe mode 2 verifies)
      const uint4* src = reinterpret_cast<const uint4*>(p.q + static_cast<long long>(q_row) * p.ldq + head * 128);
#pragma unroll 1
      for (int hlf = 0; hlf < 2; ++hlf) {
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 v4 = (q_row < p.Lq) ? __ldg(src + hlf * 8 + i) : make_uint4(0u, 0u, 0u, 0u);
          r[4 * i] = v4.x; r[4 * i + 1] = v4.y; r[4 * i + 2] = v4.z; r[4 * i + 3] = v4.w;
        }
        tmem_st32(tQ + hlf * 32, r);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&q_ready[X]);
    }

    float m_used = -INFINITY;
    float l = 0.f;
    for (int j = kv_begin; j < kv_end; ++j) {   // global key-tile index (kv_begin even: parity of j == local parity)
      mbar_wait(&s_full[X], j & 1);
      tc_fence_after();
      const int kv_rem = p.Lk - j * 64;
      if (kv_rem < 64) {   // last, partial key tile: out-of-range columns of S become -inf (cold path)
#pragma unroll 1
        for (int c = kv_rem >> 5; c < 2; ++c) {
          uint32_t t[32];
          tmem_ld32(tS + c * 32, t);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= kv_rem) t[i] = 0xff800000u;
          tmem_st32(tS + c * 32, t);
        }
        tmem_st_wait();
      }
      uint32_t s[2][32];
      tmem_ld32(tS + 0, s[0]);
      tmem_ld32(tS + 32, s[1]);
      tmem_ld_wait();
      float mxa[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) mxa[a] = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) mxa[i & 7] = fmaxf(mxa[i & 7], __uint_as_float(s[c][i]));
      const float mx = fmaxf(fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])),
                             fmaxf(fmaxf(mxa[4], mxa[5]), fmaxf(mxa[6], mxa[7])));
      const float ms = mx * sc;
      if (j == kv_begin) {
        m_used = ms;
      } else {
        const bool need = ms > m_used + 8.0f;   // lazy rescale: only when a row max grew by more than 2^8
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = fmaxf(m_used, ms);
          const float alpha = fast_exp2(m_used - m_new);
          l *= alpha;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld32(tO + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tO + c * 32, o);
          }
          tmem_st_wait();
          m_used = m_new;
        }
      }
      const uint64_t sc2 = f2_pack(sc, sc);
      const uint64_t negm2 = f2_pack(-m_used, -m_used);
      uint64_t ls2[2] = {0ull, 0ull};
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int c0 = 2 * i;
        const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(s[c0 >> 5][c0 & 31]), __uint_as_float(s[(c0 + 1) >> 5][(c0 + 1) & 31])),
                                   sc2, negm2);
        uint64_t p2;
        float p0, p1;
        if ((EMU > 0) && (i % EMU == EMU - 1)) {
          p2 = exp2_poly2(x2);
          f2_unpack(p2, p0, p1);
        } else {
          float x0, x1;
          f2_unpack(x2, x0, x1);
          p0 = fast_exp2(x0);
          p1 = fast_exp2(x1);
          p2 = f2_pack(p0, p1);
        }
        ls2[i & 1] = f2_add(ls2[i & 1], p2);
        pk[i] = pack_bf16x2(p0, p1);
        if (HALVES && i == 15) {   // keys 0..31 are ready: 16 packed columns
          uint32_t h0[16];
#pragma unroll
          for (int t = 0; t < 16; ++t) h0[t] = pk[t];
          tmem_st16(tS, h0);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&p_ready[2 * X]);
        }
      }
      if (HALVES) {
        uint32_t h1[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) h1[t] = pk[16 + t];
        tmem_st16(tS + 16, h1);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_ready[2 * X + 1]);
      } else {
        tmem_st32(tS, pk);     // P_X: 64 keys = 32 packed columns over the first half of S_X
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_ready[2 * X]);
      }
      {
        float a0, a1, b0, b1;
        f2_unpack(ls2[0], a0, a1);
        f2_unpack(ls2[1], b0, b1);
        l += (a0 + a1) + (b0 + b1);
      }
    }

    // epilogue: O / l -> bf16 -> global [Lq, heads*128] (or the owner rank's receive buffer)
    mbar_wait(&o_done[X], 0);
    tc_fence_after();
    if (static_cast<int>(blockIdx.x) >= p.full_units) {
      // KV segment of a tail unit: leave (O, m, l) for attention64_combine_kernel
      const long long prow = static_cast<long long>(blockIdx.x - p.full_units) * 256 + X * 128 + row_in_tile;
      float* wo = p.ws_o + prow * 128;
      *reinterpret_cast<float2*>(p.ws_ml + prow * 2) = make_float2(m_used, l);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t o[32];
        tmem_ld32(tO + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<uint4*>(wo + c * 32 + 4 * i) = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
      }
    } else {
    const float inv = 1.0f / l;
    __nv_bfloat16* orow = p.out + static_cast<long long>(q_row) * p.ldo + head * 128;
    if (p.sp_world > 1) {
      const int owner = q_row / p.sp_Lp;
      const int t = q_row - owner * p.sp_Lp;
      if (owner < p.sp_world)
        orow = p.out_peers[owner] + (static_cast<long long>(p.sp_rank) * p.sp_Lp + t) * p.ldo + head * 128;
    }
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      tmem_ld32(tO + c * 32, o);
      tmem_ld_wait();
      if (q_row < p.Lq) {
        uint4* o4 = reinterpret_cast<uint4*>(orow + c * 32);
        if (p.accumulate) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 prev = o4[i];
            const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(&prev);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = __bfloat1622float2(ph[k]);
              o[8 * i + 2 * k] = __float_as_uint(__uint_as_float(o[8 * i + 2 * k]) + f.x * l);
              o[8 * i + 2 * k + 1] = __float_as_uint(__uint_as_float(o[8 * i + 2 * k + 1]) + f.y * l);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
          o4[i] = w;
        }
      }
    }
    }  // whole unit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// merge of the KV-segment partials of the tail units + the normal epilogue (same math as attention_combine_kernel)
__global__ void __launch_bounds__(256) attention64_combine_kernel(const Att64Params p, int tail_units) {
  const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gw >= tail_units * 256) return;
  const int tu = gw >> 8, r = gw & 255;
  const int unit = p.full_units + tu;
  const int head = unit / p.nq;
  const int q_row = (unit - head * p.nq) * 256 + r;
  if (q_row >= p.Lq) return;
  float M = -INFINITY;
  for (int sgm = 0; sgm < p.ns; ++sgm) M = fmaxf(M, p.ws_ml[((static_cast<long long>(tu) * p.ns + sgm) * 256 + r) * 2]);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float lsum = 0.f;
  for (int sgm = 0; sgm < p.ns; ++sgm) {
    const long long prow = (static_cast<long long>(tu) * p.ns + sgm) * 256 + r;
    const float2 ml = *reinterpret_cast<const float2*>(p.ws_ml + prow * 2);
    const float w = exp2f(ml.x - M);
    const float4 o = *reinterpret_cast<const float4*>(p.ws_o + prow * 128 + lane * 4);
    acc.x += o.x * w; acc.y += o.y * w; acc.z += o.z * w; acc.w += o.w * w;
    lsum += ml.y * w;
  }
  const float inv = 1.0f / lsum;
  __nv_bfloat16* orow = p.out + static_cast<long long>(q_row) * p.ldo + head * 128;
  if (p.sp_world > 1) {
    const int owner = q_row / p.sp_Lp;
    const int t = q_row - owner * p.sp_Lp;
    if (owner < p.sp_world)
      orow = p.out_peers[owner] + (static_cast<long long>(p.sp_rank) * p.sp_Lp + t) * p.ldo + head * 128;
  }
  *reinterpret_cast<uint2*>(orow + lane * 4) = make_uint2(pack_bf16x2(acc.x * inv, acc.y * inv), pack_bf16x2(acc.z * inv, acc.w * inv));
}

static int split_workspace64(size_t ctas, float** ws_o, float** ws_ml, cudaStream_t stream) {
  static float* buf[16] = {nullptr};
  static size_t cap[16] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return YB_ERR_LAUNCH;
  const size_t need = ctas * 256 * 130 * sizeof(float);
  if (cap[dev] < need) {
    if (buf[dev]) {
      cudaStreamSynchronize(stream);
      cudaFree(buf[dev]);
      buf[dev] = nullptr;
      cap[dev] = 0;
    }
    if (cudaMalloc(&buf[dev], need) != cudaSuccess) {
      (void)cudaGetLastError();
      return YB_ERR_LAUNCH;
    }
    cap[dev] = need;
  }
  *ws_o = buf[dev];
  *ws_ml = buf[dev] + ctas * 256 * 128;
  return YB_OK;
}

template <int EMU, bool HALVES>
static int launch64(const CUtensorMap& tmK, const CUtensorMap& tmV, Att64Params p, int heads, int flags, cudaStream_t stream) {
  auto kern = attention64_kernel<EMU, HALVES>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, A64_SMEM_BYTES) != cudaSuccess) {
      (void)cudaGetLastError();
      return YB_ERR_LAUNCH;
    }
    attr_set = true;
  }
  // same planner as the product kernel (yb_attention_plan counts 128-key tiles: hand it an Lk with the same tile count)
  int plan[4];
  int prc = yb_attention_plan(p.Lq, p.nkv * 128, heads, sm_count(), flags & ~YB_ATT_Q64, plan);
  if (prc) return prc;
  p.nq = (p.Lq + 255) / 256;
  p.full_units = plan[0];
  p.ns = plan[2];
  p.ws_o = p.ws_ml = nullptr;
  const int tail = plan[1];
  if (tail > 0) {
    prc = split_workspace64(static_cast<size_t>(tail) * p.ns, &p.ws_o, &p.ws_ml, stream);
    if (prc) return prc;
  }
  kern<<<p.full_units + tail * p.ns, A64_THREADS, A64_SMEM_BYTES, stream>>>(tmK, tmV, p);
  int rc = check_launch("attention64");
  if (rc || tail == 0) return rc;
  attention64_combine_kernel<<<tail * 32, 256, 0, stream>>>(p, tail);
  return check_launch("attention64_combine");
}

// Called by yb_attention_ex when YB_ATT_Q64 is set (attention.cu). out_peers / world / rank / Lp as for yb_attention_sp
// (world <= 1: plain output).
int attention64_launch(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* out,
                       long long ldo, int Lq, int Lk, int heads, float scale, int flags, void* const* out_peers, int world,
                       int rank, int Lp, cudaStream_t stream) {
  if ((ldq % 8) || (reinterpret_cast<uintptr_t>(q) & 0xF)) return YB_ERR_ALIGNMENT;
  CUtensorMap tmK, tmV;
  const uint64_t cols = static_cast<uint64_t>(heads) * 128;
  int rc = make_tmap_bf16_2d(&tmK, k, Lk, cols, ldk, 64, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmV, v, Lk, cols, ldv, 64, 64);
  if (rc) return rc;
  Att64Params p;
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.ldq = ldq;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.Lq = Lq;
  p.Lk = Lk;
  p.nkv = (Lk + 63) / 64;
  p.accumulate = (flags & YB_ATT_ACCUMULATE) ? 1 : 0;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.sp_world = world > 1 ? world : 1;
  p.sp_rank = rank;
  p.sp_Lp = Lp;
  for (int i = 0; i < 8; ++i) p.out_peers[i] = (out_peers && i < world) ? static_cast<__nv_bfloat16*>(out_peers[i]) : nullptr;
  const char* hv = getenv("YB_ATT64_HALVES");
  if (hv != nullptr && hv[0] == '1') {
    switch ((flags >> 2) & 3) {
      case 1: return launch64<4, true>(tmK, tmV, p, heads, flags, stream);
      case 2: return launch64<3, true>(tmK, tmV, p, heads, flags, stream);
      case 3: return launch64<2, true>(tmK, tmV, p, heads, flags, stream);
      default: return launch64<0, true>(tmK, tmV, p, heads, flags, stream);
    }
  }
  switch ((flags >> 2) & 3) {
    case 1: return launch64<4, false>(tmK, tmV, p, heads, flags, stream);
    case 2: return launch64<3, false>(tmK, tmV, p, heads, flags, stream);
    case 3: return launch64<2, false>(tmK, tmV, p, heads, flags, stream);
    default: return launch64<0, false>(tmK, tmV, p, heads, flags, stream);
  }
}

}  // namespace yb
