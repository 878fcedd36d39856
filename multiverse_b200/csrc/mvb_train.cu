// Backward of the ConvLSTM cell (BPTT step), for Trainer (code/pred_models.py:1636-1742; the
// reference gets these from tf.gradients :1698 through tf.contrib.rnn.ConvLSTMCell).
//
//   lstm_gates_bwd   pointwise: (dh_t, dc_t, gates_t, c_{t-1}, c_t) -> dG_t (pre-activation gate
//                    gradients, bf16 planes, packed column order), dc_{t-1}, dbias
//   cell dgrad       dxh[r, c]  = sum_tap sum_g dG[r - shift(tap), g] * W[tap, c, g]
//                    = one tcgen05 implicit GEMM  [R, 9*1024] x [9*1024, cpad]   (same halo trick
//                    as the forward: a tap is a constant row shift of the dG matrix)
//   cell wgrad       dW[tap, c, g] = sum_r xh[r + shift(tap), c] * dG[r, g]
//                    = tcgen05 GEMM  dG^T [1024, R] x xh^T_tap [cpad, R]^T over 9 tap-shifted
//                    transposes of the activations; 8 x 18 output tiles = one wave of CTAs,
//                    each looping over all R rows (K) and adding into the fp32 dW accumulator
// Both GEMMs use the forward kernel's operand-plane scheme (P bf16 planes, products i+j<P) and
// the same TMA / mbarrier / TMEM pipeline.  dgrad tiles: two N tiles of cpad/2 when the x block is
// needed, one N = 256 tile (h block only) when it is not (regression encoder).  Measured on B200: MMA
// time is proportional to N (no granule penalty at N = 144); a separate 32-wide x tile is TMA-bound and
// costs 30 % of an h tile; a cta_group::2 variant gave no gain (the kernel is not smem-bandwidth bound).
// Algorithmic FLOPs: dgrad = wgrad = forward (2*R*9*cpad*1024 each).
#include "mvb_common.cuh"
#include "mvb_kernels.h"
#include <stdlib.h>

namespace mvb {

constexpr int G_BLOCK_M = 128;
constexpr int G_BLOCK_K = 32;
constexpr int G_UMMA_K = 16;
constexpr int G_MAX_BN = 256;
constexpr int G_EPI_WARPS = 4;
constexpr int G_THREADS = 128 + 32 * G_EPI_WARPS;
constexpr int G_A_PLANE = G_BLOCK_M * G_BLOCK_K * 2;   // 8 KB
constexpr int G_B_PLANE = G_MAX_BN * G_BLOCK_K * 2;    // 16 KB (smaller N tiles use part of it)
constexpr uint32_t G_SW64_LAYOUT = 4;
constexpr uint32_t G_SW64_SBO = 512;

enum { MODE_DGRAD = 0, MODE_WGRAD = 1, MODE_WGRAD_MN = 2 };

// MT = number of 128-row M sub-tiles of one CTA tile.  MT = 2 halves the B traffic per MMA (the backward GEMMs
// are L2->smem feed bound at MT = 1: 79 / 69 B/clk/SM against ~60 the forward sustains); its two accumulators
// fill TMEM's 512 columns, so there is one accumulator stage and the (light) epilogue is not overlapped.
template <int P, int MT> struct GemmCfg {
  static constexpr int A_BYTES = MT * P * G_A_PLANE;
  static constexpr int STAGE_BYTES = A_BYTES + P * G_B_PLANE;
  static constexpr int STAGES = (227 * 1024 - 2048) / STAGE_BYTES > 8 ? 8 : (227 * 1024 - 2048) / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int NACC = MT == 1 ? 2 : 1;
};

struct GemmParams {
  float* out;          // dgrad: [R, cpad] fp32;  wgrad: [1024, 9*cpad] fp32 accumulator (+=)
  long long R;         // halo rows
  int H, W;
  int cpad, bn;        // N tile of the wgrad modes
  int cxp;             // dgrad: width of the x block
  int need_x;          // dgrad: 1 -> two N tiles of cpad/2 covering [0, cpad); 0 -> one N tile [cxp, cxp+256) (h only)
  int num_kb;          // k-blocks per tile
  long long num_m_tiles;
  int num_n_tiles;
  // MODE_WGRAD_MN: operands are read MN-major straight from the row-major activations
  int ubn, nb;         // channels / 32-channel blocks of one (tap, chunk) unit
  int n_per_tap;       // units per tap (cpad / ubn)
  int upt;             // units per N tile (bn = upt * ubn): two 96-wide units are paired into N = 192
  int n_units;         // 9 * n_per_tap
  int ksplit;          // K (= halo rows) is split over this many work items, each with its own fp32 slab
  uint32_t lbo, sbo;   // UMMA descriptor strides of the MN-major SWIZZLE_64B tiles
};

template <int P, int MODE, int MT>
__global__ void __launch_bounds__(G_THREADS, 1)
pgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const GemmParams prm) {
  using Cfg = GemmCfg<P, MT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Grid g = make_grid(prm.H, prm.W);
  const long long num_tiles = prm.num_m_tiles * prm.num_n_tiles;     // num_m_tiles counts CTA tiles of MT*128 rows
  const uint32_t stage_tx = (uint32_t)(Cfg::A_BYTES + P * prm.bn * G_BLOCK_K * 2);
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(prm.bn >> 3) << 17) |
                         ((uint32_t)(G_BLOCK_M >> 4) << 24) |
                         (MODE == MODE_WGRAD_MN ? ((1u << 15) | (1u << 16)) : 0u);   // A, B MN-major
  // byte offset of (M sub-tile j, plane pa) inside a stage's A region
  auto a_off = [&](int j, int pa) -> uint32_t {
    return MODE == MODE_WGRAD_MN ? (uint32_t)(pa * MT * G_A_PLANE + j * G_A_PLANE)    // one TMA box [P][4*MT blocks]
                                 : (uint32_t)((j * P + pa) * G_A_PLANE);              // MT boxes of [P][128 rows]
  };

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], G_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0;
    for (long long t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const long long mt = t / prm.num_n_tiles;
      const int ntile = (int)(t % prm.num_n_tiles);
      for (int kb = 0; kb < prm.num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        mbar_expect_tx(&full_bar[stage], stage_tx);
        if (MODE == MODE_DGRAD) {
          // A = dG[rows - shift(tap), 32 gate columns];  B = Wd[N channels, tap*1024 + 32 gate columns]
          const int q = kb / 9, tap = kb - q * 9;
          const int shift = (tap / 3 - 1) * g.Wp + (tap % 3 - 1);
          for (int j = 0; j < MT; ++j)
            tma_load_3d(sa + a_off(j, 0), &tmA, &full_bar[stage], q * G_BLOCK_K,
                        (int)((mt * MT + j) * G_BLOCK_M - shift), 0);
          tma_load_3d(sb, &tmB, &full_bar[stage], tap * kGates + q * G_BLOCK_K,
                      prm.need_x ? ntile * prm.bn : prm.cxp, 0);
        } else if (MODE == MODE_WGRAD) {
          // A = dG^T[128 gate rows, 32 halo rows];  B = tap-shifted xh^T[tap][bn channels, 32 halo rows]
          const int tap = ntile >> 1, half = ntile & 1;
          tma_load_3d(sa, &tmA, &full_bar[stage], kb * G_BLOCK_K, (int)(mt * G_BLOCK_M), 0);
          tma_load_3d(sb, &tmB, &full_bar[stage], kb * G_BLOCK_K, tap * prm.cpad + half * prm.bn, 0);
        } else {
          // MN-major: A = dG[32 halo rows (K), 4*MT blocks of 32 gate columns]; B = `upt` units, each
          // xh[32 halo rows + shift(tap), nb blocks of 32 channels]; ntile = (unit group, k-split)
          const int ks = ntile % prm.ksplit, grp = ntile / prm.ksplit;
          const int k0 = (ks * prm.num_kb + kb) * G_BLOCK_K;
          tma_load_4d(sa, &tmA, &full_bar[stage], 0, k0, (int)mt * 4 * MT, 0);
          for (int j = 0; j < prm.upt; ++j) {
            int u = grp * prm.upt + j;
            if (u >= prm.n_units) u = prm.n_units - 1;     // odd tail: duplicate, discarded by the epilogue
            const int tap = u / prm.n_per_tap, chunk = u - tap * prm.n_per_tap;
            const int shift = (tap / 3 - 1) * g.Wp + (tap % 3 - 1);
            for (int p = 0; p < P; ++p)
              tma_load_4d(sb + p * (prm.bn * 64) + j * (prm.ubn * 64), &tmB, &full_bar[stage], 0, k0 + shift,
                          chunk * prm.nb, p);
          }
        }
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    int stage = 0; uint32_t phase = 0;
    long long it = 0;
    for (long long t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int as = (int)(it % Cfg::NACC);
      const uint32_t aphase = (uint32_t)((it / Cfg::NACC) & 1);
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * MT * 256;
      const uint32_t b_plane = (uint32_t)prm.bn * G_BLOCK_K * 2;   // TMA packs planes back to back
      for (int kb = 0; kb < prm.num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          uint32_t first = (kb == 0) ? 0u : 1u;
#pragma unroll
          for (int pa = 0; pa < P; ++pa) {
#pragma unroll
            for (int pb = 0; pb < P - pa; ++pb) {
#pragma unroll
              for (int k = 0; k < G_BLOCK_K / G_UMMA_K; ++k) {
                uint64_t ad, bd;
                if (MODE == MODE_WGRAD_MN) {
                  // one UMMA consumes 16 K rows = two 8-row groups (sbo apart) of every 32-wide MN block
                  ad = make_smem_desc(sa + a_off(j, pa) + k * 2 * prm.sbo, prm.sbo, G_SW64_LAYOUT, prm.lbo);
                  bd = make_smem_desc(sb + pb * b_plane + k * 2 * prm.sbo, prm.sbo, G_SW64_LAYOUT, prm.lbo);
                } else {
                  ad = make_smem_desc(sa + a_off(j, pa) + k * G_UMMA_K * 2, G_SW64_SBO, G_SW64_LAYOUT);
                  bd = make_smem_desc(sb + pb * b_plane + k * G_UMMA_K * 2, G_SW64_SBO, G_SW64_LAYOUT);
                }
                umma_bf16(d_tmem + j * 256, ad, bd, idesc, first);
                first = 1u;
              }
            }
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tfull_bar[as]);
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> fp32 global =====================
    const int wq = warp & 3;
    long long it = 0;
    for (long long t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int as = (int)(it % Cfg::NACC);
      const uint32_t aphase = (uint32_t)((it / Cfg::NACC) & 1);
      const long long mt = t / prm.num_n_tiles;
      const int ntile = (int)(t % prm.num_n_tiles);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < MT; ++j) {
        const long long row = (mt * MT + j) * G_BLOCK_M + wq * 32 + lane;
        bool valid;
        float* dst;
        if (MODE == MODE_DGRAD) {
          valid = row < prm.R;
          if (valid) {
            const int rem = (int)(row % g.S);
            const int y = rem / g.Wp, x = rem - y * g.Wp;
            valid = (x < g.W) && (y < g.H);
          }
          dst = prm.out + row * prm.cpad + (prm.need_x ? ntile * prm.bn : prm.cxp);
        } else if (MODE == MODE_WGRAD) {
          valid = row < kGates;
          const int tap = ntile >> 1, half = ntile & 1;
          dst = prm.out + row * (9LL * prm.cpad) + tap * prm.cpad + half * prm.bn;
        } else {
          valid = row < kGates;
          dst = prm.out + (long long)(ntile % prm.ksplit) * kGates * 9LL * prm.cpad + row * (9LL * prm.cpad);
        }
        const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + (as * MT + j) * 256;
        for (int c0 = 0; c0 < prm.bn; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(t_row + c0, v);
          tmem_ld_wait();
          bool ok = valid;
          float4* d4 = reinterpret_cast<float4*>(dst + c0);
          if (MODE == MODE_WGRAD_MN) {
            // column c0 of the tile -> (unit, channel): dW[tap][chunk*ubn + c]; this (tile, k-split) owns its slab
            const int ju = c0 / prm.ubn;
            const int u = (ntile / prm.ksplit) * prm.upt + ju;
            ok = valid && u < prm.n_units;
            const int tap = u / prm.n_per_tap, chunk = u - tap * prm.n_per_tap;
            d4 = reinterpret_cast<float4*>(dst + tap * prm.cpad + chunk * prm.ubn + (c0 - ju * prm.ubn));
          }
          if (ok) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4 o = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                     __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
              if (MODE != MODE_DGRAD) { const float4 old = d4[q]; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
              d4[q] = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ----------------------------------------------------------------------------------
// pointwise LSTM backward
//   gates [R,1024] (activated i,j,f,o; packed column order tile*256 + gate*64 + j), c_prev, c [R,256]
//   dc_total = dc_in + dh*o*(1-tanh(c)^2);  dG = {di_pre, dj_pre, df_pre, do_pre};  dc_prev = dc_total*f
// block = 32 channel groups (8 channels each) x 8 rows
// ----------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(256)
lstm_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                const float* __restrict__ c_new, const float* __restrict__ dh,
                const float* __restrict__ dc_in, __nv_bfloat16* __restrict__ dg_planes,
                long long plane_stride, float* __restrict__ dc_prev, float* __restrict__ dbias,
                long long NS, Grid g) {
  const int cg = threadIdx.x & 31;     // channel group: channels [8cg, 8cg+8)
  const int rl = threadIdx.x >> 5;     // row lane 0..7
  const int ch0 = cg * 8;
  const int colbase = (ch0 / 64) * 256 + (ch0 % 64);
  const int hw = g.H * g.W;
  float bsum[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int k = 0; k < 8; ++k) bsum[a][k] = 0.f;
  for (long long pix = (long long)blockIdx.x * 8 + rl; pix < NS * hw; pix += (long long)gridDim.x * 8) {
    const long long s = pix / hw;
    const int p = (int)(pix - s * hw);
    const long long row = s * g.S + (long long)(p / g.W) * g.Wp + (p % g.W);
    float gt[4][8], cp[8], cn[8], dhv[8], dcv[8];
    auto ld8 = [](const float* ptr, float (&o)[8]) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(ptr)), b = __ldg(reinterpret_cast<const float4*>(ptr) + 1);
      o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    };
#pragma unroll
    for (int a = 0; a < 4; ++a) ld8(gates + row * kGates + colbase + a * 64, gt[a]);
    if (c_prev) ld8(c_prev + row * kHidden + ch0, cp);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) cp[k] = 0.f;
    }
    ld8(c_new + row * kHidden + ch0, cn);
    ld8(dh + row * kHidden + ch0, dhv);
    if (dc_in) ld8(dc_in + row * kHidden + ch0, dcv);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) dcv[k] = 0.f;
    }
    float dgv[4][8], dcp[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float ai = gt[0][k], aj = gt[1][k], af = gt[2][k], ao = gt[3][k];
      const float th = tanh_acc(cn[k]);
      const float dct = dcv[k] + dhv[k] * ao * (1.f - th * th);
      dgv[0][k] = dct * aj * ai * (1.f - ai);
      dgv[1][k] = dct * ai * (1.f - aj * aj);
      dgv[2][k] = dct * cp[k] * af * (1.f - af);
      dgv[3][k] = dhv[k] * th * ao * (1.f - ao);
      dcp[k] = dct * af;
#pragma unroll
      for (int a = 0; a < 4; ++a) bsum[a][k] += dgv[a][k];
    }
    float4* dc4 = reinterpret_cast<float4*>(dc_prev + row * kHidden + ch0);
    dc4[0] = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
    dc4[1] = make_float4(dcp[4], dcp[5], dcp[6], dcp[7]);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      uint32_t pk[P][4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        __nv_bfloat16 x0[P], x1[P];
        split_planes<P>(dgv[a][2 * v], x0);
        split_planes<P>(dgv[a][2 * v + 1], x1);
#pragma unroll
        for (int q = 0; q < P; ++q) pk[q][v] = pack_bf16x2(x0[q], x1[q]);
      }
#pragma unroll
      for (int q = 0; q < P; ++q)
        *reinterpret_cast<uint4*>(dg_planes + q * plane_stride + row * kGates + colbase + a * 64) =
            make_uint4(pk[q][0], pk[q][1], pk[q][2], pk[q][3]);
    }
  }
  // dbias: reduce the 8 row lanes through shared memory, one atomic per column per block
  __shared__ float red[8][32][33];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int k = 0; k < 8; ++k) red[rl][cg][a * 8 + k] = bsum[a][k];
  __syncthreads();
  if (rl == 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float sacc = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) sacc += red[r][cg][a * 8 + k];
        atomicAdd(dbias + colbase + a * 64 + k, sacc);
      }
  }
}

// src [P][R][C] bf16 -> dst [P][T][C][Rp] bf16 (64x64 tiles through shared memory).
// T == 1: plain transpose.  T == 9: one copy per 3x3 tap with the tap's row shift applied,
// dst[p][tap][c][r] = src[p][r + shift(tap)][c] (zero outside) - TMA needs 16-byte aligned inner
// coordinates, so the wgrad GEMM cannot shift along its contiguous K (= row) axis itself.
__global__ void __launch_bounds__(256)
transpose_planes_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                        long long R, int C, long long Rp, int taps, int Wp) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int p = blockIdx.z / taps, tap = blockIdx.z % taps;
  const long long shift = taps == 9 ? (long long)(tap / 3 - 1) * Wp + (tap % 3 - 1) : 0;
  const long long r0 = (long long)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const __nv_bfloat16* s = src + (long long)p * R * C;
  __nv_bfloat16* d = dst + ((long long)p * taps + tap) * C * Rp;
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int rr = i / 64, cc = i % 64;
    const long long sr = r0 + rr + shift;
    tile[rr][cc] = (sr >= 0 && sr < R && c0 + cc < C) ? s[sr * C + c0 + cc] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int cc = i / 64, rr = i % 64;
    if (c0 + cc < C && r0 + rr < Rp) d[(long long)(c0 + cc) * Rp + r0 + rr] = tile[rr][cc];
  }
}

// dgrad weights: Wd planes [P][cpad][9*1024], Wd[kc][tap*1024 + n_packed] = W_tf[tap][cin(kc)][col(n_packed)]
template <int P>
__global__ void pack_dgrad_kernel(const float* __restrict__ kernel, __nv_bfloat16* __restrict__ wd,
                                  int cx, int cxp, int cpad) {
  const long long ktot = 9LL * kGates;
  const long long total = (long long)cpad * ktot;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int kc = (int)(i / ktot);
    const int k = (int)(i - (long long)kc * ktot);
    const int tap = k / kGates, n = k - tap * kGates;
    const int tile = n / 256, gate = (n % 256) / 64, j = n % 64;
    const int col = gate * kHidden + tile * 64 + j;
    int cin = -1;
    if (kc < cx) cin = kc;
    else if (kc >= cxp) cin = cx + (kc - cxp);
    const float v = (cin >= 0) ? kernel[((long long)tap * (cx + kHidden) + cin) * kGates + col] : 0.f;
    __nv_bfloat16 pl[P];
    split_planes<P>(v, pl);
#pragma unroll
    for (int p = 0; p < P; ++p) wd[(long long)p * total + i] = pl[p];
  }
}

// dWp [1024][9*cpad] fp32 (packed) -> dkernel [3,3,cx+256,1024] (TF layout), dbias packed -> TF order
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, const float* __restrict__ dbp,
                                    float* __restrict__ dkernel, float* __restrict__ dbiases, int cx,
                                    int cxp, int cpad, int comp, int accumulate, int slabs) {
  const int cin_tot = cx + kHidden;
  const long long total = 9LL * cin_tot * kGates;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % kGates);
    const int cin = (int)((i / kGates) % cin_tot);
    const int tap = (int)(i / ((long long)kGates * cin_tot));
    const int gate = col / kHidden, ch = col % kHidden;
    const int n = (ch / 64) * 256 + gate * 64 + (ch % 64);
    const int kc = cin < cx ? cin : cxp + (cin - cx);
    float v = 0.f;
    for (int sl = 0; sl < slabs; ++sl) {
      const float* d = dwp + (long long)sl * kGates * 9LL * cpad + (long long)n * (9LL * cpad) + tap * cpad;
      v += d[kc];
      if (comp && cin < cx) v += d[cx + cin];   // + residual block of the compensated x block
    }
    dkernel[i] = accumulate ? dkernel[i] + v : v;
    if (tap == 0 && cin == 0) dbiases[col] = accumulate ? dbiases[col] + dbp[n] : dbp[n];
  }
}

template <int P, int MODE, int MT>
static int launch_pgemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& prm, int num_sms,
                        cudaStream_t stream) {
  using Cfg = GemmCfg<P, MT>;
  static SmemOptIn opt;
  MVB_CHECK_CUDA(smem_opt_in(opt, pgemm_kernel<P, MODE, MT>, Cfg::SMEM_BYTES));
  const long long tiles = prm.num_m_tiles * prm.num_n_tiles;
  const int grid = (int)(tiles < num_sms ? tiles : num_sms);
  pgemm_kernel<P, MODE, MT><<<grid, G_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, prm);
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

static int num_sms_of_device(int* out) {
  int dev = 0;
  MVB_CHECK_CUDA(cudaGetDevice(&dev));
  MVB_CHECK_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
  return MVB_OK;
}

int cell_dgrad(const void* dg_planes, const void* wd_planes, float* dxh, long long NS, int H, int W,
               int cpad, int P, int need_x, cudaStream_t stream) {
  MVB_REQUIRE(P >= 1 && P <= 3, "cell_dgrad: planes P=%d", P);
  MVB_REQUIRE(dg_planes && wd_planes && dxh && NS > 0, "cell_dgrad: bad args");
  const int cxp = cpad - kHidden;
  MVB_REQUIRE(cpad % 32 == 0 && cxp >= 32 && cxp <= 256 && cxp % 16 == 0, "cell_dgrad: cpad=%d unsupported", cpad);
  const Grid g = make_grid(H, W);
  const long long R = NS * g.S;
  CUtensorMap tmA, tmB;
  int rc = encode_tmap_3d_bf16(&tmA, dg_planes, kGates, (uint64_t)R, P, kGates * 2ull, (uint64_t)R * kGates * 2,
                               G_BLOCK_K, G_BLOCK_M, P, 64);
  if (rc) return rc;
  const uint64_t ktot = 9ull * kGates;
  // with the x block: two N tiles of cpad/2 (144 / 160); h only: one N tile of 256.  (Measured: MMA time is
  // proportional to N, and a separate 32-wide x tile is TMA-bound and costs 30 % of an h tile.)
  const int bn = need_x ? cpad / 2 : 256;
  MVB_REQUIRE(bn % 16 == 0, "cell_dgrad: cpad=%d unsupported", cpad);
  rc = encode_tmap_3d_bf16(&tmB, wd_planes, ktot, (uint64_t)cpad, P, ktot * 2, ktot * cpad * 2, G_BLOCK_K, bn, P, 64);
  if (rc) return rc;
  GemmParams prm = {};
  prm.out = dxh; prm.R = R; prm.H = H; prm.W = W; prm.cpad = cpad; prm.bn = bn; prm.cxp = cxp; prm.need_x = need_x;
  prm.num_kb = 9 * (kGates / G_BLOCK_K);
  // with the x block (N = 144 / 160) a CTA tile is 256 rows (two accumulators sharing every B tile);
  // the h-only N = 256 tile already has the forward kernel's operand intensity
  const int mt_sub = need_x ? 2 : 1;
  prm.num_m_tiles = (R + G_BLOCK_M * mt_sub - 1) / (G_BLOCK_M * mt_sub);
  prm.num_n_tiles = need_x ? 2 : 1;
  int sms = 0;
  if ((rc = num_sms_of_device(&sms))) return rc;
  if (need_x) {
    switch (P) {
      case 1: return launch_pgemm<1, MODE_DGRAD, 2>(tmA, tmB, prm, sms, stream);
      case 2: return launch_pgemm<2, MODE_DGRAD, 2>(tmA, tmB, prm, sms, stream);
      default: return launch_pgemm<3, MODE_DGRAD, 2>(tmA, tmB, prm, sms, stream);
    }
  }
  switch (P) {
    case 1: return launch_pgemm<1, MODE_DGRAD, 1>(tmA, tmB, prm, sms, stream);
    case 2: return launch_pgemm<2, MODE_DGRAD, 1>(tmA, tmB, prm, sms, stream);
    default: return launch_pgemm<3, MODE_DGRAD, 1>(tmA, tmB, prm, sms, stream);
  }
}

int cell_wgrad(const void* dgT_planes, const void* xhT_planes, float* dwp, long long NS, int H, int W,
               int cpad, long long Rp, int P, cudaStream_t stream) {
  MVB_REQUIRE(P >= 1 && P <= 3, "cell_wgrad: planes P=%d", P);
  MVB_REQUIRE(dgT_planes && xhT_planes && dwp && NS > 0, "cell_wgrad: bad args");
  MVB_REQUIRE(cpad % 32 == 0 && (cpad / 2) % 16 == 0 && cpad / 2 <= G_MAX_BN, "cell_wgrad: cpad=%d unsupported", cpad);
  const Grid g = make_grid(H, W);
  const long long R = NS * g.S;
  MVB_REQUIRE(Rp >= R && Rp % 8 == 0, "cell_wgrad: Rp=%lld must be >= R and a multiple of 8", Rp);
  CUtensorMap tmA, tmB;
  int rc = encode_tmap_3d_bf16(&tmA, dgT_planes, (uint64_t)R, kGates, P, (uint64_t)Rp * 2, (uint64_t)Rp * kGates * 2,
                               G_BLOCK_K, G_BLOCK_M, P, 64);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&tmB, xhT_planes, (uint64_t)R, 9ull * cpad, P, (uint64_t)Rp * 2, (uint64_t)Rp * cpad * 18,
                           G_BLOCK_K, cpad / 2, P, 64);
  if (rc) return rc;
  GemmParams prm = {};
  prm.out = dwp; prm.R = R; prm.H = H; prm.W = W; prm.cpad = cpad; prm.bn = cpad / 2;
  prm.num_kb = (int)((R + G_BLOCK_K - 1) / G_BLOCK_K); prm.num_m_tiles = kGates / G_BLOCK_M; prm.num_n_tiles = 18;
  int sms = 0;
  if ((rc = num_sms_of_device(&sms))) return rc;
  switch (P) {
    case 1: return launch_pgemm<1, MODE_WGRAD, 1>(tmA, tmB, prm, sms, stream);
    case 2: return launch_pgemm<2, MODE_WGRAD, 1>(tmA, tmB, prm, sms, stream);
    default: return launch_pgemm<3, MODE_WGRAD, 1>(tmA, tmB, prm, sms, stream);
  }
}

int cell_wgrad_mn_slabs(int cpad) { return cpad == 288 ? 5 : 2; }

int cell_wgrad_mn(const void* dg_planes, const void* xh_planes, float* dwp, long long NS, int H, int W,
                  int cpad, int P, cudaStream_t stream) {
  MVB_REQUIRE(P >= 1 && P <= 3, "cell_wgrad_mn: planes P=%d", P);
  MVB_REQUIRE(dg_planes && xh_planes && dwp && NS > 0, "cell_wgrad_mn: bad args");
  MVB_REQUIRE(cpad == 288 || cpad == 320, "cell_wgrad_mn: cpad=%d unsupported", cpad);
  const Grid g = make_grid(H, W);
  const long long R = NS * g.S;
  GemmParams prm = {};
  prm.ubn = cpad == 288 ? 96 : 160;
  prm.upt = cpad == 288 ? 2 : 1;          // pair two 96-wide units: N = 192 keeps the MMA off the smem-bandwidth limit
  prm.bn = prm.ubn * prm.upt;
  prm.nb = prm.ubn / 32; prm.n_per_tap = cpad / prm.ubn; prm.n_units = 9 * prm.n_per_tap;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[4] = {32, (uint64_t)R, kGates / 32, (uint64_t)P};
    const uint64_t st[3] = {kGates * 2ull, 64, (uint64_t)R * kGates * 2};
    const uint32_t box[4] = {32, G_BLOCK_K, 8, (uint32_t)P};      // MT = 2: 256 gate columns per CTA tile
    int rc = encode_tmap_4d_bf16(&tmA, dg_planes, dims, st, box, 64);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {32, (uint64_t)R, (uint64_t)cpad / 32, (uint64_t)P};
    const uint64_t st[3] = {(uint64_t)cpad * 2, 64, (uint64_t)R * cpad * 2};
    const uint32_t box[4] = {32, G_BLOCK_K, (uint32_t)prm.nb, 1};
    int rc = encode_tmap_4d_bf16(&tmB, xh_planes, dims, st, box, 64);
    if (rc) return rc;
  }
  prm.out = dwp; prm.R = R; prm.H = H; prm.W = W; prm.cpad = cpad;
  const long long kb_total = (R + G_BLOCK_K - 1) / G_BLOCK_K;
  // work items = 4 M tiles (256 gate columns) x unit groups x k-splits, sized to fill whole waves of 148 CTAs:
  // cpad 288: 4 x 14 x 5 = 280 (1.9 waves); cpad 320: 4 x 18 x 2 = 144
  prm.ksplit = cell_wgrad_mn_slabs(cpad);
  prm.num_kb = (int)((kb_total + prm.ksplit - 1) / prm.ksplit);
  prm.num_m_tiles = kGates / (2 * G_BLOCK_M);
  prm.num_n_tiles = ((prm.n_units + prm.upt - 1) / prm.upt) * prm.ksplit;
  prm.lbo = 32 * 64;   // bytes between 32-wide MN blocks ([32 K rows][64 B] each)
  prm.sbo = 8 * 64;    // bytes between groups of 8 K rows
  int sms = 0;
  int rc = num_sms_of_device(&sms);
  if (rc) return rc;
  switch (P) {
    case 1: return launch_pgemm<1, MODE_WGRAD_MN, 2>(tmA, tmB, prm, sms, stream);
    case 2: return launch_pgemm<2, MODE_WGRAD_MN, 2>(tmA, tmB, prm, sms, stream);
    default: return launch_pgemm<3, MODE_WGRAD_MN, 2>(tmA, tmB, prm, sms, stream);
  }
}

int lstm_gates_bwd(const float* gates, const float* c_prev, const float* c_new, const float* dh,
                   const float* dc_in, void* dg_planes, long long plane_stride, float* dc_prev,
                   float* dbias_packed, long long NS, int H, int W, int P, cudaStream_t stream) {
  MVB_REQUIRE(P >= 1 && P <= 3, "lstm_gates_bwd: planes P=%d", P);
  MVB_REQUIRE(gates && c_new && dh && dg_planes && dc_prev && dbias_packed && NS > 0, "lstm_gates_bwd: bad args");
  const Grid g = make_grid(H, W);
  const long long pix = NS * H * W;
  const int blocks = (int)((pix + 7) / 8 < sm_count() * 8 ? (pix + 7) / 8 : sm_count() * 8);
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dg_planes);
  switch (P) {
    case 1: lstm_bwd_kernel<1><<<blocks, 256, 0, stream>>>(gates, c_prev, c_new, dh, dc_in, d, plane_stride, dc_prev, dbias_packed, NS, g); break;
    case 2: lstm_bwd_kernel<2><<<blocks, 256, 0, stream>>>(gates, c_prev, c_new, dh, dc_in, d, plane_stride, dc_prev, dbias_packed, NS, g); break;
    default: lstm_bwd_kernel<3><<<blocks, 256, 0, stream>>>(gates, c_prev, c_new, dh, dc_in, d, plane_stride, dc_prev, dbias_packed, NS, g); break;
  }
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

int transpose_planes(const void* src, void* dst, long long R, int C, long long Rp, int P, int taps,
                     int Wp, cudaStream_t stream) {
  MVB_REQUIRE(src && dst && R > 0 && C > 0 && Rp >= R && P >= 1 && (taps == 1 || taps == 9), "transpose_planes: bad args");
  dim3 grid((unsigned)((Rp + 63) / 64), (unsigned)((C + 63) / 64), (unsigned)(P * taps));
  transpose_planes_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src),
                                                    reinterpret_cast<__nv_bfloat16*>(dst), R, C, Rp, taps, Wp);
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

int pack_cell_weights_dgrad(const float* kernel, void* wd_planes, int cx, int P, cudaStream_t stream) {
  MVB_REQUIRE(P >= 1 && P <= 3 && kernel && wd_planes && cx >= 1, "pack_cell_weights_dgrad: bad args");
  const int cxp = (cx + 31) / 32 * 32, cpad = cxp + kHidden;
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(wd_planes);
  switch (P) {
    case 1: pack_dgrad_kernel<1><<<1184, 256, 0, stream>>>(kernel, d, cx, cxp, cpad); break;
    case 2: pack_dgrad_kernel<2><<<1184, 256, 0, stream>>>(kernel, d, cx, cxp, cpad); break;
    default: pack_dgrad_kernel<3><<<1184, 256, 0, stream>>>(kernel, d, cx, cxp, cpad); break;
  }
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

int unpack_cell_wgrad(const float* dwp, const float* dbias_packed, float* dkernel, float* dbiases, int cx,
                      int comp, int accumulate, int slabs, cudaStream_t stream) {
  MVB_REQUIRE(dwp && dbias_packed && dkernel && dbiases && cx >= 1, "unpack_cell_wgrad: bad args");
  const int cxp = (cx + 31) / 32 * 32, cpad = cxp + kHidden;
  MVB_REQUIRE(slabs >= 1, "unpack_cell_wgrad: slabs=%d", slabs);
  unpack_wgrad_kernel<<<1184, 256, 0, stream>>>(dwp, dbias_packed, dkernel, dbiases, cx, cxp, cpad, comp, accumulate, slabs);
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

}  // namespace mvb
