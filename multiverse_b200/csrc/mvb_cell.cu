// K-cell: the fused ConvLSTM cell of the Multiverse encoder/decoder.
//
// Replaces, per call, what the reference runs as tf.contrib.rnn.ConvLSTMCell.call
// (built at code/pred_models.py:189-202 and :236-249, driven by dynamic_rnn :212/:232
// and raw_rnn :455/:678):  concat([x,h]) -> conv3x3 SAME -> +biases -> split i,j,f,o ->
// c' = sigmoid(f+forget_bias)*c + sigmoid(i)*tanh(j);  h' = tanh(c')*sigmoid(o).
//
// Formulation: one persistent warp-specialised tcgen05 GEMM
//     G[R, 1024] = A[R, 9*Cpad] * Bt[1024, 9*Cpad]^T
// over the halo layout (mvb_common.cuh): the A k-block for tap t / channel chunk q is
// the TMA box  rows [m0+shift(t), +128) x channels [32q, +32)  of the activation matrix,
// so the 3x3 im2col never exists in memory.  fp32 parity on bf16 tensor cores comes
// from operand planes: x = x0+x1(+x2) with every plane bf16 (mvb::split_planes), and
// the products a_i*b_j with i+j < P are all accumulated into the same fp32 TMEM tile
// (P=2 -> 3 MMAs, error ~2^-17; P=3 -> 6 MMAs, ~2^-24; P=1 -> plain bf16).
// Output columns are gate-interleaved (tile of 256 = 4 gates x 64 channels) so the
// epilogue warps own i,j,f,o of a channel and emit (c', h') directly - the gate
// pre-activations never reach HBM.
//
// Warp roles (384 threads, 1 CTA/SM, persistent over tiles):
//   warp 0   TMA producer      warp 1   MMA issuer (one thread)
//   warp 2   TMEM allocator    warp 3   idle
//   warps 4-11  epilogue: two warpgroups split the 64 channels of a tile; TMEM holds
//               two 256-column accumulators so tile i+1's MMAs overlap tile i's epilogue.
#include "mvb_common.cuh"
#include "mvb_kernels.h"
#include <stdlib.h>

namespace mvb {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int XPAD = 32;      // the x block is zero-padded to a multiple of 32 channels (cpad = roundup(cx,32) + 256)
constexpr int CHUNK = 64;     // channels per K chunk: 64 16-bit elements = 128 B = one SWIZZLE_128B row
constexpr int ROW_BYTES = 128;
constexpr int UMMA_K = 16;
constexpr int TILE_CH = 64;   // hidden channels per N tile
constexpr int N_TILES = kGates / BLOCK_N;  // 4
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + 32 * NUM_EPI_WARPS;
constexpr int B_SLOT_BYTES = BLOCK_N * ROW_BYTES;  // 32 KB: 256 weight rows x 128 B
constexpr uint32_t SW128_LAYOUT = 2;
constexpr uint32_t SW128_SBO = 8 * ROW_BYTES;      // 1024 B between 8-row groups

// Shared-memory rings.  What bounded the round-1 kernel was not the tensor pipe but the operand feed: the TMA unit
// writes about ONE BOX ROW PER CLOCK into shared memory whatever the row's width (measured with the MMAs and the
// epilogue switched off: 275 / 549 / 824 rows per k-block -> 3.2 / 6.0 / 8.6 ms, 32- and 64-byte rows alike), and
// the round-1 stages were made of 32- and 64-byte rows, nine shifted copies of every activation row among them.
//   B ring: slots of 256 rows x 128 B = 64 channels of ONE plane of the weight tile (SWIZZLE_128B): per (chunk,
//           tap) P slots (bf16 planes) or 2 (f16f8: the fp16 plane, then both e4m3 planes interleaved in one row).
//   A ring: one stage per 64-channel chunk: rows [m0 - (Wp+1), m0 + 128 + (Wp+1)) of every activation plane,
//           rounded up to a multiple of 8 rows (RA8).  The nine taps of the chunk read THE SAME stage through UMMA
//           descriptors that start (dy Wp + dx) rows into it: tcgen05 applies the swizzle to absolute shared-memory
//           address bits, so a K-major operand may start at any row of a TMA-written swizzled tile (probed on B200
//           for SWIZZLE_32B / 64B / 128B: tools/umma_rowshift_probe.cu).
// Rows written per 32 channels and tap: 147 (round 1: 768 bf16 x 2, 1152 f16f8); L2 -> SM bytes 18.4 KB (48 KB).
template <int P> struct CellCfg {
  static constexpr int A_STAGES = 2;
  static constexpr int MAX_RA8 = 256;           // TMA box limit: 128 + 2 (W + 2) <= 256  ->  W <= 62
  static constexpr int a_stage_bytes(int ra8) { return P * ra8 * ROW_BYTES; }
  static constexpr int b_slots(int ra8) {       // 4 slots when they fit beside the A ring, else 3
    return (4 * B_SLOT_BYTES + A_STAGES * a_stage_bytes(ra8) + 2048 <= 227 * 1024) ? 4 : 3;
  }
  static constexpr int smem_bytes(int ra8) {
    return b_slots(ra8) * B_SLOT_BYTES + A_STAGES * a_stage_bytes(ra8) + 1024 /*align*/ + 512 /*barriers*/;
  }
};

struct CellParams {
  const float* bias;        // [1024] packed (tile, gate, channel) order
  const float* col_scale;   // f16f8 only: [1024] 2^-S of every packed column (the weights are stored times 2^S)
  const float* c_in;        // [R_src, 256] or nullptr (zero state)
  const int* row_map;       // [NS] source sample-row of c_in, or nullptr (identity)
  float* c_out;             // [R, 256]
  float* h32_out;           // [R, 256] or nullptr
  float* gates_out;         // [R, 1024] activated gates (packed column order) for training, or nullptr
  // "x-fold" (class decoder only): the input is grid_emb(one_hot(id)), i.e. tanh(b) everywhere except
  // the 3x3 cells around id, so its whole contribution to the pre-activations is a table look-up:
  // xf_B[border class of the cell][1024] (bias folded in) + xf_T2[border class of id][5x5 offset][1024]
  // for the <=25 cells around id.  The x chunk is then skipped in the K loop (skip_x).
  const float* xf_B;        // [9][1024] packed column order, or nullptr
  const float* xf_T2;       // [9][25][1024]
  const int* xf_ids;        // [NS] arg-max cell of every sample row
  int skip_x;               // x-fold: the x chunk of the K loop is skipped
  // dense raw x block of two channels, added in fp32 in the epilogue instead of going through the tensor cores
  // (regression encoder: the +-1.9e3 pixel offsets need all 24 bits, which neither operand format carries in 2 passes):
  // sparse x block (class encoder: scene features at ONE cell per sample row): per-sample table rows
  // xs_tab[sample][tap][1024] = features(label cell) . W[tap], added to the <= 9 cells around the label
  const float* xs_tab;      // [NS, 9, 1024] fp32 packed column order, or nullptr
  const int* xs_label;      // [NS]
  const float* xr_in;       // [NS, H, W, 2] fp32 NHWC (no halo), or nullptr
  const float* xr_W;        // [9 taps * 2 channels][1024] fp32, packed column order
  int order;                // work order, see work_index()
  int abl;                  // debug ablations (MVB_CELL_ABL; results are then WRONG): 1 skip the fp8 MMAs, 2 skip the
                            // 16-bit MMAs, 4 skip the epilogue's math and stores, 8 the issuer does not wait for operand
                            // data, 16 the producer loads nothing (use with 8)
  float* preact_out;        // [R, 1024] raw accumulators (packed column order) instead of the state update: first stage of
                            // the fan-out step (fanout_children_kernel turns every parent row into its K children)
  int hp_mixed;             // hp_out is written in the f16f8 format (else P bf16 planes)
  __nv_bfloat16* hp_out;    // [P][R][cpad_out] plane base or nullptr
  long long hp_plane_stride;  // elements between planes of hp_out
  int cpad_out;             // row pitch of hp_out (elements)
  int ch_off_out;           // channel offset of the h block inside hp_out rows
  long long R;              // total halo rows
  int H, W;
  int cpad;                 // K channels per tap (multiple of 32)
  float forget_bias;
};

// The state update of one element from its four gate pre-activations, with the roundings spelled out so that the
// cell epilogue and the fan-out kernel produce the same bits:  c' = sigmoid(f + forget_bias) c + sigmoid(i) tanh(j);
// h' = tanh(c') sigmoid(o).
struct GateOut { float ai, aj, af, ao, c, h; };
__device__ __forceinline__ GateOut lstm_update(float xi, float xj, float xf, float xo, float cprev, float forget_bias) {
  GateOut r;
  r.ai = sigmoid_acc(xi); r.aj = tanh_acc(xj); r.af = sigmoid_acc(__fadd_rn(xf, forget_bias)); r.ao = sigmoid_acc(xo);
  r.c = __fmaf_rn(r.af, cprev, __fmul_rn(r.ai, r.aj));
  r.h = __fmul_rn(tanh_acc(r.c), r.ao);
  return r;
}
// pre-activation from the accumulator: acc * (column scale, f16f8 only) + (bias + x-fold table row)
template <int FMT>
__device__ __forceinline__ float preact(float acc, float scale, float q) {
  return FMT ? __fmaf_rn(acc, scale, q) : __fadd_rn(acc, q);
}

// UMMA instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=256.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((BLOCK_N >> 3) << 17) |
                            ((BLOCK_M >> 4) << 24);
// same with A=B=fp16 for kind::f16 (format code 0), which is also A=B=e4m3 for kind::f8f6f4 (format code 0)
constexpr uint32_t kIdescF16 = (1u << 4) | ((BLOCK_N >> 3) << 17) | ((BLOCK_M >> 4) << 24);

// MC = true: clusters of two CTAs work on two M tiles of the same N tile in lock step; each loads half of every B
// (weight) tile and TMA-multicasts it to both, so the L2 -> shared-memory traffic per CTA and stage drops from
// 48 KB to 32 KB (P = 2).  A stage may be refilled once BOTH CTAs' MMAs have consumed it (empty barrier count 2,
// commits multicast to the pair).
// FMT = 1: f16f8 operands (P must be 2: same stage bytes).  tmA / tmB then describe the fp16 regions (one
// "plane") and tmA8 / tmB8 the two fp8 planes; per 32-channel stage the issuer sends two kind::f16 MMAs (K = 16
// each) and two kind::f8f6f4 MMAs (K = 32 each) into the same accumulator: 4 dispatches instead of 6.
// CG = 2: the pair instead issues ONE cta_group::2 MMA of 256 rows per dispatch from CTA 0: every CTA keeps only ITS
// half (128 rows) of each weight slot - the tensor cores of the pair read both halves in place - so the same 128 KB
// ring holds 8 slots instead of 4 (twice the prefetch distance: the L2 -> shared-memory latency under load, not
// bandwidth, is what left the tensor pipe idle with 4) and no byte of B is written into two shared memories.
template <int P, int CG, int FMT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
cell_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmA8, const __grid_constant__ CUtensorMap tmB8,
                const CellParams prm) {
  static_assert(FMT == 0 || P == 2, "the f16f8 format occupies the bytes of two bf16 planes");
  using Cfg = CellCfg<P>;
  constexpr bool MC = CG == 1, CG2 = CG == 2, PAIR = CG != 0;
  constexpr int SLOT_BYTES = CG2 ? B_SLOT_BYTES / 2 : B_SLOT_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int ra8 = (BLOCK_M + 2 * (prm.W + 2) + 7) & ~7;     // rows of an A stage
  const int a_stage_bytes = Cfg::a_stage_bytes(ra8);
  const int b_slots = Cfg::b_slots(ra8) * (CG2 ? 2 : 1);
  uint8_t* smem_a = smem + b_slots * SLOT_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_a + Cfg::A_STAGES * a_stage_bytes);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* afull_bar = empty_bar + 8;
  uint64_t* aempty_bar = afull_bar + Cfg::A_STAGES;
  uint64_t* tfull_bar = aempty_bar + Cfg::A_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const Grid g = make_grid(prm.H, prm.W);
  // K chunks of 64 channels: the x chunk [0, 64) - of which only the cxp channels of the x block are multiplied -
  // then the four chunks of the h block [cxp + 64 j, +64).  fp8 rows (f16f8): [x: e0 (cxp) | e1 (cxp)] then per h
  // chunk [e0 (64) | e1 (64)], 2 * cpad bytes per row (mvb_common.cuh f8_off).
  const int cxp = prm.cpad - kHidden;
  const int q_begin = prm.skip_x ? 1 : 0;
  constexpr int NQ = 1 + kHidden / CHUNK;      // 5
  constexpr int NS = FMT ? 2 : P;              // B slots per (chunk, tap)
  const long long num_m_tiles = (prm.R + BLOCK_M - 1) / BLOCK_M;
  // work index w -> (m tile, n tile).  MC: the pair shares w; rank r takes m tile 2*(w / N_TILES) + r.
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const long long num_tiles = (PAIR ? (num_m_tiles + 1) / 2 : num_m_tiles) * N_TILES;
  const long long w_begin = PAIR ? (long long)(blockIdx.x >> 1) : (long long)blockIdx.x;
  const long long w_step = PAIR ? (long long)(gridDim.x >> 1) : (long long)gridDim.x;
  // iteration it of this CTA (pair) -> work index.  order 1 (default): the N tiles of an M tile run back to back on
  // the same CTA (pair), so its operand rows are re-read from L2 by the SM that fetched them; order 0: strided
  // (the N tiles of an M tile run concurrently on neighbouring CTAs).
  auto work_index = [&](long long it) -> long long {
    return prm.order ? (w_begin + (it / N_TILES) * w_step) * N_TILES + it % N_TILES : w_begin + it * w_step;
  };
  auto tile_m0 = [&](long long w) -> long long { return ((w / N_TILES) * (PAIR ? 2 : 1) + rank) * BLOCK_M; };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (FMT == 1) { prefetch_tmap(&tmA8); prefetch_tmap(&tmB8); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < b_slots; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], MC ? 2 : 1); }
    for (int s = 0; s < Cfg::A_STAGES; ++s) { mbar_init(&afull_bar[s], 1); mbar_init(&aempty_bar[s], 1); }
    // CG2: the issuer (CTA 0) waits for the epilogue warps of BOTH CTAs before it reuses an accumulator
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], NUM_EPI_WARPS * (CG2 ? 2 : 1)); }
    fence_barrier_init();
  }
  if (warp == 2) { if (CG2) tmem_alloc_2sm(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();     // the peer's barriers (and, CG2, its TMEM) exist before anything is sent to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int slot = 0, astage = 0; uint32_t phase = 0, aphase = 0;
    for (long long it = 0, t; (t = work_index(it)) < num_tiles; ++it) {
      const long long m0 = tile_m0(t);
      const int n0 = (int)(t % N_TILES) * BLOCK_N;
      // chunk-major K order: the x block - whose terms can be orders of magnitude larger than the h terms (raw
      // pixel offsets in the regression encoder) - is accumulated first, so the small h products are never added
      // onto a large transient partial sum.
      if (prm.abl & 16) continue;
      for (int q = q_begin; q < NQ; ++q) {
        const int c16 = q == 0 ? 0 : cxp + (q - 1) * CHUNK;              // 16-bit channel coordinate
        const int c8 = q == 0 ? 0 : 2 * cxp + (q - 1) * 2 * CHUNK;       // fp8 byte coordinate
        mbar_wait(&aempty_bar[astage], aphase ^ 1);
        uint8_t* sa = smem_a + astage * a_stage_bytes;
        if (CG2) {
          // both CTAs' bytes are reported to CTA 0's barrier: its MMAs read the A tiles of both
          if (rank == 0) mbar_expect_tx(&afull_bar[astage], 2 * a_stage_bytes);
          tma_load_3d_2sm(sa, &tmA, &afull_bar[astage], c16, (int)(m0 - g.Wp - 1), 0);
          if (FMT == 1) tma_load_3d_2sm(sa + ra8 * ROW_BYTES, &tmA8, &afull_bar[astage], c8, (int)(m0 - g.Wp - 1), 0);
        } else {
          mbar_expect_tx(&afull_bar[astage], a_stage_bytes);
          tma_load_3d(sa, &tmA, &afull_bar[astage], c16, (int)(m0 - g.Wp - 1), 0);
          if (FMT == 1) tma_load_3d(sa + ra8 * ROW_BYTES, &tmA8, &afull_bar[astage], c8, (int)(m0 - g.Wp - 1), 0);
        }
        if (++astage == Cfg::A_STAGES) { astage = 0; aphase ^= 1; }
        for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
          for (int sl = 0; sl < NS; ++sl) {
            mbar_wait(&empty_bar[slot], phase ^ 1);
            uint8_t* sb = smem + slot * SLOT_BYTES;
            if (!CG2 || rank == 0) mbar_expect_tx(&full_bar[slot], B_SLOT_BYTES);
            const bool f8 = FMT == 1 && sl == 1;
            const CUtensorMap* tm = f8 ? &tmB8 : &tmB;
            const int kcol = f8 ? tap * 2 * prm.cpad + c8 : tap * prm.cpad + c16;
            const int plane = FMT == 1 ? 0 : sl;
            // MC: this CTA's half (128 rows) of the slot, delivered to both CTAs of the pair
            if (CG2) tma_load_3d_2sm(sb, tm, &full_bar[slot], kcol, n0 + (int)rank * (BLOCK_N / 2), plane);
            else if (MC) tma_load_3d_mc(sb + rank * (B_SLOT_BYTES / 2), tm, &full_bar[slot], kcol,
                                        n0 + (int)rank * (BLOCK_N / 2), plane, (uint16_t)3);
            else tma_load_3d(sb, tm, &full_bar[slot], kcol, n0, plane);
            if (++slot == b_slots) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && lane == 0 && (!CG2 || rank == 0)) {
    // ===================== MMA issuer =====================
    // One thread; its instruction stream is the critical path at 128 cycles per dispatch (measured with the loads and
    // the epilogue switched off: the round-2 first draft, with run-time K loops and 64-bit descriptor builds, reached
    // only 79 % of the dispatch floor).  So: compile-time trip counts for the h chunks, one IADD per operand and
    // dispatch (umma_lohi), the accumulate flag a compile-time constant except for the first dispatch of a tile.
    constexpr uint32_t kIdM = CG2 ? ((2u * BLOCK_M) >> 4) << 24 : 0u;     // cta_group::2: M = 256
    constexpr uint32_t kIdBf16 = CG2 ? ((kIdesc & 0x00FFFFFFu) | kIdM) : kIdesc;
    constexpr uint32_t kIdF16 = CG2 ? ((kIdescF16 & 0x00FFFFFFu) | kIdM) : kIdescF16;
    constexpr uint32_t kHi = smem_desc_hi(SW128_SBO, SW128_LAYOUT);
    const bool do16 = !(prm.abl & 2), do8 = !(prm.abl & 1), wait_data = !(prm.abl & 8);
    const uint32_t a_plane_lo = (uint32_t)(ra8 * ROW_BYTES) >> 4;
    int slot = 0, astage = 0; uint32_t phase = 0, aphase = 0;
    long long it = 0;
    for (long long t; (t = work_index(it)) < num_tiles; ++it) {
      const int as = (int)(it & 1);
      const uint32_t tphase = (uint32_t)((it >> 1) & 1);
      mbar_wait(&tempty_bar[as], tphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BLOCK_N;
      bool fresh = true;                     // the tile's first dispatch overwrites the accumulator
      for (int q = q_begin; q < NQ; ++q) {
        if (wait_data) mbar_wait(&afull_bar[astage], aphase);
        const uint32_t sa_lo = smem_u32(smem_a + astage * a_stage_bytes) >> 4;
        for (int tap = 0; tap < 9; ++tap) {
          // the tap's A tile: the stage's rows starting (dy-1) Wp + (dx-1) + (Wp+1) = dy Wp + dx rows in
          const uint32_t a_lo = sa_lo + (uint32_t)((tap / 3) * g.Wp + (tap % 3)) * (ROW_BYTES >> 4);
#pragma unroll
          for (int sl = 0; sl < NS; ++sl) {
            if (wait_data) mbar_wait(&full_bar[slot], phase);
            tc_fence_after();
            const uint32_t b_lo = smem_u32(smem + slot * SLOT_BYTES) >> 4;
            if (q > 0) {
              // ---- h chunk: 64 channels = 4 K16 steps per 16-bit plane pair, 2 K32 steps per e4m3 plane ----
              if (FMT == 1 && sl == 0) {
                if (do16) {
                  if (fresh) umma_lohi<0, CG2, false>(d_tmem, a_lo, b_lo, kHi, kIdF16);
                  else umma_lohi<0, CG2, true>(d_tmem, a_lo, b_lo, kHi, kIdF16);
                  umma_lohi<0, CG2, true>(d_tmem, a_lo + 2, b_lo + 2, kHi, kIdF16);
                  umma_lohi<0, CG2, true>(d_tmem, a_lo + 4, b_lo + 4, kHi, kIdF16);
                  umma_lohi<0, CG2, true>(d_tmem, a_lo + 6, b_lo + 6, kHi, kIdF16);
                  fresh = false;
                }
              } else if (FMT == 1) {
                if (do8) {
                  const uint32_t a8 = a_lo + a_plane_lo;          // [e0 (64 B) | e1 (64 B)] per row
                  if (fresh) umma_lohi<1, CG2, false>(d_tmem, a8, b_lo, kHi, kIdF16);
                  else umma_lohi<1, CG2, true>(d_tmem, a8, b_lo, kHi, kIdF16);
                  umma_lohi<1, CG2, true>(d_tmem, a8 + 2, b_lo + 2, kHi, kIdF16);
                  umma_lohi<1, CG2, true>(d_tmem, a8 + 4, b_lo + 4, kHi, kIdF16);
                  umma_lohi<1, CG2, true>(d_tmem, a8 + 6, b_lo + 6, kHi, kIdF16);
                  fresh = false;
                }
              } else if (do16) {
                // B plane sl against the A planes pa with pa + sl < P
#pragma unroll
                for (int pa = 0; pa < P - sl; ++pa) {
                  const uint32_t ap = a_lo + pa * a_plane_lo;
                  if (fresh) umma_lohi<0, CG2, false>(d_tmem, ap, b_lo, kHi, kIdBf16);
                  else umma_lohi<0, CG2, true>(d_tmem, ap, b_lo, kHi, kIdBf16);
                  umma_lohi<0, CG2, true>(d_tmem, ap + 2, b_lo + 2, kHi, kIdBf16);
                  umma_lohi<0, CG2, true>(d_tmem, ap + 4, b_lo + 4, kHi, kIdBf16);
                  umma_lohi<0, CG2, true>(d_tmem, ap + 6, b_lo + 6, kHi, kIdBf16);
                  fresh = false;
                }
              }
            } else {
              // ---- x chunk (only cells whose input is not folded): cxp = 32 or 64 channels, run-time trip counts ----
              const int ks16 = cxp / UMMA_K, ks8 = cxp / 32;
              const uint32_t poff = (uint32_t)cxp >> 4;              // e1 sits cxp bytes after e0 in an fp8 row
              if (FMT == 1 && sl == 0) {
                for (int k = 0; k < ks16 && do16; ++k) {
                  if (fresh) umma_lohi<0, CG2, false>(d_tmem, a_lo + 2 * k, b_lo + 2 * k, kHi, kIdF16);
                  else umma_lohi<0, CG2, true>(d_tmem, a_lo + 2 * k, b_lo + 2 * k, kHi, kIdF16);
                  fresh = false;
                }
              } else if (FMT == 1) {
                for (int pk = 0; pk < 2 * ks8 && do8; ++pk) {
                  const uint32_t o = (pk / ks8) * poff + (pk % ks8) * 2;
                  if (fresh) umma_lohi<1, CG2, false>(d_tmem, a_lo + a_plane_lo + o, b_lo + o, kHi, kIdF16);
                  else umma_lohi<1, CG2, true>(d_tmem, a_lo + a_plane_lo + o, b_lo + o, kHi, kIdF16);
                  fresh = false;
                }
              } else {
                for (int pa = 0; pa < P - sl; ++pa)
                  for (int k = 0; k < ks16 && do16; ++k) {
                    if (fresh) umma_lohi<0, CG2, false>(d_tmem, a_lo + pa * a_plane_lo + 2 * k, b_lo + 2 * k, kHi, kIdBf16);
                    else umma_lohi<0, CG2, true>(d_tmem, a_lo + pa * a_plane_lo + 2 * k, b_lo + 2 * k, kHi, kIdBf16);
                    fresh = false;
                  }
              }
            }
            if (CG2) umma_commit_2sm(&empty_bar[slot]);
            else if (MC) umma_commit_mc(&empty_bar[slot], (uint16_t)3);
            else umma_commit(&empty_bar[slot]);
            if (++slot == b_slots) { slot = 0; phase ^= 1; }
          }
        }
        if (CG2) umma_commit_2sm(&aempty_bar[astage]);
        else umma_commit(&aempty_bar[astage]);      // this CTA's nine taps have consumed the A stage
        if (++astage == Cfg::A_STAGES) { astage = 0; aphase ^= 1; }
      }
      if (CG2) umma_commit_2sm(&tfull_bar[as]);
      else umma_commit(&tfull_bar[as]);
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int wq = warp & 3;                 // TMEM lane quarter this warp may touch
    const int cgp = (warp - 4) >> 2;         // column group (0/1): channels [32*cgp, +32)
    long long it = 0;
    for (long long t; (t = work_index(it)) < num_tiles; ++it) {
      const int as = (int)(it & 1);
      const uint32_t aphase = (uint32_t)((it >> 1) & 1);
      const long long m0 = tile_m0(t);
      const int nt = (int)(t % N_TILES);
      const long long row = m0 + wq * 32 + lane;
      bool valid = row < prm.R;
      long long src_row = row;
      const float* xfb = nullptr;      // x-fold: table row of this cell's border class (bias included)
      int py = 0, px = 0, prem = 0;    // cell position and row offset inside its sample row
      long long psmp = 0;
      if (valid) {
        const long long smp = row / g.S;
        const int rem = (int)(row - smp * g.S);
        const int y = rem / g.Wp, x = rem - y * g.Wp;
        valid = (x < g.W) && (y < g.H);
        if (valid && prm.row_map) src_row = (long long)prm.row_map[smp] * g.S + rem;
        if (valid && prm.xf_B) {
          const int cy = y == 0 ? 0 : (y == g.H - 1 ? 2 : 1), cx = x == 0 ? 0 : (x == g.W - 1 ? 2 : 1);
          xfb = prm.xf_B + (cy * 3 + cx) * kGates;
        }
        py = y; px = x; psmp = smp; prem = rem;
      }
      // x-fold table row of this cell for the sample row whose arg-max cell is `a` (nullptr outside its 5x5)
      auto xft_of = [&](int a) -> const float* {
        const int ay = a / g.W, ax = a - ay * g.W;
        const int ry = py - ay, rx = px - ax;
        if (ry < -2 || ry > 2 || rx < -2 || rx > 2) return nullptr;
        const int acy = ay == 0 ? 0 : (ay == g.H - 1 ? 2 : 1), acx = ax == 0 ? 0 : (ax == g.W - 1 ? 2 : 1);
        return prm.xf_T2 + ((long long)(acy * 3 + acx) * 25 + (ry + 2) * 5 + (rx + 2)) * kGates;
      };
      float xv[18];                    // dense x path: the 3x3 neighbourhood of this cell's two raw input channels
      if (prm.xr_W) {
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) {
          const int yy = py + tp / 3 - 1, xx = px + tp % 3 - 1;
          const bool ok = valid && yy >= 0 && yy < g.H && xx >= 0 && xx < g.W;
          const float2 v = ok ? __ldg(reinterpret_cast<const float2*>(prm.xr_in + ((psmp * g.H + yy) * g.W + xx) * 2))
                              : make_float2(0.f, 0.f);
          xv[2 * tp] = v.x; xv[2 * tp + 1] = v.y;
        }
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      if (prm.abl & 4) valid = false;
      const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + as * BLOCK_N;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int j0 = cgp * 32 + cc * 16;            // channel offset inside the tile
        const int ch0 = nt * TILE_CH + j0;            // hidden channel
        uint32_t gi[16], gj[16], gf[16], go[16];
        tmem_ld16(t_row + 0 * TILE_CH + j0, gi);
        tmem_ld16(t_row + 1 * TILE_CH + j0, gj);
        tmem_ld16(t_row + 2 * TILE_CH + j0, gf);
        tmem_ld16(t_row + 3 * TILE_CH + j0, go);
        float cprev[16];
        if (valid && prm.c_in) {
          const float4* cp = reinterpret_cast<const float4*>(prm.c_in + src_row * kHidden + ch0);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float4 q4 = __ldg(cp + v);
            cprev[4 * v] = q4.x; cprev[4 * v + 1] = q4.y; cprev[4 * v + 2] = q4.z; cprev[4 * v + 3] = q4.w;
          }
        } else {
#pragma unroll
          for (int v = 0; v < 16; ++v) cprev[v] = 0.f;
        }
        tmem_ld_wait();
        if (valid && prm.preact_out) {
          float* gp = prm.preact_out + row * kGates + nt * BLOCK_N + j0;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            reinterpret_cast<uint4*>(gp + 0 * TILE_CH)[v] = make_uint4(gi[4 * v], gi[4 * v + 1], gi[4 * v + 2], gi[4 * v + 3]);
            reinterpret_cast<uint4*>(gp + 1 * TILE_CH)[v] = make_uint4(gj[4 * v], gj[4 * v + 1], gj[4 * v + 2], gj[4 * v + 3]);
            reinterpret_cast<uint4*>(gp + 2 * TILE_CH)[v] = make_uint4(gf[4 * v], gf[4 * v + 1], gf[4 * v + 2], gf[4 * v + 3]);
            reinterpret_cast<uint4*>(gp + 3 * TILE_CH)[v] = make_uint4(go[4 * v], go[4 * v + 1], go[4 * v + 2], go[4 * v + 3]);
          }
        } else if (valid) {
          const float* bptr = (xfb ? xfb : prm.bias) + nt * BLOCK_N + j0;
          const float* xft = prm.xf_B ? xft_of(prm.xf_ids[psmp]) : nullptr;
          if (prm.xs_tab) {          // out[p] += in[p + off(tap)] . W[tap] with the input at the label cell only
            const int l = prm.xs_label[psmp];
            if (l >= 0 && l < g.H * g.W) {
              const int ly = l / g.W, dy = ly - py, dx = (l - ly * g.W) - px;
              if (dy >= -1 && dy <= 1 && dx >= -1 && dx <= 1)
                xft = prm.xs_tab + ((long long)psmp * 9 + (dy + 1) * 3 + (dx + 1)) * kGates;
            }
          }
          const float* tptr = xft ? xft + nt * BLOCK_N + j0 : nullptr;
          const float* sptr = FMT == 1 ? prm.col_scale + nt * BLOCK_N + j0 : bptr;
          float cn[16], hn[16];
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            // bias (or bias-folded table), the x-fold table row of this cell and the column scales: one 128-bit load
            // per gate and 4 columns
            const float4* b4 = reinterpret_cast<const float4*>(bptr) + v4;
            const float4* s4 = reinterpret_cast<const float4*>(sptr) + v4;
            float4 qi = __ldg(b4 + 0 * (TILE_CH / 4)), qj = __ldg(b4 + 1 * (TILE_CH / 4)),
                   qf = __ldg(b4 + 2 * (TILE_CH / 4)), qo = __ldg(b4 + 3 * (TILE_CH / 4));
            if (tptr) {
              const float4* t4 = reinterpret_cast<const float4*>(tptr) + v4;
              const float4 ti = __ldg(t4 + 0 * (TILE_CH / 4)), tj = __ldg(t4 + 1 * (TILE_CH / 4)),
                           tf = __ldg(t4 + 2 * (TILE_CH / 4)), to = __ldg(t4 + 3 * (TILE_CH / 4));
              qi.x = __fadd_rn(qi.x, ti.x); qi.y = __fadd_rn(qi.y, ti.y); qi.z = __fadd_rn(qi.z, ti.z); qi.w = __fadd_rn(qi.w, ti.w);
              qj.x = __fadd_rn(qj.x, tj.x); qj.y = __fadd_rn(qj.y, tj.y); qj.z = __fadd_rn(qj.z, tj.z); qj.w = __fadd_rn(qj.w, tj.w);
              qf.x = __fadd_rn(qf.x, tf.x); qf.y = __fadd_rn(qf.y, tf.y); qf.z = __fadd_rn(qf.z, tf.z); qf.w = __fadd_rn(qf.w, tf.w);
              qo.x = __fadd_rn(qo.x, to.x); qo.y = __fadd_rn(qo.y, to.y); qo.z = __fadd_rn(qo.z, to.z); qo.w = __fadd_rn(qo.w, to.w);
            }
            if (prm.xr_W) {           // + sum over taps and the two channels of x * W, fp32 (warp-uniform weight loads)
              const float* wb = prm.xr_W + nt * BLOCK_N + j0 + 4 * v4;
#pragma unroll
              for (int k = 0; k < 18; ++k) {
                const float4 wi = __ldg(reinterpret_cast<const float4*>(wb + k * kGates + 0 * TILE_CH)),
                             wj = __ldg(reinterpret_cast<const float4*>(wb + k * kGates + 1 * TILE_CH)),
                             wf = __ldg(reinterpret_cast<const float4*>(wb + k * kGates + 2 * TILE_CH)),
                             wo = __ldg(reinterpret_cast<const float4*>(wb + k * kGates + 3 * TILE_CH));
                const float xk = xv[k];
                qi.x = __fmaf_rn(xk, wi.x, qi.x); qi.y = __fmaf_rn(xk, wi.y, qi.y); qi.z = __fmaf_rn(xk, wi.z, qi.z); qi.w = __fmaf_rn(xk, wi.w, qi.w);
                qj.x = __fmaf_rn(xk, wj.x, qj.x); qj.y = __fmaf_rn(xk, wj.y, qj.y); qj.z = __fmaf_rn(xk, wj.z, qj.z); qj.w = __fmaf_rn(xk, wj.w, qj.w);
                qf.x = __fmaf_rn(xk, wf.x, qf.x); qf.y = __fmaf_rn(xk, wf.y, qf.y); qf.z = __fmaf_rn(xk, wf.z, qf.z); qf.w = __fmaf_rn(xk, wf.w, qf.w);
                qo.x = __fmaf_rn(xk, wo.x, qo.x); qo.y = __fmaf_rn(xk, wo.y, qo.y); qo.z = __fmaf_rn(xk, wo.z, qo.z); qo.w = __fmaf_rn(xk, wo.w, qo.w);
              }
            }
            float4 si = qi, sj = qj, sf = qf, so = qo;
            if (FMT == 1) {
              si = __ldg(s4 + 0 * (TILE_CH / 4)); sj = __ldg(s4 + 1 * (TILE_CH / 4));
              sf = __ldg(s4 + 2 * (TILE_CH / 4)); so = __ldg(s4 + 3 * (TILE_CH / 4));
            }
            const float bi[4] = {qi.x, qi.y, qi.z, qi.w}, bj[4] = {qj.x, qj.y, qj.z, qj.w},
                        bf[4] = {qf.x, qf.y, qf.z, qf.w}, bo[4] = {qo.x, qo.y, qo.z, qo.w};
            const float ci[4] = {si.x, si.y, si.z, si.w}, cj[4] = {sj.x, sj.y, sj.z, sj.w},
                        cf[4] = {sf.x, sf.y, sf.z, sf.w}, co4[4] = {so.x, so.y, so.z, so.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int v = 4 * v4 + e;
              const GateOut r = lstm_update(preact<FMT>(__uint_as_float(gi[v]), ci[e], bi[e]),
                                            preact<FMT>(__uint_as_float(gj[v]), cj[e], bj[e]),
                                            preact<FMT>(__uint_as_float(gf[v]), cf[e], bf[e]),
                                            preact<FMT>(__uint_as_float(go[v]), co4[e], bo[e]), cprev[v], prm.forget_bias);
              cn[v] = r.c;
              hn[v] = r.h;
              if (prm.gates_out) {   // reuse the accumulator registers as staging for the stores below
                gi[v] = __float_as_uint(r.ai); gj[v] = __float_as_uint(r.aj);
                gf[v] = __float_as_uint(r.af); go[v] = __float_as_uint(r.ao);
              }
            }
          }
          const long long orow = row;
          if (prm.gates_out) {
            float* gp = prm.gates_out + orow * kGates + nt * BLOCK_N + j0;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              reinterpret_cast<uint4*>(gp + 0 * TILE_CH)[v] = make_uint4(gi[4 * v], gi[4 * v + 1], gi[4 * v + 2], gi[4 * v + 3]);
              reinterpret_cast<uint4*>(gp + 1 * TILE_CH)[v] = make_uint4(gj[4 * v], gj[4 * v + 1], gj[4 * v + 2], gj[4 * v + 3]);
              reinterpret_cast<uint4*>(gp + 2 * TILE_CH)[v] = make_uint4(gf[4 * v], gf[4 * v + 1], gf[4 * v + 2], gf[4 * v + 3]);
              reinterpret_cast<uint4*>(gp + 3 * TILE_CH)[v] = make_uint4(go[4 * v], go[4 * v + 1], go[4 * v + 2], go[4 * v + 3]);
            }
          }
          float4* co = reinterpret_cast<float4*>(prm.c_out + orow * kHidden + ch0);
#pragma unroll
          for (int v = 0; v < 4; ++v) co[v] = make_float4(cn[4 * v], cn[4 * v + 1], cn[4 * v + 2], cn[4 * v + 3]);
          if (prm.h32_out) {
            float4* ho = reinterpret_cast<float4*>(prm.h32_out + orow * kHidden + ch0);
#pragma unroll
            for (int v = 0; v < 4; ++v) ho[v] = make_float4(hn[4 * v], hn[4 * v + 1], hn[4 * v + 2], hn[4 * v + 3]);
          }
          if (prm.hp_out && prm.hp_mixed) {
            const float (&h0)[8] = *reinterpret_cast<const float (*)[8]>(&hn[0]);
            const float (&h1)[8] = *reinterpret_cast<const float (*)[8]>(&hn[8]);
            store_f16f8_x8(prm.hp_out, prm.hp_plane_stride, orow, prm.ch_off_out + ch0, prm.cpad_out, h0);
            store_f16f8_x8(prm.hp_out, prm.hp_plane_stride, orow, prm.ch_off_out + ch0 + 8, prm.cpad_out, h1);
          } else if (prm.hp_out) {
            uint32_t pk[P][8];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              __nv_bfloat16 a[P], b[P];
              split_planes<P>(hn[2 * v], a);
              split_planes<P>(hn[2 * v + 1], b);
#pragma unroll
              for (int p = 0; p < P; ++p) pk[p][v] = pack_bf16x2(a[p], b[p]);
            }
#pragma unroll
            for (int p = 0; p < P; ++p) {
              uint4* po = reinterpret_cast<uint4*>(prm.hp_out + p * prm.hp_plane_stride +
                                                   orow * prm.cpad_out + prm.ch_off_out + ch0);
              po[0] = make_uint4(pk[p][0], pk[p][1], pk[p][2], pk[p][3]);
              po[1] = make_uint4(pk[p][4], pk[p][5], pk[p][6], pk[p][7]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (CG2) mbar_arrive_cta0(&tempty_bar[as]); else mbar_arrive(&tempty_bar[as]); }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();     // the peer may still send into this CTA's smem / barriers / TMEM
  if (warp == 2) { if (CG2) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

// ----------------------------------------------------------------------------------
// Fan-out step of the beam decoder (the first K-row step: every child's parent is its sample's single t0 row, so the
// K children share the graph-attended h, the GEMM and c, and differ only in the folded table rows of their selected
// cell).  Stage 1 = the cell kernel with preact_out (one GEMM per PARENT row, raw accumulators to HBM: 4 KB per cell);
// stage 2 = this kernel: one warp per parent cell keeps the 1024 accumulators, the bias-folded table row and c in
// registers and emits the K children (c', h') - HBM-bound on its 2 KB of stores per child cell, where the round-1
// in-epilogue fan-out ran the K passes on 8 warps per SM (25 ms against a 2.5 ms HBM bound).
// Bit-identical to the K-times tiled launch: same preact() / lstm_update() roundings.
// ----------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(256)
fanout_children_kernel(const float* __restrict__ acc, const float* __restrict__ col_scale,
                       const float* __restrict__ xf_B, const float* __restrict__ xf_T2, const int* __restrict__ ids,
                       const float* __restrict__ c_in, float* __restrict__ c_out, float* __restrict__ h32_out,
                       long long NS, int K, Grid g, float forget_bias) {
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int hw = g.H * g.W;
  if (wid >= NS * hw) return;
  const long long smp = wid / hw;
  const int cell = (int)(wid - smp * hw);
  const int y = cell / g.W, x = cell - y * g.W;
  const int rem = y * g.Wp + x;
  const long long prow = smp * g.S + rem;
  const int ch0 = lane * 8;                                  // this lane's 8 hidden channels
  const int col0 = (ch0 / TILE_CH) * BLOCK_N + ch0 % TILE_CH;   // packed column of gate 0 (gate g: + g * 64)
  const int cy = y == 0 ? 0 : (y == g.H - 1 ? 2 : 1), cx = x == 0 ? 0 : (x == g.W - 1 ? 2 : 1);
  float a[4][8], q0[4][8], sc[4][8], cp[8];
#pragma unroll
  for (int gt = 0; gt < 4; ++gt) {
    const float4* ap = reinterpret_cast<const float4*>(acc + prow * kGates + col0 + gt * TILE_CH);
    const float4* bp = reinterpret_cast<const float4*>(xf_B + (cy * 3 + cx) * kGates + col0 + gt * TILE_CH);
    const float4* sp = reinterpret_cast<const float4*>((FMT ? col_scale : xf_B) + col0 + gt * TILE_CH);
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const float4 av = __ldg(ap + v), bv = __ldg(bp + v), sv = __ldg(sp + v);
      a[gt][4 * v] = av.x; a[gt][4 * v + 1] = av.y; a[gt][4 * v + 2] = av.z; a[gt][4 * v + 3] = av.w;
      q0[gt][4 * v] = bv.x; q0[gt][4 * v + 1] = bv.y; q0[gt][4 * v + 2] = bv.z; q0[gt][4 * v + 3] = bv.w;
      sc[gt][4 * v] = sv.x; sc[gt][4 * v + 1] = sv.y; sc[gt][4 * v + 2] = sv.z; sc[gt][4 * v + 3] = sv.w;
    }
  }
  {
    const float4* cq = reinterpret_cast<const float4*>(c_in + prow * kHidden + ch0);
    const float4 c0 = __ldg(cq), c1 = __ldg(cq + 1);
    cp[0] = c0.x; cp[1] = c0.y; cp[2] = c0.z; cp[3] = c0.w; cp[4] = c1.x; cp[5] = c1.y; cp[6] = c1.z; cp[7] = c1.w;
  }
  for (int k = 0; k < K; ++k) {
    const long long osmp = smp * K + k;
    const int id = ids[osmp];
    const int ay = id / g.W, ax = id - ay * g.W;
    const int ry = y - ay, rx = x - ax;
    const float* tp = nullptr;      // x-fold table row of this cell for the child's selected cell (inside its 5x5)
    if (ry >= -2 && ry <= 2 && rx >= -2 && rx <= 2) {
      const int acy = ay == 0 ? 0 : (ay == g.H - 1 ? 2 : 1), acx = ax == 0 ? 0 : (ax == g.W - 1 ? 2 : 1);
      tp = xf_T2 + ((long long)(acy * 3 + acx) * 25 + (ry + 2) * 5 + (rx + 2)) * kGates + col0;
    }
    float cn[8], hn[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      float q[4];
#pragma unroll
      for (int gt = 0; gt < 4; ++gt) q[gt] = tp ? __fadd_rn(q0[gt][v], __ldg(tp + gt * TILE_CH + v)) : q0[gt][v];
      const GateOut r = lstm_update(preact<FMT>(a[0][v], sc[0][v], q[0]), preact<FMT>(a[1][v], sc[1][v], q[1]),
                                    preact<FMT>(a[2][v], sc[2][v], q[2]), preact<FMT>(a[3][v], sc[3][v], q[3]),
                                    cp[v], forget_bias);
      cn[v] = r.c; hn[v] = r.h;
    }
    const long long orow = osmp * g.S + rem;
    float4* co = reinterpret_cast<float4*>(c_out + orow * kHidden + ch0);
    co[0] = make_float4(cn[0], cn[1], cn[2], cn[3]); co[1] = make_float4(cn[4], cn[5], cn[6], cn[7]);
    float4* ho = reinterpret_cast<float4*>(h32_out + orow * kHidden + ch0);
    ho[0] = make_float4(hn[0], hn[1], hn[2], hn[3]); ho[1] = make_float4(hn[4], hn[5], hn[6], hn[7]);
  }
}

// ----------------------------------------------------------------------------------
// weight packing:  TF kernel [3,3,Cx+256,1024] (HWIO, gate order i,j,f,o) + biases
//   -> planes bf16 [P][1024][9*cpad]  (row = tile*256 + gate*64 + j,  k = tap*cpad + kc)
//   -> bias fp32 [1024] in the same row order
// kc < cx: input channel kc;  cx <= kc < cxp: zero;  kc >= cxp: hidden channel kc-cxp.
// ----------------------------------------------------------------------------------
template <int P>
__global__ void pack_weights_kernel(const float* __restrict__ kernel, const float* __restrict__ biases,
                                    __nv_bfloat16* __restrict__ wp, float* __restrict__ bias_packed,
                                    int cx, int cxp, int cpad, int comp) {
  const long long ktot = 9LL * cpad;
  const long long total = (long long)kGates * ktot;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / ktot);
    const int k = (int)(i - (long long)n * ktot);
    const int tap = k / cpad, kcn = k - tap * cpad;
    const int tile = n / BLOCK_N, gate = (n % BLOCK_N) / TILE_CH, j = n % TILE_CH;
    const int col = gate * kHidden + tile * TILE_CH + j;
    int cin = -1, blk = 0;
    if (kcn >= cxp) cin = cx + (kcn - cxp);
    else if (!comp) { if (kcn < cx) cin = kcn; }
    else if (kcn < 4 * cx) { blk = kcn / cx; cin = kcn - blk * cx; }
    const float v = (cin >= 0) ? kernel[((long long)tap * (cx + kHidden) + cin) * kGates + col] : 0.f;
    __nv_bfloat16 pl[P];
    split_planes<P>(v, pl);
    if (comp && kcn < cxp && P == 2) {
      // compensated x block (see nhwc_to_planes_comp): [W | W | W-w0-w1 | w1]
      if (blk == 2) {
        float r = v;
#pragma unroll
        for (int p = 0; p < P; ++p) r -= __bfloat162float(pl[p]);
        split_planes<P>(r, pl);
      } else if (blk == 3) {
        pl[0] = pl[P - 1];
        pl[P - 1] = __float2bfloat16_rn(0.f);
      }
    }
#pragma unroll
    for (int p = 0; p < P; ++p) wp[(long long)p * total + i] = pl[p];
    if (k == 0) bias_packed[n] = biases[col];
  }
}

// x-fold tables of a class-decoder cell (packed column order n = tile*256 + gate*64 + j):
//   X0[e] = tanh(be[e]);  delta[k][e] = tanh(be[e] + We[8-k][e]) - X0[e]   (k = 3x3 position around the arg-max;
//   the one-hot at a reaches cell q = a + k through tap a - q, i.e. tap index 8 - k)
//   B[cls][n]        = bias[n] + sum_{tap valid at a cell of border class cls} X0 . W[tap][:E][n]
//   T2[acls][r][n]   = sum_{tap: q = p + off(tap) in 3x3(a), q inside the grid} delta[k(q)] . W[tap][:E][n],  r = p - a
__global__ void xfold_tables_kernel(const float* __restrict__ kernel, const float* __restrict__ biases,
                                    const float* __restrict__ We, const float* __restrict__ be, int E,
                                    float* __restrict__ Bt, float* __restrict__ T2) {
  const int total = (9 + 9 * 25) * kGates;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i % kGates, item = i / kGates;
    const int tile = n / BLOCK_N, gate = (n % BLOCK_N) / TILE_CH, j = n % TILE_CH;
    const int col = gate * kHidden + tile * TILE_CH + j;
    const int cin_tot = E + kHidden;
    float acc = 0.f;
    if (item < 9) {
      const int cy = item / 3, cx = item % 3;
      acc = biases[col];
      for (int t = 0; t < 9; ++t) {
        const int oy = t / 3 - 1, ox = t % 3 - 1;
        if ((cy == 0 && oy < 0) || (cy == 2 && oy > 0) || (cx == 0 && ox < 0) || (cx == 2 && ox > 0)) continue;
        for (int e = 0; e < E; ++e) acc = fmaf(tanhf(be[e]), kernel[((long long)t * cin_tot + e) * kGates + col], acc);
      }
      Bt[item * kGates + n] = acc;
    } else {
      const int it2 = item - 9, acls = it2 / 25, rr = it2 % 25;
      const int acy = acls / 3, acx = acls % 3;
      const int ry = rr / 5 - 2, rx = rr % 5 - 2;
      for (int t = 0; t < 9; ++t) {
        const int ky = ry + t / 3 - 1, kx = rx + t % 3 - 1;      // q - a
        if (ky < -1 || ky > 1 || kx < -1 || kx > 1) continue;
        if ((acy == 0 && ky < 0) || (acy == 2 && ky > 0) || (acx == 0 && kx < 0) || (acx == 2 && kx > 0)) continue;
        const int k = (ky + 1) * 3 + (kx + 1);
        for (int e = 0; e < E; ++e) {
          const float d = tanhf(be[e] + We[(8 - k) * E + e]) - tanhf(be[e]);
          acc = fmaf(d, kernel[((long long)t * cin_tot + e) * kGates + col], acc);
        }
      }
      T2[(long long)it2 * kGates + n] = acc;
    }
  }
}

// weights of the dense x path: rows (tap, channel) of the TF kernel [3,3,2+256,1024] in the packed column order
__global__ void xdense_weights_kernel(const float* __restrict__ kernel, float* __restrict__ out) {
  const int total = 18 * kGates;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i % kGates, k = i / kGates, tap = k >> 1, ch = k & 1;
    const int tile = n / BLOCK_N, gate = (n % BLOCK_N) / TILE_CH, j = n % TILE_CH;
    const int col = gate * kHidden + tile * TILE_CH + j;
    out[i] = kernel[((long long)tap * (2 + kHidden) + ch) * kGates + col];
  }
}
int cell_xdense_weights(const float* kernel, float* out, cudaStream_t stream) {
  MVB_REQUIRE(kernel && out, "cell_xdense_weights: bad args");
  xdense_weights_kernel<<<72, 256, 0, stream>>>(kernel, out);
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

// weights of the sparse x path: rows (tap, channel) of the x block of the TF kernel [3,3,cx+256,1024], packed columns
__global__ void xsparse_weights_kernel(const float* __restrict__ kernel, int cx, float* __restrict__ out) {
  const long long total = 9ll * cx * kGates;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % kGates);
    const long long k = i / kGates;
    const int tap = (int)(k / cx), ch = (int)(k % cx);
    const int tile = n / BLOCK_N, gate = (n % BLOCK_N) / TILE_CH, j = n % TILE_CH;
    const int col = gate * kHidden + tile * TILE_CH + j;
    out[i] = kernel[((long long)tap * (cx + kHidden) + ch) * kGates + col];
  }
}
int cell_xsparse_weights(const float* kernel, int cx, float* out, cudaStream_t stream) {
  MVB_REQUIRE(kernel && out && cx > 0, "cell_xsparse_weights: bad args");
  xsparse_weights_kernel<<<sm_count() * 4, 256, 0, stream>>>(kernel, cx, out);
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

// xs_tab[s][tap][n] = sum_ch feat[frame[s]][label[s]][ch] * Wx[tap][ch][n]   (64 scene channels).
// One block per (tap, 256-column tile, sample chunk): the weight tile [64][256] sits in shared memory, thread = column.
__global__ void __launch_bounds__(256)
xsparse_table_kernel(const float* __restrict__ scene_conv, const int* __restrict__ frame_idx,
                     const int* __restrict__ label, const float* __restrict__ Wx, float* __restrict__ tab,
                     long long NS, int hw, int chunks) {
  extern __shared__ float wsm[];                      // [64][256]
  const int tap = blockIdx.x / N_TILES, nt = blockIdx.x % N_TILES;
  for (int i = threadIdx.x; i < 64 * 256; i += 256)
    wsm[i] = Wx[((long long)tap * 64 + i / 256) * kGates + nt * 256 + (i % 256)];
  __syncthreads();
  const int col = threadIdx.x;
  for (long long s = blockIdx.y; s < NS; s += chunks) {
    const int l = label[s];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (l >= 0 && l < hw) {
      const float4* f4 = reinterpret_cast<const float4*>(scene_conv + ((long long)frame_idx[s] * hw + l) * 64);
#pragma unroll 4
      for (int c4 = 0; c4 < 16; ++c4) {
        const float4 f = __ldg(f4 + c4);              // warp-uniform
        acc[0] = fmaf(f.x, wsm[(4 * c4 + 0) * 256 + col], acc[0]);
        acc[1] = fmaf(f.y, wsm[(4 * c4 + 1) * 256 + col], acc[1]);
        acc[2] = fmaf(f.z, wsm[(4 * c4 + 2) * 256 + col], acc[2]);
        acc[3] = fmaf(f.w, wsm[(4 * c4 + 3) * 256 + col], acc[3]);
      }
    }
    tab[(s * 9 + tap) * kGates + nt * 256 + col] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  }
}
int cell_xsparse_table(const float* scene_conv, const int* frame_idx, const int* label, const float* Wx, float* tab,
                       long long NS, int H, int W, cudaStream_t stream) {
  MVB_REQUIRE(scene_conv && frame_idx && label && Wx && tab && NS > 0, "cell_xsparse_table: bad args");
  static SmemOptIn opt;
  MVB_CHECK_CUDA(smem_opt_in(opt, xsparse_table_kernel, 64 * 256 * (int)sizeof(float)));
  const int chunks = (int)(NS < 8 ? NS : 8);
  xsparse_table_kernel<<<dim3(9 * N_TILES, chunks), 256, 64 * 256 * sizeof(float), stream>>>(
      scene_conv, frame_idx, label, Wx, tab, NS, H * W, chunks);
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

int cell_xfold_tables(const float* kernel, const float* biases, const float* We, const float* be, int E,
                      float* Bt, float* T2, cudaStream_t stream) {
  MVB_REQUIRE(kernel && biases && We && be && Bt && T2 && E > 0, "cell_xfold_tables: bad args");
  xfold_tables_kernel<<<234, 256, 0, stream>>>(kernel, biases, We, be, E, Bt, T2);
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

// variant of the last launch_cell() of this process: planes code * 2 + multicast (tests assert which kernel ran)
static int g_last_variant = -1;
static unsigned long long g_variants_seen = 0;      // bit (format index * 2 + pair): formats 1, 2, 3 planes, f16f8
int cell_last_variant() { return g_last_variant; }
unsigned long long cell_variants_seen(int reset) {
  const unsigned long long v = g_variants_seen;
  if (reset) g_variants_seen = 0;
  return v;
}
static void note_variant(int planes, int pair) {
  g_last_variant = planes * 2 + pair;
  const int fi = planes == kPlanesF16F8 ? 3 : planes - 1;
  g_variants_seen |= 1ull << (fi * 2 + pair);
}

struct CellMaps { CUtensorMap A, B, Bh, A8, B8, B8h; };

// Work order of a launch (CellParams::order, work_index()).  1: the four N tiles of an M tile (pair) run back to back
// on the same CTA (pair) - measured on the K=20 beam step: DRAM reads 1.05x algorithmic instead of 1.39x, 3 % faster.
// 0: work items strided over the CTAs.  A launch of few M tiles (the encoders and the greedy decoders of a small
// shard: 32 trajectories of 36x18 = 176 M tiles on 148 CTAs) is bound by its longest CTA instead: back to back the
// busiest CTA runs 2 x 4 items, strided ceil(704 / 148) = 5.  Strided whenever that makespan is shorter and the
// launch is small enough for its operands to stay in L2.
static int pick_order(int forced, long long units, long long ctas) {
  if (forced == 0 || forced == 1) return forced;
  const long long back_to_back = ((units + ctas - 1) / ctas) * N_TILES, strided = (units * N_TILES + ctas - 1) / ctas;
  return (strided < back_to_back && units < 4 * ctas) ? 0 : 1;
}

template <int P, int FMT>
static int launch_cell(const CellMaps& tm, const CellParams& prm_in, int num_sms, bool multicast, cudaStream_t stream) {
  using Cfg = CellCfg<P>;
  CellParams prm = prm_in;
  static SmemOptIn opt_plain, opt_mc;
  const int ra8 = (BLOCK_M + 2 * (prm.W + 2) + 7) & ~7;
  const int smem_bytes = Cfg::smem_bytes(ra8);
  MVB_REQUIRE(ra8 <= Cfg::MAX_RA8 && smem_bytes <= 227 * 1024, "cell_fwd: grid width W=%d too large (A stage of %d rows, %d B shared memory)", prm.W, ra8, smem_bytes);
  static SmemOptIn opt_cg2;
  MVB_CHECK_CUDA(smem_opt_in(opt_plain, cell_fwd_kernel<P, 0, FMT>, smem_bytes));
  MVB_CHECK_CUDA(smem_opt_in(opt_mc, cell_fwd_kernel<P, 1, FMT>, smem_bytes));
  MVB_CHECK_CUDA(smem_opt_in(opt_cg2, cell_fwd_kernel<P, 2, FMT>, smem_bytes));
  // pair mode: 2 = cta_group::2 (default), 1 = two cta_group::1 CTAs with weight multicast (MVB_CELL_PAIR=1)
  static const int pair_mode = [] { const char* e = getenv("MVB_CELL_PAIR"); return e ? atoi(e) : 2; }();
  const long long m_tiles = (prm.R + BLOCK_M - 1) / BLOCK_M;
  if (multicast && m_tiles >= 2 * (long long)num_sms) {
    prm.order = pick_order(prm_in.order, (m_tiles + 1) / 2, num_sms / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(num_sms / 2 * 2)); cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (pair_mode == 2) MVB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, cell_fwd_kernel<P, 2, FMT>, tm.A, tm.Bh, tm.A8, tm.B8h, prm));
    else MVB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, cell_fwd_kernel<P, 1, FMT>, tm.A, tm.Bh, tm.A8, tm.B8h, prm));
    count_launch(1);
    note_variant(FMT ? kPlanesF16F8 : P, 1);
    return MVB_OK;
  }
  const long long num_tiles = m_tiles * N_TILES;
  const int grid = (int)(num_tiles < num_sms ? num_tiles : num_sms);
  prm.order = pick_order(prm_in.order, m_tiles, grid);
  cell_fwd_kernel<P, 0, FMT><<<grid, NUM_THREADS, smem_bytes, stream>>>(tm.A, tm.B, tm.A8, tm.B8, prm);
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  note_variant(FMT ? kPlanesF16F8 : P, 0);
  return MVB_OK;
}

int cell_fwd(const void* xh_planes, const void* w_planes, const float* bias, const float* c_in,
             const int* row_map, float* c_out, float* h32_out, void* hp_out, long long hp_plane_stride,
             int cpad_out, int ch_off_out, long long NS, int H, int W, int cpad, int P,
             float forget_bias, float* gates_out, const float* xf_B, const float* xf_T2, const int* xf_ids,
             int fanout, cudaStream_t stream, const float* xr_in, const float* xr_W, const float* xs_tab,
             const int* xs_label) {
  // planes = format of the inputs and weights | (format of hp_out << 8), the latter only when it differs
  const int P_out = (P >> 8) ? (P >> 8) : (P & 0xFF);
  P &= 0xFF;
  const bool mixed = P == kPlanesF16F8;
  MVB_REQUIRE(P_out == P || P_out == kPlanesF16F8 || (mixed && P_out == 2), "cell_fwd: output planes %d with input planes %d", P_out, P);
  MVB_REQUIRE((P >= 1 && P <= 3) || mixed, "cell_fwd: planes P=%d not in {1,2,3,%d}", P, kPlanesF16F8);
  MVB_REQUIRE(!mixed || !gates_out || fanout > 1, "cell_fwd: the f16f8 format is an inference format (no gates_out)");
  MVB_REQUIRE(fanout <= 1 || (xf_B && xf_T2 && xf_ids && gates_out && c_in && h32_out && !row_map && !hp_out),
              "cell_fwd: fanout=%d needs the x-fold tables, c_in, h32_out, a [R,1024] fp32 workspace and no row_map / hp_out", fanout);
  MVB_REQUIRE(cpad == kHidden + XPAD || cpad == kHidden + 2 * XPAD, "cell_fwd: cpad=%d must be 288 or 320 (x block of 32 or 64 channels)", cpad);
  MVB_REQUIRE(NS > 0 && H > 0 && W > 0, "cell_fwd: bad sizes NS=%lld H=%d W=%d", NS, H, W);
  MVB_REQUIRE(xh_planes && w_planes && bias && c_out, "cell_fwd: null pointer");
  if (hp_out) MVB_REQUIRE(cpad_out % 8 == 0 && ch_off_out % 8 == 0, "cell_fwd: hp_out pitch/offset must be multiples of 8");
  const Grid g = make_grid(H, W);
  const long long R = NS * g.S;
  MVB_REQUIRE(R + 2LL * g.Wp + 256 < 0x7fffffffLL, "cell_fwd: too many rows (%lld) for int32 TMA coordinates", R);

  // weight-tile multicast across CTA pairs is on by default (MVB_CELL_MULTICAST=0 turns it off for A/B runs)
  static const bool multicast = [] { const char* e = getenv("MVB_CELL_MULTICAST"); return !(e && e[0] == '0'); }();
  CellMaps tm;
  const int P16 = mixed ? 1 : P;      // 16-bit "planes" the A / B maps describe
  const uint32_t ra8 = (uint32_t)((BLOCK_M + 2 * (W + 2) + 7) & ~7);      // rows of an A stage (see CellCfg)
  MVB_REQUIRE(ra8 <= 256, "cell_fwd: grid width W=%d too large for the halo'd A stage (%u rows > 256)", W, ra8);
  int rc = encode_tmap_3d_bf16(&tm.A, xh_planes, (uint64_t)cpad, (uint64_t)R, (uint64_t)P16,
                               (uint64_t)cpad * 2, (uint64_t)R * cpad * 2, CHUNK, ra8, P16, 128);
  if (rc) return rc;
  const uint64_t ktot = 9ull * cpad;
  rc = encode_tmap_3d_bf16(&tm.B, w_planes, ktot, (uint64_t)kGates, (uint64_t)P16, ktot * 2,
                           ktot * kGates * 2, CHUNK, BLOCK_N, 1, 128);          // one plane of a tile = one slot
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&tm.Bh, w_planes, ktot, (uint64_t)kGates, (uint64_t)P16, ktot * 2,
                           ktot * kGates * 2, CHUNK, BLOCK_N / 2, 1, 128);      // half of it (CTA pairs)
  if (rc) return rc;
  tm.A8 = tm.A; tm.B8 = tm.B; tm.B8h = tm.Bh;
  const float* col_scale = nullptr;
  if (mixed) {
    // [fp16 region][fp8 region: rows of 2*cpad bytes, both e4m3 planes interleaved per chunk (f8_off)]
    // ([+ 1024 fp32 column scales] after the weights)
    const uint8_t* a8 = reinterpret_cast<const uint8_t*>(xh_planes) + 2ull * R * cpad;
    const uint8_t* b8 = reinterpret_cast<const uint8_t*>(w_planes) + 2ull * kGates * ktot;
    rc = encode_tmap_3d_u8(&tm.A8, a8, 2ull * cpad, (uint64_t)R, 1, 2ull * cpad, 2ull * R * cpad, ROW_BYTES, ra8, 1, 128);
    if (rc) return rc;
    rc = encode_tmap_3d_u8(&tm.B8, b8, 2 * ktot, (uint64_t)kGates, 1, 2 * ktot, 2 * ktot * kGates, ROW_BYTES, BLOCK_N, 1, 128);
    if (rc) return rc;
    rc = encode_tmap_3d_u8(&tm.B8h, b8, 2 * ktot, (uint64_t)kGates, 1, 2 * ktot, 2 * ktot * kGates, ROW_BYTES, BLOCK_N / 2, 1, 128);
    if (rc) return rc;
    col_scale = reinterpret_cast<const float*>(b8 + 2ull * kGates * ktot);
  }

  CellParams prm;
  prm.bias = bias; prm.col_scale = col_scale; prm.c_in = c_in; prm.row_map = row_map; prm.c_out = c_out; prm.h32_out = h32_out;
  prm.gates_out = fanout > 1 ? nullptr : gates_out;
  prm.preact_out = fanout > 1 ? gates_out : nullptr;      // fan-out: stage 1 stores the raw accumulators there
  prm.xf_B = xf_B; prm.xf_T2 = xf_T2; prm.xf_ids = xf_ids;
  prm.skip_x = 0;
  // work order: chosen per launch in launch_cell (MVB_CELL_ORDER=0|1 forces one)
  static const int order = [] { const char* e = getenv("MVB_CELL_ORDER"); return e ? atoi(e) : -1; }();
  prm.order = order;
  static const int abl = [] { const char* e = getenv("MVB_CELL_ABL"); return e ? atoi(e) : 0; }();
  prm.abl = (abl & 16) ? (abl | 8) : abl;      // "load nothing" without "do not wait for data" would hang the issuer
  if (xf_B) {
    MVB_REQUIRE(xf_T2 && xf_ids && H >= 3 && W >= 3, "cell_fwd: x-fold needs its tables, ids and a grid of at least 3x3");
    prm.skip_x = 1;
  }
  prm.xs_tab = xs_tab; prm.xs_label = xs_label;
  if (xs_tab) {
    MVB_REQUIRE(xs_label && !xf_B && !xr_W && !row_map && fanout <= 1, "cell_fwd: the sparse x path needs its labels and excludes x-fold, the dense x path, row maps and fan-out");
    prm.skip_x = 1;
  }
  prm.xr_in = xr_in; prm.xr_W = xr_W;
  if (xr_W) {
    MVB_REQUIRE(xr_in && !xf_B && !row_map && fanout <= 1, "cell_fwd: the dense x path needs its input and excludes x-fold, row maps and fan-out");
    prm.skip_x = 1;
  }
  prm.hp_mixed = P_out == kPlanesF16F8;
  prm.hp_out = reinterpret_cast<__nv_bfloat16*>(hp_out);
  prm.hp_plane_stride = hp_plane_stride; prm.cpad_out = cpad_out; prm.ch_off_out = ch_off_out;
  prm.R = R; prm.H = H; prm.W = W; prm.cpad = cpad; prm.forget_bias = forget_bias;

  int dev = 0, num_sms = 0;
  MVB_CHECK_CUDA(cudaGetDevice(&dev));
  MVB_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  switch (P) {
    case 1: rc = launch_cell<1, 0>(tm, prm, num_sms, multicast, stream); break;
    case 2: rc = launch_cell<2, 0>(tm, prm, num_sms, multicast, stream); break;
    case kPlanesF16F8: rc = launch_cell<2, 1>(tm, prm, num_sms, multicast, stream); break;
    default: rc = launch_cell<3, 0>(tm, prm, num_sms, multicast, stream); break;
  }
  if (rc || fanout <= 1) return rc;
  // fan-out stage 2: every parent row -> its K children (c_out / h32_out hold NS * fanout sample rows)
  const long long warps = NS * H * W;
  const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
  if (mixed) fanout_children_kernel<1><<<blocks, 256, 0, stream>>>(gates_out, col_scale, xf_B, xf_T2, xf_ids, c_in, c_out,
                                                                    h32_out, NS, fanout, g, forget_bias);
  else fanout_children_kernel<0><<<blocks, 256, 0, stream>>>(gates_out, col_scale, xf_B, xf_T2, xf_ids, c_in, c_out,
                                                             h32_out, NS, fanout, g, forget_bias);
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

// ----------------------------------------------------------------------------------
// f16f8 weight packing (see mvb_common.cuh): per packed column n a power-of-two scale 2^S with
// max|w| * 2^S in [2^13, 2^14), so that b0 = fp16(w 2^S) is a normal number for every weight down to 2^-27 of the
// column maximum, the residual b1 = w 2^S - b0 (|b1| <= 4) and b0 2^-12 (<= 4) sit in e4m3's normal range.
//   w16 [1024][9*cpad] fp16 = b0;  w8 [1024][9][2*cpad bytes] = e4m3(b1) and e4m3(b0 * 2^-12), interleaved per chunk
//   like the activation rows (f8_off);  col_scale [1024] = 2^-S
// ----------------------------------------------------------------------------------
__device__ __forceinline__ float packed_weight(const float* __restrict__ kernel, int n, int k, int cx, int cxp,
                                               int cpad) {
  const int tap = k / cpad, kcn = k - tap * cpad;
  const int tile = n / BLOCK_N, gate = (n % BLOCK_N) / TILE_CH, j = n % TILE_CH;
  const int col = gate * kHidden + tile * TILE_CH + j;
  int cin = -1;
  if (kcn >= cxp) cin = cx + (kcn - cxp);
  else if (kcn < cx) cin = kcn;
  return (cin >= 0) ? kernel[((long long)tap * (cx + kHidden) + cin) * kGates + col] : 0.f;
}

__global__ void colscale_kernel(const float* __restrict__ kernel, float* __restrict__ col_scale, int cx) {
  // one warp per packed column: max |w| over the 9 * (cx + 256) weights that feed it
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= kGates) return;
  const int tile = n / BLOCK_N, gate = (n % BLOCK_N) / TILE_CH, j = n % TILE_CH;
  const int col = gate * kHidden + tile * TILE_CH + j;
  float m = 0.f;
  for (int r = lane; r < 9 * (cx + kHidden); r += 32) m = fmaxf(m, fabsf(kernel[(long long)r * kGates + col]));
  m = warp_max(m);
  if (lane == 0) {
    int e = 0;
    if (m > 0.f) frexpf(m, &e);            // m = f * 2^e, f in [0.5, 1)  ->  floor(log2 m) = e - 1
    const int S = m > 0.f ? 13 - (e - 1) : 0;
    col_scale[n] = ldexpf(1.0f, -S);
  }
}

__global__ void pack_weights_f16f8_kernel(const float* __restrict__ kernel, const float* __restrict__ biases,
                                          __half* __restrict__ w16, uint8_t* __restrict__ w8,
                                          const float* __restrict__ col_scale, float* __restrict__ bias_packed,
                                          int cx, int cxp, int cpad) {
  const long long ktot = 9LL * cpad;
  const long long total = (long long)kGates * ktot;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / ktot);
    const int k = (int)(i - (long long)n * ktot);
    const float v = packed_weight(kernel, n, k, cx, cxp, cpad) / col_scale[n];     // exact: power of two
    const __half b0 = __float2half_rn(v);
    const float f0 = __half2float(b0);
    w16[i] = b0;
    const int tap = k / cpad, kcn = k - tap * cpad;
    uint8_t* w8row = w8 + ((long long)n * 9 + tap) * 2 * cpad;
    w8row[f8_off(kcn, 0, cpad)] = to_e4m3(v - f0);
    w8row[f8_off(kcn, 1, cpad)] = to_e4m3(f0 * (1.0f / kF8ResidualScale));
    if (k == 0) {
      const int tile = n / BLOCK_N, gate = (n % BLOCK_N) / TILE_CH, j = n % TILE_CH;
      bias_packed[n] = biases[gate * kHidden + tile * TILE_CH + j];
    }
  }
}


int pack_cell_weights(const float* kernel, const float* biases, void* w_planes, float* bias_packed,
                      int cx, int P, int comp, cudaStream_t stream) {
  MVB_REQUIRE((P >= 1 && P <= 3) || P == kPlanesF16F8, "pack_cell_weights: planes P=%d not in {1,2,3,%d}", P,
              kPlanesF16F8);
  MVB_REQUIRE(cx >= 1, "pack_cell_weights: cx=%d", cx);
  const int cxp = (cx + XPAD - 1) / XPAD * XPAD;
  const int cpad = cxp + kHidden;
  if (P == kPlanesF16F8) {
    MVB_REQUIRE(!comp, "pack_cell_weights: the compensated x block exists for bf16 planes only");
    const long long total = (long long)kGates * 9 * cpad;
    __half* w16 = reinterpret_cast<__half*>(w_planes);
    uint8_t* w8 = reinterpret_cast<uint8_t*>(w_planes) + 2 * total;
    float* col_scale = reinterpret_cast<float*>(w8 + 2 * total);
    colscale_kernel<<<kGates / 8, 256, 0, stream>>>(kernel, col_scale, cx);
    pack_weights_f16f8_kernel<<<1184, 256, 0, stream>>>(kernel, biases, w16, w8, col_scale, bias_packed, cx, cxp, cpad);
    MVB_CHECK_CUDA(cudaGetLastError());
    count_launch(2);
    return MVB_OK;
  }
  MVB_REQUIRE(!comp || (P == 2 && 4 * cx <= cxp), "pack_cell_weights: compensated x block needs planes=2 and 4*cx <= %d", cxp);
  __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(w_planes);
  const int threads = 256, blocks = 1184;
  switch (P) {
    case 1: pack_weights_kernel<1><<<blocks, threads, 0, stream>>>(kernel, biases, wp, bias_packed, cx, cxp, cpad, comp); break;
    case 2: pack_weights_kernel<2><<<blocks, threads, 0, stream>>>(kernel, biases, wp, bias_packed, cx, cxp, cpad, comp); break;
    default: pack_weights_kernel<3><<<blocks, threads, 0, stream>>>(kernel, biases, wp, bias_packed, cx, cxp, cpad, comp); break;
  }
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

}  // namespace mvb
