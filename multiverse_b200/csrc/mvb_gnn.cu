// K-gnn: the parameter-free graph attention of the class decoder.
//
// Reference: Model.gnn_edge (code/pred_models.py:808-858) builds the dense [HW,HW] cosine matrix
// of F = l2_normalize([h ; mean_t scene_conv]); gnn_mask_edge (:885-909) adds -1e30 everywhere
// except the 3x3 neighbourhood (self included, borders clipped); gnn_node (:860-882) soft-maxes
// and multiplies by h; the caller adds the residual (:378, :651).  Only the <=9 unmasked entries
// of each row survive the softmax in fp32 (exp(-1e30) == 0), so this kernel evaluates exactly
// that 3x3 band:  h'_p = h_p + sum_q softmax_q(F^_p . F^_q) h_q.
//
// One warp walks one image row of one sample row with a 3x3 register window (8 h + 2 scene values
// per lane per cell) plus one column in flight: moving one cell to the right loads only the 3 new cells of a column,
// so every h row is fetched 3x (not 9x) through L1/L2, squared norms are computed once per loaded
// cell, and the centre-left dot product is the previous step's centre-right one.  Per cell: 7 dot
// products + 3 norms reduced by warp shuffles, a <=9-way softmax, 72 FMAs of weighted sum, and the
// result is written straight as the bf16 operand planes of the next cell step.
// HBM-bound by design: 4*HW*(256+64) bytes read, 2*P*HW*256 written per sample row.
#include "mvb_common.cuh"
#include "mvb_kernels.h"

namespace mvb {

constexpr int GNN_WARPS = 4;

struct GnnCol {         // one window column: rows y-1, y, y+1
  float h[3][8];
  float s[3][2];
  float n[3];           // squared norm of [h ; s] (warp-reduced)
};

// Raw loads of one window column, issued one step ahead of their use: with load-then-use the kernel was bound
// by the latency of these loads (ncu: 2.7 warps per issue stalled on the long scoreboard at 12 warps/SM).
__device__ __forceinline__ void gnn_load_raw(GnnCol& c, const float* __restrict__ h32,
                                             const float* __restrict__ scene, long long hrow0,
                                             long long srow0, int x, int W, int Wp, const bool (&rok)[3],
                                             int lane, int y) {
  const bool cok = (x >= 0) && (x < W);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    if (cok && rok[r]) {
      const float4* p4 = reinterpret_cast<const float4*>(h32 + (hrow0 + (long long)(y + r - 1) * Wp + x) * kHidden + lane * 8);
      const float4 a = __ldg(p4), b = __ldg(p4 + 1);
      c.h[r][0] = a.x; c.h[r][1] = a.y; c.h[r][2] = a.z; c.h[r][3] = a.w;
      c.h[r][4] = b.x; c.h[r][5] = b.y; c.h[r][6] = b.z; c.h[r][7] = b.w;
      if (scene) {
        const float2 sv = __ldg(reinterpret_cast<const float2*>(scene + (srow0 + (long long)(y + r - 1) * W + x) * 64 + lane * 2));
        c.s[r][0] = sv.x; c.s[r][1] = sv.y;
      } else {
        c.s[r][0] = 0.f; c.s[r][1] = 0.f;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) c.h[r][k] = 0.f;
      c.s[r][0] = 0.f; c.s[r][1] = 0.f;
    }
  }
}

// Squared norms of a loaded column (warp-reduced).
__device__ __forceinline__ void gnn_norms(GnnCol& c) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) q = fmaf(c.h[r][k], c.h[r][k], q);
    q = fmaf(c.s[r][0], c.s[r][0], q);
    q = fmaf(c.s[r][1], c.s[r][1], q);
    c.n[r] = q;
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) c.n[r] = warp_sum(c.n[r]);
}

__device__ __forceinline__ float gnn_dot(const GnnCol& a, int ra, const GnnCol& b, int rb) {
  float d = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) d = fmaf(a.h[ra][k], b.h[rb][k], d);
  d = fmaf(a.s[ra][0], b.s[rb][0], d);
  d = fmaf(a.s[ra][1], b.s[rb][1], d);
  return d;
}

// MIX: the output is written in the f16f8 operand format (mvb_common.cuh) instead of P bf16 planes
template <int P, bool MIX = false>
__global__ void __launch_bounds__(GNN_WARPS * 32)
gnn_kernel(const float* __restrict__ h32, const int* __restrict__ row_map,
           const float* __restrict__ scene_mean, int beam, __nv_bfloat16* __restrict__ hp_out,
           long long plane_stride, int cpad_out, int ch_off, long long NS, Grid g) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * GNN_WARPS + (threadIdx.x >> 5);
  if (wid >= NS * g.H) return;
  const long long s = wid / g.H;
  const int y = (int)(wid - s * g.H);
  const long long ss = row_map ? (long long)row_map[s] : s;
  const long long hrow0 = ss * g.S;
  const long long srow0 = (s / beam) * (long long)g.H * g.W;
  const bool rok[3] = {y > 0, true, y < g.H - 1};

  // Rotating 4-column register window (no register moves: the loop is unrolled by four and the roles
  // L / C / R / N(ext, in flight) rotate through the four structs): every column is requested one step before
  // its first use.  Measured at 2 560 rows of 36x18: 3 columns (load, then use) 1.86 ms; 4 columns 1.41 ms;
  // 5 columns (two steps ahead, 236 registers) 1.54 ms; an L1 prefetch instead of the fourth column 2.16 ms;
  // 4 columns + folded (transpose) reduction of the ten warp sums 1.54 ms (longer dependent shuffle chain);
  // forcing 3 CTAs/SM (168 registers, spills) 1.59 ms.
  GnnCol W0, W1, W2, W3;
  gnn_load_raw(W0, h32, scene_mean, hrow0, srow0, -1, g.W, g.Wp, rok, lane, y);   // zeros
  gnn_load_raw(W1, h32, scene_mean, hrow0, srow0, 0, g.W, g.Wp, rok, lane, y);
  gnn_load_raw(W2, h32, scene_mean, hrow0, srow0, 1, g.W, g.Wp, rok, lane, y);
  gnn_norms(W0);
  gnn_norms(W1);
  float d_cl = 0.f;   // dot(centre, left-centre), carried from the previous cell
  auto step = [&](const GnnCol& L, const GnnCol& C, GnnCol& R, GnnCol& N, int x) {
    gnn_load_raw(N, h32, scene_mean, hrow0, srow0, x + 2, g.W, g.Wp, rok, lane, y);   // used by the next step
    gnn_norms(R);
    // dots of the centre cell (C,1) with its 8 neighbours; self = squared norm
    float d[9];
    d[0] = gnn_dot(C, 1, L, 0); d[1] = gnn_dot(C, 1, C, 0); d[2] = gnn_dot(C, 1, R, 0);
    d[5] = gnn_dot(C, 1, R, 1);
    d[6] = gnn_dot(C, 1, L, 2); d[7] = gnn_dot(C, 1, C, 2); d[8] = gnn_dot(C, 1, R, 2);
    d[0] = warp_sum(d[0]); d[1] = warp_sum(d[1]); d[2] = warp_sum(d[2]); d[5] = warp_sum(d[5]);
    d[6] = warp_sum(d[6]); d[7] = warp_sum(d[7]); d[8] = warp_sum(d[8]);
    d[3] = d_cl;
    d[4] = C.n[1];
    d_cl = d[5];
    const bool cokL = x > 0, cokR = x < g.W - 1;
    const bool ok[9] = {rok[0] && cokL, rok[0], rok[0] && cokR, cokL, true, cokR,
                        rok[2] && cokL, rok[2], rok[2] && cokR};
    const float nq[9] = {L.n[0], C.n[0], R.n[0], L.n[1], C.n[1], R.n[1], L.n[2], C.n[2], R.n[2]};
    // tf.nn.l2_normalize: x * rsqrt(max(sum x^2, 1e-12))
    const float inv_p = rsqrtf(fmaxf(C.n[1], 1e-12f));
    float e[9], m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      e[k] = d[k] * inv_p * rsqrtf(fmaxf(nq[k], 1e-12f));
      if (ok[k]) m = fmaxf(m, e[k]);
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) { e[k] = ok[k] ? __expf(e[k] - m) : 0.f; sum += e[k]; }
    const float inv_sum = 1.0f / sum;
    float o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] = C.h[1][c];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float aL = e[r * 3 + 0] * inv_sum, aC = e[r * 3 + 1] * inv_sum, aR = e[r * 3 + 2] * inv_sum;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        o[c] = fmaf(aL, L.h[r][c], o[c]);
        o[c] = fmaf(aC, C.h[r][c], o[c]);
        o[c] = fmaf(aR, R.h[r][c], o[c]);
      }
    }
    const long long orow = s * g.S + (long long)y * g.Wp + x;
    if (MIX) {
      store_f16f8_x8(hp_out, plane_stride, orow, ch_off + lane * 8, cpad_out, o);
      return;
    }
    uint32_t pk[P][4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      __nv_bfloat16 a[P], b[P];
      split_planes<P>(o[2 * v], a);
      split_planes<P>(o[2 * v + 1], b);
#pragma unroll
      for (int p = 0; p < P; ++p) pk[p][v] = pack_bf16x2(a[p], b[p]);
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
      uint4* po = reinterpret_cast<uint4*>(hp_out + p * plane_stride + orow * cpad_out + ch_off + lane * 8);
      *po = make_uint4(pk[p][0], pk[p][1], pk[p][2], pk[p][3]);
    }
  };
  for (int x = 0; x < g.W; x += 4) {
    step(W0, W1, W2, W3, x);
    if (x + 1 < g.W) step(W1, W2, W3, W0, x + 1);
    if (x + 2 < g.W) step(W2, W3, W0, W1, x + 2);
    if (x + 3 < g.W) step(W3, W0, W1, W2, x + 3);
  }
}

int gnn_attend_fwd(const float* h32, const int* row_map, const float* scene_mean, int beam,
                   void* hp_out, long long hp_plane_stride, int cpad_out, int ch_off_out,
                   long long NS, int H, int W, int P, cudaStream_t stream) {
  MVB_REQUIRE((P >= 1 && P <= 3) || P == kPlanesF16F8, "gnn_attend_fwd: planes P=%d", P);
  MVB_REQUIRE(h32 && hp_out && NS > 0 && beam >= 1, "gnn_attend_fwd: bad args");
  MVB_REQUIRE(cpad_out % 8 == 0 && ch_off_out % 8 == 0, "gnn_attend_fwd: pitch/offset must be multiples of 8");
  const Grid g = make_grid(H, W);
  const long long warps = NS * H;
  const unsigned blocks = (unsigned)((warps + GNN_WARPS - 1) / GNN_WARPS);
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(hp_out);
  switch (P) {
    case 1: gnn_kernel<1><<<blocks, GNN_WARPS * 32, 0, stream>>>(h32, row_map, scene_mean, beam, d, hp_plane_stride, cpad_out, ch_off_out, NS, g); break;
    case 2: gnn_kernel<2><<<blocks, GNN_WARPS * 32, 0, stream>>>(h32, row_map, scene_mean, beam, d, hp_plane_stride, cpad_out, ch_off_out, NS, g); break;
    case kPlanesF16F8: gnn_kernel<2, true><<<blocks, GNN_WARPS * 32, 0, stream>>>(h32, row_map, scene_mean, beam, d, hp_plane_stride, cpad_out, ch_off_out, NS, g); break;
    default: gnn_kernel<3><<<blocks, GNN_WARPS * 32, 0, stream>>>(h32, row_map, scene_mean, beam, d, hp_plane_stride, cpad_out, ch_off_out, NS, g); break;
  }
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

}  // namespace mvb
