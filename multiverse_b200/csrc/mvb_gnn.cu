// K-gnn: the parameter-free graph attention of the class decoder.
//
// Reference: Model.gnn_edge (code/pred_models.py:808-858) builds the dense [HW,HW] cosine matrix
// of F = l2_normalize([h ; mean_t scene_conv]); gnn_mask_edge (:885-909) adds -1e30 everywhere
// except the 3x3 neighbourhood (self included, borders clipped); gnn_node (:860-882) soft-maxes
// and multiplies by h; the caller adds the residual (:378, :651).  Only the <=9 unmasked entries
// of each row survive the softmax in fp32 (exp(-1e30) == 0), so this kernel evaluates exactly
// that 3x3 band:  h'_p = h_p + sum_q softmax_q(F^_p . F^_q) h_q.
//
// One warp per cell: the 9 neighbour rows of h (256 fp32) and scene_mean (64 fp32) are held in
// registers (8 + 2 values per lane each), 9 dot products + 9 squared norms are reduced with warp
// shuffles, and the result is written straight as the bf16 operand planes of the next cell step.
// HBM-bound: 4*HW*(256+64) bytes read, 2*P*HW*256 written per sample row.
#include "mvb_common.cuh"
#include "mvb_kernels.h"

namespace mvb {

constexpr int GNN_WARPS = 8;

template <int P>
__global__ void __launch_bounds__(GNN_WARPS * 32)
gnn_kernel(const float* __restrict__ h32, const int* __restrict__ row_map,
           const float* __restrict__ scene_mean, int beam, __nv_bfloat16* __restrict__ hp_out,
           long long plane_stride, int cpad_out, int ch_off, long long NS, Grid g) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * GNN_WARPS + (threadIdx.x >> 5);
  const int hw = g.H * g.W;
  if (wid >= NS * hw) return;
  const long long s = wid / hw;
  const int pix = (int)(wid - s * hw);
  const int y = pix / g.W, x = pix - y * g.W;
  const long long ss = row_map ? (long long)row_map[s] : s;
  const long long n = s / beam;

  float hq[9][8];
  float dot[9], nrm[9];
  bool ok[9];
  float hp[8], sp[2] = {0.f, 0.f};
  {
    const float4* p4 = reinterpret_cast<const float4*>(h32 + (ss * g.S + (long long)y * g.Wp + x) * kHidden + lane * 8);
    const float4 a = __ldg(p4), b = __ldg(p4 + 1);
    hp[0] = a.x; hp[1] = a.y; hp[2] = a.z; hp[3] = a.w; hp[4] = b.x; hp[5] = b.y; hp[6] = b.z; hp[7] = b.w;
    if (scene_mean) {
      const float2 c = __ldg(reinterpret_cast<const float2*>(scene_mean + ((n * g.H + y) * g.W + x) * 64 + lane * 2));
      sp[0] = c.x; sp[1] = c.y;
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
    ok[k] = (yy >= 0) && (yy < g.H) && (xx >= 0) && (xx < g.W);
    float d = 0.f, q = 0.f;
    if (ok[k]) {
      const float4* p4 = reinterpret_cast<const float4*>(h32 + (ss * g.S + (long long)yy * g.Wp + xx) * kHidden + lane * 8);
      const float4 a = __ldg(p4), b = __ldg(p4 + 1);
      hq[k][0] = a.x; hq[k][1] = a.y; hq[k][2] = a.z; hq[k][3] = a.w;
      hq[k][4] = b.x; hq[k][5] = b.y; hq[k][6] = b.z; hq[k][7] = b.w;
#pragma unroll
      for (int c = 0; c < 8; ++c) { d = fmaf(hp[c], hq[k][c], d); q = fmaf(hq[k][c], hq[k][c], q); }
      if (scene_mean) {
        const float2 sq = __ldg(reinterpret_cast<const float2*>(scene_mean + ((n * g.H + yy) * g.W + xx) * 64 + lane * 2));
        d = fmaf(sp[0], sq.x, d); d = fmaf(sp[1], sq.y, d);
        q = fmaf(sq.x, sq.x, q); q = fmaf(sq.y, sq.y, q);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) hq[k][c] = 0.f;
    }
    dot[k] = d; nrm[k] = q;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) { dot[k] = warp_sum(dot[k]); nrm[k] = warp_sum(nrm[k]); }
  // tf.nn.l2_normalize: x * rsqrt(max(sum x^2, 1e-12))
  const float inv_p = rsqrtf(fmaxf(nrm[4], 1e-12f));
  float e[9], m = -INFINITY;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    e[k] = dot[k] * inv_p * rsqrtf(fmaxf(nrm[k], 1e-12f));
    if (ok[k]) m = fmaxf(m, e[k]);
  }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) { e[k] = ok[k] ? __expf(e[k] - m) : 0.f; sum += e[k]; }
  const float inv_sum = 1.0f / sum;
  float o[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) o[c] = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float a = e[k] * inv_sum;
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] = fmaf(a, hq[k][c], o[c]);
  }
  uint32_t pk[P][4];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    __nv_bfloat16 a[P], b[P];
    split_planes<P>(hp[2 * v] + o[2 * v], a);
    split_planes<P>(hp[2 * v + 1] + o[2 * v + 1], b);
#pragma unroll
    for (int p = 0; p < P; ++p) pk[p][v] = pack_bf16x2(a[p], b[p]);
  }
  const long long orow = s * g.S + (long long)y * g.Wp + x;
#pragma unroll
  for (int p = 0; p < P; ++p) {
    uint4* po = reinterpret_cast<uint4*>(hp_out + p * plane_stride + orow * cpad_out + ch_off + lane * 8);
    *po = make_uint4(pk[p][0], pk[p][1], pk[p][2], pk[p][3]);
  }
}

int gnn_attend_fwd(const float* h32, const int* row_map, const float* scene_mean, int beam,
                   void* hp_out, long long hp_plane_stride, int cpad_out, int ch_off_out,
                   long long NS, int H, int W, int P, cudaStream_t stream) {
  MVB_REQUIRE(P >= 1 && P <= 3, "gnn_attend_fwd: planes P=%d", P);
  MVB_REQUIRE(h32 && hp_out && NS > 0 && beam >= 1, "gnn_attend_fwd: bad args");
  MVB_REQUIRE(cpad_out % 8 == 0 && ch_off_out % 8 == 0, "gnn_attend_fwd: pitch/offset must be multiples of 8");
  const Grid g = make_grid(H, W);
  const long long warps = NS * H * W;
  const unsigned blocks = (unsigned)((warps + GNN_WARPS - 1) / GNN_WARPS);
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(hp_out);
  switch (P) {
    case 1: gnn_kernel<1><<<blocks, GNN_WARPS * 32, 0, stream>>>(h32, row_map, scene_mean, beam, d, hp_plane_stride, cpad_out, ch_off_out, NS, g); break;
    case 2: gnn_kernel<2><<<blocks, GNN_WARPS * 32, 0, stream>>>(h32, row_map, scene_mean, beam, d, hp_plane_stride, cpad_out, ch_off_out, NS, g); break;
    default: gnn_kernel<3><<<blocks, GNN_WARPS * 32, 0, stream>>>(h32, row_map, scene_mean, beam, d, hp_plane_stride, cpad_out, ch_off_out, NS, g); break;
  }
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

}  // namespace mvb
