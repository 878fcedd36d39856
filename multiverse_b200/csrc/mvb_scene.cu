// Scene CNN of code/pred_models.py:146-165: two `tanh(conv3x3 stride 2 SAME + b)` layers over the
// per-frame one-hot segmentation (helper conv2d, :1333-1373), and the time-mean of the result that
// gnn_edge (:826-828) concatenates to the hidden state.  Runs once per batch on the UNIQUE frames
// (scene_feat [F,...]); callers index the result with obs_scene instead of materialising
// embedding_lookup's [N*T,...] copy (:148-152).  <0.1 % of the path's FLOPs; direct convolution.
#include "mvb_common.cuh"
#include "mvb_kernels.h"

namespace mvb {

constexpr int SC_PIX = 4;  // output pixels per block

// block = Cout x SC_PIX threads; thread (oc, p) computes one output value.
__global__ void scene_conv_kernel(const float* __restrict__ in, const float* __restrict__ Wt,
                                  const float* __restrict__ b, float* __restrict__ out,
                                  long long F, int IH, int IW, int OH, int OW, int pad_t, int pad_l,
                                  int Cin, int Cout) {
  extern __shared__ float patch[];  // [SC_PIX][9][Cin]
  const int oc = threadIdx.x % Cout;
  const int p = threadIdx.x / Cout;
  const long long total_pix = F * OH * OW;
  const long long pix0 = (long long)blockIdx.x * SC_PIX;
  // cooperative patch load (zero for SAME padding)
  for (int i = threadIdx.x; i < SC_PIX * 9 * Cin; i += blockDim.x) {
    const int ci = i % Cin;
    const int tap = (i / Cin) % 9;
    const int pp = i / (9 * Cin);
    const long long pix = pix0 + pp;
    float v = 0.f;
    if (pix < total_pix) {
      const int ox = (int)(pix % OW);
      const int oy = (int)((pix / OW) % OH);
      const long long f = pix / ((long long)OW * OH);
      const int iy = oy * 2 - pad_t + tap / 3;
      const int ix = ox * 2 - pad_l + tap % 3;
      if (iy >= 0 && iy < IH && ix >= 0 && ix < IW) v = in[((f * IH + iy) * IW + ix) * Cin + ci];
    }
    patch[i] = v;
  }
  __syncthreads();
  const long long pix = pix0 + p;
  if (pix >= total_pix) return;
  const float* pt = patch + p * 9 * Cin;
  float acc = 0.f;
  for (int k = 0; k < 9 * Cin; ++k) acc = fmaf(pt[k], __ldg(Wt + (long long)k * Cout + oc), acc);
  out[pix * Cout + oc] = tanhf(acc + b[oc]);
}

__global__ void scene_time_mean_kernel(const float* __restrict__ sc, const int* __restrict__ fidx,
                                       float* __restrict__ out, long long N, int T, long long HWC) {
  const long long total = N * HWC;
  const float inv = 1.0f / (float)T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HWC, e = i - n * HWC;
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += sc[(long long)fidx[n * T + t] * HWC + e];
    out[i] = s * inv;
  }
}

int scene_conv_fwd(const float* in, const float* W, const float* b, float* out, long long F, int IH,
                   int IW, int Cin, int Cout, cudaStream_t stream) {
  MVB_REQUIRE(in && W && b && out && F > 0, "scene_conv_fwd: bad args");
  MVB_REQUIRE(Cout * SC_PIX <= 1024 && Cout > 0 && Cin > 0, "scene_conv_fwd: Cout=%d too large", Cout);
  // TF SAME, k=3, stride 2 (oracle/multiverse_ref.py:same_pad)
  const int OH = (IH + 1) / 2, OW = (IW + 1) / 2;
  const int tot_h = (OH - 1) * 2 + 3 - IH > 0 ? (OH - 1) * 2 + 3 - IH : 0;
  const int tot_w = (OW - 1) * 2 + 3 - IW > 0 ? (OW - 1) * 2 + 3 - IW : 0;
  const long long total_pix = F * OH * OW;
  const unsigned blocks = (unsigned)((total_pix + SC_PIX - 1) / SC_PIX);
  const size_t smem = (size_t)SC_PIX * 9 * Cin * sizeof(float);
  scene_conv_kernel<<<blocks, Cout * SC_PIX, smem, stream>>>(in, W, b, out, F, IH, IW, OH, OW,
                                                             tot_h / 2, tot_w / 2, Cin, Cout);
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

int scene_time_mean(const float* scene_conv, const int* frame_idx, float* out, long long N, int T,
                    long long HWC, cudaStream_t stream) {
  MVB_REQUIRE(scene_conv && frame_idx && out && N > 0 && T > 0 && HWC > 0, "scene_time_mean: bad args");
  const long long total = N * HWC;
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads < sm_count() * 16 ? (total + threads - 1) / threads : sm_count() * 16);
  scene_time_mean_kernel<<<blocks, threads, 0, stream>>>(scene_conv, frame_idx, out, N, T, HWC);
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

}  // namespace mvb
