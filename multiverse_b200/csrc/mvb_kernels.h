// Internal C++ launchers (one per kernel group); the extern "C" wrappers live in mvb_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mvb {

void count_launch(int n);

// mvb_cell.cu
int cell_fwd(const void* xh_planes, const void* w_planes, const float* bias, const float* c_in,
             const int* row_map, float* c_out, float* h32_out, void* hp_out,
             long long hp_plane_stride, int cpad_out, int ch_off_out, long long NS, int H, int W,
             int cpad, int P, float forget_bias, float* gates_out, const float* xf_B, const float* xf_T2,
             const int* xf_ids, int fanout, cudaStream_t stream, const float* xr_in = nullptr,
             const float* xr_W = nullptr, const float* xs_tab = nullptr, const int* xs_label = nullptr);
int cell_xdense_weights(const float* kernel, float* out, cudaStream_t stream);
int cell_xsparse_weights(const float* kernel, int cx, float* out, cudaStream_t stream);
int cell_xsparse_table(const float* scene_conv, const int* frame_idx, const int* label, const float* Wx, float* tab,
                       long long NS, int H, int W, cudaStream_t stream);
int cell_last_variant();
unsigned long long cell_variants_seen(int reset);
int cell_xfold_tables(const float* kernel, const float* biases, const float* We, const float* be, int E,
                      float* Bt, float* T2, cudaStream_t stream);
int pack_cell_weights(const float* kernel, const float* biases, void* w_planes, float* bias_packed,
                      int cx, int P, int comp, cudaStream_t stream);

// mvb_layout.cu
int nhwc_to_planes(const float* src, void* dst_planes, long long plane_stride, int cpad, int ch_off,
                   long long NS, int H, int W, int C, int P, int comp, cudaStream_t stream);
int traj_to_grid(const double* traj, const double* centers, double h_gap, double w_gap, int* labels,
                 float* regress, long long NT, int H, int W, cudaStream_t stream);
int nhwc_halo_copy(const float* src, float* dst, long long NS, int H, int W, int C, int to_nhwc,
                   cudaStream_t stream);
int enc_class_input(const float* scene_conv, const int* frame_idx, const int* label,
                    const int* prev_label, void* xh_planes, long long plane_stride, int cpad,
                    long long NS, int H, int W, int P, cudaStream_t stream);

// mvb_scene.cu
int scene_conv_fwd(const float* in, const float* W, const float* b, float* out, long long F, int IH,
                   int IW, int Cin, int Cout, cudaStream_t stream);
int scene_time_mean(const float* scene_conv, const int* frame_idx, float* out, long long N, int T,
                    long long HWC, cudaStream_t stream);

// mvb_gnn.cu
int gnn_attend_fwd(const float* h32, const int* row_map, const float* scene_mean, int beam,
                   void* hp_out, long long hp_plane_stride, int cpad_out, int ch_off_out,
                   long long NS, int H, int W, int P, cudaStream_t stream);

// mvb_head.cu
int head_fwd(const float* h32, const float* Wo, int Pout, float* out, int* ids_out, const float* We,
             const float* be, int E, void* xh_next, long long plane_stride, int cpad, long long NS,
             int H, int W, int P, cudaStream_t stream);
int emb_onehot_fwd(const int* ids, const float* We, const float* be, int E, void* xh_next,
                   long long plane_stride, int cpad, long long NS, int H, int W, int P,
                   cudaStream_t stream);
int emb_dense_fwd(const float* x, const float* We, const float* be, int E, void* xh_next,
                  long long plane_stride, int cpad, long long NS, int H, int W, int P,
                  cudaStream_t stream);

// mvb_beam.cu
int decode_trajectories(const int* ids, const float* offs, const float* centers, float* out, long long N,
                        int K, int Tp, int V, cudaStream_t stream);
int beam_step(const float* logits, const float* score_in, float* score_out, int* ids_out,
              int* parents_out, int* row_map_out, long long N, int B, int V, int first_step,
              int zero_scores, int diverse, float log_gamma, cudaStream_t stream);
int beam_backtrace(const int* step_ids, const int* step_parents, const float* step_logits,
                   int* out_ids, float* out_logits, long long N, int B, int Tp, int V,
                   cudaStream_t stream);

// mvb_metrics.cu
int min_ade_fde(const float* pred, const float* gt, const int* gt_len, double* ade_err, int* ade_idx, double* fde,
                int* fde_idx, long long N, int G, int K, int Tp, int Tg, cudaStream_t stream);
int beam_nll(const float* logits, const float* logprobs, const int* gt_idx, const int* steps, double* nll, int* count,
             long long N, int K, int Tp, int V, int J, int G, cudaStream_t stream);

// mvb_train.cu
int cell_dgrad(const void* dg_planes, const void* wd_planes, float* dxh, long long NS, int H, int W,
               int cpad, int P, int need_x, cudaStream_t stream);
int cell_wgrad(const void* dgT_planes, const void* xhT_planes, float* dwp, long long NS, int H, int W,
               int cpad, long long Rp, int P, cudaStream_t stream);
int cell_wgrad_mn(const void* dg_planes, const void* xh_planes, float* dwp, long long NS, int H, int W,
                  int cpad, int P, cudaStream_t stream);
int lstm_gates_bwd(const float* gates, const float* c_prev, const float* c_new, const float* dh,
                   const float* dc_in, void* dg_planes, long long plane_stride, float* dc_prev,
                   float* dbias_packed, long long NS, int H, int W, int P, cudaStream_t stream);
int transpose_planes(const void* src, void* dst, long long R, int C, long long Rp, int P, int taps,
                     int Wp, cudaStream_t stream);
int pack_cell_weights_dgrad(const float* kernel, void* wd_planes, int cx, int P, cudaStream_t stream);
int unpack_cell_wgrad(const float* dwp, const float* dbias_packed, float* dkernel, float* dbiases,
                      int cx, int comp, int accumulate, int slabs, cudaStream_t stream);
int cell_wgrad_mn_slabs(int cpad);

// mvb_train2.cu
int loss_fwd_bwd(const float* logits, const int* labels, float* dlogits, long long rows, int V,
                 float cls_scale, const float* reg, const float* target, float* dreg, long long nreg,
                 float reg_scale, float* loss_out, cudaStream_t stream);
int head_bwd(const float* h32, const float* dout, const float* Wo, int Pout, float* dWo, float* dh,
             int accumulate_dh, long long NS, int H, int W, cudaStream_t stream);
int emb_bwd(const float* dxh, int cpad, const int* ids, const float* in_map, const float* We,
            const float* be, int E, int Pout, float* dWe, float* dbe, float* d_in, int accumulate_din,
            long long NS, int H, int W, cudaStream_t stream);
int gnn_bwd(const float* h32, const float* scene_mean, const float* gout, float* work, float* dh,
            int accumulate_dh, float* dscene_mean, long long NS, int H, int W, cudaStream_t stream);
int scene_conv_bwd(const float* in, const float* W, const float* out, const float* dout, float* dW,
                   float* db, float* din, long long F, int IH, int IW, int Cin, int Cout,
                   cudaStream_t stream);
int enc_class_input_bwd(const float* dxh, int cpad, const int* frame_idx, const int* label,
                        float* dscene, long long NS, int H, int W, cudaStream_t stream);
int scene_mean_bwd(const float* dmean, const int* frame_idx, float* dscene, long long N, int T,
                   long long HWC, cudaStream_t stream);
int clip_update(float* w, const float* grad, float* s1, float* s2, long long n, int kind, float lr, float p1, float p2,
                float eps, float clip, float wd, float gscale, cudaStream_t stream);
int adv_step(const float* x, const float* adv, const float* grad, float* out, float eps, float step, long long n,
             cudaStream_t stream);
int mix(const float* a, const float* b, float* out, float w, long long n, cudaStream_t stream);
int ce_rows(const float* logits, const int* labels, float* loss, long long rows, int V, cudaStream_t stream);
int enc_class_input_mix(const float* scene_conv, const int* frame_idx, const int* label, const int* label2, float beta,
                        void* xh_planes, long long plane_stride, int cpad, long long NS, int H, int W, int P,
                        cudaStream_t stream);
int enc_class_input_mix_bwd(const float* dxh, int cpad, const int* frame_idx, const int* label, const int* label2,
                            float beta, float* dscene, long long NS, int H, int W, cudaStream_t stream);
int clip_adadelta(float* w, const float* grad, float* acc, float* acc_upd, long long n, float lr,
                  float rho, float eps, float clip, float wd, float gscale, cudaStream_t stream);

}  // namespace mvb
