// K-beam: the per-step selection of the K-way multi-future rollout and its back-trace.
//
// Reference: decoder_loop_fn of Model.grid_decoder_beam_search (code/pred_models.py:547-606):
// log_softmax (:557) + running score (:560) + optional diverse penalty add_div_penalty
// (:1197-1223: rank of every entry inside its own beam row, obtained there with a full
// top_k(k=V) sort + invert_permutation; here by counting - rank = #greater + #equal-with-lower-
// index, which is what a stable descending sort yields), flatten to B*V (beam 0 only while
// time <= 1, :569-573), top_k(B, sorted) (:578; ties -> lower index), score reset while
// time <= fix_num_timestep (:581-584), ids = idx % V, parents = idx // V (:588-591).
// The (c,h) gather by parent (:611-623, gather_helper :1225-1251) is not a copy in this library:
// the kernel emits row_map = n*B + parent and the next K-gnn / K-cell launch reads its state
// through it.  Back-trace: tf.while_loop at :722-764.
//
// One CTA per sample; B*V fp32 candidates live in shared memory.  Latency-bound glue (<1 % of a
// rollout step), kept bit-faithful to fp32 TF arithmetic (no FMA contraction on the penalty).
#include <cstdlib>
#include "mvb_common.cuh"
#include "mvb_kernels.h"

namespace mvb {

constexpr int BEAM_THREADS = 256;

__device__ __forceinline__ void block_argmax(float& v, int& i, float* red_v, int* red_i) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
  if (lane == 0) { red_v[warp] = v; red_i[warp] = i; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bv = red_v[0]; int bi = red_i[0];
    for (int w = 1; w < BEAM_THREADS / 32; ++w)
      if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bi)) { bv = red_v[w]; bi = red_i[w]; }
    red_v[0] = bv; red_i[0] = bi;
  }
  __syncthreads();
  v = red_v[0]; i = red_i[0];
  __syncthreads();
}

__global__ void __launch_bounds__(BEAM_THREADS)
beam_step_kernel(const float* __restrict__ logits, const float* __restrict__ score_in,
                 float* __restrict__ score_out, int* __restrict__ ids_out,
                 int* __restrict__ parents_out, int* __restrict__ row_map_out, int B, int V,
                 int first_step, int zero_scores, int diverse, float log_gamma) {
  extern __shared__ float sm[];
  float* lp = sm;            // [B][V] log-probs (+score)
  float* cand = sm + (size_t)B * V;  // [B][V] candidates (penalised)
  __shared__ float red_v[BEAM_THREADS / 32];
  __shared__ int red_i[BEAM_THREADS / 32];
  __shared__ float row_stat[2];
  const long long n = blockIdx.x;
  const int rows = first_step ? 1 : B;   // all beams are identical at time 1 (:570-573)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // 1. log_softmax per beam row + running score
  for (int b = warp; b < rows; b += BEAM_THREADS / 32) {
    const float* lg = logits + (n * B + b) * V;
    float m = -INFINITY;
    for (int v = lane; v < V; v += 32) m = fmaxf(m, lg[v]);
    m = warp_max(m);
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += expf(lg[v] - m);
    s = warp_sum(s);
    const float lse = logf(s);
    const float sc = score_in ? score_in[n * B + b] : 0.f;
    for (int v = lane; v < V; v += 32) lp[b * V + v] = __fadd_rn(__fsub_rn(__fsub_rn(lg[v], m), lse), sc);
  }
  __syncthreads();
  // 2. diverse penalty: + log(gamma) * rank within the row
  for (int i = threadIdx.x; i < rows * V; i += blockDim.x) {
    float val = lp[i];
    if (diverse) {
      const int b = i / V, v = i - b * V;
      const float* r = lp + b * V;
      int rank = 0;
      for (int u = 0; u < V; ++u) {
        const float o = r[u];
        rank += (o > val) || (o == val && u < v);
      }
      val = __fadd_rn(val, __fmul_rn(log_gamma, (float)rank));
    }
    cand[i] = val;
  }
  __syncthreads();
  // 3. top-B, descending, ties -> lower flat index
  const int ncand = rows * V;
  for (int k = 0; k < B; ++k) {
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < ncand; i += blockDim.x) {
      const float c = cand[i];
      if (c > bv) { bv = c; bi = i; }
    }
    block_argmax(bv, bi, red_v, red_i);
    if (bi == 0x7fffffff) bi = 0;
    if (threadIdx.x == 0) {
      cand[bi] = -INFINITY;
      // exhausted candidates (B > ncand) cannot happen: V >= B in every config
      const int parent = bi / V;
      score_out[n * B + k] = zero_scores ? 0.f : bv;
      ids_out[n * B + k] = bi - parent * V;
      parents_out[n * B + k] = parent;
      row_map_out[n * B + k] = (int)(n * B) + parent;
    }
    __syncthreads();
  }
  (void)row_stat;
}

// Same selection in O(B*V) per beam row instead of the O(V^2) rank count.  With log(gamma) <= 0 (every published
// setting: gamma = 0.01, or no penalty) the penalised value lp - |log gamma|*rank is non-increasing in the rank
// order of its row, and equal values keep index order, so an entry of rank >= B is preceded by B entries of its
// own row in the global order and can never be selected: the global top-B is the top-B of the rows' top-B lists.
// Each warp extracts the top-B of its rows by B arg-max sweeps over the row in shared memory (the k-th sweep's
// winner has rank k - the same rank the count gives, ties to the lower index); warp 0 then picks the B best of
// the <= B*B candidates, ties to the lower flat index.  Bit-identical outputs to beam_step_kernel.
__device__ __forceinline__ void warp_argmax(float& v, int& i, int& aux) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    const int oa = __shfl_xor_sync(0xffffffffu, aux, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; aux = oa; }
  }
}

__global__ void __launch_bounds__(BEAM_THREADS)
beam_step_topk_kernel(const float* __restrict__ logits, const float* __restrict__ score_in,
                      float* __restrict__ score_out, int* __restrict__ ids_out,
                      int* __restrict__ parents_out, int* __restrict__ row_map_out, int B, int V,
                      int first_step, int zero_scores, int diverse, float log_gamma) {
  extern __shared__ float sm[];
  float* lp = sm;                                   // [rows][V] log-probs + score
  float* cv = sm + (size_t)B * V;                   // [rows][B] candidate values (penalised)
  int* ci = reinterpret_cast<int*>(cv + B * B);     // [rows][B] flat indices b*V + v
  const long long n = blockIdx.x;
  const int rows = first_step ? 1 : B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = warp; b < rows; b += BEAM_THREADS / 32) {
    const float* lg = logits + (n * B + b) * V;
    float* r = lp + (size_t)b * V;
    float m = -INFINITY;
    for (int v = lane; v < V; v += 32) m = fmaxf(m, lg[v]);
    m = warp_max(m);
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += expf(lg[v] - m);
    s = warp_sum(s);
    const float lse = logf(s);
    const float sc = score_in ? score_in[n * B + b] : 0.f;
    for (int v = lane; v < V; v += 32) r[v] = __fadd_rn(__fsub_rn(__fsub_rn(lg[v], m), lse), sc);
    __syncwarp();
    for (int k = 0; k < B; ++k) {
      float bv = -INFINITY; int bi = 0x7fffffff, aux = 0;
      for (int v = lane; v < V; v += 32) {
        const float c = r[v];
        if (c > bv) { bv = c; bi = v; }
      }
      warp_argmax(bv, bi, aux);
      if (lane == 0) {
        if (bi == 0x7fffffff) bi = 0;
        cv[b * B + k] = diverse ? __fadd_rn(bv, __fmul_rn(log_gamma, (float)k)) : bv;
        ci[b * B + k] = b * V + bi;
        r[bi] = -INFINITY;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == 0) {
    const int ncand = rows * B;
    for (int k = 0; k < B; ++k) {
      float bv = -INFINITY; int bi = 0x7fffffff, pos = 0;
      for (int i = lane; i < ncand; i += 32) {
        const float c = cv[i];
        const int f = ci[i];
        if (c > bv || (c == bv && f < bi)) { bv = c; bi = f; pos = i; }
      }
      warp_argmax(bv, bi, pos);
      if (lane == 0) {
        if (bi == 0x7fffffff) bi = 0;
        cv[pos] = -INFINITY;
        ci[pos] = 0x7fffffff;
        const int parent = bi / V;
        score_out[n * B + k] = zero_scores ? 0.f : bv;
        ids_out[n * B + k] = bi - parent * V;
        parents_out[n * B + k] = parent;
        row_map_out[n * B + k] = (int)(n * B) + parent;
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(256)
beam_backtrace_kernel(const int* __restrict__ step_ids, const int* __restrict__ step_parents,
                      const float* __restrict__ step_logits, int* __restrict__ out_ids,
                      float* __restrict__ out_logits, long long N, int B, int Tp, int V) {
  extern __shared__ int src[];  // [Tp] source beam of each step for this (n, b)
  const long long nb = blockIdx.x;
  const long long n = nb / B;
  const int b = (int)(nb - n * B);
  if (threadIdx.x == 0) {
    int p = b;                                     // initial parents = range(B), :714-716
    for (int tau = Tp - 1; tau >= 0; --tau) {
      const long long o = ((long long)tau * N + n) * B + p;
      src[tau] = p;
      out_ids[(n * B + b) * Tp + tau] = step_ids[o];
      p = step_parents[o];
    }
  }
  __syncthreads();
  for (int tau = 0; tau < Tp; ++tau) {
    const float* s = step_logits + (((long long)tau * N + n) * B + src[tau]) * V;
    float* d = out_logits + ((n * B + b) * (long long)Tp + tau) * V;
    for (int v = threadIdx.x; v < V; v += blockDim.x) d[v] = s[v];
  }
}

// Post-decode: trajectory point = centre[cell] + offset[cell] for the K selected cells,
// what the caller does on the host with the fetched [N,K,Tp,HW] logits and [N,Tp,HW,2] offsets
// (code/multifuture_inference.py:504-517, code/pred_utils.py:460-492).  ids [N,K,Tp]; offs [Tp,N,HW,2];
// centers [HW,2] -> out [N,K,Tp,2]: 1.9 KB per trajectory leave the device instead of 680 KB.
__global__ void decode_traj_kernel(const int* __restrict__ ids, const float* __restrict__ offs,
                                   const float* __restrict__ centers, float* __restrict__ out,
                                   long long N, int K, int Tp, int V) {
  const long long total = N * K * Tp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % Tp);
    const long long n = i / ((long long)Tp * K);
    const int id = ids[i];
    const float2 c = *reinterpret_cast<const float2*>(centers + 2 * id);
    const float2 o = *reinterpret_cast<const float2*>(offs + (((long long)t * N + n) * V + id) * 2);
    *reinterpret_cast<float2*>(out + 2 * i) = make_float2(c.x + o.x, c.y + o.y);
  }
}

int decode_trajectories(const int* ids, const float* offs, const float* centers, float* out, long long N,
                        int K, int Tp, int V, cudaStream_t stream) {
  MVB_REQUIRE(ids && offs && centers && out && N > 0 && K > 0 && Tp > 0 && V > 0, "decode_trajectories: bad args");
  const long long total = N * K * Tp;
  const int blocks = (int)((total + 255) / 256 < sm_count() * 8 ? (total + 255) / 256 : sm_count() * 8);
  decode_traj_kernel<<<blocks, 256, 0, stream>>>(ids, offs, centers, out, N, K, Tp, V);
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

int beam_step(const float* logits, const float* score_in, float* score_out, int* ids_out,
              int* parents_out, int* row_map_out, long long N, int B, int V, int first_step,
              int zero_scores, int diverse, float log_gamma, cudaStream_t stream) {
  MVB_REQUIRE(logits && score_out && ids_out && parents_out && row_map_out, "beam_step: null pointer");
  MVB_REQUIRE(N > 0 && B >= 1 && V >= B, "beam_step: bad sizes N=%lld B=%d V=%d", N, B, V);
  const char* full = getenv("MVB_BEAM_FULL_RANK");   // tests: force the O(V^2) rank-count kernel
  if ((!diverse || log_gamma <= 0.f) && !(full && full[0] == '1')) {
    const size_t smem_t = sizeof(float) * ((size_t)B * V + 2 * (size_t)B * B);
    MVB_REQUIRE(smem_t <= 227 * 1024, "beam_step: B*V=%d too large for shared memory", B * V);
    static SmemOptIn opt_t;
    if (smem_t > 48 * 1024) MVB_CHECK_CUDA(smem_opt_in(opt_t, beam_step_topk_kernel, smem_t));
    beam_step_topk_kernel<<<(unsigned)N, BEAM_THREADS, smem_t, stream>>>(logits, score_in, score_out, ids_out,
                                                                         parents_out, row_map_out, B, V, first_step,
                                                                         zero_scores, diverse, log_gamma);
    MVB_CHECK_CUDA(cudaGetLastError());
    count_launch(1);
    return MVB_OK;
  }
  // log(gamma) > 0 rewards high ranks: no per-row bound on the winners, use the full rank count
  const size_t smem = sizeof(float) * 2 * (size_t)B * V;
  MVB_REQUIRE(smem <= 227 * 1024, "beam_step: B*V=%d too large for shared memory", B * V);
  static SmemOptIn opt;
  if (smem > 48 * 1024) MVB_CHECK_CUDA(smem_opt_in(opt, beam_step_kernel, smem));
  beam_step_kernel<<<(unsigned)N, BEAM_THREADS, smem, stream>>>(logits, score_in, score_out, ids_out,
                                                                parents_out, row_map_out, B, V, first_step,
                                                                zero_scores, diverse, log_gamma);
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

int beam_backtrace(const int* step_ids, const int* step_parents, const float* step_logits,
                   int* out_ids, float* out_logits, long long N, int B, int Tp, int V,
                   cudaStream_t stream) {
  MVB_REQUIRE(step_ids && step_parents && step_logits && out_ids && out_logits, "beam_backtrace: null pointer");
  MVB_REQUIRE(N > 0 && B >= 1 && Tp >= 1 && V >= 1, "beam_backtrace: bad sizes");
  beam_backtrace_kernel<<<(unsigned)(N * B), 256, sizeof(int) * Tp, stream>>>(
      step_ids, step_parents, step_logits, out_ids, out_logits, N, B, Tp, V);
  MVB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return MVB_OK;
}

}  // namespace mvb
