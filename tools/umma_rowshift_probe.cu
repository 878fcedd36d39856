// Hardware probe (not product code): may a tcgen05 K-major shared-memory operand START AT AN ARBITRARY ROW of a
// TMA-written swizzled tile?  If yes, the 9 taps of the ConvLSTM's 3x3 stencil can all read ONE halo'd A tile per
// channel chunk (rows [m0-Wp-1, m0+128+Wp+1)) instead of nine separately fetched shifted copies.
//
// For each swizzle (32/64/128 B rows), row offset r and descriptor base_offset mode, one M128 x N32 x K16 bf16 MMA
// with B = identity slice reads A[r+m][16*ks + n] into D[m][n]; the host checks which element arrived.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I multiverse_b200/csrc tools/umma_rowshift_probe.cu -o /tmp/probe
#include "mvb_common.cuh"
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace mvb;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_fn() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return (EncodeTiledFn)p;
}

namespace mvb {
void set_error(const char*, ...) {}
const char* get_error() { return ""; }
}

constexpr int ROWS_A = 256, N_B = 32;

__device__ __forceinline__ uint64_t desc_bo(uint32_t saddr, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  uint64_t d = make_smem_desc(saddr, sbo, layout);
  d |= (uint64_t)(base_off & 7u) << 49;
  return d;
}

// row_bytes: 32 / 64 / 128 (= swizzle span).  mode 0: base_offset 0; mode 1: (addr >> 7) & 7.
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out,
             int row_bytes, int r, int ks, int mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sA = smem;                          // 256 rows x row_bytes
  uint8_t* sB = smem + 32768;                  // 32 rows x row_bytes
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 40960);
  uint64_t* mbar = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mbar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (ROWS_A + N_B) * row_bytes);
    tma_load_3d(sA, &tmA, bar, 0, 0, 0);
    tma_load_3d(sB, &tmB, bar, 0, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t layout = row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : 6u;
    const uint32_t sbo = 8 * row_bytes;
    const uint32_t a_addr = smem_u32(sA) + r * row_bytes + ks * 32;
    const uint32_t b_addr = smem_u32(sB) + ks * 32;
    const uint32_t bo = mode ? ((a_addr >> 7) & 7u) : 0u;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((N_B >> 3) << 17) | ((128 >> 4) << 24);
    umma_bf16(tmem, desc_bo(a_addr, sbo, layout, bo), desc_bo(b_addr, sbo, layout, 0), idesc, 0u);
    umma_commit(mbar);
  }
  __syncthreads();
  mbar_wait(mbar, 0);
  tc_fence_after();
  uint32_t v[16];
  for (int c = 0; c < 2; ++c) {
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c * 16, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(warp * 32 + (threadIdx.x & 31)) * N_B + c * 16 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

int main() {
  EncodeTiledFn enc = get_fn();
  if (!enc) { printf("no encode fn\n"); return 1; }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  float* d_out; cudaMalloc(&d_out, 128 * N_B * 4);
  std::vector<float> out(128 * N_B);
  const int rbs[3] = {64, 128, 32};
  for (int which = 0; which < 2; ++which) {       // 0: A[row][k] = row;  1: A[row][k] = k
    for (int rb : rbs) {
      const int kel = rb / 2;                      // bf16 elements per row
      std::vector<__nv_bfloat16> hA(ROWS_A * kel), hB(N_B * kel);
      for (int i = 0; i < ROWS_A; ++i) for (int k = 0; k < kel; ++k)
        hA[i * kel + k] = __float2bfloat16(which == 0 ? (float)i : (float)k);
      for (int n = 0; n < N_B; ++n) for (int k = 0; k < kel; ++k)
        hB[n * kel + k] = __float2bfloat16((k % 16 == n % 16 && n < 16) ? 1.f : 0.f);   // identity per 16-wide K slice
      __nv_bfloat16 *dA, *dB;
      cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2);
      cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
      cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
      CUtensorMapSwizzle sw = rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
      CUtensorMap tmA, tmB;
      cuuint64_t dimsA[3] = {(cuuint64_t)kel, ROWS_A, 1}, dimsB[3] = {(cuuint64_t)kel, N_B, 1};
      cuuint64_t strA[2] = {(cuuint64_t)rb, (cuuint64_t)rb * ROWS_A}, strB[2] = {(cuuint64_t)rb, (cuuint64_t)rb * N_B};
      cuuint32_t boxA[3] = {(cuuint32_t)kel, ROWS_A, 1}, boxB[3] = {(cuuint32_t)kel, N_B, 1}, es[3] = {1, 1, 1};
      CUresult r1 = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dA, dimsA, strA, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      CUresult r2 = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dB, dimsB, strB, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r1 || r2) { printf("encode failed %d %d\n", (int)r1, (int)r2); return 1; }
      const int shifts[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 18, 19, 20, 37, 38, 64, 100, 127};
      for (int mode = 0; mode < 2; ++mode) {
        int bad_cfg = 0;
        for (int r : shifts) for (int ks = 0; ks < kel / 16; ++ks) {
          cudaMemset(d_out, 0xff, 128 * N_B * 4);
          probe_kernel<<<1, 128, 48 * 1024>>>(tmA, tmB, d_out, rb, r, ks, mode);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
          int bad = 0, first_m = -1, first_n = -1; float got = 0, want = 0;
          for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
            const float w = which == 0 ? (float)(r + m) : (float)(16 * ks + n);
            if (out[m * N_B + n] != w) { if (!bad) { first_m = m; first_n = n; got = out[m * N_B + n]; want = w; } ++bad; }
          }
          if (bad) { ++bad_cfg; printf("  which=%d rb=%d mode=%d r=%d ks=%d: %d wrong (first m=%d n=%d got %g want %g)\n", which, rb, mode, r, ks, bad, first_m, first_n, got, want); }
        }
        printf("which=%d row_bytes=%d base_offset_mode=%d: %s\n", which, rb, mode, bad_cfg ? "MISMATCHES" : "ALL OK");
      }
      cudaFree(dA); cudaFree(dB);
    }
  }
  return 0;
}
