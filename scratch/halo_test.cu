// Experiment: can a tcgen05 SS-mode MMA read its A operand from a SHIFTED WINDOW of a TMA-loaded halo tile
// (128-byte swizzled pixel rows, start address not 1024-aligned, 8-row group stride = PW * 128 bytes)?
// D = A_window x I  ->  D[m][c] = X[y0 + h + dy][x0 + w + dx][c],  m = h * 8 + w.
#include <cstdio>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../gtsfm_b200/csrc/common.cuh"
#include "../gtsfm_b200/csrc/tma.cuh"
#include "../gtsfm_b200/csrc/conv_ps.cuh"

constexpr int HH = 40, WW = 40, X0 = 8, Y0 = 8;

__device__ __forceinline__ uint64_t desc_var(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

struct Maps {
  CUtensorMap a, b;
};

// mode 0: base_offset 0;  mode 1: base_offset = (start >> 7) & 7
__global__ void __launch_bounds__(128, 1) k_halo(const __grid_constant__ Maps maps, int PW, int mode, float* out /*[9][128][64]*/) {
  extern __shared__ unsigned char raw_[];
  const uint32_t raw = tc::smem_u32(raw_);
  const uint32_t smem0 = (raw + 1023u) & ~1023u;
  __shared__ uint64_t full, done;
  __shared__ uint32_t slot;
  const int t = threadIdx.x, warp = t >> 5;
  const uint32_t sA = smem0, sB = smem0 + 48 * 1024;
  if (t == 0) {
    tc::mbar_init(&full, 1), tc::mbar_init(&done, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&slot, 64);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = slot;
  if (t == 0) {
    tc::mbar_expect_tx(&full, 18 * PW * 128 + 64 * 128);
    tc::tma_load_3d(sA, &maps.a, &full, 0, X0 - 1, Y0 - 1);
    tc::tma_load_2d(sB, &maps.b, &full, 0, 0);
  }
  tc::mbar_wait(&full, 0);
  __syncthreads();
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    if (warp == 0) {
      const uint32_t start = sA + ((dy + 1) * PW + (dx + 1)) * 128;
      const uint64_t dA = desc_var(start, PW * 128, mode ? (start >> 7) & 7 : 0), dB = tc::smem_desc_sw128(sB);
      const uint32_t idesc = tc::idesc_f16(128, 64);
      for (int ks = 0; ks < 4; ++ks) tc::umma_f16_w(tmem, dA + ks * 2, dB + ks * 2, idesc, ks ? 1u : 0u);
      tc::umma_commit_w(&done);
    }
    tc::mbar_wait(&done, tap & 1);
    tc::fence_after_sync();
    float v[32];
    for (int cc = 0; cc < 2; ++cc) {
      tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cc * 32, v);
      for (int j = 0; j < 32; ++j) out[((size_t)tap * 128 + t) * 64 + cc * 32 + j] = v[j];
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
  }
  if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<__half> X((size_t)HH * WW * 64), I(64 * 64);
  auto xv = [](int y, int x, int c) { return (c & 1) ? (float)c : (float)((y * WW + x) % 1024); };
  for (int y = 0; y < HH; ++y)
    for (int x = 0; x < WW; ++x)
      for (int c = 0; c < 64; ++c) X[((size_t)y * WW + x) * 64 + c] = __float2half(xv(y, x, c));
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 64; ++k) I[n * 64 + k] = __float2half(n == k ? 1.f : 0.f);
  __half *dX, *dI;
  float* dO;
  cudaMalloc(&dX, X.size() * 2), cudaMalloc(&dI, I.size() * 2), cudaMalloc(&dO, 9 * 128 * 64 * 4);
  cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice), cudaMemcpy(dI, I.data(), I.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = 48 * 1024 + 8 * 1024 + 1024;
  cudaFuncSetAttribute(k_halo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  PFN_encodeTiled enc = tma_encoder();
  if (!enc) return printf("no encoder\n"), 1;
  for (int PW : {10, 16, 12, 24}) {
    Maps m;
    cuuint64_t dims[3] = {64, (cuuint64_t)WW, (cuuint64_t)HH};
    cuuint64_t strides[2] = {64 * 2, (cuuint64_t)WW * 64 * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)PW, 18};
    cuuint32_t estr[3] = {1, 1, 1};
    if (18 * PW * 128 > 48 * 1024) continue;
    CUresult r = enc(&m.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, dX, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS || !tma_map_2d(&m.b, dI, 64, 64, 64, 64)) return printf("encode failed %d\n", (int)r), 1;
    for (int mode = 0; mode < 2; ++mode) {
      cudaMemset(dO, 0, 9 * 128 * 64 * 4);
      k_halo<<<1, 128, smem>>>(m, PW, mode, dO);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> O(9 * 128 * 64);
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      printf("PW=%2d base_offset mode %d (%s): mismatches per tap:", PW, mode, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
      for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        int bad = 0;
        for (int mrow = 0; mrow < 128; ++mrow)
          for (int c = 0; c < 64; ++c)
            if (O[((size_t)tap * 128 + mrow) * 64 + c] != xv(Y0 + mrow / 8 + dy, X0 + mrow % 8 + dx, c)) ++bad;
        printf(" %d", bad);
      }
      printf("\n");
    }
  }
  return 0;
}
