// Micro-benchmark: cycles per tcgen05.mma (M = 128, K = 16, fp16) as a function of N, the number of accumulators the
// instruction stream rotates over, and the A-operand source (shared memory vs TMEM).   nvcc -arch=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#include "../gtsfm_b200/csrc/tc.cuh"
#include "../gtsfm_b200/csrc/gemm_tma.cuh"

template <int N, int NACC, bool TS, bool MN = false, int NOISE = 0, int CE = 0>
__global__ void __launch_bounds__(128, 1) k_mma(int iters, long long* out) {
  extern __shared__ unsigned char raw_[];
  const uint32_t raw = tc::smem_u32(raw_);
  const uint32_t smem0 = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[4];
  if (threadIdx.x == 0) for (int i = 0; i < 4; ++i) tc::mbar_init(&bar2[i], 1);
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  if (threadIdx.x == 0) done = 0;
  // zero operands (avoid NaN side effects)
  for (int i = threadIdx.x; i < (16384 + 256 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw_ + (smem0 - raw))[i] = 0;
  if (threadIdx.x == 0) tc::mbar_init(&bar, 1), tc::fence_mbar_init();
  if (threadIdx.x < 32) tc::tmem_alloc(&slot, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t id = tc::idesc_f16(128, N) | (MN ? tc::IDESC_B_MN_MAJOR : 0u);
    const uint64_t dA = tc::smem_desc_sw128(smem0), dB = tc::smem_desc_sw128(smem0 + 16384);
    const uint32_t tA = tmem + 480;  // 32 columns of A
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t acc = tmem + ((it * 4 + ks) % NACC) * N;
        const uint64_t advB = MN ? (uint64_t)(ks * 128) : (uint64_t)(ks * 2);
        if (TS) tc::umma_f16_ts(acc, tA + ks * 8, dB + advB, id, 1u);
        else tc::umma_f16(acc, dA + ks * 2, dB + advB, id, 1u);
      }
      if (CE && (it % CE) == CE - 1) tc::umma_commit(&bar2[(it / CE) & 3]);  // one commit per 4 * CE MMAs, nobody waits
    }
    long long t1 = clock64();
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0, out[1] = t2 - t0;
    done = 1;
  } else if (NOISE && threadIdx.x >= 32) {
    // other warps hammer TMEM (their own lane quarter, columns 256..383) like the softmax warpgroups do
    const uint32_t ta = tmem + 256 + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
    float a[64];
    uint32_t w[32];
    for (int i = 0; i < 32; ++i) w[i] = i;
    while (!done) {
      if (NOISE & 1) tc::tmem_ld64(ta, a), w[0] += __float_as_uint(a[5]);
      if (NOISE & 2) tc::tmem_st32(ta + 64, w), tc::tmem_st32(ta + 96, w), tc::tmem_st_wait();
    }
    if (w[0] == 0x12345) out[1] = 0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}


__device__ __forceinline__ void umma_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(tc::smem_u32(bar)) : "memory");
}

// the whole warp runs the issue loop (operands warp-uniform), one elected lane issues
template <int N, int CE, int UNI>
__global__ void __launch_bounds__(128, 1) k_mma_w(int iters, long long* out) {
  extern __shared__ unsigned char raw_[];
  const uint32_t raw = tc::smem_u32(raw_);
  const uint32_t smem0 = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[4];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 256 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw_ + (smem0 - raw))[i] = 0;
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    for (int i = 0; i < 4; ++i) tc::mbar_init(&bar2[i], 1);
    tc::fence_mbar_init();
  }
  if (threadIdx.x < 32) tc::tmem_alloc(&slot, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  uint32_t tmem = slot;
  if (UNI) tmem = __shfl_sync(0xffffffffu, tmem, 0);
  if (threadIdx.x < 32) {
    const uint32_t id = tc::idesc_f16(128, N);
    const uint64_t dB = tc::smem_desc_sw128(UNI == 2 ? 0x400u + 16384u : smem0 + 16384);
    const uint32_t tA = tmem + 480;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_ts_elect(tmem, tA + ks * 8, dB + ks * 2, id, 1u);
      if (CE && (it % CE) == CE - 1) commit_elect(&bar2[(it / CE) & 3]);
    }
    long long t1 = clock64();
    commit_elect(&bar);
    tc::mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0, out[1] = t2 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}
template <int N, int CE, int UNI>
void runw(int blocks, long long* d) {
  const int iters = 2048;
  const size_t smem = 16384 + 256 * 128 + 1024;
  cudaFuncSetAttribute(k_mma_w<N, CE, UNI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) k_mma_w<N, CE, UNI><<<blocks, 128, smem>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double n = iters * 4.0;
  printf("warp-uniform issue N=%3d commit/%d uni=%d : issue %.1f cyc/mma, complete %.1f cyc/mma (ideal %d) %s\n", N, CE * 4, UNI, h[0] / n, h[1] / n,
         N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int N, int NACC, bool TS, bool MN = false, int NOISE = 0, int CE = 0>
void run(int blocks, long long* d) {
  const int iters = 2048;
  const size_t smem = 16384 + 256 * 128 + 1024;
  cudaFuncSetAttribute(k_mma<N, NACC, TS, MN, NOISE, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) k_mma<N, NACC, TS, MN, NOISE, CE><<<blocks, 128, smem>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double n = iters * 4.0;
  printf("N=%3d acc=%d A=%s B=%s noise=%d commit/%d blocks=%3d : issue %.1f cyc/mma, complete %.1f cyc/mma (ideal %d) %s\n", N, NACC, TS ? "tmem" : "smem", MN ? "MN" : "K", NOISE, CE * 4, blocks,
         h[0] / n, h[1] / n, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  for (int blocks : {148}) {
    run<64, 1, true>(blocks, d);
    run<64, 1, true, false, 0, 3>(blocks, d);
    runw<64, 0, 0>(blocks, d);
    runw<64, 3, 0>(blocks, d);
    runw<64, 0, 1>(blocks, d);
    runw<64, 3, 1>(blocks, d);
    runw<64, 0, 2>(blocks, d);
    runw<64, 3, 2>(blocks, d);
    runw<128, 3, 1>(blocks, d);
  }
  return 0;
}
