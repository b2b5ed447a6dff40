// Stand-alone timing harness for k_flash_ts: random planes, both problems N x N, prints per-phase cycle sums of CTA 0
// and the kernel time.   nvcc -DB2_ATTN_TIMING -arch=sm_100a   (scratch tool, not part of the library)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../gtsfm_b200/csrc/attn_ts.cuh"

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 5000;
  const int nsplit = argc > 2 ? atoi(argv[2]) : 4;
  const size_t plane = (size_t)4 * N * 64;
  std::vector<__half> h(plane);
  __half* bufs[12];  // per problem: q, k, v, each hi + lo contiguous
  srand(1);
  for (int i = 0; i < 6; ++i) {
    cudaMalloc(&bufs[i], plane * 2 * sizeof(__half));
    for (int pl = 0; pl < 2; ++pl) {
      for (size_t j = 0; j < plane; ++j) h[j] = __float2half(((rand() % 2001) - 1000) * (pl ? 1e-7f : 1e-3f));
      cudaMemcpy(bufs[i] + pl * plane, h.data(), plane * sizeof(__half), cudaMemcpyHostToDevice);
    }
  }
  AttnTsMaps maps;
  AttnTsArgs a{};
  long long* timing;
  cudaMalloc(&timing, 32 * 8);
  cudaMemset(timing, 0, 32 * 8);
  int* err;
  cudaMalloc(&err, 4);
  cudaMemset(err, 0, 4);
  for (int i = 0; i < 2; ++i) {
    __half *q = bufs[3 * i], *k = bufs[3 * i + 1], *v = bufs[3 * i + 2];
    bool ok = tma_map_2d(&maps.kh[i], k, (uint64_t)4 * N, 64, 64, AW_KV) && tma_map_2d(&maps.kl[i], k + plane, (uint64_t)4 * N, 64, 64, AW_KV) &&
              tma_map_2d(&maps.vh[i], v, (uint64_t)4 * N, 64, 64, AW_KV) && tma_map_2d(&maps.vl[i], v + plane, (uint64_t)4 * N, 64, 64, AW_KV);
    if (!ok) { printf("tma map failed\n"); return 1; }
    AttnTsProblem& p = a.p[i];
    p.Qh = q, p.Ql = q + plane, p.Nq = N, p.Nk = N;
    cudaMalloc(&p.Oh, (size_t)N * 256 * 2 * 2);
    p.Ol = p.Oh + (size_t)N * 256;
    cudaMalloc(&p.Opart, (size_t)nsplit * N * 256 * 4);
    cudaMalloc(&p.ml, (size_t)nsplit * 4 * N * 2 * 4);
  }
  a.scale = 0.125f, a.nsplit = nsplit, a.err_flag = err, a.timing = timing;
  cudaFuncSetAttribute(k_flash_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AS_SMEM);
  const int qt = (N + 255) / 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_flash_ts<<<dim3(qt, 4, 2 * nsplit), AS_THREADS, AS_SMEM>>>(maps, a);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 4.0 * 2 * 2 * 64 * 2.0 * N * N;
    printf("N=%d nsplit=%d grid=%d: %.1f us  %.1f TFLOP/s %s\n", N, nsplit, qt * 4 * 2 * nsplit, ms * 1e3, fl / ms * 1e-9, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  long long t[32];
  int herr;
  cudaMemcpy(t, timing, sizeof(t), cudaMemcpyDeviceToHost);
  cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
  const double T = (double)t[24];
  printf("err=%d tiles=%d, cycles per tile:\n", herr, (int)t[24]);
  const char* wg[7] = {"wait s_full", "ldtm+max", "rescale", "exp+split", "sttm+wait", "wait o_full", "fence+arrive"};
  for (int q = 0; q < 2; ++q) {
    double s = 0;
    printf(" WG%d:", q);
    for (int i = 0; i < 7; ++i) printf(" %s %.0f |", wg[i], t[q * 8 + i] / T), s += t[q * 8 + i] / T;
    printf(" total %.0f\n", s);
  }
  const char* mm[6] = {"A:wait kv", "wait p0", "issue0", "B:wait kv", "wait p1", "issue1"};
  double s = 0;
  printf(" MMA:");
  for (int i = 0; i < 6; ++i) printf(" %s %.0f |", mm[i], t[16 + i] / T), s += t[16 + i] / T;
  printf(" total %.0f\n", s);
  return 0;
}
