// Test-only entry points: run one GEMM through the SIMT fp32 kernel or the tcgen05 split-fp16 kernel (parity tests).
#include <stdlib.h>

#include "common.cuh"
#include "gemm.cuh"
#include "gemm_tc.cuh"
#include "gemm_tma.cuh"

extern "C" int b2_debug_gemm_host(b2_context* ctx, int mode, const float* A, const float* B, const float* bias, float* C,
                                  int M, int N, int K) {
  if (!ctx || !A || !B || !C || M <= 0 || N <= 0 || K <= 0 || (K % 64)) return B2_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  DevBuf dA, dB, dC, dBias, dBh, dBl, dErr;
  B2_CUDA(ctx, dA.ensure((size_t)M * K * 4));
  B2_CUDA(ctx, dB.ensure((size_t)N * K * 4));
  B2_CUDA(ctx, dC.ensure((size_t)M * N * 4));
  B2_CUDA(ctx, dBias.ensure((size_t)N * 4));
  B2_CUDA(ctx, dErr.ensure(16));
  B2_CUDA(ctx, cudaMemcpyAsync(dA.p, A, (size_t)M * K * 4, cudaMemcpyHostToDevice, st));
  B2_CUDA(ctx, cudaMemcpyAsync(dB.p, B, (size_t)N * K * 4, cudaMemcpyHostToDevice, st));
  if (bias) B2_CUDA(ctx, cudaMemcpyAsync(dBias.p, bias, (size_t)N * 4, cudaMemcpyHostToDevice, st));
  B2_CUDA(ctx, cudaMemsetAsync(dErr.p, 0, 16, st));
  int rc = B2_OK;
  if (mode == 0) {
    rc = launch_gemm(ctx, st, gemm_linear(dA.as<float>(), K, K, dB.as<float>(), bias ? dBias.as<float>() : nullptr, dC.as<float>(), N, M, N));
  } else {
    // modes 1 and 2 both run the tcgen05 kernel on operands split here (A and B as fp16 hi / lo planes)
    DevBuf dAh, dAl;
    B2_CUDA(ctx, dAh.ensure((size_t)M * K * 2));
    B2_CUDA(ctx, dAl.ensure((size_t)M * K * 2));
    B2_CUDA(ctx, dBh.ensure((size_t)N * K * 2));
    B2_CUDA(ctx, dBl.ensure((size_t)N * K * 2));
    B2_LAUNCH(ctx, k_split_f32, (unsigned)(((size_t)M * K + 255) / 256), 256, 0, st, dA.as<float>(), (size_t)M * K, dAh.as<__half>(), dAl.as<__half>());
    B2_LAUNCH(ctx, k_split_f32, (unsigned)(((size_t)N * K + 255) / 256), 256, 0, st, dB.as<float>(), (size_t)N * K, dBh.as<__half>(), dBl.as<__half>());
    GemmTcArgs g{};
    g.p[0].A1h = dAh.as<__half>(), g.p[0].A1l = dAl.as<__half>(), g.lda1 = K, g.K1 = K, g.Bh = dBh.as<__half>(), g.Bl = dBl.as<__half>(), g.ldb = K;
    g.p[0].C = dC.as<float>(), g.ldc = N, g.p[0].M = M, g.N = N, g.bias = bias ? dBias.as<float>() : nullptr, g.scale = 1.f, g.err_flag = dErr.as<int>();
    DevBuf dT;
    B2_CUDA(ctx, dT.ensure(32 * 8));
    B2_CUDA(ctx, cudaMemsetAsync(dT.p, 0, 32 * 8, st));
    g.timing = dT.as<long long>();
    dim3 grid(cdiv(N, TC_N), cdiv(M, TC_M), 1);
    if (mode == 3) {
      GemmTmaMaps maps;
      bool okm = tma_map_2d(&maps.a1h[0], dAh.as<__half>(), M, K, K, TM_M) && tma_map_2d(&maps.a1l[0], dAl.as<__half>(), M, K, K, TM_M) &&
                 tma_map_2d(&maps.bh, dBh.as<__half>(), N, K, K, TM_N) && tma_map_2d(&maps.bl, dBl.as<__half>(), N, K, K, TM_N);
      if (!okm) return b2_fail(ctx, B2_ERR_CUDA, "cuTensorMapEncodeTiled failed");
      maps.a1h[1] = maps.a1h[0], maps.a1l[1] = maps.a1l[0], maps.a2h[0] = maps.a2h[1] = maps.a1h[0], maps.a2l[0] = maps.a2l[1] = maps.a1l[0];
      GemmTmaArgs q{};
      q.p[0].C = dC.as<float>(), q.p[0].M = M, q.K1 = K, q.K2 = 0, q.N = N, q.bias = bias ? dBias.as<float>() : nullptr, q.scale = 1.f;
      q.ldc = N, q.err_flag = dErr.as<int>();
      B2_CUDA(ctx, cudaFuncSetAttribute(k_gemm_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TM_GEMM_SMEM));
      b2_prof_work(ctx, "k_gemm_tma", 2.0 * M * N * K);
      B2_LAUNCH(ctx, k_gemm_tma, dim3(cdiv(N, TM_N), cdiv(M, TM_M), 1), 128, TM_GEMM_SMEM, st, maps, q);
    } else {
      B2_CUDA(ctx, cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_GEMM_SMEM));
      B2_LAUNCH(ctx, k_gemm_tc, grid, 128, TC_GEMM_SMEM, st, g);
    }
    B2_CHECK_LAUNCH(ctx);
    B2_CUDA(ctx, cudaStreamSynchronize(st));
    {
      long long hT[32];
      cudaMemcpy(hT, dT.p, sizeof(hT), cudaMemcpyDeviceToHost);
      if (getenv("B2_GEMM_TIMING")) {
        printf("gemm_tc stamps (cycles since start):");
        for (int i = 1; i < (int)hT[31] && i < 31; ++i) printf(" %lld", hT[i] - hT[0]);
        printf("\n");
      }
      dT.release();
    }
    dAh.release(), dAl.release();
  }
  int err = 0;
  if (rc == B2_OK) {
    cudaError_t e = cudaMemcpyAsync(C, dC.p, (size_t)M * N * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&err, dErr.p, 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = b2_fail(ctx, B2_ERR_CUDA, std::string("debug gemm: ") + cudaGetErrorString(e));
  }
  dA.release(), dB.release(), dC.release(), dBias.release(), dBh.release(), dBl.release(), dErr.release();
  if (rc == B2_OK && err) rc = b2_fail(ctx, B2_ERR_STATE, "tcgen05 pipeline timed out on an mbarrier (kernel bug)");
  return rc;
}
