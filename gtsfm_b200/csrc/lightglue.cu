// LightGlue matcher for sm_100a, as GTSfM drives it (features = "superpoint").
//
// Reference semantics restated (paths relative to the reference repo):
//   thirdparty/LightGlue/lightglue/lightglue.py:31-43 (bbox keypoint normalisation), :68-81 (rotary table),
//   :140-172 (SelfBlock), :175-230 (CrossBlock, shared-sim branch), :84-94 + :645-656 (confidence / early exit),
//   :636-643 + :551-566 (pruning), :265-318 (assignment + filter), :474-629 (_forward); wrapper
//   gtsfm/frontend/matcher/lightglue_matcher.py:43-112.
//
// HBM layout: residual streams [N][256] fp32 row-major; attention operands head-major [4][N][64]; rotary table
// cos/sin [N][32]; the (N0 x N1) similarity of the final assignment is materialised once in fp32.
#include <stdlib.h>

#include "common.cuh"
#include "linear.cuh"

namespace {
constexpr int LG_LAYERS = 9;
constexpr size_t LG_NFLOATS = 11851601;
constexpr int D = 256;

struct SelfW {
  float *wqkv, *bqkv, *wout, *bout, *w0, *b0, *lng, *lnb, *w3, *b3;
};
struct CrossW {
  float *wqk, *bqk, *wv, *bv, *wout, *bout, *w0, *b0, *lng, *lnb, *w3, *b3;
};
struct AssignW {
  float *wm, *bm, *wf, *bf;
};
struct ConfW {
  float *w, *b;
};
}  // namespace

struct LgSide {  // per-image workspace
  DevBuf x[2], xs[2], qkv, q, k, v, ctx, msg, h, hs, cs[2], sn[2], ind[2], conf, mat, src, md, rmax, rlog, ls, lsg, amax, aidx;
  int cur = 0;  // which of x / xs / cs / sn / ind is live
  int n = 0;
  int cap = 0;  // rows allocated; split-plane buffers keep their lo plane at +cap * width halves whatever n shrinks to
};

struct LightGlueState {
  bool loaded = false;
  int persist_ctas = 148;  // CTAs of the persistent kernels = SMs of the device minus the context's reserve_sms
  DevBuf wblob, wblob_h, wblob_l, errflag;  // fp32 weights + their split-fp16 (hi, lo * 2^11) copies for tcgen05
  bool use_tc = true;                          // B2_FORCE_SIMT=1 keeps every GEMM on the exact-fp32 SIMT kernel
  float* wr = nullptr;
  SelfW sw[LG_LAYERS];
  CrossW cw[LG_LAYERS];
  AssignW aw[LG_LAYERS];
  ConfW tw[LG_LAYERS - 1];
  float thr[LG_LAYERS];
  LgSide side[2];
  DevBuf sim, counters, bbox, outm, outs, attn_part[2], attn_ml[2];
  HostBuf hread;
};

void lg_destroy(b2_context* ctx) {
  if (!ctx->lg) return;
  LightGlueState* s = ctx->lg;
  s->wblob.release();
  s->wblob_h.release();
  s->wblob_l.release();
  s->errflag.release();
  for (auto& sd : s->side) {
    DevBuf* bufs[] = {&sd.x[0], &sd.x[1], &sd.xs[0], &sd.xs[1], &sd.hs, &sd.qkv, &sd.q, &sd.k, &sd.v, &sd.ctx, &sd.msg, &sd.h, &sd.cs[0], &sd.cs[1],
                      &sd.sn[0], &sd.sn[1], &sd.ind[0], &sd.ind[1], &sd.conf, &sd.mat, &sd.src, &sd.md, &sd.rmax,
                      &sd.rlog, &sd.ls, &sd.lsg, &sd.amax, &sd.aidx};
    for (DevBuf* b : bufs) b->release();
  }
  s->sim.release();
  for (int i = 0; i < 2; ++i) s->attn_part[i].release(), s->attn_ml[i].release();
  s->counters.release();
  s->bbox.release();
  s->outm.release();
  s->outs.release();
  s->hread.release();
  delete s;
  ctx->lg = nullptr;
}

// ------------------------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------------------------

// normalize_keypoints with size=None (lightglue.py:31-43) + LearnableFourierPositionalEncoding (:68-81).  Every block
// re-derives the bounding box (40 KB of keypoints, L2-resident) and then takes a grid-stride share of the n x 32
// (cos, sin) table.  Also initialises ind[n] = n.
__global__ void __launch_bounds__(1024) k_lg_posenc(const float* __restrict__ kp, int n, const float* __restrict__ wr /*[32][2]*/,
                                                     float* __restrict__ cs, float* __restrict__ sn, int* __restrict__ ind) {
  __shared__ float red[4][32];
  __shared__ float bb[4];
  float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float x = kp[2 * i], y = kp[2 * i + 1];
    mnx = fminf(mnx, x), mny = fminf(mny, y), mxx = fmaxf(mxx, x), mxy = fmaxf(mxy, y);
  }
  mnx = -warp_max(-mnx), mny = -warp_max(-mny), mxx = warp_max(mxx), mxy = warp_max(mxy);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[0][warp] = mnx, red[1][warp] = mny, red[2][warp] = mxx, red[3][warp] = mxy;
  __syncthreads();
  if (warp == 0) {
    float a = -warp_max(-red[0][lane]), b = -warp_max(-red[1][lane]), c = warp_max(red[2][lane]), d = warp_max(red[3][lane]);
    if (lane == 0) bb[0] = a, bb[1] = b, bb[2] = c, bb[3] = d;
  }
  __syncthreads();
  // size = 1 + max - min ; shift = size / 2 ; scale = max(size) / 2
  const float sx = 1.0f + bb[2] - bb[0], sy = 1.0f + bb[3] - bb[1];
  const float shx = sx / 2.0f, shy = sy / 2.0f, sc = fmaxf(sx, sy) / 2.0f;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
  for (int i = gtid; i < n * 32; i += gsz) {
    int p = i >> 5, f = i & 31;
    float x = (kp[2 * p] - shx) / sc, y = (kp[2 * p + 1] - shy) / sc;
    float pr = x * wr[2 * f] + y * wr[2 * f + 1];
    cs[i] = cosf(pr);
    sn[i] = sinf(pr);
  }
  for (int i = gtid; i < n; i += gsz) ind[i] = i;
}

// qkv [N][768] with feature (h*64 + j)*3 + {q,k,v} (lightglue.py:166-167) -> rotary on q,k (:58-65) -> [4][N][64],
// either as fp32 (SIMT attention) or split into fp16 hi / lo planes (tcgen05 attention; plane stride = 4*N*64 halves).
struct RotJob {  // one image's share of a two-image launch (blockIdx.y)
  const float *qkv, *cs, *sn;
  int n;
  size_t plane;
  void *qo, *ko, *vo;
};
template <bool SPLIT>
__global__ void __launch_bounds__(256) k_lg_split_rotary(RotJob j0, RotJob j1, int qk_unscaled) {
  const RotJob& jb = blockIdx.y ? j1 : j0;
  const float* __restrict__ qkv = jb.qkv;
  const float* __restrict__ cs = jb.cs;
  const float* __restrict__ sn = jb.sn;
  const int n = jb.n;
  const size_t plane = jb.plane;
  void* __restrict__ qo = jb.qo;
  void* __restrict__ ko = jb.ko;
  void* __restrict__ vo = jb.vo;
  int i = blockIdx.x * blockDim.x + threadIdx.x;  // over n * 4 * 32 pairs
  if (i >= n * 128) return;
  int p = i & 31, h = (i >> 5) & 3, r = i >> 7;
  const float* src = qkv + (size_t)r * 768 + (h * 64 + 2 * p) * 3;
  float q0 = src[0], k0 = src[1], v0 = src[2], q1 = src[3], k1 = src[4], v1 = src[5];
  float c = cs[r * 32 + p], s = sn[r * 32 + p];
  size_t o = ((size_t)h * n + r) * 64 + 2 * p;
  // (t * cos) + (rotate_half(t) * sin), rotate_half: (x1, x2) -> (-x2, x1)
  const float qa = __fadd_rn(__fmul_rn(q0, c), __fmul_rn(-q1, s)), qb = __fadd_rn(__fmul_rn(q1, c), __fmul_rn(q0, s));
  const float ka = __fadd_rn(__fmul_rn(k0, c), __fmul_rn(-k1, s)), kb = __fadd_rn(__fmul_rn(k1, c), __fmul_rn(k0, s));
  if (SPLIT) {
    __half* outs[3] = {static_cast<__half*>(qo), static_cast<__half*>(ko), static_cast<__half*>(vo)};
    const float va[3] = {qa, ka, v0}, vb[3] = {qb, kb, v1};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      __half ha, la, hb, lb;
      if ((qk_unscaled >> (j < 2 ? 0 : 1)) & 1) {  // bit 0: q, k (single-accumulator logits); bit 1: v (k_flash_ts)
        tc::split_h_unscaled(va[j], ha, la);
        tc::split_h_unscaled(vb[j], hb, lb);
      } else {
        tc::split_h(va[j], ha, la);
        tc::split_h(vb[j], hb, lb);
      }
      *reinterpret_cast<__half2*>(outs[j] + o) = __halves2half2(ha, hb);
      *reinterpret_cast<__half2*>(outs[j] + plane + o) = __halves2half2(la, lb);
    }
  } else {
    float *q = static_cast<float*>(qo), *k = static_cast<float*>(ko), *v = static_cast<float*>(vo);
    q[o] = qa, q[o + 1] = qb, k[o] = ka, k[o + 1] = kb, v[o] = v0, v[o + 1] = v1;
  }
}

// LayerNorm(512, eps 1e-5, affine) + exact GELU in place (lightglue.py:152-157). one warp per row.
// When `hi` is given the result is written as split fp16 planes (the next GEMM's A operand) instead of in place.
struct LnJob {  // one image's share of a two-image launch (blockIdx.y)
  float* h;
  int n;
  __half *hi, *lo;
};
__global__ void __launch_bounds__(256) k_lg_ln_gelu(LnJob j0, LnJob j1, const float* __restrict__ g, const float* __restrict__ b) {
  const LnJob& jb = blockIdx.y ? j1 : j0;
  float* __restrict__ h = jb.h;
  __half* __restrict__ hi = jb.hi;
  __half* __restrict__ lo = jb.lo;
  const int n = jb.n;
  int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= n) return;
  float4* row = reinterpret_cast<float4*>(h + (size_t)r * 512);
  float4 v[4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = row[lane + 32 * i];
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  float mean = warp_sum(s) / 512.0f;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a = v[i].x - mean, bq = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += a * a + bq * bq + c * c + d * d;
  }
  float rstd = 1.0f / sqrtf(warp_sum(q) / 512.0f + 1e-5f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c0 = (lane + 32 * i) * 4;
    float4 gg = *reinterpret_cast<const float4*>(g + c0), bb = *reinterpret_cast<const float4*>(b + c0);
    float e[4] = {(v[i].x - mean) * rstd * gg.x + bb.x, (v[i].y - mean) * rstd * gg.y + bb.y,
                  (v[i].z - mean) * rstd * gg.z + bb.z, (v[i].w - mean) * rstd * gg.w + bb.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) e[j] = 0.5f * e[j] * (1.0f + erff(e[j] * 0.70710678118654752440f));
    if (hi) {
      uint32_t h01, l01, h23, l23;
      tc::split2(e[0], e[1], h01, l01);
      tc::split2(e[2], e[3], h23, l23);
      *reinterpret_cast<uint2*>(hi + (size_t)r * 512 + c0) = make_uint2(h01, h23);
      *reinterpret_cast<uint2*>(lo + (size_t)r * 512 + c0) = make_uint2(l01, l23);
    } else {
      row[lane + 32 * i] = make_float4(e[0], e[1], e[2], e[3]);
    }
  }
}

// two per-row heads in one pass: sigmoid(w1.x + b1) and sigmoid(w2.x + b2) (token confidence :84-94, matchability
// :298-299).  Also raw logit of head 2 when zraw != null.  one warp per row.
struct HeadJob {  // one image's share of a two-image launch (blockIdx.y)
  const float* x;
  int n;
  const float *w2, *b2;  // head 2 may be off for one image only (pruning threshold on the keypoint count)
  float *o1, *o2, *zraw;
};
__global__ void __launch_bounds__(256) k_lg_rowheads(HeadJob j0, HeadJob j1, const float* __restrict__ w1,
                                                      const float* __restrict__ b1) {
  const HeadJob& jb = blockIdx.y ? j1 : j0;
  const float* __restrict__ x = jb.x;
  const float* __restrict__ w2 = jb.w2;
  const float* __restrict__ b2 = jb.b2;
  float* __restrict__ o1 = jb.o1;
  float* __restrict__ o2 = jb.o2;
  float* __restrict__ zraw = jb.zraw;
  const int n = jb.n;
  int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= n) return;
  const float4* row = reinterpret_cast<const float4*>(x + (size_t)r * 256);
  float4 a = row[lane], b = row[lane + 32];
  float s1 = 0.f, s2 = 0.f;
  if (w1) {
    float4 wa = reinterpret_cast<const float4*>(w1)[lane], wb = reinterpret_cast<const float4*>(w1)[lane + 32];
    s1 = a.x * wa.x + a.y * wa.y + a.z * wa.z + a.w * wa.w + b.x * wb.x + b.y * wb.y + b.z * wb.z + b.w * wb.w;
  }
  if (w2) {
    float4 wa = reinterpret_cast<const float4*>(w2)[lane], wb = reinterpret_cast<const float4*>(w2)[lane + 32];
    s2 = a.x * wa.x + a.y * wa.y + a.z * wa.z + a.w * wa.w + b.x * wb.x + b.y * wb.y + b.z * wb.z + b.w * wb.w;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) {
    if (w1) o1[r] = 1.0f / (1.0f + expf(-(s1 + b1[0])));
    if (w2) {
      float z = s2 + b2[0];
      if (o2) o2[r] = 1.0f / (1.0f + expf(-z));
      if (zraw) zraw[r] = z;
    }
  }
}

// pruning decision + ordered compaction map for one image (single block):
//   counters[0 + side] += #(conf < thr)            (check_if_stop numerator, lightglue.py:653-655)
//   keep = matchability > (1 - width_conf) | conf <= thr   (:636-643); src[pos] = old index; counters[2 + side] = #kept
struct PruneJob {  // one image's share of a two-image launch (blockIdx.y = side)
  const float *conf, *mat;
  int n;
  int* src;
};
__global__ void __launch_bounds__(1024) k_lg_prune_plan(PruneJob j0, PruneJob j1, float thr, float keep_thr, int* __restrict__ counters) {
  const int side = blockIdx.y;
  const PruneJob& jb = side ? j1 : j0;
  const float* __restrict__ conf = jb.conf;
  const float* __restrict__ mat = jb.mat;
  int* __restrict__ src = jb.src;
  const int n = jb.n;
  __shared__ int wtot[32];
  __shared__ int carry, unconf;
  if (threadIdx.x == 0) carry = 0, unconf = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    int i = base + threadIdx.x;
    bool keep = false, unc = false;
    if (i < n) {
      float c = conf[i];
      unc = c < thr;
      keep = (mat[i] > keep_thr) || (c <= thr);
    }
    unsigned km = __ballot_sync(0xffffffffu, keep), um = __ballot_sync(0xffffffffu, unc);
    if (lane == 0) wtot[warp] = __popc(km);
    if (lane == 0 && um) atomicAdd(&unconf, __popc(um));
    __syncthreads();
    if (warp == 0) {
      int w = wtot[lane], ws = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, ws, o);
        if (lane >= o) ws += u;
      }
      wtot[lane] = ws - w;
    }
    __syncthreads();
    int pos = carry + wtot[warp] + __popc(km & ((1u << lane) - 1));
    if (keep) src[pos] = i;
    __syncthreads();
    if (threadIdx.x == 1023) carry = pos + (keep ? 1 : 0);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    counters[side] = unconf;
    counters[2 + side] = carry;
  }
}

// gather rows by src map: x [n][256], cos/sin [n][32], ind [n]  (lightglue.py:556-566)
__global__ void __launch_bounds__(256) k_lg_gather(const int* __restrict__ src, const int* __restrict__ cnt,
                                                    const float* __restrict__ x, const float* __restrict__ cs,
                                                    const float* __restrict__ sn, const int* __restrict__ ind,
                                                    float* __restrict__ x2, float* __restrict__ cs2, float* __restrict__ sn2,
                                                    int* __restrict__ ind2, const __half* __restrict__ ph, size_t pstride,
                                                    __half* __restrict__ ph2, size_t pstride2) {
  int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= *cnt) return;
  int s = src[r];
  const float4* a = reinterpret_cast<const float4*>(x + (size_t)s * 256);
  float4* o = reinterpret_cast<float4*>(x2 + (size_t)r * 256);
  o[lane] = a[lane];
  o[lane + 32] = a[lane + 32];
  cs2[r * 32 + lane] = cs[s * 32 + lane];
  sn2[r * 32 + lane] = sn[s * 32 + lane];
  if (lane == 0) ind2[r] = ind[s];
  if (ph) {  // split planes of x travel with it: 256 halves = 32 lanes x 16 bytes per plane
    reinterpret_cast<uint4*>(ph2 + (size_t)r * 256)[lane] = reinterpret_cast<const uint4*>(ph + (size_t)s * 256)[lane];
    reinterpret_cast<uint4*>(ph2 + pstride2 + (size_t)r * 256)[lane] = reinterpret_cast<const uint4*>(ph + pstride + (size_t)s * 256)[lane];
  }
}

// log-softmax statistics of sim rows: max and log(sum(exp(x - max)))  (F.log_softmax, lightglue.py:271). warp per row.
__device__ __forceinline__ float logsigmoid(float z) {  // F.logsigmoid: min(z, 0) - log1p(exp(-|z|))
  return fminf(z, 0.f) - log1pf(expf(-fabsf(z)));
}

// Row statistics of log_softmax(sim, dim 2) (:271): warp per row.  Also tabulates logsigmoid(z) of the row's
// matchability logit so the arg-max passes do not re-evaluate it per matrix element.
__global__ void __launch_bounds__(256) k_lg_row_stats(const float* __restrict__ sim, int m, int n, float* __restrict__ rmax,
                                                       float* __restrict__ rlog, const float* __restrict__ z,
                                                       float* __restrict__ lsg) {
  int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= m) return;
  const float* row = sim + (size_t)r * n;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, row[j]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s += expf(row[j] - mx);
  s = warp_sum(s);
  if (lane == 0) rmax[r] = mx, rlog[r] = logf(s), lsg[r] = logsigmoid(z[r]);
}
// same along columns (log_softmax of sim^T, :272): a block owns 32 columns; its 32 warps stride over the rows (each row
// access is one coalesced 128-byte line, four independent loads in flight per warp - the pass is load-latency bound),
// keeping an online (max, sum) pair that is merged across warps at the end.
constexpr int LG_COL_WARPS = 32, LG_COL_UNROLL = 4;
__global__ void __launch_bounds__(1024) k_lg_col_stats(const float* __restrict__ sim, int m, int n, float* __restrict__ cmax,
                                                        float* __restrict__ clog, const float* __restrict__ z,
                                                        float* __restrict__ lsg) {
  __shared__ float sm[LG_COL_WARPS][32], ss[LG_COL_WARPS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  float mx = -INFINITY, s = 0.f;
  if (j < n) {
    for (int i0 = warp; i0 < m; i0 += LG_COL_WARPS * LG_COL_UNROLL) {
      float x[LG_COL_UNROLL];
#pragma unroll
      for (int u = 0; u < LG_COL_UNROLL; ++u) {
        const int i = i0 + u * LG_COL_WARPS;
        x[u] = i < m ? sim[(size_t)i * n + j] : -INFINITY;
      }
#pragma unroll
      for (int u = 0; u < LG_COL_UNROLL; ++u) {
        if (x[u] > mx) {
          s = s * expf(mx - x[u]) + 1.0f;
          mx = x[u];
        } else if (x[u] > -INFINITY) {
          s += expf(x[u] - mx);
        }
      }
    }
  }
  sm[warp][lane] = mx, ss[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && j < n) {
    float M = sm[0][lane];
    for (int w = 1; w < LG_COL_WARPS; ++w) M = fmaxf(M, sm[w][lane]);
    float S = 0.f;
    for (int w = 0; w < LG_COL_WARPS; ++w)
      if (ss[w][lane] > 0.f) S += ss[w][lane] * expf(sm[w][lane] - M);
    cmax[j] = M, clog[j] = logf(S), lsg[j] = logsigmoid(z[j]);
  }
}

// scores[i][j] = (log_softmax_rows + log_softmax_cols) + (logsigmoid(z0_i) + logsigmoid(z1_j))  (:269-274);
// row arg-max (first maximum) per i. warp per row.  l0 / l1 = the tabulated logsigmoid terms.
__global__ void __launch_bounds__(256) k_lg_row_argmax(const float* __restrict__ sim, int m, int n,
                                                        const float* __restrict__ rmax, const float* __restrict__ rlog,
                                                        const float* __restrict__ cmax, const float* __restrict__ clog,
                                                        const float* __restrict__ l0g, const float* __restrict__ l1g,
                                                        float* __restrict__ best, int* __restrict__ arg) {
  int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= m) return;
  const float* row = sim + (size_t)r * n;
  const float rm = rmax[r], rl = rlog[r], l0 = l0g[r];
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = lane; j < n; j += 32) {
    float x = row[j];
    float sc = (((x - rm) - rl) + ((x - cmax[j]) - clog[j])) + (l0 + l1g[j]);
    if (sc > bv) bv = sc, bi = j;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
  }
  if (lane == 0) best[r] = bv, arg[r] = bi;
}
__global__ void __launch_bounds__(1024) k_lg_col_argmax(const float* __restrict__ sim, int m, int n,
                                                         const float* __restrict__ rmax, const float* __restrict__ rlog,
                                                         const float* __restrict__ cmax, const float* __restrict__ clog,
                                                         const float* __restrict__ l0g, const float* __restrict__ l1g,
                                                         int* __restrict__ arg) {
  __shared__ float sv[LG_COL_WARPS][32];
  __shared__ int si[LG_COL_WARPS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  if (j < n) {
    const float cm = cmax[j], cl = clog[j], l1 = l1g[j];
    for (int i0 = warp; i0 < m; i0 += LG_COL_WARPS * LG_COL_UNROLL) {
      float x[LG_COL_UNROLL], rm[LG_COL_UNROLL], rl[LG_COL_UNROLL], l0[LG_COL_UNROLL];
#pragma unroll
      for (int u = 0; u < LG_COL_UNROLL; ++u) {
        const int i = i0 + u * LG_COL_WARPS;
        const bool in = i < m;
        x[u] = in ? sim[(size_t)i * n + j] : 0.f;
        rm[u] = in ? rmax[i] : 0.f, rl[u] = in ? rlog[i] : 0.f, l0[u] = in ? l0g[i] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < LG_COL_UNROLL; ++u) {
        const int i = i0 + u * LG_COL_WARPS;
        if (i < m) {
          const float sc = (((x[u] - rm[u]) - rl[u]) + ((x[u] - cm) - cl)) + (l0[u] + l1);
          if (sc > bv) bv = sc, bi = i;  // rows ascend within a warp: first maximum kept
        }
      }
    }
  }
  sv[warp][lane] = bv, si[warp][lane] = bi;
  __syncthreads();
  if (warp == 0 && j < n) {
    for (int w = 1; w < LG_COL_WARPS; ++w) {
      float ov = sv[w][lane];
      int oi = si[w][lane];
      if (ov > bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
    }
    arg[j] = bi == 0x7fffffff ? 0 : bi;
  }
}

// filter_matches (:302-318) + index mapping through ind0 / ind1 (:598-602) + ordered compaction (single block).
__global__ void __launch_bounds__(1024) k_lg_filter(const float* __restrict__ best0, const int* __restrict__ a0,
                                                     const int* __restrict__ a1, int m, float th, const int* __restrict__ ind0,
                                                     const int* __restrict__ ind1, long long* __restrict__ out,
                                                     float* __restrict__ outs, int* __restrict__ count) {
  __shared__ int wtot[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < m; base += 1024) {
    int i = base + threadIdx.x;
    bool valid = false;
    float ms = 0.f;
    int j = 0;
    if (i < m) {
      j = a0[i];
      bool mutual = a1[j] == i;
      ms = mutual ? expf(best0[i]) : 0.f;
      valid = mutual && ms > th;
    }
    unsigned vm = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) wtot[warp] = __popc(vm);
    __syncthreads();
    if (warp == 0) {
      int w = wtot[lane], ws = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, ws, o);
        if (lane >= o) ws += u;
      }
      wtot[lane] = ws - w;
    }
    __syncthreads();
    int pos = carry + wtot[warp] + __popc(vm & ((1u << lane) - 1));
    if (valid) {
      out[2 * (size_t)pos] = ind0[i];
      out[2 * (size_t)pos + 1] = ind1[j];
      if (outs) outs[pos] = ms;
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry = pos + (valid ? 1 : 0);
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry;
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------

extern "C" int b2_lightglue_set_weights(b2_context* ctx, const float* blob, size_t n_floats) {
  if (!ctx || !blob) return B2_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n_floats != LG_NFLOATS)
    return b2_fail(ctx, B2_ERR_ARG, "lightglue blob must hold 11851601 floats, got " + std::to_string(n_floats));
  cudaSetDevice(ctx->device);
  if (!ctx->lg) ctx->lg = new LightGlueState();
  LightGlueState* s = ctx->lg;
  // device copy with every tensor 256-byte aligned
  std::vector<size_t> sizes;
  sizes.push_back(64);  // posenc.Wr
  for (int i = 0; i < LG_LAYERS; ++i) {
    const size_t self_sz[] = {768 * 256, 768, 256 * 256, 256, 512 * 512, 512, 512, 512, 256 * 512, 256};
    const size_t cross_sz[] = {256 * 256, 256, 256 * 256, 256, 256 * 256, 256, 512 * 512, 512, 512, 512, 256 * 512, 256};
    for (size_t z : self_sz) sizes.push_back(z);
    for (size_t z : cross_sz) sizes.push_back(z);
  }
  for (int i = 0; i < LG_LAYERS; ++i) {
    const size_t asz[] = {256, 1, 256 * 256, 256};
    for (size_t z : asz) sizes.push_back(z);
  }
  for (int i = 0; i < LG_LAYERS - 1; ++i) sizes.push_back(256), sizes.push_back(1);
  size_t total = 0, src_total = 0;
  std::vector<size_t> doff;
  for (size_t z : sizes) {
    doff.push_back(total);
    total += (z + 63) / 64 * 64;
    src_total += z;
  }
  if (src_total != LG_NFLOATS) return b2_fail(ctx, B2_ERR_STATE, "internal lightglue layout mismatch");
  std::vector<float> host(total, 0.f);
  size_t so = 0;
  for (size_t i = 0; i < sizes.size(); ++i) {
    memcpy(host.data() + doff[i], blob + so, sizes[i] * sizeof(float));
    so += sizes[i];
  }
  B2_CUDA(ctx, s->wblob.ensure(total * sizeof(float)));
  B2_CUDA(ctx, cudaMemcpy(s->wblob.p, host.data(), total * sizeof(float), cudaMemcpyHostToDevice));
  float* base = s->wblob.as<float>();
  size_t ti = 0;
  auto next = [&]() { return base + doff[ti++]; };
  s->wr = next();
  for (int i = 0; i < LG_LAYERS; ++i) {
    SelfW& a = s->sw[i];
    a.wqkv = next(), a.bqkv = next(), a.wout = next(), a.bout = next(), a.w0 = next(), a.b0 = next(), a.lng = next(),
    a.lnb = next(), a.w3 = next(), a.b3 = next();
    CrossW& c = s->cw[i];
    c.wqk = next(), c.bqk = next(), c.wv = next(), c.bv = next(), c.wout = next(), c.bout = next(), c.w0 = next(),
    c.b0 = next(), c.lng = next(), c.lnb = next(), c.w3 = next(), c.b3 = next();
  }
  for (int i = 0; i < LG_LAYERS; ++i) {
    AssignW& a = s->aw[i];
    a.wm = next(), a.bm = next(), a.wf = next(), a.bf = next();
  }
  for (int i = 0; i < LG_LAYERS - 1; ++i) s->tw[i].w = next(), s->tw[i].b = next();
  // confidence_threshold buffer (lightglue.py:631-634) evaluated in double then stored as float32 like torch.Tensor([...])
  for (int i = 0; i < LG_LAYERS; ++i) {
    double t = 0.8 + 0.1 * exp(-4.0 * i / LG_LAYERS);
    t = t < 0 ? 0 : (t > 1 ? 1 : t);
    s->thr[i] = (float)t;
  }
  B2_CUDA(ctx, s->wblob_h.ensure(total * sizeof(__half)));
  B2_CUDA(ctx, s->wblob_l.ensure(total * sizeof(__half)));
  B2_CUDA(ctx, s->errflag.ensure(16));
  B2_CUDA(ctx, cudaMemset(s->errflag.p, 0, 16));
  B2_LAUNCH(ctx, k_split_f32, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)0, s->wblob.as<float>(), total,
            s->wblob_h.as<__half>(), s->wblob_l.as<__half>());
  B2_CHECK_LAUNCH(ctx);
  B2_CUDA(ctx, cudaDeviceSynchronize());
  B2_CUDA(ctx, cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_GEMM_SMEM));
  B2_CUDA(ctx, cudaFuncSetAttribute(k_gemm_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TM_GEMM_SMEM));
  B2_CUDA(ctx, cudaFuncSetAttribute(k_gemm_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GW_SMEM));
  B2_CUDA(ctx, cudaFuncSetAttribute(k_flash_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AT_SMEM));
  B2_CUDA(ctx, cudaFuncSetAttribute(k_flash_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AW_SMEM));
  B2_CUDA(ctx, cudaFuncSetAttribute(k_flash_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AS_SMEM));
  B2_CUDA(ctx, cudaFuncSetAttribute(k_flash_ps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AS_SMEM));
  s->use_tc = !b2_force_simt(ctx);
  B2_CUDA(ctx, cudaFuncSetAttribute(k_flash_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FA_SMEM));
  B2_CUDA(ctx, s->hread.ensure(64));
  B2_CUDA(ctx, s->counters.ensure(64));
  s->loaded = true;
  return B2_OK;
}

static inline TcWeights lg_tw(LightGlueState* s) {
  TcWeights t{s->wblob.as<float>(), s->wblob_h.as<__half>(), s->wblob_l.as<__half>(), s->errflag.as<int>(), s->use_tc};
  const char* e = getenv("B2_NO_TMA");
  t.use_tma = !(e && e[0] == '1');
  t.attn_part = s->attn_part, t.attn_ml = s->attn_ml, t.sm_count = s->persist_ctas;
  return t;
}
static int lg_linear(b2_context* ctx, cudaStream_t st, LightGlueState* s, const LinArgs& a, const LinArgs* b = nullptr) {
  return run_linear(ctx, st, lg_tw(s), a, b);
}
static int lg_flash2(b2_context* ctx, cudaStream_t st, LightGlueState* s, const FlashJob& a, const FlashJob& b, float scale) {
  return run_flash2(ctx, st, lg_tw(s), a, b, scale);
}

static int lg_side_alloc(b2_context* ctx, LgSide& sd, int n) {
  const size_t N = (size_t)(n > 0 ? n : 1);
  for (int i = 0; i < 2; ++i) {
    B2_CUDA(ctx, sd.x[i].ensure(N * 256 * 4));
    B2_CUDA(ctx, sd.xs[i].ensure(N * 256 * 4));
    B2_CUDA(ctx, sd.cs[i].ensure(N * 32 * 4));
    B2_CUDA(ctx, sd.sn[i].ensure(N * 32 * 4));
    B2_CUDA(ctx, sd.ind[i].ensure(N * 4));
  }
  B2_CUDA(ctx, sd.qkv.ensure(N * 768 * 4));
  B2_CUDA(ctx, sd.q.ensure(N * 256 * 4));
  B2_CUDA(ctx, sd.k.ensure(N * 256 * 4));
  B2_CUDA(ctx, sd.v.ensure(N * 256 * 4));
  B2_CUDA(ctx, sd.ctx.ensure(N * 256 * 4));
  B2_CUDA(ctx, sd.msg.ensure(N * 256 * 4));
  B2_CUDA(ctx, sd.h.ensure(N * 512 * 4));
  B2_CUDA(ctx, sd.hs.ensure(N * 512 * 4));
  B2_CUDA(ctx, sd.md.ensure(N * 256 * 4));
  DevBuf* small[] = {&sd.conf, &sd.mat, &sd.src, &sd.rmax, &sd.rlog, &sd.ls, &sd.lsg, &sd.amax, &sd.aidx};
  for (DevBuf* b : small) B2_CUDA(ctx, b->ensure(N * 4));
  sd.cur = 0;
  sd.n = n;
  sd.cap = n;
  return B2_OK;
}

// x + ffn(cat[x, msg])  (lightglue.py:152-157,172,228-229): Linear(512,512) -> LN -> GELU -> Linear(512,256) + x, preceded
// by the attention output projection (out_proj / to_out); both images of the pair go through each GEMM together.
static int lg_out_and_ffn(b2_context* ctx, cudaStream_t st, LightGlueState* s, const float* wout, const float* bout,
                          const float* w0, const float* b0, const float* lng, const float* lnb, const float* w3, const float* b3) {
  int rc;
  LinArgs o[2], f0[2], f3[2];
  for (int i = 0; i < 2; ++i) {
    LgSide& sd = s->side[i];
    const size_t e = (size_t)sd.cap * 256;
    float* x = sd.x[sd.cur].as<float>();
    LinArgs& a = o[i];
    a.a1f = sd.ctx.as<float>(), a.a1p = planes_of(sd.ctx, e), a.lda1 = 256, a.K1 = 256, a.w = wout, a.ldb = 256, a.bias = bout;
    a.cf = sd.msg.as<float>(), a.ldc = 256, a.cp = planes_of(sd.msg, e), a.ldch = 256, a.M = sd.n, a.N = 256;
    LinArgs& f = f0[i];
    f.a1f = x, f.a1p = planes_of(sd.xs[sd.cur], e), f.lda1 = 256, f.K1 = 256;
    f.a2f = sd.msg.as<float>(), f.a2p = planes_of(sd.msg, e), f.lda2 = 256, f.K2 = 256;
    f.w = w0, f.ldb = 512, f.bias = b0, f.cf = sd.h.as<float>(), f.ldc = 512, f.tc_want_f32 = true, f.M = sd.n, f.N = 512;
    LinArgs& c = f3[i];
    c.a1f = sd.h.as<float>(), c.a1p = planes_of(sd.hs, (size_t)sd.cap * 512), c.lda1 = 512, c.K1 = 512, c.w = w3, c.ldb = 512, c.bias = b3;
    c.resid = x, c.ldr = 256;  // in place: every element is read (as residual) and written by the same thread
    c.cf = x, c.ldc = 256, c.tc_want_f32 = true, c.cp = planes_of(sd.xs[sd.cur], e), c.ldch = 256, c.M = sd.n, c.N = 256;
  }
  if ((rc = lg_linear(ctx, st, s, o[0], &o[1]))) return rc;
  if ((rc = lg_linear(ctx, st, s, f0[0], &f0[1]))) return rc;
  {
    LnJob lj[2];
    int mx = 0;
    for (int i = 0; i < 2; ++i) {
      LgSide& sd = s->side[i];
      const Pl hs = planes_of(sd.hs, (size_t)sd.cap * 512);
      lj[i] = {sd.h.as<float>(), sd.n, s->use_tc ? hs.hi : (__half*)nullptr, s->use_tc ? hs.lo : (__half*)nullptr};
      mx = sd.n > mx ? sd.n : mx;
    }
    if (mx > 0) {
      B2_LAUNCH(ctx, k_lg_ln_gelu, dim3(cdiv(mx, 8), 2), 256, 0, st, lj[0], lj[1], lng, lnb);
      B2_CHECK_LAUNCH(ctx);
    }
  }
  return lg_linear(ctx, st, s, f3[0], &f3[1]);
}

static int lg_self_layer(b2_context* ctx, cudaStream_t st, LightGlueState* s, int layer) {
  const SelfW& w = s->sw[layer];
  int rc;
  LinArgs q[2];
  for (int i = 0; i < 2; ++i) {
    LgSide& sd = s->side[i];
    LinArgs& a = q[i];
    a.a1f = sd.x[sd.cur].as<float>(), a.a1p = planes_of(sd.xs[sd.cur], (size_t)sd.cap * 256), a.lda1 = 256, a.K1 = 256;
    a.w = w.wqkv, a.ldb = 256, a.bias = w.bqkv, a.cf = sd.qkv.as<float>(), a.ldc = 768, a.tc_want_f32 = true, a.M = sd.n, a.N = 768;
  }
  if ((rc = lg_linear(ctx, st, s, q[0], &q[1]))) return rc;
  {
    RotJob rj[2];
    int mx = 0;
    for (int i = 0; i < 2; ++i) {
      LgSide& sd = s->side[i];
      rj[i] = {sd.qkv.as<float>(), sd.cs[sd.cur].as<float>(), sd.sn[sd.cur].as<float>(), sd.n, (size_t)sd.cap * 256, sd.q.p, sd.k.p, sd.v.p};
      mx = sd.n > mx ? sd.n : mx;
    }
    if (mx > 0) {
      if (s->use_tc)
        B2_LAUNCH(ctx, k_lg_split_rotary<true>, dim3(cdiv(mx * 128, 256), 2), 256, 0, st, rj[0], rj[1],
                  (attn_qk_unscaled(lg_tw(s)) ? 1 : 0) | (attn_v_unscaled(lg_tw(s)) ? 2 : 0));
      else
        B2_LAUNCH(ctx, k_lg_split_rotary<false>, dim3(cdiv(mx * 128, 256), 2), 256, 0, st, rj[0], rj[1], 0);
      B2_CHECK_LAUNCH(ctx);
    }
  }
  LgSide &a = s->side[0], &b = s->side[1];
  FlashJob ja{&a.q, &a.k, &a.v, &a.ctx, a.n, a.n, a.cap, a.cap}, jb{&b.q, &b.k, &b.v, &b.ctx, b.n, b.n, b.cap, b.cap};
  if ((rc = lg_flash2(ctx, st, s, ja, jb, 0.125f))) return rc;
  return lg_out_and_ffn(ctx, st, s, w.wout, w.bout, w.w0, w.b0, w.lng, w.lnb, w.w3, w.b3);
}

static int lg_cross_block(b2_context* ctx, cudaStream_t st, LightGlueState* s, int layer) {
  const CrossW& w = s->cw[layer];
  int rc;
  for (int which = 0; which < 2; ++which) {  // to_qk -> sd.q, to_v -> sd.v, both head-major [4][n][64]
    LinArgs p[2];
    for (int i = 0; i < 2; ++i) {
      LgSide& sd = s->side[i];
      const size_t e = (size_t)sd.cap * 256;
      LinArgs& a = p[i];
      a.a1f = sd.x[sd.cur].as<float>(), a.a1p = planes_of(sd.xs[sd.cur], e), a.lda1 = 256, a.K1 = 256;
      a.w = which ? w.wv : w.wqk, a.ldb = 256, a.bias = which ? w.bv : w.bqk;
      DevBuf& dst = which ? sd.v : sd.q;
      a.cf = dst.as<float>(), a.cp = planes_of(dst, e), a.head_major = 1, a.M = sd.n, a.N = 256;
      a.lo_unscaled = (which == 0 ? attn_qk_unscaled(lg_tw(s)) : attn_v_unscaled(lg_tw(s))) ? 1 : 0;  // attention operands
    }
    if ((rc = lg_linear(ctx, st, s, p[0], &p[1]))) return rc;
  }
  // m0 = softmax(s * qk0 qk1^T) v1 ; m1 = softmax(s * qk1 qk0^T) v0 with s = 64^-0.5 (the reference scales each
  // operand by 64^-0.25, lightglue.py:216-221); both directions in one launch
  LgSide &a = s->side[0], &b = s->side[1];
  FlashJob ja{&a.q, &b.q, &b.v, &a.ctx, a.n, b.n, a.cap, b.cap}, jb{&b.q, &a.q, &a.v, &b.ctx, b.n, a.n, b.cap, a.cap};
  if ((rc = lg_flash2(ctx, st, s, ja, jb, 0.125f))) return rc;
  return lg_out_and_ffn(ctx, st, s, w.wout, w.bout, w.w0, w.b0, w.lng, w.lnb, w.w3, w.b3);
}

static int lg_match_impl(b2_context* ctx, const float* kp0, const float* desc0, int n0, const float* kp1,
                         const float* desc1, int n1, const b2_lightglue_params* prm, long long* out_matches,
                         float* out_scores, int* out_k, int* out_stop, cudaStream_t st) {
  LightGlueState* s = ctx->lg;
  if (!s || !s->loaded) return b2_fail(ctx, B2_ERR_STATE, "lightglue weights not set");
  s->persist_ctas = ctx->sm_count - ctx->reserve_sms > 0 ? ctx->sm_count - ctx->reserve_sms : 1;
  *out_k = 0;
  *out_stop = 1;
  if (n0 <= 0 || n1 <= 0) return B2_OK;  // lightglue.py:568-588 (no keypoints -> empty matches)
  int rc;
  const float* kps[2] = {kp0, kp1};
  const float* descs[2] = {desc0, desc1};
  const int ns[2] = {n0, n1};
  for (int i = 0; i < 2; ++i) {
    LgSide& sd = s->side[i];
    if ((rc = lg_side_alloc(ctx, sd, ns[i]))) return rc;
    B2_CUDA(ctx, cudaMemcpyAsync(sd.x[0].p, descs[i], (size_t)ns[i] * 256 * 4, cudaMemcpyDeviceToDevice, st));
    if (s->use_tc) {
      const Pl xp = planes_of(sd.xs[0], (size_t)sd.cap * 256);
      B2_LAUNCH(ctx, k_split_f32, (unsigned)cdiv(ns[i] * 256, 256), 256, 0, st, descs[i], (size_t)ns[i] * 256, xp.hi, xp.lo);
      B2_CHECK_LAUNCH(ctx);
    }
    B2_LAUNCH(ctx, k_lg_posenc, ns[i] * 32 <= 1024 ? 1 : (cdiv(ns[i] * 32, 1024) < 148 ? cdiv(ns[i] * 32, 1024) : 148), 1024, 0, st, kps[i], ns[i], s->wr, sd.cs[0].as<float>(), sd.sn[0].as<float>(), sd.ind[0].as<int>());
    B2_CHECK_LAUNCH(ctx);
  }
  const bool do_stop = prm->depth_confidence > 0.0, do_prune = prm->width_confidence > 0.0;
  const float keep_thr = (float)(1.0 - prm->width_confidence);  // scores > float32(1 - width_confidence)
  int* counters = s->counters.as<int>();
  int* hread = s->hread.as<int>();
  int layer = 0;
  for (layer = 0; layer < LG_LAYERS; ++layer) {
    LgSide &a = s->side[0], &b = s->side[1];
    if (a.n == 0 || b.n == 0) break;
    if ((rc = lg_self_layer(ctx, st, s, layer))) return rc;
    if ((rc = lg_cross_block(ctx, st, s, layer))) return rc;
    if (layer == LG_LAYERS - 1) break;
    if (!do_stop && !do_prune) continue;
    bool prune_side[2];
    HeadJob hj[2];
    PruneJob pj[2];
    int mxn = 0;
    for (int i = 0; i < 2; ++i) {
      LgSide& sd = s->side[i];
      prune_side[i] = do_prune && sd.n > prm->prune_min_kpts;
      hj[i] = {sd.x[sd.cur].as<float>(), sd.n, prune_side[i] ? s->aw[layer].wm : nullptr, prune_side[i] ? s->aw[layer].bm : nullptr,
               sd.conf.as<float>(), sd.mat.as<float>(), nullptr};
      pj[i] = {sd.conf.as<float>(), sd.mat.as<float>(), sd.n, sd.src.as<int>()};
      mxn = sd.n > mxn ? sd.n : mxn;
    }
    B2_LAUNCH(ctx, k_lg_rowheads, dim3(cdiv(mxn, 8), 2), 256, 0, st, hj[0], hj[1], do_stop ? s->tw[layer].w : nullptr,
              do_stop ? s->tw[layer].b : nullptr);
    B2_CHECK_LAUNCH(ctx);
    for (int i = 0; i < 2; ++i) {
      LgSide& sd = s->side[i];
      if (!do_stop) B2_CUDA(ctx, cudaMemsetAsync(sd.conf.p, 0, (size_t)sd.n * 4, st));  // confidences None -> never "<= thr"
      if (!prune_side[i]) B2_CUDA(ctx, cudaMemsetAsync(sd.mat.p, 0x7f, (size_t)sd.n * 4, st));  // huge positive: keep all
    }
    B2_LAUNCH(ctx, k_lg_prune_plan, dim3(1, 2), 1024, 0, st, pj[0], pj[1], do_stop ? s->thr[layer] : -1.0f, keep_thr, counters);
    B2_CHECK_LAUNCH(ctx);
    B2_CUDA(ctx, cudaMemcpyAsync(hread, counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    B2_CUDA(ctx, cudaStreamSynchronize(st));
    if (do_stop) {
      // check_if_stop (lightglue.py:645-656) in float32: 1 - (#unconfident / (m + n)) > depth_confidence
      float ratio = 1.0f - (float)(hread[0] + hread[1]) / (float)(n0 + n1);
      if (ratio > (float)prm->depth_confidence) break;
    }
    for (int i = 0; i < 2; ++i) {
      LgSide& sd = s->side[i];
      if (!prune_side[i] || hread[2 + i] == sd.n) continue;  // nothing pruned: the buffers stay as they are
      const float* x = sd.x[sd.cur].as<float>();
      {
        int nxt = sd.cur ^ 1;
        B2_LAUNCH(ctx, k_lg_gather, cdiv(sd.n, 8), 256, 0, st, sd.src.as<int>(), counters + 2 + i, x,
                  sd.cs[sd.cur].as<float>(), sd.sn[sd.cur].as<float>(), sd.ind[sd.cur].as<int>(), sd.x[nxt].as<float>(),
                  sd.cs[nxt].as<float>(), sd.sn[nxt].as<float>(), sd.ind[nxt].as<int>(),
                  s->use_tc ? sd.xs[sd.cur].as<__half>() : (const __half*)nullptr, (size_t)sd.cap * 256, sd.xs[nxt].as<__half>(),
                  (size_t)sd.cap * 256);
        B2_CHECK_LAUNCH(ctx);
      }
      sd.cur ^= 1;
      sd.n = hread[2 + i];
    }
  }
  if (layer == LG_LAYERS) layer = LG_LAYERS - 1;
  *out_stop = layer + 1;
  LgSide &a = s->side[0], &b = s->side[1];
  if (a.n == 0 || b.n == 0) return B2_OK;
  // MatchAssignment (lightglue.py:280-296) at the stopping layer
  const AssignW& aw = s->aw[layer];
  for (int i = 0; i < 2; ++i) {
    LgSide& sd = s->side[i];
    const float* x = sd.x[sd.cur].as<float>();
    LinArgs g;
    g.a1f = x, g.a1p = planes_of(sd.xs[sd.cur], (size_t)sd.cap * 256), g.lda1 = 256, g.K1 = 256, g.w = aw.wf, g.ldb = 256;
    g.bias = aw.bf, g.scale = 0.25f;  // / 256 ** 0.25
    g.cf = sd.md.as<float>(), g.ldc = 256, g.cp = planes_of(sd.md, (size_t)sd.cap * 256), g.ldch = 256, g.M = sd.n, g.N = 256;
    if ((rc = lg_linear(ctx, st, s, g))) return rc;  // (the two images may stop with different sizes: separate launches)
    {
      HeadJob hj1{x, sd.n, aw.wm, aw.bm, nullptr, nullptr, sd.ls.as<float>()};
      B2_LAUNCH(ctx, k_lg_rowheads, dim3(cdiv(sd.n, 8), 1), 256, 0, st, hj1, hj1, (const float*)nullptr, (const float*)nullptr);
    }
    B2_CHECK_LAUNCH(ctx);
  }
  B2_CUDA(ctx, s->sim.ensure((size_t)a.n * b.n * 4));
  LinArgs gs;
  gs.a1f = a.md.as<float>(), gs.a1p = planes_of(a.md, (size_t)a.cap * 256), gs.lda1 = 256, gs.K1 = 256;
  gs.bf = b.md.as<float>(), gs.bp = planes_of(b.md, (size_t)b.cap * 256), gs.ldb = 256;
  gs.cf = s->sim.as<float>(), gs.ldc = b.n, gs.tc_want_f32 = true, gs.M = a.n, gs.N = b.n;
  if ((rc = lg_linear(ctx, st, s, gs))) return rc;
  const float* sim = s->sim.as<float>();
  B2_LAUNCH(ctx, k_lg_row_stats, cdiv(a.n, 8), 256, 0, st, sim, a.n, b.n, a.rmax.as<float>(), a.rlog.as<float>(), a.ls.as<float>(),
            a.lsg.as<float>());
  B2_CHECK_LAUNCH(ctx);
  B2_LAUNCH(ctx, k_lg_col_stats, cdiv(b.n, 32), 1024, 0, st, sim, a.n, b.n, b.rmax.as<float>(), b.rlog.as<float>(), b.ls.as<float>(),
            b.lsg.as<float>());
  B2_CHECK_LAUNCH(ctx);
  B2_LAUNCH(ctx, k_lg_row_argmax, cdiv(a.n, 8), 256, 0, st, sim, a.n, b.n, a.rmax.as<float>(), a.rlog.as<float>(),
            b.rmax.as<float>(), b.rlog.as<float>(), a.lsg.as<float>(), b.lsg.as<float>(), a.amax.as<float>(), a.aidx.as<int>());
  B2_CHECK_LAUNCH(ctx);
  B2_LAUNCH(ctx, k_lg_col_argmax, cdiv(b.n, 32), 1024, 0, st, sim, a.n, b.n, a.rmax.as<float>(), a.rlog.as<float>(),
            b.rmax.as<float>(), b.rlog.as<float>(), a.lsg.as<float>(), b.lsg.as<float>(), b.aidx.as<int>());
  B2_CHECK_LAUNCH(ctx);
  B2_LAUNCH(ctx, k_lg_filter, 1, 1024, 0, st, a.amax.as<float>(), a.aidx.as<int>(), b.aidx.as<int>(), a.n,
            (float)prm->filter_threshold, a.ind[a.cur].as<int>(), b.ind[b.cur].as<int>(), out_matches, out_scores, counters + 4);
  B2_CHECK_LAUNCH(ctx);
  B2_CUDA(ctx, cudaMemcpyAsync(hread + 4, counters + 4, sizeof(int), cudaMemcpyDeviceToHost, st));
  B2_CUDA(ctx, cudaStreamSynchronize(st));
  *out_k = hread[4];
  if (s->use_tc) {
    int err = 0;
    B2_CUDA(ctx, cudaMemcpyAsync(&err, s->errflag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    B2_CUDA(ctx, cudaStreamSynchronize(st));
    if (err) return b2_fail(ctx, B2_ERR_STATE, "tcgen05 pipeline timed out on an mbarrier (kernel bug)");
  }
  ctx->debug["lg_desc0"] = {a.x[a.cur].as<float>(), (int64_t)a.n * 256};
  ctx->debug["lg_desc1"] = {b.x[b.cur].as<float>(), (int64_t)b.n * 256};
  return B2_OK;
}

extern "C" int b2_lightglue_match_dev(b2_context* ctx, const float* kp0, const float* desc0, int n0, const float* kp1,
                                      const float* desc1, int n1, const b2_lightglue_params* params, int64_t* out_matches,
                                      float* out_scores, int* out_k, int* out_stop_layer, void* stream) {
  if (!ctx || !params || !out_k || !out_stop_layer || n0 < 0 || n1 < 0) return B2_ERR_ARG;
  if (n0 > 0 && n1 > 0 && (!kp0 || !desc0 || !kp1 || !desc1 || !out_matches)) return B2_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  return lg_match_impl(ctx, kp0, desc0, n0, kp1, desc1, n1, params, (long long*)out_matches, out_scores, out_k,
                       out_stop_layer, (cudaStream_t)stream);
}


// Content hash of a host feature array: EVERY byte takes part (four interleaved 64-bit multiply-xorshift lanes over 8-byte
// words, then the tail), so an in-place edit anywhere in the array changes the signature.  ~10 GB/s on one host core.
static uint64_t b2_feat_signature(const void* host, size_t bytes) {
  const unsigned char* p = static_cast<const unsigned char*>(host);
  const uint64_t K0 = 0x9E3779B97F4A7C15ull, K1 = 0xC2B2AE3D27D4EB4Full;
  uint64_t h[4] = {K0 ^ bytes, K1 + bytes, K0 * 3 + bytes, K1 * 5 ^ bytes};
  size_t i = 0;
  for (; i + 32 <= bytes; i += 32) {
    uint64_t w[4];
    memcpy(w, p + i, 32);
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      h[l] = (h[l] ^ w[l]) * K0;
      h[l] ^= h[l] >> 29;
    }
  }
  uint64_t tail = 0;
  for (int sh = 0; i < bytes; ++i, sh += 8) {
    tail ^= (uint64_t)p[i] << (sh & 63);
    if ((sh & 63) == 56) h[0] = (h[0] ^ tail) * K1, h[0] ^= h[0] >> 31, tail = 0;
  }
  h[1] = (h[1] ^ tail) * K1;
  uint64_t r = h[0];
  for (int l = 1; l < 4; ++l) r = (r ^ (h[l] + K0 + (r << 6) + (r >> 2))) * K1, r ^= r >> 32;
  return r;
}
// Device address of a host array: from the cache when the same (pointer, size, signature) was uploaded before, else
// copied (into an LRU cache slot, or into `fallback` when the cache is off).
static int b2_upload_cached(b2_context* ctx, const void* host, size_t bytes, DevBuf* fallback, cudaStream_t st, const void** dev) {
  if (ctx->fcache_on < 0) {  // OFF unless asked for: b2_set_option("feature_cache", 1) or B2_FEATURE_CACHE=1
    const char* e = getenv("B2_FEATURE_CACHE");
    ctx->fcache_on = (e && e[0] == '1') ? 1 : 0;
  }
  if (!ctx->fcache_on || bytes < 4) {
    B2_CUDA(ctx, fallback->ensure(bytes));
    B2_CUDA(ctx, cudaMemcpyAsync(fallback->p, host, bytes, cudaMemcpyHostToDevice, st));
    ctx->h2d_bytes += bytes;
    *dev = fallback->p;
    return B2_OK;
  }
  const uint64_t sig = b2_feat_signature(host, bytes);
  FeatCacheEntry* lru = &ctx->fcache[0];
  for (FeatCacheEntry& e : ctx->fcache) {
    if (e.host == host && e.bytes == bytes && e.sig == sig) {
      e.stamp = ++ctx->fstamp;
      *dev = e.buf.p;
      return B2_OK;
    }
    if (e.stamp < lru->stamp) lru = &e;
  }
  B2_CUDA(ctx, lru->buf.ensure(bytes));  // (cudaFree inside ensure synchronises: no kernel still reads the old block)
  B2_CUDA(ctx, cudaMemcpyAsync(lru->buf.p, host, bytes, cudaMemcpyHostToDevice, st));
  ctx->h2d_bytes += bytes;
  lru->host = host, lru->bytes = bytes, lru->sig = sig, lru->stamp = ++ctx->fstamp;
  *dev = lru->buf.p;
  return B2_OK;
}

extern "C" int b2_lightglue_match_host(b2_context* ctx, const float* kp0, const float* desc0, int n0, const float* kp1,
                                       const float* desc1, int n1, const b2_lightglue_params* params,
                                       int64_t* out_matches, float* out_scores, int* out_k, int* out_stop_layer) {
  if (!ctx || !params || !out_k || !out_stop_layer || n0 < 0 || n1 < 0) return B2_ERR_ARG;
  *out_k = 0;
  *out_stop_layer = 1;
  if (n0 == 0 || n1 == 0) return B2_OK;
  if (!kp0 || !desc0 || !kp1 || !desc1 || !out_matches) return B2_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const int mk = n0 < n1 ? n0 : n1;
  B2_CUDA(ctx, ctx->stage_d[5].ensure((size_t)mk * 2 * 8));
  B2_CUDA(ctx, ctx->stage_d[6].ensure((size_t)mk * 4));
  const float* dkp[2];
  const float* ddesc[2];
  {
    const float* hk[2] = {kp0, kp1};
    const float* hd[2] = {desc0, desc1};
    const int nn[2] = {n0, n1};
    for (int i = 0; i < 2; ++i) {
      const void *a = nullptr, *b = nullptr;
      int rc2;
      if ((rc2 = b2_upload_cached(ctx, hk[i], (size_t)nn[i] * 2 * 4, &ctx->stage_d[2 * i], st, &a))) return rc2;
      if ((rc2 = b2_upload_cached(ctx, hd[i], (size_t)nn[i] * 256 * 4, &ctx->stage_d[2 * i + 1], st, &b))) return rc2;
      dkp[i] = static_cast<const float*>(a), ddesc[i] = static_cast<const float*>(b);
    }
  }
  int rc = lg_match_impl(ctx, dkp[0], ddesc[0], n0, dkp[1], ddesc[1], n1, params, ctx->stage_d[5].as<long long>(),
                         ctx->stage_d[6].as<float>(), out_k, out_stop_layer, st);
  if (rc) return rc;
  if (*out_k > 0) {
    B2_CUDA(ctx, cudaMemcpyAsync(out_matches, ctx->stage_d[5].p, (size_t)*out_k * 2 * 8, cudaMemcpyDeviceToHost, st));
    if (out_scores) B2_CUDA(ctx, cudaMemcpyAsync(out_scores, ctx->stage_d[6].p, (size_t)*out_k * 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(ctx, cudaStreamSynchronize(st));
  }
  return B2_OK;
}
