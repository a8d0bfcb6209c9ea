// Fused (query, latent)-pair BACKWARD on the tensor cores (EnfPrecision::ENF_PREC_BF16, d = 128).
//
// Three kernels, all (field, latent)-major (a CTA owns one latent (b, z) at a time and walks that field's query tiles of
// 128 rows), so per-latent operands load once, per-latent gradients reduce on chip, and weight-gradient accumulators
// live in TMEM:
//
//   A  (this file) value path, top: stream the `that` operand tiles the forward stashed -> m_h -> n_h,
//                          softmax backward (ds), LayerNorm/gelu backward -> dm_h,
//                          wgrad dW3[b,z,h] += that^T dm_h (TMEM), dgrad dthat = sum_h dm_h W3_h^T (TMEM),
//                          out: dthat (fp16), ds, dW3, db3.
//   B  (enf_pairs_tc_bwd_v.cu) value path, bottom: gamma_v -> h1v -> tpre/that; dtpre = LNbwd(dthat) gelu'(tpre);
//                          dW', dW1_v, biases; d gamma_v -> du_v.
//   C  (enf_pairs_tc_bwd_q.cu) query path: gamma_q -> h1q; dzq = scale sum_h ds_h U_h [h1q>0]; dU, dkappa, dW1_q;
//                          d gamma_q -> du; window / invariant backward -> dLam record, dsigma.
//
// TMEM (512 columns) is what forces the split: the five weight-gradient accumulators of the chain need 640
// columns.  Gradient operands are fp16, so all cotangents are pre-scaled by a power of two `gs` that brings
// max|dnbar| to 16; accumulators are multiplied by 1/gs when they are flushed.
#include "enf_pairs_tc_common.cuh"

namespace {

using namespace tcp;

template <int D, int H> struct BwdCfg {
  static constexpr int NQ = D / 32;
  static constexpr int NT = ROWS * NQ;
  static constexpr uint32_t WIMG = D * D * 2;
  static constexpr uint32_t WBLK = D * 128;
  static constexpr uint32_t ABLK = ROWS * 128;
  static constexpr uint32_t ATILE = (D / 64) * ABLK;
  static constexpr int HD = D / 2;
};

__device__ __forceinline__ void load_scale(const float* gmax, float& gs, float& inv_gs) {
  float m = *gmax;
  int e = 0;
  if (m > 0.f && isfinite(m)) { frexpf(m, &e); }       // m = f * 2^e, f in [0.5, 1)
  gs = (m > 0.f && isfinite(m)) ? ldexpf(1.f, 4 - e) : 1.f;
  inv_gs = 1.f / gs;
}

__device__ __forceinline__ void st_half32(__half* dst, const float* v) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int t = 0; t < 4; ++t) h2[t] = __floats2half2_rn(v[q * 8 + 2 * t], v[q * 8 + 2 * t + 1]);
    d4[q] = u;
  }
}
// Dg[b,c,h] = dnbar . nbar ; gmax = max |dnbar|
__global__ void __launch_bounds__(256) bwd_prep_kernel(const float* __restrict__ dnbar, const float* __restrict__ nbar,
                                                       int64_t rows, int D, float* __restrict__ Dg, float* gmax) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float mx = 0.f;
  for (int64_t r = warp; r < rows; r += nwarps) {
    float s = 0.f;
    for (int j = lane; j < D; j += 32) {
      float a = dnbar[r * D + j];
      s = fmaf(a, nbar[r * D + j], s);
      mx = fmaxf(mx, fabsf(a));
    }
    s = warp_sum(s);
    if (lane == 0) Dg[r] = s;
  }
  mx = warp_max(mx);
  if (lane == 0) atomicMax(reinterpret_cast<int*>(gmax), __float_as_int(mx));     // non-negative floats order as ints
}

// =================================================================================================
// kernel A
// =================================================================================================
// The forward stashed the fp16 operand tile of `that` for every (field, latent, query tile); A streams those tiles
// back with the bulk-copy engine (double buffered) instead of recomputing gamma_v -> h1v -> that, so a tile costs
//     G4_h   m_h = that W3_h                       (one per head; the next tile's first one is issued a tile ahead)
//     E_h    n_h, softmax backward (ds), LayerNorm / gelu backward -> dm_h (fp16 tile, one buffer per head)
//     wgrad  dW3_h += that^T dm_h (TMEM, CTA lifetime)      dgrad  dthat (+)= dm_h W3_h^T
// TMEM: two working regions + H weight-gradient accumulators.  The regions swap roles every tile: the dgrad lands in
// the region the first head's epilogue has already consumed, the region freed by the last head's epilogue receives
// the NEXT tile's first G4, so only the last head's dgrad is ever waited for.
template <int D, int H> struct ACfg {
  using B = BwdCfg<D, H>;
  static constexpr uint32_t OFF_W3 = 0;                              // [H] W3 images
  static constexpr uint32_t OFF_T = H * B::WIMG;                     // [2] that tiles
  static constexpr uint32_t OFF_DM = OFF_T + 2 * B::ATILE;           // [H] dm tiles
  static constexpr uint32_t OFF_F = OFF_DM + H * B::ATILE;
  static constexpr int F_TOTAL = H * D /*b3*/ + 2 * B::NQ * ROWS * 2 /*exchange*/ + H * D /*db3*/ + 2 * 3 * ROWS * H /*row scalars*/;
  static constexpr uint32_t SMEM_BYTES = OFF_F + F_TOTAL * 4 + 128 + 1024;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <int D, int H>
__global__ void __launch_bounds__(BwdCfg<D, H>::NT, 1) pairs_bwd_tc_a_kernel(EnfPairTcBwdParams P) {
  using C = BwdCfg<D, H>;
  using A = ACfg<D, H>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; pointer stays in the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* sW3 = base + A::OFF_W3;
  uint8_t* sT = base + A::OFF_T;
  uint8_t* sDm = base + A::OFF_DM;
  float* f = reinterpret_cast<float*>(base + A::OFF_F);
  float* s_b3 = f; f += H * D;
  float* s_exch = f; f += 2 * C::NQ * ROWS * 2;
  float* s_db3 = f; f += H * D;
  float* s_rs = f; f += 2 * 3 * ROWS * H;             // [2 tiles][logit | lse | Dg][ROWS][H], filled one tile ahead by cp.async
  uint64_t* bars = reinterpret_cast<uint64_t*>(f);
  uint64_t *bar_w = bars, *bar_t = bars + 1 /*[2]*/, *bar_g4 = bars + 3 /*[2]*/, *bar_d = bars + 5, *bar_gb = bars + 6;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 7);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lq = warp & 3, cq = warp >> 2;
  const int row = lq * 32 + lane, col0 = cq * 32;
  const int64_t bz = blockIdx.x;
  const int b = (int)(bz / P.Z), z = (int)(bz % P.Z);
  const int ntiles = (P.C + ROWS - 1) / ROWS;
  const uint8_t* timg = P.that_img + (size_t)bz * ntiles * C::ATILE;

  if (tid == 0) {
    for (int i = 0; i < 7; ++i) tc::mbar_init(bars + i, 1);
    tc::mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc<512>(s_tmem);
  for (int e = tid; e < H * D; e += C::NT) { s_b3[e] = P.b3[bz * H * D + e]; s_db3[e] = 0.f; }
  float gs, inv_gs;
  load_scale(P.gmax, gs, inv_gs);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = *s_tmem;
  const uint32_t tW3 = tm + 2 * D;
  const uint32_t lane_off = (uint32_t)(lq * 32) << 16;
  const uint32_t my_t = lane_off + col0;
  const uint32_t aW3 = tc::smem_u32(sW3), aT = tc::smem_u32(sT), aDm = tc::smem_u32(sDm);
  if (tid == 0) {
    tc::mbar_expect_tx(bar_w, H * C::WIMG);
    for (int h = 0; h < H; ++h) tc::bulk_g2s(sW3 + h * C::WIMG, P.img_W3 + (bz * H + h) * C::WIMG, C::WIMG, bar_w);
    tc::mbar_expect_tx(&bar_t[0], C::ATILE);
    tc::bulk_g2s(sT, timg, C::ATILE, &bar_t[0]);
  }
  int xw = 0;
  // per-row scalars of tile `ct` (logits, log-sum-exp, Dg of my row, all heads) -> shared, without passing through registers
  auto prefetch_rows = [&](int ct) {
    const int c = ct * ROWS + row;
    if (cq == 0 && c < P.C) {
      const int64_t q = (int64_t)b * P.C + c;
      float* dst = s_rs + (ct & 1) * 3 * ROWS * H + row * H;
      tc::cp_async<4 * H>(dst, P.slog + (q * P.Z + z) * H);
      tc::cp_async<4 * H>(dst + ROWS * H, P.lse + q * H);
      tc::cp_async<4 * H>(dst + 2 * ROWS * H, P.Dg + q * H);
    }
  };
  prefetch_rows(0);
  tc::cp_async_wait_all();
  __syncthreads();

  for (int ct = 0; ct < ntiles; ++ct) {
    const uint32_t par = ct & 1;
    const int e = ct & 1;                              // that buffer of this tile; region roles: first = tR[e], other = tR[e ^ 1]
    const uint32_t tF = tm + e * D, tS = tm + (e ^ 1) * D;
    const uint32_t aTc = aT + e * C::ATILE;
    const int c0 = ct * ROWS;
    const bool valid = c0 + row < P.C;
    const int64_t bc = (int64_t)b * P.C + c0 + row;
    if (ct + 1 < ntiles) prefetch_rows(ct + 1);
    if (ct > 0) {                                      // every MMA of the previous tile is done with that[e ^ 1] and the dm tiles
      tc::mbar_wait(bar_gb, par ^ 1);
      tc::tc_fence_after();
    }
    if (tid == 0) {
      if (ct + 1 < ntiles) {
        tc::mbar_expect_tx(&bar_t[e ^ 1], C::ATILE);
        tc::bulk_g2s(sT + (e ^ 1) * C::ATILE, timg + (size_t)(ct + 1) * C::ATILE, C::ATILE, &bar_t[e ^ 1]);
      }
      if (ct == 0) {
        tc::mbar_wait(bar_w, 0);
        tc::mbar_wait(&bar_t[0], 0);
        tc::tc_fence_after();
        issue_gemm<D>(tF, aTc, aW3, C::ABLK, C::WBLK);
        tc::mma_commit(&bar_g4[0]);
      }
      if (H > 1) {
        issue_gemm<D>(tS, aTc, aW3 + C::WIMG, C::ABLK, C::WBLK);
        tc::mma_commit(&bar_g4[1]);
      }
    }
    float v[32];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float att = 0.f, Dh = 0.f;
      if (valid) {
        const float* rs = s_rs + e * 3 * ROWS * H + row * H + h;
        att = __expf(rs[0] - rs[ROWS * H]);
        Dh = rs[2 * ROWS * H];
      }
      const float atts = att * gs;                     // cotangents are carried scaled by gs
      float dnb[32];
      {
        const float4* src = reinterpret_cast<const float4*>(P.dnbar + (bc * H + h) * D + col0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 a = valid ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          dnb[4 * q] = a.x; dnb[4 * q + 1] = a.y; dnb[4 * q + 2] = a.z; dnb[4 * q + 3] = a.w;
        }
      }
      tc::mbar_wait(&bar_g4[h], par);
      tc::tc_fence_after();
      tc::tmem_ld32((h == 0 ? tF : tS) + my_t, v);
      tc::tmem_ld_wait();
      float dg[32];
      float st[2] = {0.f, 0.f};
#pragma unroll
      for (int j4 = 0; j4 < 32; j4 += 4) {
        const float4 bb = *reinterpret_cast<const float4*>(s_b3 + h * D + col0 + j4);
        const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float g;
          gelu_fast_both(v[j4 + t] + bv[t], g, dg[j4 + t]);
          v[j4 + t] = g; st[0] += g; st[1] = fmaf(g, g, st[1]);
        }
      }
      row_exchange<C::NQ, 2>(s_exch, xw, cq, row, lq, st);
      const float mu = st[0] * (1.f / D);
      const float rstd = rsqrtf(fmaxf(st[1] * (1.f / D) - mu * mu, 0.f) + 1e-6f);
      const float nm = -mu * rstd;
      float dd[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = fmaf(v[j], rstd, nm);               // n
        dd[0] = fmaf(dnb[j], v[j], dd[0]);
        dd[1] += dnb[j];
      }
      row_exchange<C::NQ, 2>(s_exch, xw, cq, row, lq, dd);
      const float dsh = atts * (dd[0] - Dh);
      const float mean1 = atts * dd[1] * (1.f / D), mean2 = atts * dd[0] * (1.f / D);
#pragma unroll
      for (int j = 0; j < 32; ++j) dnb[j] = (fmaf(atts, dnb[j], -mean1) - v[j] * mean2) * (rstd * dg[j]);     // dm
      uint8_t* sDh = sDm + h * C::ATILE;
#pragma unroll
      for (int c8 = 0; c8 < 32; c8 += 8) tc::st_row8_bf16(sDh, C::ABLK, row, col0 + c8, dnb + c8);
      {
        float cs = warp_colsum32(dnb, lane);
        atomicAdd(&s_db3[h * D + col0 + lane], cs);
      }
      if (cq == 0 && valid) P.ds[((bz * P.C) + c0 + row) * H + h] = dsh;
      tc::tc_fence_before();
      tc::fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        tc::tc_fence_after();
        const uint32_t aDh = aDm + h * C::ATILE, aWh = aW3 + h * C::WIMG;
        if (h + 1 < H) {
          issue_wgrad<D>(tW3 + h * D, aTc, aDh, C::ABLK, ct > 0);
          issue_dgrad<D>(tF, aDh, aWh, C::ABLK, C::WBLK, h > 0);
        } else {
          issue_dgrad<D>(tF, aDh, aWh, C::ABLK, C::WBLK, h > 0);
          tc::mma_commit(bar_d);
          issue_wgrad<D>(tW3 + h * D, aTc, aDh, C::ABLK, ct > 0);
          tc::mma_commit(bar_gb);
          if (ct + 1 < ntiles) {                       // tS is consumed: the next tile's first G4 goes there now
            tc::mbar_wait(&bar_t[e ^ 1], ((ct + 1) >> 1) & 1);
            tc::tc_fence_after();
            issue_gemm<D>(tS, aT + (e ^ 1) * C::ATILE, aW3, C::ABLK, C::WBLK);
            tc::mma_commit(&bar_g4[0]);
          }
        }
      }
    }
    // dthat = sum_h dm_h W3_h^T -> global (fp16, scaled)
    tc::mbar_wait(bar_d, par);
    tc::tc_fence_after();
    tc::tmem_ld32(tF + my_t, v);
    tc::tmem_ld_wait();
    if (valid) st_half32(P.dthat + ((bz * P.C) + c0 + row) * D + col0, v);
    tc::cp_async_wait_all();                  // next tile's row scalars have landed (visible to all after the barrier)
    tc::tc_fence_before();
    __syncthreads();                          // tF has been read by everyone: the next tile's second G4 may overwrite it
  }
  // flush dW3[b,z,h] (lane = input feature, column = output feature) and db3
  tc::mbar_wait(bar_gb, (ntiles - 1) & 1);
  tc::tc_fence_after();
#pragma unroll
  for (int h = 0; h < H; ++h) {
    float w[32];
    tc::tmem_ld32(tW3 + h * D + my_t, w);
    tc::tmem_ld_wait();
    float* o = P.g_W3 + ((bz * H + h) * D + row) * D + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *reinterpret_cast<float4*>(o + j) = make_float4(w[j] * inv_gs, w[j + 1] * inv_gs, w[j + 2] * inv_gs, w[j + 3] * inv_gs);
  }
  for (int e = tid; e < H * D; e += C::NT) P.g_b3[bz * H * D + e] = s_db3[e] * inv_gs;
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tm);
}

template <int D, int H>
int launch_bwd(cudaStream_t st, const EnfPairTcBwdParams& p) {
  using C = BwdCfg<D, H>;
  const int64_t BC = (int64_t)p.B * p.C;
  int blocks = (int)((BC * H * 32 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  bwd_prep_kernel<<<blocks, 256, 0, st>>>(p.dnbar, p.nbar, BC * H, D, const_cast<float*>(p.Dg), const_cast<float*>(p.gmax));
  size_t smem_a = ACfg<D, H>::SMEM_BYTES;
  if (cudaFuncSetAttribute(pairs_bwd_tc_a_kernel<D, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a) != cudaSuccess) return -1;
  const unsigned grid = (unsigned)(p.B * p.Z);
  pairs_bwd_tc_a_kernel<D, H><<<grid, C::NT, smem_a, st>>>(p);
  if (enf_launch_pairs_bwd_tc_v(st, D, p) < 0) return -1;
  if (enf_launch_pairs_bwd_tc_q(st, D, H, p) < 0) return -1;
  return 4;
}

}  // namespace

bool enf_pairs_bwd_tc_supported(int d, int H) { return d == 128 && (H == 1 || H == 2); }

int enf_launch_pairs_bwd_tc(cudaStream_t st, int d, int H, const EnfPairTcBwdParams& p) {
  if (d == 128 && H == 2) return launch_bwd<128, 2>(st, p);
  if (d == 128 && H == 1) return launch_bwd<128, 1>(st, p);
  return -1;
}
