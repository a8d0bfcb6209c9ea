// Fused (query, latent)-pair BACKWARD on the tensor cores (EnfPrecision::ENF_PREC_BF16, d = 128).
//
// Two kernels, both (field, latent)-major: one CTA owns one latent (b, z) and walks that field's query tiles
// (128 rows each), so every per-latent operand (W3[b,z,h] images, U, b3, the pose record) is loaded once per
// CTA, per-latent gradients reduce on chip, and weight-gradient accumulators live in TMEM for the CTA's life:
//
//   A  value path, top:    stream the `that` operand tiles the forward stashed -> m_h -> n_h,
//                          softmax backward (ds), LayerNorm/gelu backward -> dm_h,
//                          wgrad dW3[b,z,h] += that^T dm_h (TMEM), dgrad dthat = sum_h dm_h W3_h^T (TMEM),
//                          out: dthat (fp16), ds, dW3, db3.
//   B  value path, bottom + query path:
//                          recompute gamma_v -> h1v -> tpre/that;  dtpre = LNbwd(dthat) gelu'(tpre);
//                          wgrad dW' , dgrad -> dzv = (.)[h1v>0];  wgrad dW1_v, dgrad -> d gamma_v -> du;
//                          gamma_q -> h1q; dzq = scale sum_h ds_h U_h [h1q>0]; dU, dkappa; wgrad dW1_q,
//                          dgrad -> d gamma_q -> du; window/invariant backward -> dLam record, dsigma.
//
// TMEM (512 columns) is what forces the split: the five weight-gradient accumulators of the chain need 640
// columns.  A holds dW3_0, dW3_1 (256) + two working tiles; B holds dW', dW1_v, dW1_q (384) + one working tile.
// Gradient operands are fp16, so all cotangents are pre-scaled by a power of two `gs` that brings max|dnbar| to 16;
// accumulators are multiplied by 1/gs when they are flushed.
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

__device__ __forceinline__ void ld_half32(const __half* src, float* out) {
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u = __ldg(s4 + q);
    const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int t = 0; t < 4; ++t) { float2 f = __half22float2(h2[t]); out[q * 8 + 2 * t] = f.x; out[q * 8 + 2 * t + 1] = f.y; }
  }
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
// read one 16-bit operand element back from a swizzled activation tile
__device__ __forceinline__ float tile_elem(const uint8_t* tile, uint32_t ablk, int row, int col) {
  const uint8_t* p = tile + (col >> 6) * ablk + tc::swz_chunk_off(row, (col & 63) >> 3) + (col & 7) * 2;
  return __half2float(*reinterpret_cast<const __half*>(p));
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

// =================================================================================================
// kernel B
// =================================================================================================
template <int D, int H, bool QPATH>
__global__ void __launch_bounds__(BwdCfg<D, H>::NT, 1) pairs_bwd_tc_b_kernel(EnfPairTcBwdParams P) {
  using C = BwdCfg<D, H>;
  constexpr int HD = C::HD;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; pointer stays in the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* sW1q = base;
  uint8_t* sW1v = base + C::WIMG;
  uint8_t* sWp = base + 2 * C::WIMG;
  uint8_t* sB0 = base + 3 * C::WIMG;
  uint8_t* sB1 = sB0 + C::ATILE;
  uint8_t* sB2 = sB1 + C::ATILE;
  float* f = reinterpret_cast<float*>(sB2 + C::ATILE);
  float* s_xi = f; f += ROWS * 8;
  float* s_lam = f; f += 64;
  float* s_uz = f; f += H * D;
  float* s_bias = f; f += 3 * D;                      // b1q | b1v | bp
  float* s_om = f; f += 2 * 6 * HD;                   // omega_q | omega_v (x 2 pi)
  float* s_exch = f; f += 2 * C::NQ * ROWS * 2;
  float* s_du = f; f += ROWS * 8;
  float* s_dq = f; f += ROWS * 8;
  float* s_dsg = f; f += ROWS;
  float* s_dU = f; f += H * D;
  float* s_db = f; f += 3 * D;                        // db1q | db1v | dbp
  float* s_dlam = f; f += 64;
  float* s_misc = f; f += 8;                          // dkappa[H], dsigma at [4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(f);
  uint64_t* bar_w = bars;
  uint64_t* bar_m = bars + 1;                         // [6]
  uint64_t* bar_lo = bars + 7;                        // [2]: W1_v low image -> B1, W1_q low image -> B2
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 10);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lq = warp & 3, cq = warp >> 2;
  const int row = lq * 32 + lane, col0 = cq * 32;
  const int64_t bz = blockIdx.x;
  const int b = (int)(bz / P.Z);
  const float scale = rsqrtf((float)D);

  if (tid == 0) {
    for (int i = 0; i < 10; ++i) tc::mbar_init(bars + i, 1);
    tc::mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc<512>(s_tmem);
  if (tid < ENF_LAM_SIZE) { s_lam[tid] = P.lam[bz * ENF_LAM_SIZE + tid]; s_dlam[tid] = 0.f; }
  if (tid < 8) s_misc[tid] = 0.f;
  for (int e = tid; e < D; e += C::NT) { s_bias[e] = P.q_b1[e]; s_bias[D + e] = P.v_b1[e]; s_bias[2 * D + e] = P.bp[e]; }
  for (int e = tid; e < 3 * D; e += C::NT) s_db[e] = 0.f;
  for (int e = tid; e < H * D; e += C::NT) { s_uz[e] = P.U[bz * H * D + e]; s_dU[e] = 0.f; }
  for (int e = tid; e < 6 * HD; e += C::NT) {
    s_om[e] = e < P.I * HD ? 6.283185307179586f * P.q_omega[e] : 0.f;
    s_om[6 * HD + e] = e < P.I * HD ? 6.283185307179586f * P.v_omega[e] : 0.f;
  }
  float gs, inv_gs;
  load_scale(P.gmax, gs, inv_gs);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = *s_tmem;
  const uint32_t tT = tm, tWp = tm + D, tW1v = tm + 2 * D, tW1q = tm + 3 * D;
  const uint32_t my_t = ((uint32_t)(lq * 32) << 16) + col0;
  if (tid == 0) {
    tc::mbar_expect_tx(bar_w, 3 * C::WIMG);
    tc::bulk_g2s(sW1q, P.img_q_w1, C::WIMG, bar_w);
    tc::bulk_g2s(sW1v, P.img_v_w1, C::WIMG, bar_w);
    tc::bulk_g2s(sWp, P.img_Wp, C::WIMG, bar_w);
    tc::mbar_expect_tx(&bar_lo[0], C::WIMG);
    tc::bulk_g2s(sB1, P.img_v_w1_lo, C::WIMG, &bar_lo[0]);
  }
  const uint32_t aB0 = tc::smem_u32(sB0), aB1 = tc::smem_u32(sB1), aB2 = tc::smem_u32(sB2);
  const uint32_t aW1q = tc::smem_u32(sW1q), aW1v = tc::smem_u32(sW1v), aWp = tc::smem_u32(sWp);
  int xw = 0;
  const int ntiles = (P.C + ROWS - 1) / ROWS;
  const float sigma = P.sigma ? P.sigma[bz] : 1.f;
  const bool sin_cols = col0 < HD;                    // my 32 columns are sin (true) or cos (false) features
  const int jj0 = sin_cols ? col0 : col0 - HD;

  for (int ct = 0; ct < ntiles; ++ct) {
    const uint32_t par = ct & 1;
    const int c0 = ct * ROWS;
    const bool valid = c0 + row < P.C;
    const int64_t pr = bz * P.C + c0 + row;              // (b, z, c) pair index
    for (int e = tid; e < ROWS * 8; e += C::NT) {
      int r = e >> 3;
      s_xi[e] = (c0 + r < P.C) ? P.xi[(int64_t)b * P.xi_bs + (int64_t)(c0 + r) * 8 + (e & 7)] : 0.f;
      s_du[e] = 0.f;
    }
    __syncthreads();
    const Rec rec = pair_record(P, s_lam, s_xi + row * 8, sigma);
    float v[32];
    // ---------------- value path, bottom --------------------------------------------------------------
    // the two relu layers are evaluated with a two-term 16-bit split (hi*hi + lo*hi + hi*lo): their masks decide
    // whole gradient entries, and fp16-level noise in the pre-activation flips enough of them to matter
    rff_to_tile_split<D>(rec, P.I, s_om + 6 * HD, sB0, sB2, C::ABLK, row, col0);
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      if (ct == 0) tc::mbar_wait(bar_w, 0);
      tc::mbar_wait(&bar_lo[0], par);                 // W1_v low image has landed in B1
      tc::tc_fence_after();
      issue_gemm<D>(tT, aB0, aW1v, C::ABLK, C::WBLK);
      issue_gemm<D>(tT, aB2, aW1v, C::ABLK, C::WBLK, 1);
      issue_gemm<D>(tT, aB0, aB1, C::ABLK, C::WBLK, 1);
      tc::mma_commit(&bar_m[0]);
    }
    tc::mbar_wait(&bar_m[0], par);
    tc::tc_fence_after();
    tc::tmem_ld32(tT + my_t, v);
    tc::tmem_ld_wait();
    uint32_t mask = 0;
#pragma unroll
    for (int c8 = 0; c8 < 32; c8 += 8) {
      float o[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        o[t] = fmaxf(v[c8 + t] + s_bias[D + col0 + c8 + t], 0.f);
        mask |= (o[t] > 0.f ? 1u : 0u) << (c8 + t);
      }
      tc::st_row8_bf16(sB1, C::ABLK, row, col0 + c8, o);
    }
    tc::tc_fence_before();
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      issue_gemm<D>(tT, aB1, aWp, C::ABLK, C::WBLK);
      tc::mma_commit(&bar_m[1]);
    }
    tc::mbar_wait(&bar_m[1], par);
    tc::tc_fence_after();
    tc::tmem_ld32(tT + my_t, v);
    tc::tmem_ld_wait();
    {
      float dg[32];
      float st[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float g;
        gelu_fast_both(v[j] + s_bias[2 * D + col0 + j], g, dg[j]);
        v[j] = g; st[0] += g; st[1] = fmaf(g, g, st[1]);
      }
      row_exchange<C::NQ, 2>(s_exch, xw, cq, row, lq, st);
      const float mu = st[0] * (1.f / D);
      const float rstd = rsqrtf(fmaxf(st[1] * (1.f / D) - mu * mu, 0.f) + 1e-6f);
      const float nm = -mu * rstd;
      float dth[32];
      if (valid) ld_half32(P.dthat + pr * D + col0, dth);
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) dth[j] = 0.f;
      }
      float dd[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = fmaf(v[j], rstd, nm);               // that
        dd[0] += dth[j];
        dd[1] = fmaf(dth[j], v[j], dd[1]);
      }
      row_exchange<C::NQ, 2>(s_exch, xw, cq, row, lq, dd);
      const float m1 = dd[0] * (1.f / D), m2 = dd[1] * (1.f / D);
#pragma unroll
      for (int j = 0; j < 32; ++j) dth[j] = rstd * (dth[j] - m1 - v[j] * m2) * dg[j];      // dtpre
#pragma unroll
      for (int c8 = 0; c8 < 32; c8 += 8) tc::st_row8_bf16(sB2, C::ABLK, row, col0 + c8, dth + c8);
      float cs = warp_colsum32(dth, lane);
      atomicAdd(&s_db[2 * D + col0 + lane], cs);
    }
    tc::tc_fence_before();
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      issue_wgrad<D>(tWp, aB1, aB2, C::ABLK, ct > 0);
      issue_dgrad<D>(tT, aB2, aWp, C::ABLK, C::WBLK, 0);
      tc::mma_commit(&bar_m[2]);
    }
    tc::mbar_wait(&bar_m[2], par);
    tc::tc_fence_after();
    if (QPATH && tid == 0) {                           // B2 (dtpre) is free again: fetch the W1_q low image into it
      tc::mbar_expect_tx(&bar_lo[1], C::WIMG);
      tc::bulk_g2s(sB2, P.img_q_w1_lo, C::WIMG, &bar_lo[1]);
    }
    tc::tmem_ld32(tT + my_t, v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = ((mask >> j) & 1u) ? v[j] : 0.f;              // dzv
#pragma unroll
    for (int c8 = 0; c8 < 32; c8 += 8) tc::st_row8_bf16(sB1, C::ABLK, row, col0 + c8, v + c8);
    {
      float cs = warp_colsum32(v, lane);
      atomicAdd(&s_db[D + col0 + lane], cs);
    }
    tc::tc_fence_before();
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      issue_wgrad<D>(tW1v, aB0, aB1, C::ABLK, ct > 0);
      issue_dgrad<D>(tT, aB1, aW1v, C::ABLK, C::WBLK, 0);
      tc::mma_commit(&bar_m[3]);
    }
    tc::mbar_wait(&bar_m[3], par);
    tc::tc_fence_after();
    tc::tmem_ld32(tT + my_t, v);                       // d gamma_v, my 32 columns
    tc::tmem_ld_wait();
    {
      float du[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        // d proj_jj += cos * dsin   (sin columns)   or   -sin * dcos   (cos columns)
        float partner = tile_elem(sB0, C::ABLK, row, sin_cols ? col0 + j + HD : col0 + j - HD);
        float dproj = sin_cols ? partner * v[j] : -partner * v[j];
#pragma unroll
        for (int i = 0; i < 6; ++i) if (i < P.I) du[i] = fmaf(dproj, s_om[6 * HD + i * HD + jj0 + j], du[i]);
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) if (i < P.I) atomicAdd(&s_du[row * 8 + i], du[i]);
    }
    tc::tc_fence_before();
    __syncthreads();                                   // everyone is done reading gamma_v from B0
    if (!QPATH) {                                      // kernel C (enf_pairs_tc_bwd_q.cu) takes it from here
      if (cq == 0 && valid) {
        float4* dst = reinterpret_cast<float4*>(P.duv + pr * 8);
        dst[0] = make_float4(s_du[row * 8], s_du[row * 8 + 1], s_du[row * 8 + 2], s_du[row * 8 + 3]);
        dst[1] = make_float4(s_du[row * 8 + 4], s_du[row * 8 + 5], s_du[row * 8 + 6], s_du[row * 8 + 7]);
      }
      if (tid == 0 && ct + 1 < ntiles) {               // B1 (dzv) is free: every MMA reading it has completed
        tc::mbar_expect_tx(&bar_lo[0], C::WIMG);
        tc::bulk_g2s(sB1, P.img_v_w1_lo, C::WIMG, &bar_lo[0]);
      }
      __syncthreads();
      continue;
    }
    // ---------------- query path -----------------------------------------------------------------------------
    rff_to_tile_split<D>(rec, P.I, s_om, sB0, sB1, C::ABLK, row, col0);      // B1 (dzv) is free: all its MMAs completed
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::mbar_wait(&bar_lo[1], par);
      tc::tc_fence_after();
      issue_gemm<D>(tT, aB0, aW1q, C::ABLK, C::WBLK);
      issue_gemm<D>(tT, aB1, aW1q, C::ABLK, C::WBLK, 1);
      issue_gemm<D>(tT, aB0, aB2, C::ABLK, C::WBLK, 1);
      tc::mma_commit(&bar_m[4]);
    }
    float dsv[H];
#pragma unroll
    for (int h = 0; h < H; ++h) dsv[h] = valid ? P.ds[pr * H + h] : 0.f;
    tc::mbar_wait(&bar_m[4], par);
    tc::tc_fence_after();
    if (tid == 0 && ct + 1 < ntiles) {                 // B1 is free until the next tile's E2: prefetch the W1_v low image
      tc::mbar_expect_tx(&bar_lo[0], C::WIMG);
      tc::bulk_g2s(sB1, P.img_v_w1_lo, C::WIMG, &bar_lo[0]);
    }
    tc::tmem_ld32(tT + my_t, v);
    tc::tmem_ld_wait();
    {
      float dz[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = fmaxf(v[j] + s_bias[col0 + j], 0.f);                 // h1q
        float a = 0.f;
#pragma unroll
        for (int h = 0; h < H; ++h) a = fmaf(dsv[h], s_uz[h * D + col0 + j], a);
        dz[j] = v[j] > 0.f ? scale * a : 0.f;                        // dzq
      }
#pragma unroll
      for (int c8 = 0; c8 < 32; c8 += 8) tc::st_row8_bf16(sB2, C::ABLK, row, col0 + c8, dz + c8);
      float cs = warp_colsum32(dz, lane);
      atomicAdd(&s_db[col0 + lane], cs);
#pragma unroll
      for (int h = 0; h < H; ++h) {
        float t[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = dsv[h] * v[j];
        float cu = warp_colsum32(t, lane);
        atomicAdd(&s_dU[h * D + col0 + lane], scale * cu);
        if (cq == 0) {
          float sk = warp_sum(dsv[h]);
          if (lane == 0) atomicAdd(&s_misc[h], scale * sk);
        }
      }
    }
    tc::tc_fence_before();
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      issue_wgrad<D>(tW1q, aB0, aB2, C::ABLK, ct > 0);
      issue_dgrad<D>(tT, aB2, aW1q, C::ABLK, C::WBLK, 0);
      tc::mma_commit(&bar_m[5]);
    }
    tc::mbar_wait(&bar_m[5], par);
    tc::tc_fence_after();
    tc::tmem_ld32(tT + my_t, v);                       // d gamma_q
    tc::tmem_ld_wait();
    {
      float du[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float partner = tile_elem(sB0, C::ABLK, row, sin_cols ? col0 + j + HD : col0 + j - HD);
        float dproj = sin_cols ? partner * v[j] : -partner * v[j];
#pragma unroll
        for (int i = 0; i < 6; ++i) if (i < P.I) du[i] = fmaf(dproj, s_om[i * HD + jj0 + j], du[i]);
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) if (i < P.I) atomicAdd(&s_du[row * 8 + i], du[i]);
    }
    tc::tc_fence_before();
    __syncthreads();
    // ---------------- invariants / window backward (one thread per row) ---------------------------------------
    if (cq == 0) {
      float dq[ENF_R_LAM];
#pragma unroll
      for (int r = 0; r < ENF_R_LAM; ++r) dq[r] = 0.f;
      float du[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) du[i] = i < P.I ? s_du[row * 8 + i] : 0.f;
      float dw = 0.f;
#pragma unroll
      for (int h = 0; h < H; ++h) dw += dsv[h];
      float dsg = 0.f;
      if (P.win_kind != ENF_WIN_NONE) {
        const float inv_s2 = 1.f / (sigma * sigma);
        if (P.win_kind == ENF_WIN_NP) {
          dq[P.I] = -dw * inv_s2;
          dsg = dw * (-2.f * rec.w / sigma);
        } else if (P.win_kind == ENF_WIN_PER) {
          du[0] += dw * 2.f * rec.u[0] * inv_s2;
          du[1] += dw * 2.f * rec.u[1] * inv_s2;
          dsg = dw * (-2.f * rec.w / sigma);
        } else {
          float cl = fminf(fmaxf(rec.c, -1.f + 1e-6f), 1.f - 1e-6f);
          float ac = acosf(cl);
          float dc = (rec.c > -1.f + 1e-6f && rec.c < 1.f - 1e-6f) ? dw * rec.w * ac * inv_s2 * rsqrtf(1.f - cl * cl) : 0.f;
          if (P.win_row >= 0) dq[P.I] = dc; else du[0] += dc;
          dsg = dw * rec.w * ac * ac * inv_s2 / sigma;
        }
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        if (r < P.I) {
          if (P.row_kind == ENF_ROW_SQDIST_SQRT) dq[r] = rec.u[r] > 0.f ? du[r] / (2.f * rec.u[r]) : 0.f;
          else dq[r] = du[r];
        }
      }
#pragma unroll
      for (int r = 0; r < ENF_R_LAM; ++r) s_dq[row * 8 + r] = dq[r];
      s_dsg[row] = dsg;
    }
    __syncthreads();
    if (tid < ENF_LAM_SIZE) {
      int r = tid / ENF_F_XI, ff = tid % ENF_F_XI;
      float a = 0.f;
      for (int t = 0; t < ROWS; ++t) a = fmaf(s_dq[t * 8 + r], s_xi[t * 8 + ff], a);
      s_dlam[tid] += a;
    } else if (tid == 64) {
      float a = 0.f;
      for (int t = 0; t < ROWS; ++t) a += s_dsg[t];
      s_misc[4] += a;
    }
    __syncthreads();           // the reduction reads s_xi / s_dq, which the next tile overwrites
  }
  // ---- flush ---------------------------------------------------------------------------------------------------
  __syncthreads();
  tc::tc_fence_after();
  {
    float* dst[3] = {P.g_Wp, P.g_v_w1, P.g_q_w1};
    const uint32_t src[3] = {tWp, tW1v, tW1q};
#pragma unroll
    for (int k = 0; k < (QPATH ? 3 : 2); ++k) {
      float v[32];
      tc::tmem_ld32(src[k] + my_t, v);
      tc::tmem_ld_wait();
      float* o = dst[k] + (size_t)row * D + col0;
#pragma unroll
      for (int j = 0; j < 32; ++j) atomicAdd(o + j, v[j] * inv_gs);
    }
  }
  for (int e = tid; e < D; e += C::NT) {
    if (QPATH) atomicAdd(P.g_q_b1 + e, s_db[e] * inv_gs);
    atomicAdd(P.g_v_b1 + e, s_db[D + e] * inv_gs);
    atomicAdd(P.g_bp + e, s_db[2 * D + e] * inv_gs);
  }
  if (QPATH) {
    for (int e = tid; e < H * D; e += C::NT) P.g_U[bz * H * D + e] = s_dU[e] * inv_gs;
    if (tid < H) P.g_kappa[bz * H + tid] = s_misc[tid] * inv_gs;
    if (tid < ENF_LAM_SIZE) P.g_lam[bz * ENF_LAM_SIZE + tid] = s_dlam[tid] * inv_gs;
    if (tid == 64 && P.win_kind != ENF_WIN_NONE) P.g_sigma[bz] = s_misc[4] * inv_gs;
  }
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
  size_t smem_b = 3 * C::WIMG + 3 * C::ATILE +
                  (ROWS * 8 + 64 + H * D + 3 * D + 12 * C::HD + 2 * C::NQ * ROWS * 2 + ROWS * 8 * 2 + ROWS + H * D + 3 * D + 64 + 8) * 4 + 128 + 1024;
  if (cudaFuncSetAttribute(pairs_bwd_tc_a_kernel<D, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(pairs_bwd_tc_b_kernel<D, H, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b) != cudaSuccess) return -1;
  const unsigned grid = (unsigned)(p.B * p.Z);
  pairs_bwd_tc_a_kernel<D, H><<<grid, C::NT, smem_a, st>>>(p);
  pairs_bwd_tc_b_kernel<D, H, false><<<grid, C::NT, smem_b, st>>>(p);
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
