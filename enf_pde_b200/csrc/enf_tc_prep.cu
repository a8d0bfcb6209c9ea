// Operand-image preparation for the tensor-core kernels, and a single-tile tcgen05 GEMM used by the
// tests to pin the descriptor / swizzle conventions of enf_tc.cuh on real hardware.
#include "enf_common.cuh"
#include "enf_tc.cuh"

namespace {

// Wt [batch][N][K] fp32 (k contiguous)  ->  img [batch][K/64][N][64] bf16, 128B-swizzled rows (the exact bytes a
// K-major B operand tile occupies in shared memory, so the hot kernels fetch it with one bulk copy).
// cw [batch][N] (optional) = sum_k of the bf16-ROUNDED values (column sums of W as the MMA sees it).
__global__ void __launch_bounds__(256) weight_image_kernel(const float* __restrict__ Wt, uint8_t* __restrict__ img,
                                                           float* __restrict__ cw, int N, int K, int residual) {
  const int64_t b = blockIdx.x;
  const float* src = Wt + b * (int64_t)N * K;
  // image bytes: [max(K / 64, 1)][N rows][128 B]; K = 32 keeps the 128-byte row pitch and fills half of it
  uint8_t* dst = img + b * (int64_t)N * (K < 64 ? 64 : K) * 2;
  const int cpr = K / 8;                       // 16-byte chunks per row (4, 8 or 16: divides the warp)
  for (int e = threadIdx.x; e < N * cpr; e += blockDim.x) {
    int n = e / cpr, ch = e % cpr;
    int k0 = ch * 8;
    const float4* s4 = reinterpret_cast<const float4*>(src + (int64_t)n * K + k0);
    float4 a = __ldg(s4), c = __ldg(s4 + 1);
    float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    if (residual) {            // image of W - round16(W): the low part of a two-term 16-bit split
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] -= tc::round_operand(v[i]);
    }
    uint4 q;
    q.x = tc::pack_bf16(v[0], v[1]); q.y = tc::pack_bf16(v[2], v[3]); q.z = tc::pack_bf16(v[4], v[5]); q.w = tc::pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst + (size_t)(k0 >> 6) * N * 128 + tc::swz_chunk_off(n, (k0 & 63) >> 3)) = q;
    if (cw) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += tc::round_operand(v[i]);
      for (int o = cpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (ch == 0) cw[b * N + n] = s;
    }
  }
}

// Same image, built straight from the UNtransposed weight W [batch][K][N] (n contiguous, Flax layout): a thread owns
// column n and 8 consecutive k, so a warp's loads are 8 coalesced 128-byte rows; the image is assembled in shared
// memory and written out linearly.  Saves the fp32 transpose round trip of the per-latent W3 (268 MB each way at ns64).
__global__ void __launch_bounds__(256) weight_image_T_kernel(const float* __restrict__ W, uint8_t* __restrict__ img, int N, int K) {
  extern __shared__ uint4 s_img[];
  const int64_t b = blockIdx.x;
  const float* src = W + b * (int64_t)N * K;
  const int KP = K < 64 ? 64 : K;              // K = 32: rows keep the 128-byte pitch (half used; the rest is zeroed)
  if (K < 64) for (int e = threadIdx.x; e < N * KP / 8; e += blockDim.x) s_img[e] = make_uint4(0u, 0u, 0u, 0u);
  if (K < 64) __syncthreads();
  const int cpr = K / 8;
  for (int e = threadIdx.x; e < N * cpr; e += blockDim.x) {
    const int n = e % N, k0 = (e / N) * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(src + (int64_t)(k0 + i) * N + n);
    uint4 q;
    q.x = tc::pack_bf16(v[0], v[1]); q.y = tc::pack_bf16(v[2], v[3]); q.z = tc::pack_bf16(v[4], v[5]); q.w = tc::pack_bf16(v[6], v[7]);
    s_img[((size_t)(k0 >> 6) * N * 128 + tc::swz_chunk_off(n, (k0 & 63) >> 3)) >> 4] = q;
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(img + b * (int64_t)N * KP * 2);
  for (int e = threadIdx.x; e < N * KP / 8; e += blockDim.x) dst[e] = s_img[e];
}

// ---- single-tile test GEMM: 128 rows, D features (D = 64 or 128) ---------------------------------------
// mode 0: Dout[r][n] = sum_k X[r][k] * Wt[n][k]          (A K-major from registers, B K-major from the image)
// mode 1: Dout[r][k] = sum_n X[r][n] * Wt[n][k]          (B read MN-major from the SAME image: dgrad)
// mode 2: Dout[i][j] = sum_r X[r][i] * Y[r][j]           (A and B MN-major activation tiles: wgrad)
template <int D>
__global__ void __launch_bounds__(128) tc_gemm_test_kernel(int mode, const float* __restrict__ X,
                                                           const float* __restrict__ Y, const uint8_t* __restrict__ img,
                                                           float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; pointer stays in the shared address space (LDS/STS, not generic LD/ST)
  constexpr uint32_t BLK = 128 * 128;            // bytes of one [128 rows][64] block
  constexpr uint32_t TILE = (D / 64) * BLK;
  uint8_t* tA = base;
  uint8_t* tB = base + TILE;
  __shared__ uint64_t bar_w, bar_mma;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { tc::mbar_init(&bar_w, 1); tc::mbar_init(&bar_mma, 1); tc::mbar_fence_init(); }
  if (warp == 0) tc::tmem_alloc<128>(&tmem_base);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;

  // activation tiles: thread = row
  for (int c0 = 0; c0 < D; c0 += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = X[tid * D + c0 + i];
    tc::st_row8_bf16(tA, BLK, tid, c0, v);
    if (mode == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = Y[tid * D + c0 + i];
      tc::st_row8_bf16(tB, BLK, tid, c0, v);
    }
  }
  if (mode != 2 && tid == 0) {
    tc::mbar_expect_tx(&bar_w, (uint32_t)(D * D * 2));
    tc::bulk_g2s(tB, img, (uint32_t)(D * D * 2), &bar_w);       // image = [D/64][D rows][64]: block bytes = D*128
  }
  tc::fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    if (mode != 2) tc::mbar_wait(&bar_w, 0);
    tc::tc_fence_after();
    const uint32_t a0 = tc::smem_u32(tA), b0 = tc::smem_u32(tB);
    const uint32_t WBLK = D * 128;                // block bytes of the weight image (D rows)
    if (mode == 0) {
      const uint32_t idesc = tc::make_idesc(128, D, tc::kOperandFmt, 0, 0);
      for (int kk = 0; kk < D / 16; ++kk)
        tc::mma_f16(tm, tc::desc_kmajor(a0 + (kk >> 2) * BLK + (kk & 3) * 32), tc::desc_kmajor(b0 + (kk >> 2) * WBLK + (kk & 3) * 32),
                    idesc, kk > 0);
    } else if (mode == 1) {
      const uint32_t idesc = tc::make_idesc(128, D, tc::kOperandFmt, 0, 1);
      for (int kk = 0; kk < D / 16; ++kk)       // reduction over image rows n: 16 rows per step
        tc::mma_f16(tm, tc::desc_kmajor(a0 + (kk >> 2) * BLK + (kk & 3) * 32), tc::desc_mnmajor(b0 + kk * 2048, WBLK), idesc, kk > 0);
    } else {
      const uint32_t idesc = tc::make_idesc(D, D, tc::kOperandFmt, 1, 1);      // M = D features of X, N = D features of Y, K = 128 rows
      for (int kk = 0; kk < 128 / 16; ++kk)
        tc::mma_f16(tm, tc::desc_mnmajor(a0 + kk * 2048, BLK), tc::desc_mnmajor(b0 + kk * 2048, BLK), idesc, kk > 0);
    }
    tc::mma_commit(&bar_mma);
  }
  tc::mbar_wait(&bar_mma, 0);
  tc::tc_fence_after();
  // accumulator row of this thread: an M = 128 accumulator keeps row r in TMEM lane r; an M = 64 accumulator (mode 2 at
  // D = 64: dW is 64 x 64) keeps row r in lane 32 (r / 16) + r % 16 -- 16 rows in the lower half of every lane quadrant
  const int lane = tid & 31;
  const bool m64 = (mode == 2) && D == 64;
  const int out_row = m64 ? (lane < 16 ? warp * 16 + lane : -1) : tid;
  for (int c0 = 0; c0 < D; c0 += 32) {
    float v[32];
    tc::tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_ld_wait();
    if (out_row >= 0)
      for (int i = 0; i < 32; ++i) out[out_row * D + c0 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<128>(tm);
}

// ---- bring-up of num_hidden = 32 (K = N = 32 < the 64-element swizzle row): which MMA shapes read a half-used 128-byte row?
// Tiles keep the 128-byte row pitch of the D >= 64 kernels; features 32..63 of every row are zero.
//   mode 0: out[r][n] = sum_k X[r][k] Wt[n][k]     K-major A and B, N = 32, two K-steps
//   mode 1: out[r][k] = sum_n X[r][n] Wt[n][k]     B = the image read MN-major with N = 32 (half a swizzle row)
//   mode 3: same with N = 64 (reads the zero half too; output columns 32..63 are zeros)
//   mode 2: out[i][j] = sum_r X[r][i] Y[r][j]      A MN-major with M = 64 (rows 32..63 of the result are zeros), B MN-major N = 32
//   mode 4: same with N = 64
__global__ void __launch_bounds__(128) tc_gemm_test32_kernel(int mode, const float* __restrict__ X, const float* __restrict__ Y,
                                                             const uint8_t* __restrict__ img, float* __restrict__ out) {
  constexpr int D = 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr uint32_t BLK = 128 * 128;
  uint8_t* tA = base;
  uint8_t* tB = base + BLK;
  __shared__ uint64_t bar_w, bar_mma;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { tc::mbar_init(&bar_w, 1); tc::mbar_init(&bar_mma, 1); tc::mbar_fence_init(); }
  if (warp == 0) tc::tmem_alloc<128>(&tmem_base);
  for (int e = tid; e < 2 * (int)BLK / 16; e += 128) reinterpret_cast<uint4*>(base)[e] = make_uint4(0u, 0u, 0u, 0u);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  const bool wgrad = mode == 2 || mode == 4;
  for (int c0 = 0; c0 < D; c0 += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = X[tid * D + c0 + i];
    tc::st_row8_bf16(tA, BLK, tid, c0, v);
    if (wgrad) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = Y[tid * D + c0 + i];
      tc::st_row8_bf16(tB, BLK, tid, c0, v);
    }
  }
  if (!wgrad && tid == 0) {
    tc::mbar_expect_tx(&bar_w, (uint32_t)(D * 128));
    tc::bulk_g2s(tB, img, (uint32_t)(D * 128), &bar_w);          // image = [32 rows][128 bytes], 64 of them used
  }
  tc::fence_proxy_async();
  __syncthreads();
  const int NN = (mode == 3 || mode == 4) ? 64 : 32;
  if (tid == 0) {
    if (!wgrad) tc::mbar_wait(&bar_w, 0);
    tc::tc_fence_after();
    const uint32_t a0 = tc::smem_u32(tA), b0 = tc::smem_u32(tB);
    const uint32_t WBLK = D * 128;
    if (mode == 0) {
      const uint32_t idesc = tc::make_idesc(128, 32, tc::kOperandFmt, 0, 0);
      for (int kk = 0; kk < 2; ++kk) tc::mma_f16(tm, tc::desc_kmajor(a0 + kk * 32), tc::desc_kmajor(b0 + kk * 32), idesc, kk > 0);
    } else if (!wgrad) {
      const uint32_t idesc = tc::make_idesc(128, NN, tc::kOperandFmt, 0, 1);
      for (int kk = 0; kk < 2; ++kk) tc::mma_f16(tm, tc::desc_kmajor(a0 + kk * 32), tc::desc_mnmajor(b0 + kk * 2048, WBLK), idesc, kk > 0);
    } else {
      const uint32_t idesc = tc::make_idesc(64, NN, tc::kOperandFmt, 1, 1);
      for (int kk = 0; kk < 128 / 16; ++kk) tc::mma_f16(tm, tc::desc_mnmajor(a0 + kk * 2048, BLK), tc::desc_mnmajor(b0 + kk * 2048, BLK), idesc, kk > 0);
    }
    tc::mma_commit(&bar_mma);
  }
  tc::mbar_wait(&bar_mma, 0);
  tc::tc_fence_after();
  const int out_row = wgrad ? (lane < 16 ? warp * 16 + lane : -1) : tid;       // M = 64: row r in lane 32 (r / 16) + r % 16
  for (int c0 = 0; c0 < NN; c0 += 32) {
    float v[32];
    tc::tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_ld_wait();
    if (out_row >= 0)
      for (int i = 0; i < 32; ++i) out[out_row * 64 + c0 + i] = v[i];            // out: [128][64]
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<128>(tm);
}

}  // namespace

int enf_launch_weight_image(cudaStream_t st, const float* Wt, void* img, float* cw, int N, int K, int batch, int residual) {
  weight_image_kernel<<<batch, 256, 0, st>>>(Wt, (uint8_t*)img, cw, N, K, residual);
  return 1;
}

int enf_launch_weight_image_T(cudaStream_t st, const float* W, void* img, int N, int K, int batch) {
  const size_t smem = (size_t)N * (K < 64 ? 64 : K) * 2;
  if (smem > 48 * 1024) return -1;
  weight_image_T_kernel<<<batch, 256, smem, st>>>(W, (uint8_t*)img, N, K);
  return 1;
}

extern "C" int enf_debug_tc_gemm(int mode, int D, const float* X, const float* Y, float* out, void* scratch, enf_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 32) {       // bring-up shapes of num_hidden = 32; out is [128][64]; scratch >= 4096 bytes
    if (mode != 2 && mode != 4) {
      if (cudaMemsetAsync(scratch, 0, 32 * 128, st) != cudaSuccess) return ENF_ERR_CUDA;
      enf_launch_weight_image(st, Y, scratch, nullptr, 32, 32, 1, 0);
    }
    const size_t smem32 = 2 * 128 * 128 + 1024;
    tc_gemm_test32_kernel<<<1, 128, smem32, st>>>(mode, X, Y, (const uint8_t*)scratch, out);
    return cudaGetLastError() == cudaSuccess ? 0 : ENF_ERR_CUDA;
  }
  if (D != 64 && D != 128) return ENF_ERR_UNSUPPORTED;
  if (mode != 2) enf_launch_weight_image(st, Y, scratch, nullptr, D, D, 1, 0);
  size_t smem = 2 * (size_t)(D / 64) * 128 * 128 + 1024;
  if (D == 128) {
    cudaFuncSetAttribute(tc_gemm_test_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_gemm_test_kernel<128><<<1, 128, smem, st>>>(mode, X, Y, (const uint8_t*)scratch, out);
  } else {
    cudaFuncSetAttribute(tc_gemm_test_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_gemm_test_kernel<64><<<1, 128, smem, st>>>(mode, X, Y, (const uint8_t*)scratch, out);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : ENF_ERR_CUDA;
}
