// Tensor-core GEMMs for the per-query tail (Q), the W3 product (L) and their backward, used in the
// tensor-core precision mode.  TF32 operands (kind::tf32: fp32 containers, 10-bit mantissa, fp32 range, so
// gradients need no scaling), fp32 accumulation in TMEM.  Operands go global -> shared memory with TMA tensor
// maps (128-byte swizzle) and are consumed by tcgen05.mma straight from there: no conversion pass, no register
// staging.  These stages are HBM bound (one read of the activations, one write of the result), so the kernels
// are persistent, warp specialised (TMA producer / MMA issuer / 4 epilogue warps) and multi-stage.
//
//   F  C[M,N] = A[M,K] * B[K,N] (+bias) (*gelu'(aux)), optional second output gelu(C)      M huge, N,K <= 256
//      A row-major (K-major operand).  B either [K][N] row-major (MN-major operand; forward layers)
//      or the transposed view of an [N][K] row-major weight (K-major operand; dgrad).
//      Double-buffered TMEM accumulator: the epilogue of tile i overlaps the MMAs of tile i+1.
//   W  C[K1,N] += A[Mr,K1]^T * G[Mr,N]      weight gradients: reduction over Mr = B*C rows (huge)
//      both operands MN-major; every CTA owns a contiguous slice of the reduction, keeps the K1 x N
//      accumulator in TMEM for its whole life and adds it to C once (vector atomics).
//
// Precision.  A tf32 MMA truncates its fp32 operands to 10 mantissa bits; a plain tf32 product costs the tail
// 1-2.5e-3 of parity on its own (measured with tests' rounding-injection model), so
//   F runs the 3-term split  A B ~= A_hi B_hi + A_lo B_hi + A_hi B_lo  (A_hi = what the hardware keeps of A,
//     A_lo = A - A_hi computed in shared memory by four converter warps, B_lo precomputed per weight):
//     fp32-level accuracy, three MMAs per k-step, still at the HBM bound;
//   W rounds both operands to nearest in place in shared memory (unbiased single tf32 product; the
//     reduction over >= 1e5 rows averages the rounding noise).
#include <cuda.h>

#include "enf_common.cuh"
#include "enf_tc.cuh"

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// 2-D map of a row-major fp32 matrix [rows][cols] (row stride ld floats); box = 32 floats x box_rows
bool make_map_2d(CUtensorMap* m, const float* base, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t es[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// 3-D map of the same matrix seen as [cols/32 atoms][rows][32 floats]: a box {32, box_rows, natoms} lands in
// shared memory as natoms blocks of [box_rows][128 bytes] = the MN-major operand layout of tcgen05.mma.
// MN-major tf32 operands have ONE legal shared-memory layout: 128-byte rows whose four 32-byte chunks are XOR-swizzled
// with (row & 3) (UMMA layout type SWIZZLE_128B_BASE32B; TMA mode SWIZZLE_128B_ATOM_32B), 4 K-rows per swizzle atom.
bool make_map_3d(CUtensorMap* m, const float* base, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {32, rows, cols / 32};
  cuuint64_t strides[2] = {ld * 4, 128};
  cuuint32_t box[3] = {32, box_rows, (cuuint32_t)(cols / 32)};
  cuuint32_t es[3] = {1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// descriptor of an MN-major tf32 operand: blocks of [K rows][32 floats] per 32-float MN atom, `lbo` bytes apart;
// 4-row K groups 512 bytes apart; layout type 1 = SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint64_t desc_mn_tf32(uint32_t saddr, uint32_t lbo) {
  return (uint64_t(1) << 46) | (uint64_t(1) << 61) | uint64_t((saddr >> 4) & 0x3FFF) | (uint64_t((lbo >> 4) & 0x3FFF) << 16) |
         (uint64_t(512 >> 4) << 32);
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(tc::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(tc::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float c0 = 0.7978845608028654f, c1 = 0.7978845608028654f * 0.044715f;
  float t = tanh_fast(x * fmaf(c1, x * x, c0));
  float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  const float c0 = 0.7978845608028654f, c1 = 0.7978845608028654f * 0.044715f;
  float x2 = x * x;
  float t = tanh_fast(x * fmaf(c1, x2, c0));
  return fmaf(0.5f * x * fmaf(-t, t, 1.f), fmaf(3.f * c1, x2, c0), fmaf(0.5f, t, 0.5f));
}

constexpr int kThreads = 320;            // warp 0: TMA producer, 1: MMA issuer, 2..5: epilogue, 6..9: operand converters
constexpr uint32_t kABytesF = 128 * 128; // A stage of kernel F: 128 rows x 32 floats

struct FArgs {
  int M, N, K;
  int stages;
  float* C; int64_t ldc;
  const float* bias;
  float* C2;            // optional: gelu(C)
  const float* aux;     // optional: C *= gelu'(aux), aux laid out like C
  int ntiles;
};

// ------------------------------------------------------------------------------------------------------------
template <int BMAJ>
__global__ void __launch_bounds__(kThreads, 1) gemm_tf32_f_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB,
                                                                 const __grid_constant__ CUtensorMap tmBlo, FArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; pointer stays in the shared address space (LDS/STS, not generic LD/ST)
  const uint32_t b_bytes = (uint32_t)g.N * 128;
  const uint32_t stage_bytes = 2 * kABytesF + 2 * b_bytes;        // A | A_lo | B | B_lo
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)g.stages * stage_bytes);
  uint64_t* full = bars;                 // [stages] TMA landed
  uint64_t* empty = bars + 8;            // [stages] MMAs that read the stage completed
  uint64_t* conv = bars + 16;            // [stages] A_lo written
  uint64_t* tfull = bars + 24;           // [2]
  uint64_t* tempty = bars + 26;          // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 28);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < g.stages; ++i) { tc::mbar_init(full + i, 1); tc::mbar_init(empty + i, 1); tc::mbar_init(conv + i, 4); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(tfull + i, 1); tc::mbar_init(tempty + i, 4); }
    tc::mbar_fence_init();
    prefetch_map(&tmA); prefetch_map(&tmB); prefetch_map(&tmBlo);
  }
  if (warp == 1) tc::tmem_alloc<512>(s_tmem);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = *s_tmem;
  const int nkb = (g.K + 31) / 32;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
        const int m0 = tile * 128;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % g.stages;
          const uint32_t ph = (it / g.stages) & 1;
          tc::mbar_wait(empty + s, ph ^ 1);
          uint8_t* sa = base + (size_t)s * stage_bytes;
          uint8_t* sb = sa + 2 * kABytesF;
          tc::mbar_expect_tx(full + s, kABytesF + 2 * b_bytes);
          tma_load_2d(sa, &tmA, kb * 32, m0, full + s);
          if (BMAJ == 0) {
            tma_load_2d(sb, &tmB, kb * 32, 0, full + s);
            tma_load_2d(sb + b_bytes, &tmBlo, kb * 32, 0, full + s);
          } else {
            tma_load_3d(sb, &tmB, 0, kb * 32, 0, full + s);
            tma_load_3d(sb + b_bytes, &tmBlo, 0, kb * 32, 0, full + s);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc(128, g.N, 2, 0, BMAJ);
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++ti) {
        const uint32_t acc = ti & 1;
        tc::mbar_wait(tempty + acc, ((ti >> 1) & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t d_t = tm + acc * 256;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % g.stages;
          const uint32_t ph = (it / g.stages) & 1;
          tc::mbar_wait(full + s, ph);
          tc::mbar_wait(conv + s, ph);
          tc::tc_fence_after();
          const uint32_t a_addr = tc::smem_u32(base + (size_t)s * stage_bytes);
          const uint32_t alo_addr = a_addr + kABytesF;
          const uint32_t b_addr = a_addr + 2 * kABytesF;
          const uint32_t blo_addr = b_addr + b_bytes;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = tc::desc_kmajor(a_addr + ks * 32);
            const uint64_t ald = tc::desc_kmajor(alo_addr + ks * 32);
            const uint64_t bd = BMAJ == 0 ? tc::desc_kmajor(b_addr + ks * 32) : desc_mn_tf32(b_addr + ks * 1024, 4096);
            const uint64_t bld = BMAJ == 0 ? tc::desc_kmajor(blo_addr + ks * 32) : desc_mn_tf32(blo_addr + ks * 1024, 4096);
            mma_tf32(d_t, ad, bd, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
            mma_tf32(d_t, ald, bd, idesc, 1u);
            mma_tf32(d_t, ad, bld, idesc, 1u);
          }
          tc::mma_commit(empty + s);
        }
        tc::mma_commit(tfull + acc);
      }
    }
  } else if (warp >= 6) {
    // converters: A_lo = A - trunc_tf32(A), element by element (layout agnostic: the lo tile inherits A's swizzle)
    const int ct = tid - 6 * 32;                 // 0..127
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % g.stages;
        const uint32_t ph = (it / g.stages) & 1;
        tc::mbar_wait(full + s, ph);
        const float4* src = reinterpret_cast<const float4*>(base + (size_t)s * stage_bytes);
        float4* dst = reinterpret_cast<float4*>(base + (size_t)s * stage_bytes + kABytesF);
#pragma unroll
        for (int i = 0; i < (int)(kABytesF / 16 / 128); ++i) {
          float4 v = src[ct + i * 128];
          float4 o;
          o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
          o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
          o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
          o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
          dst[ct + i * 128] = o;
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(conv + s);
      }
    }
  } else {
    const int lq = warp & 3;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++ti) {
      const uint32_t acc = ti & 1;
      tc::mbar_wait(tfull + acc, (ti >> 1) & 1);
      tc::tc_fence_after();
      const int64_t row = (int64_t)tile * 128 + lq * 32 + lane;
      const bool ok = row < g.M;
      const uint32_t t_addr = tm + acc * 256 + ((uint32_t)(lq * 32) << 16);
      for (int c0 = 0; c0 < g.N; c0 += 32) {
        float v[32];
        tc::tmem_ld32(t_addr + c0, v);
        tc::tmem_ld_wait();
        if (g.bias) {
          const float4* b4 = reinterpret_cast<const float4*>(g.bias + c0);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 b = __ldg(b4 + q);
            v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
          }
        }
        if (ok) {
          const int64_t off = row * g.ldc + c0;
          if (g.aux) {
            const float4* a4 = reinterpret_cast<const float4*>(g.aux + off);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 a = __ldg(a4 + q);
              v[4 * q] *= gelu_grad_fast(a.x); v[4 * q + 1] *= gelu_grad_fast(a.y);
              v[4 * q + 2] *= gelu_grad_fast(a.z); v[4 * q + 3] *= gelu_grad_fast(a.w);
            }
          }
          float4* o4 = reinterpret_cast<float4*>(g.C + off);
#pragma unroll
          for (int q = 0; q < 8; ++q) o4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          if (g.C2) {
            float4* p4 = reinterpret_cast<float4*>(g.C2 + off);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              p4[q] = make_float4(gelu_fast(v[4 * q]), gelu_fast(v[4 * q + 1]), gelu_fast(v[4 * q + 2]), gelu_fast(v[4 * q + 3]));
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tempty + acc);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<512>(tm);
}

// ------------------------------------------------------------------------------------------------------------
struct WArgs {
  int N, stages;
  int64_t nblocks;     // 32-row blocks of the reduction
  float* C; int64_t ldc;
};

template <int MT>
__global__ void __launch_bounds__(kThreads, 1) gemm_tf32_w_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmG, WArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; pointer stays in the shared address space (LDS/STS, not generic LD/ST)
  constexpr uint32_t a_bytes = MT * 4 * 4096;
  const uint32_t g_bytes = (uint32_t)(g.N / 32) * 4096;
  const uint32_t stage_bytes = a_bytes + g_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)g.stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + 8;
  uint64_t* conv = bars + 16;
  uint64_t* done = bars + 24;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 28);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < g.stages; ++i) { tc::mbar_init(full + i, 1); tc::mbar_init(empty + i, 1); tc::mbar_init(conv + i, 4); }
    tc::mbar_init(done, 1);
    tc::mbar_fence_init();
    prefetch_map(&tmA); prefetch_map(&tmG);
  }
  if (warp == 1) tc::tmem_alloc<512>(s_tmem);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = *s_tmem;
  const int64_t lo = g.nblocks * blockIdx.x / gridDim.x, hi = g.nblocks * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t blk = lo; blk < hi; ++blk, ++it) {
        const int s = it % g.stages;
        const uint32_t ph = (it / g.stages) & 1;
        tc::mbar_wait(empty + s, ph ^ 1);
        uint8_t* sa = base + (size_t)s * stage_bytes;
        tc::mbar_expect_tx(full + s, stage_bytes);
        tma_load_3d(sa, &tmA, 0, (int)(blk * 32), 0, full + s);
        tma_load_3d(sa + a_bytes, &tmG, 0, (int)(blk * 32), 0, full + s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc(128, g.N, 2, 1, 1);
      uint32_t it = 0;
      for (int64_t blk = lo; blk < hi; ++blk, ++it) {
        const int s = it % g.stages;
        const uint32_t ph = (it / g.stages) & 1;
        tc::mbar_wait(full + s, ph);
        tc::mbar_wait(conv + s, ph);
        tc::tc_fence_after();
        const uint32_t a_addr = tc::smem_u32(base + (size_t)s * stage_bytes);
        const uint32_t g_addr = a_addr + a_bytes;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t gd = desc_mn_tf32(g_addr + ks * 1024, 4096);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            mma_tf32(tm + mt * 256, desc_mn_tf32(a_addr + mt * 4 * 4096 + ks * 1024, 4096), gd, idesc, (it > 0 || ks > 0) ? 1u : 0u);
        }
        tc::mma_commit(empty + s);
      }
      tc::mma_commit(done);
    }
  } else if (warp >= 6) {
    // converters: round both operand tiles to nearest tf32 in place (the MMA would truncate: biased)
    const int ct = tid - 6 * 32;
    const int nvec = (int)(stage_bytes / 16);
    uint32_t it = 0;
    for (int64_t blk = lo; blk < hi; ++blk, ++it) {
      const int s = it % g.stages;
      const uint32_t ph = (it / g.stages) & 1;
      tc::mbar_wait(full + s, ph);
      float4* buf = reinterpret_cast<float4*>(base + (size_t)s * stage_bytes);
      for (int i = ct; i < nvec; i += 128) {
        float4 v = buf[i];
        buf[i] = make_float4(enf_round_tf32(v.x), enf_round_tf32(v.y), enf_round_tf32(v.z), enf_round_tf32(v.w));
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(conv + s);
    }
  } else if (hi > lo) {
    const int lq = warp & 3;
    tc::mbar_wait(done, 0);
    tc::tc_fence_after();
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int64_t row = mt * 128 + lq * 32 + lane;
      for (int c0 = 0; c0 < g.N; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tm + mt * 256 + c0 + ((uint32_t)(lq * 32) << 16), v);
        tc::tmem_ld_wait();
        float4* o4 = reinterpret_cast<float4*>(g.C + row * g.ldc + c0);
#pragma unroll
        for (int q = 0; q < 8; ++q) atomicAdd(o4 + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<512>(tm);
}

constexpr size_t kSmemMax = 227 * 1024;

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// Returns the number of kernels launched (1), 0 if the shape is not one this path takes (caller falls back
// to the fp32 kernel), -1 on a configuration error.
int enf_gemm_tc(cudaStream_t st, int M, int N, int K, EnfMat A, EnfMat B, EnfMat C, const EnfGemmOpts& o) {
  if (o.batch != 1 || o.act_a || o.alpha != 1.0f) return 0;
  if (!aligned16(A.p) || !aligned16(B.p) || !aligned16(C.p)) return 0;
  // ---- F: big-M product with a small weight --------------------------------------------------------
  if (!o.accumulate && A.cs == 1 && C.cs == 1 && M >= 128 && N % 32 == 0 && N >= 32 && N <= 256 && K % 4 == 0 && K >= 32 &&
      (A.rs % 4) == 0 && (C.rs % 4) == 0 && (B.cs == 1 || B.rs == 1)) {
    if (o.bias && !aligned16(o.bias)) return 0;
    if (o.mul_gelu_grad && !aligned16(o.mul_gelu_grad)) return 0;
    if (o.gelu_out && !aligned16(o.gelu_out)) return 0;
    const int bmaj = B.cs == 1 ? 1 : 0;
    const int64_t ldb = bmaj ? B.rs : B.cs;
    if (ldb % 4) return 0;
    if (!o.b_lo || !aligned16(o.b_lo)) return 0;
    CUtensorMap tmA, tmB, tmBlo;
    if (!make_map_2d(&tmA, A.p, (uint64_t)K, (uint64_t)M, (uint64_t)A.rs, 128)) return -1;
    if (bmaj) {
      if (!make_map_3d(&tmB, B.p, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 32)) return -1;
      if (!make_map_3d(&tmBlo, o.b_lo, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 32)) return -1;
    } else {
      if (!make_map_2d(&tmB, B.p, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, (uint32_t)N)) return -1;
      if (!make_map_2d(&tmBlo, o.b_lo, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, (uint32_t)N)) return -1;
    }
    FArgs g;
    g.M = M; g.N = N; g.K = K;
    const size_t stage = 2 * kABytesF + 2 * (size_t)N * 128;
    int stages = (int)((kSmemMax - 2048) / stage);
    if (stages > 6) stages = 6;
    g.stages = stages;
    g.C = const_cast<float*>(C.p); g.ldc = C.rs; g.bias = o.bias; g.C2 = o.gelu_out; g.aux = o.mul_gelu_grad;
    g.ntiles = (M + 127) / 128;
    const size_t smem = stages * stage + 1024 + 256;
    int grid = g.ntiles < 148 ? g.ntiles : 148;
    auto kern = bmaj ? gemm_tf32_f_kernel<1> : gemm_tf32_f_kernel<0>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    kern<<<grid, kThreads, smem, st>>>(tmA, tmB, tmBlo, g);
    return 1;
  }
  // ---- W: weight gradient, reduction over a huge row count ------------------------------------------
  if (o.accumulate && !o.bias && !o.mul_gelu_grad && !o.gelu_out && A.rs == 1 && B.cs == 1 && C.cs == 1 && K >= 128 &&
      (M == 128 || M == 256) && N % 32 == 0 && N >= 32 && N <= 256 && (A.cs % 4) == 0 && (B.rs % 4) == 0 && (C.rs % 4) == 0) {
    CUtensorMap tmA, tmG;
    if (!make_map_3d(&tmA, A.p, (uint64_t)M, (uint64_t)K, (uint64_t)A.cs, 32)) return -1;
    if (!make_map_3d(&tmG, B.p, (uint64_t)N, (uint64_t)K, (uint64_t)B.rs, 32)) return -1;
    WArgs g;
    g.N = N;
    const int MT = M / 128;
    const size_t stage = (size_t)MT * 16384 + (size_t)N * 128;
    int stages = (int)((kSmemMax - 2048) / stage);
    if (stages > 6) stages = 6;
    g.stages = stages;
    g.nblocks = ((int64_t)K + 31) / 32;
    g.C = const_cast<float*>(C.p); g.ldc = C.rs;
    const size_t smem = stages * stage + 1024 + 256;
    int grid = g.nblocks < 148 ? (int)g.nblocks : 148;
    auto kern = MT == 1 ? gemm_tf32_w_kernel<1> : gemm_tf32_w_kernel<2>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    kern<<<grid, kThreads, smem, st>>>(tmA, tmG, g);
    return 1;
  }
  return 0;
}
