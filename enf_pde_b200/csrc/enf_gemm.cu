// Generic strided / batched / split-K fp32 GEMM used by the per-latent (L), per-weight (W) and
// per-query tail (Q) stages.  These stages are a few percent of the work of the fused pair kernels;
// this kernel favours generality (arbitrary strides => transposes and head-sliced views for free)
// over peak throughput.  64x64x16 tiles, 256 threads, 4x4 register tile per thread.
#include <cstdlib>

#include "enf_common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

struct GemmKArgs {
  const float* A; int64_t sam, sak, sab;
  const float* B; int64_t sbk, sbn, sbb;
  float* C; int64_t scm, scn, scb;
  const float* bias; int64_t bias_bs;
  const float* aux;
  float* C2;
  int round_out;
  int M, N, K, splitk, kchunk;
  int act_a, accumulate;
  float alpha;
};

__global__ void __launch_bounds__(256) enf_gemm_kernel(GemmKArgs g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int bz = blockIdx.z / g.splitk, ks = blockIdx.z % g.splitk;
  const float* A = g.A + (int64_t)bz * g.sab;
  const float* B = g.B + (int64_t)bz * g.sbb;
  float* C = g.C + (int64_t)bz * g.scb;
  const int kbeg = ks * g.kchunk;
  const int kend = min(g.K, kbeg + g.kchunk);
  const bool a_kfast = (g.sak == 1);     // consecutive threads walk k (row-major A) else walk m
  const bool b_nfast = (g.sbn == 1);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256;
      int kk, mm;
      if (a_kfast) { kk = e & (BK - 1); mm = e >> 4; } else { mm = e & (BM - 1); kk = e >> 6; }
      int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < g.M && gk < kend) {
        v = __ldg(A + (int64_t)gm * g.sam + (int64_t)gk * g.sak);
        if (g.act_a) v = enf_gelu(v);
      }
      As[kk][mm] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256;
      int kk, nn;
      if (b_nfast) { nn = e & (BN - 1); kk = e >> 6; } else { kk = e & (BK - 1); nn = e >> 4; }
      int gn = n0 + nn, gk = k0 + kk;
      float v = 0.f;
      if (gn < g.N && gk < kend) v = __ldg(B + (int64_t)gk * g.sbk + (int64_t)gn * g.sbn);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const float* bias = g.bias ? g.bias + (int64_t)bz * g.bias_bs : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= g.N) continue;
      float v = g.alpha * acc[i][j];
      if (bias && ks == 0) v += bias[gn];
      int64_t off = (int64_t)gm * g.scm + (int64_t)gn * g.scn;
      if (g.aux) v *= enf_gelu_grad(g.aux[(int64_t)bz * g.scb + off]);
      if (g.accumulate) atomicAdd(C + off, v); else C[off] = enf_maybe_round(v, g.round_out);
      if (g.C2) g.C2[(int64_t)bz * g.scb + off] = enf_maybe_round(enf_gelu(v), g.round_out);
    }
  }
}

}  // namespace

int enf_gemm(cudaStream_t st, int M, int N, int K, EnfMat A, EnfMat B, EnfMat C, const EnfGemmOpts& o) {
  if (M <= 0 || N <= 0 || o.batch <= 0) return 0;
  static const bool no_tc = getenv("ENF_DEBUG_NO_TC_GEMM") != nullptr;     // A/B switch for numerics debugging
  if (o.tc && !no_tc) {
    int r = enf_gemm_tc(st, M, N, K, A, B, C, o);
    if (r != 0) return r;
  }
  GemmKArgs g;
  g.A = A.p; g.sam = A.rs; g.sak = A.cs; g.sab = A.bs;
  g.B = B.p; g.sbk = B.rs; g.sbn = B.cs; g.sbb = B.bs;
  g.C = const_cast<float*>(C.p); g.scm = C.rs; g.scn = C.cs; g.scb = C.bs;
  g.bias = o.bias; g.bias_bs = o.bias_bs; g.aux = o.mul_gelu_grad; g.C2 = o.gelu_out; g.round_out = o.round_out;
  g.M = M; g.N = N; g.K = K; g.act_a = o.act_a; g.accumulate = o.accumulate; g.alpha = o.alpha;
  int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  int64_t tiles = (int64_t)tiles_m * tiles_n * o.batch;
  int splitk = 1;
  // split-K only when the result is accumulated atomically anyway and the grid would not fill the GPU
  if (o.accumulate && !o.mul_gelu_grad && tiles < 296 && K >= 512) {
    int want = (int)((592 + tiles - 1) / tiles);
    int maxs = (K + 127) / 128;
    splitk = want < maxs ? want : maxs;
    if (splitk < 1) splitk = 1;
  }
  int kchunk = (K + splitk - 1) / splitk;
  kchunk = ((kchunk + BK - 1) / BK) * BK;
  if (kchunk < BK) kchunk = BK;
  splitk = (K + kchunk - 1) / kchunk;
  if (splitk < 1) splitk = 1;
  g.splitk = splitk; g.kchunk = kchunk;
  dim3 grid(tiles_m, tiles_n, o.batch * splitk);
  enf_gemm_kernel<<<grid, 256, 0, st>>>(g);
  return 1;
}
