// XLA FFI (jax.ffi) custom-call handlers over the C ABI of include/enf_b200.h: the binding a maintainer of the
// reference (JAX / Flax) adds to call the B200 path from `EquivariantCrossAttentionNeF.apply`
// (enf/models/equivariant_cross_attention_nef.py:204-235; call sites experiments/fitting/trainers/pde_trainer.py:184,478,537).
//
// NOT part of libenf_b200.so and NOT built by csrc/Makefile: it needs jaxlib's headers, which this image does not have
// (no `jax`, no `xla/ffi/api/ffi.h` anywhere on disk), so this file has never been compiled here -- it is written against
// the documented XLA FFI C++ API (jaxlib >= 0.4.31) and guarded so that a build without the headers yields an empty
// translation unit.  Where JAX exists:
//
//   g++ -O2 -std=c++17 -shared -fPIC -I"$(python -c 'import jax.ffi; print(jax.ffi.include_dir())')"
//       -I/usr/local/cuda/include -I../../include enf_xla_ffi.cc -L.. -lenf_b200 -o ../libenf_b200_xla.so
//
// Python side: enf_pde_b200/jax_binding.py (register_ffi_target + custom_vjp), described in INTEGRATION.md section 3.
//
// Operands   fwd: x, p, a, sigma, <46 weight leaves in EnfWeights order>
//            bwd: x, p, a, sigma, workspace, d_out, <46 weight leaves in EnfWeights order>
//   x      f32[B,C,Dx]  or  f32[C,Dx] (one grid shared by all fields -> x_batch_stride = 0, pde_trainer.py:197)
//   sigma  f32[B,Z,1]   or  f32[0] when use_window == 0
// Results    fwd: out f32[B,C,O], workspace u8[enf_xattn_workspace_bytes]
//            bwd: workspace u8[...] (ALIASED to the workspace operand: jax_binding.py passes input_output_aliases={4: 0}, so
//                 the library's scratch writes go to a buffer XLA knows is written), then the 46 weight-gradient leaves,
//                 dp f32[B,Z,P], da f32[B,Z,L], dsigma f32[B,Z,1]
// Attributes: d, H, L, O, invariant_kind, use_window, precision, flags, chunk_fields (all i32) = the EnfDesc fields XLA cannot
// infer.  enf_xattn_bwd does not consume the forward state, so a vjp applied to several cotangents stays correct (XLA copies the
// donated workspace when it is still live).
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define ENF_HAVE_XLA_FFI 1
#endif
#endif

#ifdef ENF_HAVE_XLA_FFI

#include <cuda_runtime_api.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <type_traits>

#include "../../include/enf_b200.h"
#include "../../include/enf_ode_b200.h"
#include "xla/ffi/api/c_api.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

using F32 = ffi::Buffer<ffi::F32>;

ffi::Error describe(const F32& x, const F32& p, const F32& a, int32_t d, int32_t H, int32_t L, int32_t O, int32_t invariant_kind,
                    int32_t use_window, int32_t precision, int32_t flags, int32_t chunk_fields, EnfDesc* desc, int64_t* x_batch_stride) {
  std::memset(desc, 0, sizeof(*desc));
  auto xd = x.dimensions(), pd = p.dimensions(), ad = a.dimensions();
  if (pd.size() != 3 || ad.size() != 3 || (xd.size() != 2 && xd.size() != 3))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "enf_xattn: expected x[B,C,Dx] or x[C,Dx], p[B,Z,P], a[B,Z,L]");
  const bool shared = xd.size() == 2;
  desc->B = (int32_t)pd[0];
  desc->Z = (int32_t)pd[1];
  desc->C = (int32_t)xd[shared ? 0 : 1];
  desc->Dx = (int32_t)xd[shared ? 1 : 2];
  desc->d = d; desc->H = H; desc->L = L; desc->O = O;
  desc->invariant_kind = invariant_kind; desc->use_window = use_window; desc->precision = precision; desc->flags = flags;
  desc->chunk_fields = chunk_fields;
  if (ad[0] != pd[0] || ad[1] != pd[1] || ad[2] != L || (!shared && xd[0] != pd[0]))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "enf_xattn: x, p, a disagree on B / Z / latent_dim");
  if (pd[2] != enf_pose_dim(invariant_kind, desc->Dx))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "enf_xattn: pose width does not match the invariant");
  *x_batch_stride = shared ? 0 : (int64_t)desc->C * desc->Dx;
  return ffi::Error::Success();
}

// the 46 leaves arrive as the trailing operands, in EnfWeights order
template <class Struct, class Args>
ffi::Error collect_leaves(Args& args, size_t first, Struct* out, bool results) {
  auto** slots = reinterpret_cast<float**>(out);
  for (int i = 0; i < ENF_NUM_WEIGHT_LEAVES; ++i) {
    auto buf = args.template get<F32>(first + i);
    if (!buf.has_value()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "enf_xattn: weight leaf " + std::to_string(i) + " is not f32");
    if constexpr (std::is_same_v<Args, ffi::RemainingRets>) slots[i] = (*buf)->typed_data();
    else slots[i] = const_cast<float*>(buf->typed_data());
  }
  (void)results;
  return ffi::Error::Success();
}

ffi::Error status(int rc, const char* what) {
  if (rc == ENF_OK) return ffi::Error::Success();
  return ffi::Error(rc == ENF_ERR_UNSUPPORTED ? ffi::ErrorCode::kUnimplemented : ffi::ErrorCode::kInvalidArgument,
                    std::string(what) + ": " + enf_last_error());
}

ffi::Error FwdImpl(cudaStream_t stream, F32 x, F32 p, F32 a, F32 sigma, ffi::RemainingArgs leaves, ffi::ResultBuffer<ffi::F32> out,
                   ffi::ResultBuffer<ffi::U8> workspace, int32_t d, int32_t H, int32_t L, int32_t O, int32_t invariant_kind,
                   int32_t use_window, int32_t precision, int32_t flags, int32_t chunk_fields) {
  EnfDesc desc;
  int64_t xbs;
  if (auto e = describe(x, p, a, d, H, L, O, invariant_kind, use_window, precision, flags, chunk_fields, &desc, &xbs); e.failure()) return e;
  if (leaves.size() != ENF_NUM_WEIGHT_LEAVES) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "enf_xattn_fwd: expected 46 weight leaves");
  EnfWeights w;
  if (auto e = collect_leaves(leaves, 0, &w, false); e.failure()) return e;
  const size_t need = enf_xattn_workspace_bytes(&desc);
  if (need == 0) return status(ENF_ERR_BAD_DESC, "enf_xattn_workspace_bytes");
  if ((size_t)workspace->element_count() < need) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "enf_xattn_fwd: workspace result too small");
  const float* sg = (use_window && sigma.element_count() > 0) ? sigma.typed_data() : nullptr;
  return status(enf_xattn_fwd(&desc, &w, x.typed_data(), xbs, p.typed_data(), a.typed_data(), sg, out->typed_data(),
                              workspace->typed_data(), (size_t)workspace->element_count(), (enf_stream_t)stream),
                "enf_xattn_fwd");
}

// operands: x, p, a, sigma, workspace (the forward's result), d_out, then the 46 leaves.
// results:  workspace_out (aliased to the workspace operand by the caller: same device buffer, but declared as WRITTEN, which
//           an input operand is not), then 46 leaf gradients, dp, da, dsigma.
ffi::Error BwdImpl(cudaStream_t stream, F32 x, F32 p, F32 a, F32 sigma, ffi::Buffer<ffi::U8> workspace, F32 d_out,
                   ffi::RemainingArgs leaves, ffi::ResultBuffer<ffi::U8> workspace_out, ffi::RemainingRets grads, int32_t d,
                   int32_t H, int32_t L, int32_t O, int32_t invariant_kind, int32_t use_window, int32_t precision, int32_t flags,
                   int32_t chunk_fields) {
  EnfDesc desc;
  int64_t xbs;
  if (auto e = describe(x, p, a, d, H, L, O, invariant_kind, use_window, precision, flags, chunk_fields, &desc, &xbs); e.failure()) return e;
  if (leaves.size() != ENF_NUM_WEIGHT_LEAVES || grads.size() != ENF_NUM_WEIGHT_LEAVES + 3)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "enf_xattn_bwd: expected 46 leaves in, 46 + 3 gradients out");
  if (workspace_out->typed_data() != workspace.typed_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument,
                      "enf_xattn_bwd: the workspace result must alias the workspace operand (input_output_aliases={4: 0}): the "
                      "library keys the forward state on the buffer address");
  EnfWeights w;
  EnfWeightGrads gw;
  if (auto e = collect_leaves(leaves, 0, &w, false); e.failure()) return e;
  if (auto e = collect_leaves(grads, 0, &gw, true); e.failure()) return e;
  auto dp = grads.get<F32>(ENF_NUM_WEIGHT_LEAVES), da = grads.get<F32>(ENF_NUM_WEIGHT_LEAVES + 1), ds = grads.get<F32>(ENF_NUM_WEIGHT_LEAVES + 2);
  if (!dp.has_value() || !da.has_value() || !ds.has_value()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "enf_xattn_bwd: latent gradients must be f32");
  const float* sg = (use_window && sigma.element_count() > 0) ? sigma.typed_data() : nullptr;
  float* dsg = (use_window && (*ds)->element_count() > 0) ? (*ds)->typed_data() : nullptr;
  return status(enf_xattn_bwd(&desc, &w, x.typed_data(), xbs, p.typed_data(), a.typed_data(), sg, d_out.typed_data(), &gw,
                              (*dp)->typed_data(), (*da)->typed_data(), dsg, workspace_out->typed_data(),
                              (size_t)workspace_out->element_count(), (enf_stream_t)stream),
                "enf_xattn_bwd");
}

// ---- latent ODE model (include/enf_ode_b200.h): ode_model.apply(params, (p, a, window)) and its vector-Jacobian product -------
// (experiments/fitting/ode_models/ponita_ode_g.py:229-257; call sites pde_trainer.py:381,433,582).
// Operands   fwd: p f32[B,Z,P], a f32[B,Z,L], <leaves in EnfOdeWeights order: 5 + 8 * layers + 2 (+1 with an orientation)>
//            bwd: p, a, workspace, g_dp, g_da, <the same leaves>
// Results    fwd: dp_dt f32[B,Z,P], da_dt f32[B,Z,L], workspace u8[enf_ode_workspace_bytes]
//            bwd: workspace (aliased to operand 2), <leaf gradients>, gp f32[B,Z,P], ga f32[B,Z,L]
// Attributes: hidden, basis, layers, widen, degree, Dx, invariant_kind (i32).
ffi::Error describe_ode(const F32& p, const F32& a, int32_t hidden, int32_t basis, int32_t layers, int32_t widen, int32_t degree,
                        int32_t Dx, int32_t invariant_kind, EnfOdeDesc* desc, bool* has_ori) {
  std::memset(desc, 0, sizeof(*desc));
  auto pd = p.dimensions(), ad = a.dimensions();
  if (pd.size() != 3 || ad.size() != 3 || ad[0] != pd[0] || ad[1] != pd[1])
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "enf_ode: expected p[B,Z,P], a[B,Z,L]");
  desc->B = (int32_t)pd[0]; desc->Z = (int32_t)pd[1]; desc->L = (int32_t)ad[2];
  desc->hidden = hidden; desc->basis = basis; desc->layers = layers; desc->widen = widen; desc->degree = degree;
  desc->Dx = Dx; desc->invariant_kind = invariant_kind;
  if (pd[2] != enf_pose_dim(invariant_kind, Dx)) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "enf_ode: pose width does not match the invariant");
  if (layers < 1 || layers > ENF_ODE_MAX_LAYERS) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "enf_ode: layers out of range");
  *has_ori = invariant_kind == ENF_INV_PONITA;
  return ffi::Error::Success();
}

// leaves arrive in EnfOdeWeights order: kb_w0, kb_b0, kb_w1, kb_b1, stem_w, 8 per layer, ro_scalar, ro_rel [, ro_ori]
template <class Struct, class Args>
ffi::Error collect_ode_leaves(Args& args, size_t first, int layers, bool has_ori, Struct* out) {
  std::memset(out, 0, sizeof(*out));
  // both structs are dense arrays of pointers (const float* / float*): fill them slot by slot
  auto** head = reinterpret_cast<float**>(out);
  auto** layer0 = head + 5;
  auto** tail = head + 5 + 8 * ENF_ODE_MAX_LAYERS;
  static_assert(sizeof(Struct) == sizeof(float*) * ENF_ODE_NUM_WEIGHT_LEAVES, "EnfOdeWeights / EnfOdeWeightGrads are dense pointer arrays");
  const int n = 5 + 8 * layers + 2 + (has_ori ? 1 : 0);
  for (int i = 0; i < n; ++i) {
    auto buf = args.template get<F32>(first + i);
    if (!buf.has_value()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "enf_ode: leaf " + std::to_string(i) + " is not f32");
    float* ptr;
    if constexpr (std::is_same_v<Args, ffi::RemainingRets>) ptr = (*buf)->typed_data();
    else ptr = const_cast<float*>(buf->typed_data());
    if (i < 5) head[i] = ptr;
    else if (i < 5 + 8 * layers) layer0[i - 5] = ptr;       // EnfOdeLayer[] is a dense array of 8 pointers each
    else tail[i - 5 - 8 * layers] = ptr;
  }
  return ffi::Error::Success();
}

ffi::Error OdeFwdImpl(cudaStream_t stream, F32 p, F32 a, ffi::RemainingArgs leaves, ffi::ResultBuffer<ffi::F32> dp_dt,
                      ffi::ResultBuffer<ffi::F32> da_dt, ffi::ResultBuffer<ffi::U8> workspace, int32_t hidden, int32_t basis,
                      int32_t layers, int32_t widen, int32_t degree, int32_t Dx, int32_t invariant_kind) {
  EnfOdeDesc desc;
  bool has_ori;
  if (auto e = describe_ode(p, a, hidden, basis, layers, widen, degree, Dx, invariant_kind, &desc, &has_ori); e.failure()) return e;
  if ((int)leaves.size() != 5 + 8 * layers + 2 + (has_ori ? 1 : 0)) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "enf_ode_fwd: wrong number of weight leaves");
  EnfOdeWeights w;
  if (auto e = collect_ode_leaves(leaves, 0, layers, has_ori, &w); e.failure()) return e;
  const size_t need = enf_ode_workspace_bytes(&desc);
  if (need == 0) return status(ENF_ERR_BAD_DESC, "enf_ode_workspace_bytes");
  if ((size_t)workspace->element_count() < need) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "enf_ode_fwd: workspace result too small");
  return status(enf_ode_fwd(&desc, &w, p.typed_data(), a.typed_data(), dp_dt->typed_data(), da_dt->typed_data(), workspace->typed_data(),
                            (size_t)workspace->element_count(), (enf_stream_t)stream), "enf_ode_fwd");
}

ffi::Error OdeBwdImpl(cudaStream_t stream, F32 p, F32 a, ffi::Buffer<ffi::U8> workspace, F32 g_dp, F32 g_da, ffi::RemainingArgs leaves,
                      ffi::ResultBuffer<ffi::U8> workspace_out, ffi::RemainingRets grads, int32_t hidden, int32_t basis, int32_t layers,
                      int32_t widen, int32_t degree, int32_t Dx, int32_t invariant_kind) {
  EnfOdeDesc desc;
  bool has_ori;
  if (auto e = describe_ode(p, a, hidden, basis, layers, widen, degree, Dx, invariant_kind, &desc, &has_ori); e.failure()) return e;
  const int n = 5 + 8 * layers + 2 + (has_ori ? 1 : 0);
  if ((int)leaves.size() != n || (int)grads.size() != n + 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "enf_ode_bwd: wrong number of leaves / gradients");
  if (workspace_out->typed_data() != workspace.typed_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "enf_ode_bwd: the workspace result must alias the workspace operand (input_output_aliases={2: 0})");
  EnfOdeWeights w;
  EnfOdeWeightGrads gw;
  if (auto e = collect_ode_leaves(leaves, 0, layers, has_ori, &w); e.failure()) return e;
  if (auto e = collect_ode_leaves(grads, 0, layers, has_ori, &gw); e.failure()) return e;
  auto gp = grads.get<F32>(n), ga = grads.get<F32>(n + 1);
  if (!gp.has_value() || !ga.has_value()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "enf_ode_bwd: latent gradients must be f32");
  return status(enf_ode_bwd(&desc, &w, p.typed_data(), a.typed_data(), g_dp.typed_data(), g_da.typed_data(), &gw, (*gp)->typed_data(),
                            (*ga)->typed_data(), workspace_out->typed_data(), (size_t)workspace_out->element_count(), (enf_stream_t)stream),
                "enf_ode_bwd");
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(EnfXattnFwd, FwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>()   // x
                                  .Arg<F32>()   // p
                                  .Arg<F32>()   // a
                                  .Arg<F32>()   // sigma
                                  .RemainingArgs()
                                  .Ret<F32>()                    // out
                                  .Ret<ffi::Buffer<ffi::U8>>()   // workspace
                                  .Attr<int32_t>("d")
                                  .Attr<int32_t>("H")
                                  .Attr<int32_t>("L")
                                  .Attr<int32_t>("O")
                                  .Attr<int32_t>("invariant_kind")
                                  .Attr<int32_t>("use_window")
                                  .Attr<int32_t>("precision")
                                  .Attr<int32_t>("flags")
                                  .Attr<int32_t>("chunk_fields"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(EnfXattnBwd, BwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>()                    // x
                                  .Arg<F32>()                    // p
                                  .Arg<F32>()                    // a
                                  .Arg<F32>()                    // sigma
                                  .Arg<ffi::Buffer<ffi::U8>>()   // workspace of the matching forward
                                  .Arg<F32>()                    // d_out
                                  .RemainingArgs()
                                  .Ret<ffi::Buffer<ffi::U8>>()   // workspace, aliased to operand 4
                                  .RemainingRets()
                                  .Attr<int32_t>("d")
                                  .Attr<int32_t>("H")
                                  .Attr<int32_t>("L")
                                  .Attr<int32_t>("O")
                                  .Attr<int32_t>("invariant_kind")
                                  .Attr<int32_t>("use_window")
                                  .Attr<int32_t>("precision")
                                  .Attr<int32_t>("flags")
                                  .Attr<int32_t>("chunk_fields"));

#define ENF_ODE_ATTRS() \
  .Attr<int32_t>("hidden").Attr<int32_t>("basis").Attr<int32_t>("layers").Attr<int32_t>("widen").Attr<int32_t>("degree") \
  .Attr<int32_t>("Dx").Attr<int32_t>("invariant_kind")

XLA_FFI_DEFINE_HANDLER_SYMBOL(EnfOdeFwd, OdeFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>()   // p
                                  .Arg<F32>()   // a
                                  .RemainingArgs()
                                  .Ret<F32>()                    // dp/dt
                                  .Ret<F32>()                    // da/dt
                                  .Ret<ffi::Buffer<ffi::U8>>()   // workspace
                                  ENF_ODE_ATTRS());

XLA_FFI_DEFINE_HANDLER_SYMBOL(EnfOdeBwd, OdeBwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>()                    // p
                                  .Arg<F32>()                    // a
                                  .Arg<ffi::Buffer<ffi::U8>>()   // workspace of the matching forward
                                  .Arg<F32>()                    // cotangent of dp/dt
                                  .Arg<F32>()                    // cotangent of da/dt
                                  .RemainingArgs()
                                  .Ret<ffi::Buffer<ffi::U8>>()   // workspace, aliased to operand 2
                                  .RemainingRets()
                                  ENF_ODE_ATTRS());

#else   // no jaxlib headers: nothing to build (see the header comment)

#if defined(__GNUC__)
#pragma message("enf_xla_ffi.cc: xla/ffi/api/ffi.h not found -- the jax.ffi shim is not built (jaxlib headers are required)")
#endif

#endif  // ENF_HAVE_XLA_FFI
