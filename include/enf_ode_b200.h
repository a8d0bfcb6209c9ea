/*
 * enf_ode_b200.h -- C ABI of the latent ODE model and its fixed-step solver, the step either side of the
 * ENF decoder in the ODE phase of david-knigge/enf-pde.  Same library (libenf_b200.so) and conventions as enf_b200.h:
 * plain pointers, caller-owned device memory, caller-provided workspace, enqueue-only, negative ENF_ERR_* on failure
 * (enf_last_error()), no CPU fallback.
 *
 * Replaces
 *   PonitaODEGen.__call__ / PonitaGen.__call__     experiments/fitting/ode_models/ponita_ode_g.py:140-198, 229-257
 *     SepGconv / ConvBlock / PolynomialFeatures    experiments/fitting/ode_models/ponita_ode_g.py:15-87
 *     self-attention invariants                    enf/steerable_attention/invariant/__init__.py:13-45 (get_sa_invariant)
 *   its reverse-mode derivative (jax.value_and_grad at experiments/fitting/trainers/pde_trainer.py:299,328)
 *   _solve_latent_ode (Euler / RK4 roll-out)       experiments/fitting/trainers/trainer_utils/solvers.py:73-162
 * for kernel_size = "global", global_pool = False, vec_num_out = 1 (what get_model_pde builds, experiments/fitting/__init__.py:48-61).
 */
#ifndef ENF_ODE_B200_H_
#define ENF_ODE_B200_H_

#include "enf_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ENF_ODE_MAX_LAYERS 8

/* Mirrors the `node:` block of the experiment YAMLs + the nef's invariant. */
typedef struct EnfOdeDesc {
  int32_t B;               /* latent sets in the call (signals x time steps)                       */
  int32_t Z;               /* latents per set                                                      */
  int32_t L;               /* cfg.nef.latent_dim (scalar_num_out)                                  */
  int32_t hidden;          /* cfg.node.num_hidden (multiple of 4)                                  */
  int32_t basis;           /* cfg.node.basis_dim                                                   */
  int32_t layers;          /* cfg.node.num_layers (1..ENF_ODE_MAX_LAYERS)                          */
  int32_t widen;           /* cfg.node.widening_factor                                             */
  int32_t degree;          /* cfg.node.degree: PolynomialFeatures emits degree + 1 tensor powers   */
  int32_t Dx;              /* cfg.nef.num_in                                                       */
  int32_t invariant_kind;  /* EnfInvariantKind of cfg.nef.invariant_type; the SELF-attention variant is used
                              (get_sa_invariant): ENF_INV_PONITA means Ponita2D (3 invariants)      */
  int32_t reserved[6];     /* must be 0                                                            */
} EnfOdeDesc;

/* One ConvBlock (`interaction_layers_<i>`): conv/kernel/kernel (basis,hidden), conv/bias (hidden), norm/scale, norm/bias,
 * linear_1 (hidden, widen*hidden) + bias, linear_2 (widen*hidden, hidden) + bias. */
typedef struct EnfOdeLayer {
  const float *conv_k, *conv_b, *ln_g, *ln_b, *l1_w, *l1_b, *l2_w, *l2_b;
} EnfOdeLayer;

/* ode_model.init(...)['params']['ponita'] as device pointers; Dense kernels (in, out) row-major as Flax stores them. */
typedef struct EnfOdeWeights {
  const float *kb_w0, *kb_b0;        /* kernel_basis/layers_1  (F, hidden), F = sum_k I^(k+1), k = 0..degree   */
  const float *kb_w1, *kb_b1;        /* kernel_basis/layers_3  (hidden, basis)                                  */
  const float *stem_w;               /* a_stem/kernel          (L, hidden), no bias                             */
  EnfOdeLayer layer[ENF_ODE_MAX_LAYERS];
  const float *ro_scalar;            /* readout_scalar/layers_0/kernel (hidden, L [+1 with an orientation])     */
  const float *ro_rel;               /* readout_vec_rel/kernel         (I + hidden, 1)                          */
  const float *ro_ori;               /* readout_vec_ori/kernel         (I + hidden, 1); NULL without orientation */
} EnfOdeWeights;

#define ENF_ODE_NUM_WEIGHT_LEAVES (5 + 8 * ENF_ODE_MAX_LAYERS + 3)

typedef struct EnfOdeLayerGrads {
  float *conv_k, *conv_b, *ln_g, *ln_b, *l1_w, *l1_b, *l2_w, *l2_b;
} EnfOdeLayerGrads;
typedef struct EnfOdeWeightGrads {
  float *kb_w0, *kb_b0, *kb_w1, *kb_b1, *stem_w;
  EnfOdeLayerGrads layer[ENF_ODE_MAX_LAYERS];
  float *ro_scalar, *ro_rel, *ro_ori;
} EnfOdeWeightGrads;

enum EnfOdeMethod { ENF_ODE_EULER = 0, ENF_ODE_RK4 = 1 };   /* cfg.node.method: 'euler' | 'rk4' (solvers.py:141-152) */

/* Bytes of workspace enf_ode_fwd / enf_ode_bwd / enf_ode_solve need for `desc` (0 on an invalid description). */
size_t enf_ode_workspace_bytes(const EnfOdeDesc* desc);

/* (dp/dt, da/dt) = ode_model.apply(params, (p, a, sigma)); d sigma/dt is identically 0 (ponita_ode_g.py:252-257) and is not
 * written.  p [B,Z,P_raw] raw poses (angles, not cos/sin), a [B,Z,L], dp_dt [B,Z,P_raw], da_dt [B,Z,L].  The workspace keeps
 * the forward state enf_ode_bwd needs. */
int enf_ode_fwd(const EnfOdeDesc* desc, const EnfOdeWeights* w, const float* p, const float* a, float* dp_dt, float* da_dt,
                void* workspace, size_t workspace_bytes, enf_stream_t stream);

/* Vector-Jacobian product of the call above, on the workspace of a matching enf_ode_fwd: cotangents g_dp_dt / g_da_dt ->
 * dW (every leaf overwritten; NULL skips the weight gradients), gp [B,Z,P_raw], ga [B,Z,L] (overwritten). */
int enf_ode_bwd(const EnfOdeDesc* desc, const EnfOdeWeights* w, const float* p, const float* a, const float* g_dp_dt,
                const float* g_da_dt, const EnfOdeWeightGrads* dW, float* gp, float* ga, void* workspace,
                size_t workspace_bytes, enf_stream_t stream);

/* Forward roll-out _solve_latent_ode(f, (p0, a0, sigma), t0, tf, h, method) with num_steps = int((tf - t0) / h):
 * p_traj [B, num_steps + 1, Z, P_raw], a_traj [B, num_steps + 1, Z, L] (time-major swap of solvers.py:157-160 applied),
 * entry 0 = the initial state.  sigma is constant along the trajectory (its derivative is 0): the caller broadcasts it. */
int enf_ode_solve(const EnfOdeDesc* desc, const EnfOdeWeights* w, const float* p0, const float* a0, int32_t num_steps, float h,
                  int32_t method, float* p_traj, float* a_traj, void* workspace, size_t workspace_bytes, enf_stream_t stream);

/* ---- MLPODE (cfg.node.name == "mlp"; experiments/fitting/ode_models/mlp_ode.py:5-42): two 4-layer MLPs (Dense -> gelu x3 -> Dense) on
 * concat(p, a - 1) per latent; derivative_p_pos has 2 * vec_num_out = 2 components (the model is written for 2-D positional poses,
 * P == 2), derivative_a has L.  Not equivariant: the ablation baseline of the reference's factory (experiments/fitting/__init__.py:41-47). */
typedef struct EnfMlpOdeDesc {
  int32_t B, Z;            /* latent sets, latents per set                         */
  int32_t P;               /* raw pose width (must be 2: the output is added to p) */
  int32_t L;               /* latent_dim (scalar_num_out)                          */
  int32_t hidden;          /* cfg.node.num_hidden                                  */
  int32_t reserved[3];     /* must be 0                                            */
} EnfMlpOdeDesc;
/* mlp_a/layers_{0,2,4,6} and mlp_p/layers_{0,2,4,6}: kernels (in, out) row-major, biases */
typedef struct EnfMlpOdeWeights { const float *a_w[4], *a_b[4], *p_w[4], *p_b[4]; } EnfMlpOdeWeights;
typedef struct EnfMlpOdeWeightGrads { float *a_w[4], *a_b[4], *p_w[4], *p_b[4]; } EnfMlpOdeWeightGrads;

size_t enf_mlpode_workspace_bytes(const EnfMlpOdeDesc* desc);
int enf_mlpode_fwd(const EnfMlpOdeDesc* desc, const EnfMlpOdeWeights* w, const float* p, const float* a, float* dp_dt, float* da_dt,
                   void* workspace, size_t workspace_bytes, enf_stream_t stream);
/* VJP on the workspace of a matching enf_mlpode_fwd; dW NULL skips the weight gradients; every output is overwritten. */
int enf_mlpode_bwd(const EnfMlpOdeDesc* desc, const EnfMlpOdeWeights* w, const float* p, const float* a, const float* g_dp_dt,
                   const float* g_da_dt, const EnfMlpOdeWeightGrads* dW, float* gp, float* ga, void* workspace, size_t workspace_bytes,
                   enf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ENF_ODE_B200_H_ */
