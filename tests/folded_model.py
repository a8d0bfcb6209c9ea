"""Design-validation model of the FOLDED algorithm the CUDA kernels implement (test infrastructure).

This is the blueprint of `enf_pde_b200/csrc`: the same stages, the same folded quantities and the
same hand-derived backward formulas, written with torch float64 tensor ops so that the algebra
(DESIGN.md section "Folds") is checked against the unfused oracle's autograd on CPU before and
independently of any CUDA.  Stage names match the kernels:

  W  weights_fold      A_q, c_q, Wp, bp, W2g, b2g, M2g, c2g                         (per call)
  L  latents_fold      pose record Lam, ahat, k, v0, U, kappa, Weff, beff, W3, b3   (per (b, z))
  X  query_features    xi                                                            (per (b, c))
  P  pairs_fwd / bwd   the fused per-(query, latent) chain + softmax over latents    (hot kernels)
  Q  tail_fwd / bwd    M2g, out_proj, block FFN, decode MLP                          (per query)

Invariants are expressed in "bilinear record" form: u_r = post_r(row_r(Lam[z], xi[c])) with
row kinds DOT (Lam_r . xi) and SQDIST (sum_{i<n} (Lam_r,i - xi_i)^2), so the hot kernels are
agnostic of the invariant type and never evaluate per-pair trigonometry of raw angles.
"""
import math

import torch

from oracle import enf_ref as R

F_XI = 8        # width of the per-query feature record xi
R_LAM = 7       # rows of the per-latent record Lam: up to 6 invariant rows + 1 window row

ROW_DOT, ROW_SQDIST, ROW_SQDIST_SQRT = 0, 1, 2
WIN_NONE, WIN_NP, WIN_PER, WIN_SPH = 0, 1, 2, 3


def gelu(x):
    return R.gelu_tanh(x)


def gelu_grad(x):
    c = math.sqrt(2.0 / math.pi)
    t = torch.tanh(c * (x + 0.044715 * x ** 3))
    return 0.5 * (1 + t) + 0.5 * x * (1 - t * t) * c * (1 + 3 * 0.044715 * x * x)


def ln_core(x, eps=1e-6):
    mu = x.mean(-1, keepdim=True)
    var = torch.clamp((x * x).mean(-1, keepdim=True) - mu * mu, min=0)
    rstd = torch.rsqrt(var + eps)
    return (x - mu) * rstd, rstd


def ln_core_bwd(dxhat, xhat, rstd):
    """backward of xhat = (x - mean) * rstd (no affine) given cotangent of xhat."""
    m1 = dxhat.mean(-1, keepdim=True)
    m2 = (dxhat * xhat).mean(-1, keepdim=True)
    return rstd * (dxhat - m1 - xhat * m2)


# ------------------------------------------------------------------------------------------------
# invariant records
# ------------------------------------------------------------------------------------------------

def record_layout(cfg):
    """(I, row kinds list (len I), window kind, window row index or -1, n for SQDIST)."""
    t, n = cfg.invariant_type, cfg.num_in
    win = {"np": WIN_NP, "per": WIN_PER, "sph": WIN_SPH}[R.INVARIANTS[t]["window"]] if cfg.use_gaussian_window else WIN_NONE
    I = cfg.inv_dim
    rows = [ROW_DOT] * I
    nsq = n
    if t == "norm_rel_pos":
        rows = [ROW_SQDIST_SQRT]
    if t == "ponita":
        nsq = 2
    return I, rows, win, nsq


def query_features(cfg, x):
    """xi (B, C, F_XI)."""
    t, n = cfg.invariant_type, cfg.num_in
    B, C = x.shape[:2]
    xi = torch.zeros(B, C, F_XI, dtype=x.dtype)
    if t in ("rel_pos", "norm_rel_pos", "abs_pos"):
        xi[..., :n] = x[..., :n]
        xi[..., n] = 1.0
    elif t == "rel_pos_periodic":
        xi[..., 0] = torch.cos(math.pi * x[..., 0]); xi[..., 1] = torch.sin(math.pi * x[..., 0])
        xi[..., 2] = torch.cos(math.pi * x[..., 1]); xi[..., 3] = torch.sin(math.pi * x[..., 1])
    elif t == "ponita":
        xi[..., 0:2] = x[..., 0:2]
        xi[..., 2] = 1.0
    elif t == "polar_periodic":
        xi[..., 0:3] = R._sph_unit(x[..., 0], x[..., 1])
    elif t in ("latitude_periodic", "ball_lat"):
        xi[..., 0] = x[..., 1]
        xi[..., 1] = torch.cos(x[..., 0]); xi[..., 2] = torch.sin(x[..., 0])
        xi[..., 3] = 1.0
        xi[..., 4:7] = R._sph_unit(x[..., 0], x[..., 1])
        if t == "ball_lat":
            xi[..., 7] = x[..., 2]
    elif t == "ball":
        xi[..., 0:3] = R._sph_unit(x[..., 0], x[..., 1])
        xi[..., 3] = x[..., 2]
        xi[..., 4] = 1.0
    return xi


def latent_record(cfg, p):
    """Lam (B, Z, R_LAM, F_XI) from RAW poses p (angles not yet embedded). Differentiable in p (torch ops)."""
    t, n = cfg.invariant_type, cfg.num_in
    B, Z = p.shape[:2]
    I = cfg.inv_dim
    rows = [[torch.zeros(B, Z, dtype=p.dtype) for _ in range(F_XI)] for _ in range(R_LAM)]
    one = torch.ones(B, Z, dtype=p.dtype)
    W = I     # window row index
    if t == "rel_pos":
        for i in range(n):
            rows[i][i] = one; rows[i][n] = -p[..., i]
            rows[W][i] = p[..., i]
    elif t == "norm_rel_pos":
        for i in range(n):
            rows[0][i] = p[..., i]; rows[W][i] = p[..., i]
    elif t == "abs_pos":
        for i in range(n):
            rows[i][i] = one; rows[W][i] = p[..., i]
    elif t == "rel_pos_periodic":
        for i in range(2):
            cp, sp = torch.cos(math.pi * p[..., i]), torch.sin(math.pi * p[..., i])
            rows[i][2 * i] = cp; rows[i][2 * i + 1] = sp                 # cos(pi (p - x))
            rows[2 + i][2 * i] = sp; rows[2 + i][2 * i + 1] = -cp        # sin(pi (p - x))
    elif t == "ponita":
        p0, p1, th = p[..., 0], p[..., 1], p[..., 2]
        o0, o1 = torch.cos(th), torch.sin(th)
        rows[0][0] = o0; rows[0][1] = o1; rows[0][2] = -(p0 * o0 + p1 * o1)
        rows[1][0] = -o1; rows[1][1] = o0; rows[1][2] = p0 * o1 - p1 * o0
        rows[W][0] = p0; rows[W][1] = p1
    elif t == "polar_periodic":
        ph = R._sph_unit(p[..., 0], p[..., 1])
        for j in range(3):
            rows[0][j] = ph[..., j]
    elif t in ("latitude_periodic", "ball_lat"):
        phi, th = p[..., 0], p[..., 1]
        rows[0][0] = one
        rows[1][3] = th
        rows[2][1] = torch.cos(phi); rows[2][2] = torch.sin(phi)
        rows[3][1] = -torch.sin(phi); rows[3][2] = torch.cos(phi)
        if t == "ball_lat":
            rows[4][7] = one
            rows[5][3] = p[..., 3]
        ph = R._sph_unit(phi, th)
        for j in range(3):
            rows[W][4 + j] = ph[..., j]
    elif t == "ball":
        al, be, ga, rp = p[..., 0], p[..., 1], p[..., 2], p[..., 3]
        ca, sa, cb, sb, cg, sg = torch.cos(al), torch.sin(al), torch.cos(be), torch.sin(be), torch.cos(ga), torch.sin(ga)
        Rm = [[ca * cb, ca * sb * sg - sa * cg, ca * sb * cg + sa * sg],
              [sa * cb, sa * sb * sg + ca * cg, sa * sb * cg - ca * sg],
              [-sb, cb * sg, cb * cg]]
        for i in range(3):
            for j in range(3):
                rows[i][j] = Rm[i][j]
        rows[3][3] = one
        rows[4][4] = rp
        ph = R._sph_unit(al, be)
        for j in range(3):
            rows[W][j] = ph[..., j]
    return torch.stack([torch.stack(r, dim=-1) for r in rows], dim=-2)


def pair_invariants(cfg, xi, Lam, sigma):
    """u (B,C,Z,I), w (B,C,Z), plus what the backward needs."""
    I, rows, win, nsq = record_layout(cfg)
    dot = torch.einsum("bcf,bzrf->bczr", xi, Lam)                         # all DOT rows
    sq = ((Lam[:, None, :, :, :nsq] - xi[:, :, None, None, :nsq]) ** 2).sum(-1)   # all SQDIST rows
    us = []
    for r in range(I):
        if rows[r] == ROW_DOT:
            us.append(dot[..., r])
        else:
            us.append(torch.sqrt(sq[..., r]))
    u = torch.stack(us, -1)
    s = sigma[:, None, :, 0] if sigma is not None else None
    if win == WIN_NONE:
        w = torch.zeros_like(dot[..., 0])
    elif win == WIN_NP:
        w = -sq[..., I] / s ** 2
    elif win == WIN_PER:
        w = (u[..., 0] ** 2 + u[..., 1] ** 2) / s ** 2
    else:
        c = u[..., 0] if cfg.invariant_type == "polar_periodic" else dot[..., I]
        cl = torch.clamp(c, -1 + 1e-6, 1 - 1e-6)
        w = torch.exp(-torch.arccos(cl) ** 2 / (2 * s ** 2))
    return u, w, dot, sq


def pair_invariants_bwd(cfg, xi, Lam, sigma, u, w, dot, sq, du, dw):
    """Given du (B,C,Z,I), dw (B,C,Z): returns dLam (B,Z,R,F) and dsigma (B,Z,1) (sum over queries)."""
    I, rows, win, nsq = record_layout(cfg)
    B, C, Z = dw.shape
    dq = torch.zeros(B, C, Z, R_LAM, dtype=xi.dtype)     # cotangent of each row's raw value (dot or sqdist)
    du = du.clone()
    dsigma = torch.zeros(B, Z, 1, dtype=xi.dtype)
    if win != WIN_NONE:
        s = sigma[:, None, :, 0]
        if win == WIN_NP:
            dq[..., I] = -dw / s ** 2
            dsigma[..., 0] = (dw * (-2.0 * w / s)).sum(1)
        elif win == WIN_PER:
            du[..., 0] += dw * 2 * u[..., 0] / s ** 2
            du[..., 1] += dw * 2 * u[..., 1] / s ** 2
            dsigma[..., 0] = (dw * (-2.0 * w / s)).sum(1)
        else:
            c = u[..., 0] if cfg.invariant_type == "polar_periodic" else dot[..., I]
            inside = (c > -1 + 1e-6) & (c < 1 - 1e-6)
            cl = torch.clamp(c, -1 + 1e-6, 1 - 1e-6)
            ac = torch.arccos(cl)
            dc = torch.where(inside, dw * w * ac / (s ** 2 * torch.sqrt(1 - cl * cl)), torch.zeros_like(c))
            if cfg.invariant_type == "polar_periodic":
                du[..., 0] += dc
            else:
                dq[..., I] = dc
            dsigma[..., 0] = (dw * w * ac ** 2 / s ** 3).sum(1)
    for r in range(I):
        if rows[r] == ROW_DOT:
            dq[..., r] = du[..., r]
        else:
            ur = u[..., r]
            dq[..., r] = torch.where(ur > 0, du[..., r] / (2 * ur), torch.zeros_like(ur))
    # accumulate over queries.  DOT rows: dLam_r += dq_r * xi.  SQDIST rows: dLam_r,i += 2 dq_r (Lam_r,i - xi_i)
    dLam = torch.zeros_like(Lam)
    acc_xi = torch.einsum("bczr,bcf->bzrf", dq, xi)       # sum_c dq * xi   (what the kernel accumulates)
    acc_1 = dq.sum(1)                                      # sum_c dq
    for r in range(R_LAM):
        kind = rows[r] if r < I else (ROW_SQDIST if (r == I and win == WIN_NP) else ROW_DOT)
        if kind == ROW_DOT:
            dLam[:, :, r, :] = acc_xi[:, :, r, :]
        else:
            dLam[:, :, r, :nsq] = 2 * (Lam[:, :, r, :nsq] * acc_1[:, :, r, None] - acc_xi[:, :, r, :nsq])
    return dLam, dsigma


# ------------------------------------------------------------------------------------------------
# the folded model
# ------------------------------------------------------------------------------------------------

class Folded:
    def __init__(self, cfg, params):
        self.cfg = cfg
        P = params["params"]
        blk = P["cross_attention_blocks_0"]
        at = blk["attn"]
        g = lambda node, *ks: (node[ks[0]] if len(ks) == 1 else g(node[ks[0]], *ks[1:])).detach()
        self.w = dict(
            stem_w=g(P, "latent_stem", "kernel"), stem_b=g(P, "latent_stem", "bias"),
            ln_attn_g=g(blk, "layer_norm_attn", "scale"), ln_attn_b=g(blk, "layer_norm_attn", "bias"),
            q_omega=g(at, "invariant_embedding_query", "encoding", "coefficients"),
            q_w1=g(at, "invariant_embedding_query", "layers_0", "linear", "kernel"),
            q_b1=g(at, "invariant_embedding_query", "layers_0", "linear", "bias"),
            q_wf=g(at, "invariant_embedding_query", "linear_final", "kernel"),
            q_bf=g(at, "invariant_embedding_query", "linear_final", "bias"),
            v_omega=g(at, "invariant_embedding_value", "encoding", "coefficients"),
            v_w1=g(at, "invariant_embedding_value", "layers_0", "linear", "kernel"),
            v_b1=g(at, "invariant_embedding_value", "layers_0", "linear", "bias"),
            v_wf=g(at, "invariant_embedding_value", "linear_final", "kernel"),
            v_bf=g(at, "invariant_embedding_value", "linear_final", "bias"),
            wq=g(at, "inv_emb_to_q", "kernel"), bq=g(at, "inv_emb_to_q", "bias"),
            wk=g(at, "a_to_k", "kernel"), bk=g(at, "a_to_k", "bias"),
            wv=g(at, "a_to_v", "kernel"), bv=g(at, "a_to_v", "bias"),
            fv_w1=g(at, "inv_emb_to_v", "Dense_0", "kernel"), fv_b1=g(at, "inv_emb_to_v", "Dense_0", "bias"),
            fv_g=g(at, "inv_emb_to_v", "LayerNorm_0", "scale"), fv_beta=g(at, "inv_emb_to_v", "LayerNorm_0", "bias"),
            fv_w2=g(at, "inv_emb_to_v", "Dense_1", "kernel"), fv_b2=g(at, "inv_emb_to_v", "Dense_1", "bias"),
            mx_w1=g(at, "inv_emb_cond_mixer", "Dense_0", "kernel"), mx_b1=g(at, "inv_emb_cond_mixer", "Dense_0", "bias"),
            mx_g=g(at, "inv_emb_cond_mixer", "LayerNorm_0", "scale"), mx_beta=g(at, "inv_emb_cond_mixer", "LayerNorm_0", "bias"),
            mx_w2=g(at, "inv_emb_cond_mixer", "Dense_1", "kernel"), mx_b2=g(at, "inv_emb_cond_mixer", "Dense_1", "bias"),
            wo=g(at, "out_proj", "kernel"), bo=g(at, "out_proj", "bias"),
            fb_w1=g(blk, "pointwise_ffn", "Dense_0", "kernel"), fb_b1=g(blk, "pointwise_ffn", "Dense_0", "bias"),
            fb_g=g(blk, "pointwise_ffn", "LayerNorm_0", "scale"), fb_beta=g(blk, "pointwise_ffn", "LayerNorm_0", "bias"),
            fb_w2=g(blk, "pointwise_ffn", "Dense_1", "kernel"), fb_b2=g(blk, "pointwise_ffn", "Dense_1", "bias"),
            m0_w=g(P, "out_proj", "layers_0", "kernel"), m0_b=g(P, "out_proj", "layers_0", "bias"),
            m1_w=g(P, "out_proj", "layers_2", "kernel"), m1_b=g(P, "out_proj", "layers_2", "bias"),
            m2_w=g(P, "out_proj", "layers_4", "kernel"), m2_b=g(P, "out_proj", "layers_4", "bias"),
        )

    # ---- stage W ---------------------------------------------------------------------------
    def weights_fold(self):
        w = self.w
        f = {}
        f["A_q"] = w["q_wf"] @ w["wq"]                       # (d, Hd)
        f["c_q"] = w["q_bf"] @ w["wq"] + w["bq"]             # (Hd)
        f["Wp"] = w["v_wf"] @ w["fv_w1"]                     # (d, d)
        f["bp"] = w["v_bf"] @ w["fv_w1"] + w["fv_b1"]
        f["W2g"] = w["fv_g"][:, None] * w["fv_w2"]           # (d, 2Hd)
        f["b2g"] = w["fv_beta"] @ w["fv_w2"] + w["fv_b2"]
        f["M2g"] = w["mx_g"][:, None] * w["mx_w2"]           # (d, d)
        f["c2g"] = w["mx_beta"] @ w["mx_w2"] + w["mx_b2"]
        # tail fold: mixer Dense_1 (per head), out_proj and the block FFN's Dense_0 are three linear maps in a row
        H, d = self.cfg.num_heads, self.cfg.num_hidden
        f["P1"] = torch.cat([f["M2g"] @ w["wo"][h * d:(h + 1) * d] for h in range(H)], 0)      # blockdiag(M2g) wo  (Hd, Hd)
        f["b1"] = f["c2g"].repeat(H) @ w["wo"] + w["bo"]
        f["W_A"] = f["P1"] @ w["fb_w1"]
        f["b_A"] = f["b1"] @ w["fb_w1"] + w["fb_b1"]
        self.f = f
        return f

    # ---- stage L ---------------------------------------------------------------------------
    def latents_fold(self, p, a):
        cfg, w, f = self.cfg, self.w, self.f
        H, d = cfg.num_heads, cfg.num_hidden
        B, Z = a.shape[:2]
        L = {}
        L["Lam"] = latent_record(cfg, p)
        a0 = a @ w["stem_w"] + w["stem_b"]
        ahat_core, rstd = ln_core(a0)
        ahat = ahat_core * w["ln_attn_g"] + w["ln_attn_b"]
        k = (ahat @ w["wk"] + w["bk"]).reshape(B, Z, H, d)
        v0 = (ahat @ w["wv"] + w["bv"]).reshape(B, Z, H, d)
        A_q = f["A_q"].reshape(d, H, d)
        L["U"] = torch.einsum("ihj,bzhj->bzhi", A_q, k)                       # (B,Z,H,d)
        L["kappa"] = torch.einsum("hj,bzhj->bzh", f["c_q"].reshape(H, d), k)  # (B,Z,H)
        W2gam = f["W2g"][:, :H * d].reshape(d, H, d)
        W2bet = f["W2g"][:, H * d:].reshape(d, H, d)
        b2gam = f["b2g"][:H * d].reshape(H, d)
        b2bet = f["b2g"][H * d:].reshape(H, d)
        Weff = W2gam.permute(1, 0, 2)[None, None] * v0[:, :, :, None, :] + W2bet.permute(1, 0, 2)[None, None]  # (B,Z,H,d,d)
        beff = v0 * (1 + b2gam) + b2bet                                       # (B,Z,H,d)
        L["W3"] = Weff @ w["mx_w1"]                                           # (B,Z,H,d,d)
        L["b3"] = beff @ w["mx_w1"] + w["mx_b1"]                              # (B,Z,H,d)
        L.update(a0=a0, ahat_core=ahat_core, rstd=rstd, ahat=ahat, k=k, v0=v0, Weff=Weff, beff=beff)
        self.L = L
        return L

    # ---- stage P ---------------------------------------------------------------------------
    def _rff(self, u, omega):
        proj = 2 * math.pi * (u @ omega)
        return torch.cat([torch.sin(proj), torch.cos(proj)], -1)

    def pairs_fwd(self, xi, sigma):
        cfg, w, f, L = self.cfg, self.w, self.f, self.L
        H, d = cfg.num_heads, cfg.num_hidden
        scale = 1.0 / math.sqrt(d)
        S = {}
        u, win, dot, sq = pair_invariants(cfg, xi, L["Lam"], sigma)
        gq = self._rff(u, w["q_omega"]); gv = self._rff(u, w["v_omega"])
        h1q = torch.relu(gq @ w["q_w1"] + w["q_b1"])
        s = scale * (torch.einsum("bczi,bzhi->bczh", h1q, L["U"]) + L["kappa"][:, None]) + win[..., None]
        h1v = torch.relu(gv @ w["v_w1"] + w["v_b1"])
        tpre = h1v @ f["Wp"] + f["bp"]
        that, t_rstd = ln_core(gelu(tpre))
        mpre = torch.einsum("bczi,bzhij->bczhj", that, L["W3"]) + L["b3"][:, None]
        n, n_rstd = ln_core(gelu(mpre))
        m = s.max(dim=2, keepdim=True).values
        e = torch.exp(s - m)
        l = e.sum(dim=2, keepdim=True)
        att = e / l
        nbar = torch.einsum("bczh,bczhj->bchj", att, n)                       # (B,C,H,d)
        S.update(u=u, win=win, dot=dot, sq=sq, gq=gq, gv=gv, h1q=h1q, h1v=h1v, tpre=tpre, that=that, t_rstd=t_rstd,
                 mpre=mpre, n=n, n_rstd=n_rstd, att=att, nbar=nbar, lse=(m + torch.log(l))[:, :, 0])
        self.S = S
        return nbar

    # ---- stage Q ---------------------------------------------------------------------------
    def tail_fwd(self, nbar):
        cfg, w, f = self.cfg, self.w, self.f
        H, d = cfg.num_heads, cfg.num_hidden
        B, C = nbar.shape[:2]
        T = {}
        e1 = nbar.reshape(B, C, H * d) @ f["W_A"] + f["b_A"]
        e3c, e_rstd = ln_core(gelu(e1))
        e3 = e3c * w["fb_g"] + w["fb_beta"]
        fo = e3 @ w["fb_w2"] + w["fb_b2"]
        o1p = gelu(fo) @ w["m0_w"] + w["m0_b"]
        o2p = gelu(o1p) @ w["m1_w"] + w["m1_b"]
        out = gelu(o2p) @ w["m2_w"] + w["m2_b"]
        T.update(e1=e1, e3c=e3c, e_rstd=e_rstd, e3=e3, fo=fo, o1p=o1p, o2p=o2p)
        self.T = T
        return out

    def tail_bwd(self, nbar, d_out):
        cfg, w, f, T = self.cfg, self.w, self.f, self.T
        H, d = cfg.num_heads, cfg.num_hidden
        B, C = nbar.shape[:2]
        G = {}
        fl = lambda t: t.reshape(-1, t.shape[-1])
        o2 = gelu(T["o2p"])
        G["m2_w"] = fl(o2).T @ fl(d_out); G["m2_b"] = fl(d_out).sum(0)
        do2p = (d_out @ w["m2_w"].T) * gelu_grad(T["o2p"])
        G["m1_w"] = fl(gelu(T["o1p"])).T @ fl(do2p); G["m1_b"] = fl(do2p).sum(0)
        do1p = (do2p @ w["m1_w"].T) * gelu_grad(T["o1p"])
        G["m0_w"] = fl(gelu(T["fo"])).T @ fl(do1p); G["m0_b"] = fl(do1p).sum(0)
        dfo = (do1p @ w["m0_w"].T) * gelu_grad(T["fo"])
        G["fb_w2"] = fl(T["e3"]).T @ fl(dfo); G["fb_b2"] = fl(dfo).sum(0)
        de3 = dfo @ w["fb_w2"].T
        G["fb_g"] = fl(de3 * T["e3c"]).sum(0); G["fb_beta"] = fl(de3).sum(0)
        de2 = ln_core_bwd(de3 * w["fb_g"], T["e3c"], T["e_rstd"])
        de1 = de2 * gelu_grad(T["e1"])
        Gf = {}
        Gf["W_A"] = fl(nbar.reshape(B, C, H * d)).T @ fl(de1); Gf["b_A"] = fl(de1).sum(0)
        dnbar = (de1 @ f["W_A"].T).reshape(B, C, H, d)
        return dnbar, G, Gf

    # ---- stage P backward --------------------------------------------------------------------
    def pairs_bwd(self, xi, sigma, dnbar):
        cfg, w, f, L, S = self.cfg, self.w, self.f, self.L, self.S
        H, d = cfg.num_heads, cfg.num_hidden
        scale = 1.0 / math.sqrt(d)
        G, Gf, GL = {}, {}, {}
        fl = lambda t: t.reshape(-1, t.shape[-1])
        att, n = S["att"], S["n"]
        Dd = (dnbar * S["nbar"]).sum(-1)                                           # (B,C,H)
        ds = att * (torch.einsum("bchj,bczhj->bczh", dnbar, n) - Dd[:, :, None])   # (B,C,Z,H)
        dn = att[..., None] * dnbar[:, :, None]                                    # (B,C,Z,H,d)
        dmpre = ln_core_bwd(dn, n, S["n_rstd"]) * gelu_grad(S["mpre"])
        GL["W3"] = torch.einsum("bczi,bczhj->bzhij", S["that"], dmpre)
        GL["b3"] = dmpre.sum(1)
        dthat = torch.einsum("bczhj,bzhij->bczi", dmpre, L["W3"])
        dtpre = ln_core_bwd(dthat, S["that"], S["t_rstd"]) * gelu_grad(S["tpre"])
        Gf["Wp"] = fl(S["h1v"]).T @ fl(dtpre); Gf["bp"] = fl(dtpre).sum(0)
        dzv = (dtpre @ f["Wp"].T) * (S["h1v"] > 0)
        G["v_w1"] = fl(S["gv"]).T @ fl(dzv); G["v_b1"] = fl(dzv).sum(0)
        dgv = dzv @ w["v_w1"].T
        dzq = scale * torch.einsum("bczh,bzhi->bczi", ds, L["U"]) * (S["h1q"] > 0)
        GL["U"] = scale * torch.einsum("bczh,bczi->bzhi", ds, S["h1q"])
        GL["kappa"] = scale * ds.sum(1)
        G["q_w1"] = fl(S["gq"]).T @ fl(dzq); G["q_b1"] = fl(dzq).sum(0)
        dgq = dzq @ w["q_w1"].T
        hd = d // 2

        def rff_bwd(g, dg, omega):
            sin, cos = g[..., :hd], g[..., hd:]
            dproj = cos * dg[..., :hd] - sin * dg[..., hd:]
            return 2 * math.pi * (dproj @ omega.T)

        du = rff_bwd(S["gq"], dgq, w["q_omega"]) + rff_bwd(S["gv"], dgv, w["v_omega"])
        dw = ds.sum(-1)
        GL["Lam"], GL["sigma"] = pair_invariants_bwd(cfg, xi, L["Lam"], sigma, S["u"], S["win"], S["dot"], S["sq"], du, dw)
        return G, Gf, GL

    # ---- stage L backward --------------------------------------------------------------------
    def latents_bwd(self, p, a, GL):
        cfg, w, f, L = self.cfg, self.w, self.f, self.L
        H, d = cfg.num_heads, cfg.num_hidden
        B, Z = a.shape[:2]
        G, Gf = {}, {}
        fl = lambda t: t.reshape(-1, t.shape[-1])
        # W3 = Weff @ M1 ; b3 = beff @ M1 + c1
        G["mx_w1"] = fl(L["Weff"]).T @ fl(GL["W3"]) + fl(L["beff"]).T @ fl(GL["b3"])
        G["mx_b1"] = fl(GL["b3"]).sum(0)
        dWeff = GL["W3"] @ w["mx_w1"].T                       # (B,Z,H,d,d)
        dbeff = GL["b3"] @ w["mx_w1"].T                       # (B,Z,H,d)
        W2gam = f["W2g"][:, :H * d].reshape(d, H, d)
        b2gam = f["b2g"][:H * d].reshape(H, d)
        v0 = L["v0"]
        dW2bet = dWeff.sum((0, 1)).permute(1, 0, 2).reshape(d, H * d)
        dW2gam = (dWeff * v0[:, :, :, None, :]).sum((0, 1)).permute(1, 0, 2).reshape(d, H * d)
        Gf["W2g"] = torch.cat([dW2gam, dW2bet], dim=1)
        Gf["b2g"] = torch.cat([(dbeff * v0).sum((0, 1)).reshape(-1), dbeff.sum((0, 1)).reshape(-1)])
        dv0 = torch.einsum("ihj,bzhij->bzhj", W2gam, dWeff) + dbeff * (1 + b2gam)
        # U = A_q[:, h] k_h ; kappa = c_q[h] . k_h
        A_q = f["A_q"].reshape(d, H, d)
        c_q = f["c_q"].reshape(H, d)
        dk = torch.einsum("ihj,bzhi->bzhj", A_q, GL["U"]) + GL["kappa"][..., None] * c_q
        Gf["A_q"] = torch.einsum("bzhi,bzhj->ihj", GL["U"], L["k"]).reshape(d, H * d)
        Gf["c_q"] = (GL["kappa"][..., None] * L["k"]).sum((0, 1)).reshape(-1)
        dk2, dv2 = dk.reshape(B, Z, H * d), dv0.reshape(B, Z, H * d)
        G["wk"] = fl(L["ahat"]).T @ fl(dk2); G["bk"] = fl(dk2).sum(0)
        G["wv"] = fl(L["ahat"]).T @ fl(dv2); G["bv"] = fl(dv2).sum(0)
        dahat = dk2 @ w["wk"].T + dv2 @ w["wv"].T
        G["ln_attn_g"] = fl(dahat * L["ahat_core"]).sum(0); G["ln_attn_b"] = fl(dahat).sum(0)
        da0 = ln_core_bwd(dahat * w["ln_attn_g"], L["ahat_core"], L["rstd"])
        G["stem_w"] = fl(a).T @ fl(da0); G["stem_b"] = fl(da0).sum(0)
        da = da0 @ w["stem_w"].T
        # pose record: vector-Jacobian product of latent_record (closed forms live in the CUDA kernel;
        # here autograd of the same torch expression is the statement of it)
        pr = p.detach().clone().requires_grad_(True)
        Lam = latent_record(cfg, pr)
        dp, = torch.autograd.grad(Lam, pr, grad_outputs=GL["Lam"], allow_unused=True)
        if dp is None:
            dp = torch.zeros_like(p)
        return dp, da, G, Gf

    # ---- stage W backward --------------------------------------------------------------------
    def weights_unfold_bwd(self, Gf):
        w, f = self.w, self.f
        H, d = self.cfg.num_heads, self.cfg.num_hidden
        G = {}
        # tail fold: W_A = P1 fb_w1, b_A = b1 fb_w1 + fb_b1, P1 = blockdiag(M2g) wo, b1 = tile(c2g) wo + bo
        G["fb_w1"] = f["P1"].T @ Gf["W_A"] + torch.outer(f["b1"], Gf["b_A"])
        G["fb_b1"] = Gf["b_A"]
        dP1 = Gf["W_A"] @ w["fb_w1"].T
        db1 = Gf["b_A"] @ w["fb_w1"].T
        G["wo"] = torch.cat([f["M2g"].T @ dP1[h * d:(h + 1) * d] for h in range(H)], 0) + torch.outer(f["c2g"].repeat(H), db1)
        G["bo"] = db1
        Gf["M2g"] = sum(dP1[h * d:(h + 1) * d] @ w["wo"][h * d:(h + 1) * d].T for h in range(H))
        Gf["c2g"] = (db1 @ w["wo"].T).reshape(H, d).sum(0)
        G["q_wf"] = Gf["A_q"] @ w["wq"].T
        G["wq"] = w["q_wf"].T @ Gf["A_q"] + torch.outer(w["q_bf"], Gf["c_q"])
        G["q_bf"] = Gf["c_q"] @ w["wq"].T
        G["bq"] = Gf["c_q"]
        G["v_wf"] = Gf["Wp"] @ w["fv_w1"].T
        G["fv_w1"] = w["v_wf"].T @ Gf["Wp"] + torch.outer(w["v_bf"], Gf["bp"])
        G["v_bf"] = Gf["bp"] @ w["fv_w1"].T
        G["fv_b1"] = Gf["bp"]
        G["fv_g"] = (w["fv_w2"] * Gf["W2g"]).sum(1)
        G["fv_w2"] = w["fv_g"][:, None] * Gf["W2g"] + torch.outer(w["fv_beta"], Gf["b2g"])
        G["fv_beta"] = Gf["b2g"] @ w["fv_w2"].T
        G["fv_b2"] = Gf["b2g"]
        G["mx_g"] = (w["mx_w2"] * Gf["M2g"]).sum(1)
        G["mx_w2"] = w["mx_g"][:, None] * Gf["M2g"] + torch.outer(w["mx_beta"], Gf["c2g"])
        G["mx_beta"] = Gf["c2g"] @ w["mx_w2"].T
        G["mx_b2"] = Gf["c2g"]
        return G

    # ---- whole thing -------------------------------------------------------------------------
    def forward(self, x, p, a, sigma):
        self.weights_fold()
        self.latents_fold(p, a)
        self.xi = query_features(self.cfg, x)
        nbar = self.pairs_fwd(self.xi, sigma)
        return self.tail_fwd(nbar)

    def backward(self, x, p, a, sigma, d_out):
        nbar = self.S["nbar"]
        dnbar, G, Gf = self.tail_bwd(nbar, d_out)
        G2, Gf2, GL = self.pairs_bwd(self.xi, sigma, dnbar)
        dp, da, G3, Gf3 = self.latents_bwd(p, a, GL)
        Gf.update(Gf2); Gf.update(Gf3)
        G.update(G2); G.update(G3)
        G.update(self.weights_unfold_bwd(Gf))
        G["q_omega"] = torch.zeros_like(self.w["q_omega"]); G["v_omega"] = torch.zeros_like(self.w["v_omega"])
        self.G, self.Gf, self.GL = G, Gf, GL
        return G, dp, da, GL["sigma"]


# mapping from the flat leaf names used above (and by the C ABI's EnfWeights) to the Flax tree
LEAF_PATHS = {
    "stem_w": "latent_stem/kernel", "stem_b": "latent_stem/bias",
    "ln_attn_g": "cross_attention_blocks_0/layer_norm_attn/scale", "ln_attn_b": "cross_attention_blocks_0/layer_norm_attn/bias",
    "q_omega": "cross_attention_blocks_0/attn/invariant_embedding_query/encoding/coefficients",
    "q_w1": "cross_attention_blocks_0/attn/invariant_embedding_query/layers_0/linear/kernel",
    "q_b1": "cross_attention_blocks_0/attn/invariant_embedding_query/layers_0/linear/bias",
    "q_wf": "cross_attention_blocks_0/attn/invariant_embedding_query/linear_final/kernel",
    "q_bf": "cross_attention_blocks_0/attn/invariant_embedding_query/linear_final/bias",
    "v_omega": "cross_attention_blocks_0/attn/invariant_embedding_value/encoding/coefficients",
    "v_w1": "cross_attention_blocks_0/attn/invariant_embedding_value/layers_0/linear/kernel",
    "v_b1": "cross_attention_blocks_0/attn/invariant_embedding_value/layers_0/linear/bias",
    "v_wf": "cross_attention_blocks_0/attn/invariant_embedding_value/linear_final/kernel",
    "v_bf": "cross_attention_blocks_0/attn/invariant_embedding_value/linear_final/bias",
    "wq": "cross_attention_blocks_0/attn/inv_emb_to_q/kernel", "bq": "cross_attention_blocks_0/attn/inv_emb_to_q/bias",
    "wk": "cross_attention_blocks_0/attn/a_to_k/kernel", "bk": "cross_attention_blocks_0/attn/a_to_k/bias",
    "wv": "cross_attention_blocks_0/attn/a_to_v/kernel", "bv": "cross_attention_blocks_0/attn/a_to_v/bias",
    "fv_w1": "cross_attention_blocks_0/attn/inv_emb_to_v/Dense_0/kernel", "fv_b1": "cross_attention_blocks_0/attn/inv_emb_to_v/Dense_0/bias",
    "fv_g": "cross_attention_blocks_0/attn/inv_emb_to_v/LayerNorm_0/scale", "fv_beta": "cross_attention_blocks_0/attn/inv_emb_to_v/LayerNorm_0/bias",
    "fv_w2": "cross_attention_blocks_0/attn/inv_emb_to_v/Dense_1/kernel", "fv_b2": "cross_attention_blocks_0/attn/inv_emb_to_v/Dense_1/bias",
    "mx_w1": "cross_attention_blocks_0/attn/inv_emb_cond_mixer/Dense_0/kernel", "mx_b1": "cross_attention_blocks_0/attn/inv_emb_cond_mixer/Dense_0/bias",
    "mx_g": "cross_attention_blocks_0/attn/inv_emb_cond_mixer/LayerNorm_0/scale", "mx_beta": "cross_attention_blocks_0/attn/inv_emb_cond_mixer/LayerNorm_0/bias",
    "mx_w2": "cross_attention_blocks_0/attn/inv_emb_cond_mixer/Dense_1/kernel", "mx_b2": "cross_attention_blocks_0/attn/inv_emb_cond_mixer/Dense_1/bias",
    "wo": "cross_attention_blocks_0/attn/out_proj/kernel", "bo": "cross_attention_blocks_0/attn/out_proj/bias",
    "fb_w1": "cross_attention_blocks_0/pointwise_ffn/Dense_0/kernel", "fb_b1": "cross_attention_blocks_0/pointwise_ffn/Dense_0/bias",
    "fb_g": "cross_attention_blocks_0/pointwise_ffn/LayerNorm_0/scale", "fb_beta": "cross_attention_blocks_0/pointwise_ffn/LayerNorm_0/bias",
    "fb_w2": "cross_attention_blocks_0/pointwise_ffn/Dense_1/kernel", "fb_b2": "cross_attention_blocks_0/pointwise_ffn/Dense_1/bias",
    "m0_w": "out_proj/layers_0/kernel", "m0_b": "out_proj/layers_0/bias",
    "m1_w": "out_proj/layers_2/kernel", "m1_b": "out_proj/layers_2/bias",
    "m2_w": "out_proj/layers_4/kernel", "m2_b": "out_proj/layers_4/bias",
}
