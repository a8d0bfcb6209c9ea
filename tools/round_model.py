"""fp64 folded model with operand roundings injected at named sites: which rounding costs how much parity?"""
import math, sys, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import folded_model as FM
from folded_model import *
from oracle import enf_ref as R
from helpers import make_case, rel_err

def r16(x): return x.to(torch.float16).to(torch.float64)
def rbf(x): return x.to(torch.bfloat16).to(torch.float64)
def r19(x):
    i = x.float().contiguous().view(torch.int32)
    i = ((i + 0x1000) & ~0x1FFF)
    return i.view(torch.float32).double()
def t19(x):
    i = x.float().contiguous().view(torch.int32)
    i = (i & ~0x1FFF)
    return i.view(torch.float32).double()
MODES = {'f16': r16, 'tf32': r19, 'tf32t': t19, 'bf16': rbf}
ACCURATE_MASK = [True]
TANH_APPROX = [False]          # proxy of tanh.approx.f32 (max rel. error 2^-11): the tanh value rounded to 11 significant bits
def _tanh(u):
    t = torch.tanh(u)
    return r16(t) if TANH_APPROX[0] else t
def gelu(x):
    c = math.sqrt(2.0 / math.pi)
    return 0.5 * x * (1 + _tanh(c * (x + 0.044715 * x ** 3)))
def gelu_grad(x):
    c = math.sqrt(2.0 / math.pi)
    t = _tanh(c * (x + 0.044715 * x ** 3))
    return 0.5 * (1 + t) + 0.5 * x * (1 - t * t) * c * (1 + 3 * 0.044715 * x * x)

class Rounded(FM.Folded):
    def __init__(self, cfg, params, sites):
        super().__init__(cfg, params); self.sites = sites
    def q(self, x, site):
        m = self.sites.get(site)
        return MODES[m](x) if m else x
    def pairs_fwd(self, xi, sigma):
        cfg, w, f, L = self.cfg, self.w, self.f, self.L
        H, d = cfg.num_heads, cfg.num_hidden
        scale = 1.0 / math.sqrt(d); q = self.q
        S = {}
        u, win, dot, sq = pair_invariants(cfg, xi, L["Lam"], sigma)
        gq = self._rff(u, w["q_omega"]); gv = self._rff(u, w["v_omega"])
        h1q = torch.relu(q(gq,'gq') @ q(w["q_w1"],'q_w1') + w["q_b1"])
        S["mask_q"] = (gq @ w["q_w1"] + w["q_b1"]) > 0          # what the split (3-term) recompute of kernels B / C sees
        s = scale * (torch.einsum("bczi,bzhi->bczh", q(h1q,'h1q_s'), q(L["U"],'U')) + L["kappa"][:, None]) + win[..., None]
        h1v = torch.relu(q(gv,'gv') @ q(w["v_w1"],'v_w1') + w["v_b1"])
        S["mask_v"] = (gv @ w["v_w1"] + w["v_b1"]) > 0
        tpre = q(h1v,'h1v') @ q(f["Wp"],'Wp') + f["bp"]
        that, t_rstd = ln_core(gelu(tpre))
        mpre = torch.einsum("bczi,bzhij->bczhj", q(that,'that'), q(L["W3"],'W3')) + L["b3"][:, None]
        n, n_rstd = ln_core(gelu(mpre))
        m = s.max(dim=2, keepdim=True).values
        e = torch.exp(s - m); l = e.sum(dim=2, keepdim=True); att = e / l
        nbar = torch.einsum("bczh,bczhj->bchj", att, n)
        S.update(u=u, win=win, dot=dot, sq=sq, gq=gq, gv=gv, h1q=h1q, h1v=h1v, tpre=tpre, that=that, t_rstd=t_rstd,
                 mpre=mpre, n=n, n_rstd=n_rstd, att=att, nbar=nbar, lse=(m + torch.log(l))[:, :, 0])
        self.S = S
        return nbar
    # per-query tail with operand roundings: t_act (forward activations), t_w (weights), tb_g (dgrad cotangents), tw_act / tw_g
    # (wgrad operands).  Cotangents are scaled by a power of two first, as the kernels do.
    def tail_fwd(self, nbar):
        cfg, w, f = self.cfg, self.w, self.f
        H, d = cfg.num_heads, cfg.num_hidden
        B, C = nbar.shape[:2]; q = self.q
        T = {}
        e1 = q(nbar.reshape(B, C, H * d), 't_act') @ q(f["W_A"], 't_w') + f["b_A"]
        e3c, e_rstd = ln_core(gelu(e1))
        e3 = e3c * w["fb_g"] + w["fb_beta"]
        fo = q(e3, 't_act') @ q(w["fb_w2"], 't_w') + w["fb_b2"]
        o1p = q(gelu(fo), 't_act') @ q(w["m0_w"], 't_w') + w["m0_b"]
        o2p = q(gelu(o1p), 't_act') @ q(w["m1_w"], 't_w') + w["m1_b"]
        out = gelu(o2p) @ w["m2_w"] + w["m2_b"]
        T.update(e1=e1, e3c=e3c, e_rstd=e_rstd, e3=e3, fo=fo, o1p=o1p, o2p=o2p)
        self.T = T
        return out
    def tail_bwd(self, nbar, d_out):
        cfg, w, f, T = self.cfg, self.w, self.f, self.T
        H, d = cfg.num_heads, cfg.num_hidden
        B, C = nbar.shape[:2]; q = self.q
        G = {}
        fl = lambda t: t.reshape(-1, t.shape[-1])
        gsc = 2.0 ** math.floor(math.log2(64.0 / float(d_out.abs().max())))
        qg = lambda x, site: q(x * gsc, site) / gsc
        o2 = gelu(T["o2p"])
        G["m2_w"] = fl(o2).T @ fl(d_out); G["m2_b"] = fl(d_out).sum(0)
        do2p = (d_out @ w["m2_w"].T) * gelu_grad(T["o2p"])
        G["m1_w"] = fl(q(gelu(T["o1p"]), 'tw_act')).T @ fl(qg(do2p, 'tw_g')); G["m1_b"] = fl(do2p).sum(0)
        do1p = (qg(do2p, 'tb_g') @ q(w["m1_w"], 't_w').T) * gelu_grad(T["o1p"])
        G["m0_w"] = fl(q(gelu(T["fo"]), 'tw_act')).T @ fl(qg(do1p, 'tw_g')); G["m0_b"] = fl(do1p).sum(0)
        dfo = (qg(do1p, 'tb_g') @ q(w["m0_w"], 't_w').T) * gelu_grad(T["fo"])
        G["fb_w2"] = fl(q(T["e3"], 'tw_act')).T @ fl(qg(dfo, 'tw_g')); G["fb_b2"] = fl(dfo).sum(0)
        de3 = qg(dfo, 'tb_g') @ q(w["fb_w2"], 't_w').T
        G["fb_g"] = fl(de3 * T["e3c"]).sum(0); G["fb_beta"] = fl(de3).sum(0)
        de2 = ln_core_bwd(de3 * w["fb_g"], T["e3c"], T["e_rstd"])
        de1 = de2 * gelu_grad(T["e1"])
        Gf = {}
        Gf["W_A"] = fl(q(nbar.reshape(B, C, H * d), 'tw_act')).T @ fl(qg(de1, 'tw_g')); Gf["b_A"] = fl(de1).sum(0)
        dnbar = (qg(de1, 'tb_g') @ q(f["W_A"], 't_w').T).reshape(B, C, H, d)
        return dnbar, G, Gf
    def pairs_bwd(self, xi, sigma, dnbar):
        cfg, w, f, L, S = self.cfg, self.w, self.f, self.L, self.S
        H, d = cfg.num_heads, cfg.num_hidden
        scale = 1.0 / math.sqrt(d); q = self.q
        G, Gf, GL = {}, {}, {}
        fl = lambda t: t.reshape(-1, t.shape[-1])
        att, n = S["att"], S["n"]
        gsc = 16.0 / float(dnbar.abs().max())      # the kernels' power-of-two scaling, roughly
        Dd = (dnbar * S["nbar"]).sum(-1)
        ds = att * (torch.einsum("bchj,bczhj->bczh", dnbar, n) - Dd[:, :, None])
        dn = att[..., None] * dnbar[:, :, None]
        dmpre = ln_core_bwd(dn, n, S["n_rstd"]) * gelu_grad(S["mpre"])
        qg = lambda x, site: q(x * gsc, site) / gsc
        GL["W3"] = torch.einsum("bczi,bczhj->bzhij", q(S["that"],'b_that'), qg(dmpre,'b_dm'))
        GL["b3"] = dmpre.sum(1)
        dthat = torch.einsum("bczhj,bzhij->bczi", qg(dmpre,'b_dm'), q(L["W3"],'b_W3'))
        dthat = qg(dthat, 'b_dthat_store')
        dtpre = ln_core_bwd(dthat, S["that"], S["t_rstd"]) * gelu_grad(S["tpre"])
        Gf["Wp"] = fl(q(S["h1v"],'b_h1v')).T @ fl(qg(dtpre,'b_dt')); Gf["bp"] = fl(dtpre).sum(0)
        dzv = (qg(dtpre,'b_dt') @ q(f["Wp"],'b_Wp').T) * (S["mask_v"] if ACCURATE_MASK[0] else (S["h1v"] > 0))
        G["v_w1"] = fl(q(S["gv"],'b_gv')).T @ fl(qg(dzv,'b_dzv')); G["v_b1"] = fl(dzv).sum(0)
        dgv = qg(dzv,'b_dzv') @ q(w["v_w1"],'b_vw1').T
        dzq = scale * torch.einsum("bczh,bzhi->bczi", ds, L["U"]) * (S["mask_q"] if ACCURATE_MASK[0] else (S["h1q"] > 0))
        GL["U"] = scale * torch.einsum("bczh,bczi->bzhi", ds, S["h1q"])
        GL["kappa"] = scale * ds.sum(1)
        G["q_w1"] = fl(q(S["gq"],'b_gq')).T @ fl(qg(dzq,'b_dzq')); G["q_b1"] = fl(dzq).sum(0)
        dgq = qg(dzq,'b_dzq') @ q(w["q_w1"],'b_qw1').T
        hd = d // 2
        def rff_bwd(g, dg, omega):
            sin, cos = g[..., :hd], g[..., hd:]
            dproj = cos * dg[..., :hd] - sin * dg[..., hd:]
            return 2 * math.pi * (dproj @ omega.T)
        du = rff_bwd(S["gq"], dgq, w["q_omega"]) + rff_bwd(S["gv"], dgv, w["v_omega"])
        dw = ds.sum(-1)
        GL["Lam"], GL["sigma"] = pair_invariants_bwd(cfg, xi, L["Lam"], sigma, S["u"], S["win"], S["dot"], S["sq"], du, dw)
        return G, Gf, GL

def run(cfg, case, sites):
    params, x, p, a, sigma, d_out = case
    m = Rounded(cfg, params, sites)
    out = m.forward(x, p, a, sigma)
    G, dp, da, dsg = m.backward(x, p, a, sigma, d_out)
    return out, dp, da, dsg, G

FWD_PAIR = ['gq','q_w1','gv','v_w1','h1v','Wp','that','W3']
BWD_PAIR = ['b_that','b_dm','b_W3','b_dthat_store','b_h1v','b_dt','b_Wp','b_gv','b_dzv','b_vw1','b_gq','b_dzq','b_qw1']
TAIL_F = ['t_act','t_w']; TAIL_B = ['tb_g','tw_act','tw_g']
if __name__ == '__main__':
    from helpers import leaf_errs
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from test_gpu_fuzz import _draw
    W_SITES = ['q_w1','v_w1','Wp','W3','b_W3','b_Wp','b_vw1','b_qw1']
    A_SITES = [s_ for s_ in FWD_PAIR + BWD_PAIR if s_ not in W_SITES]
    exps = {
      'all pair sites fp16': {s_: 'f16' for s_ in FWD_PAIR + BWD_PAIR},
      'weights only': {s_: 'f16' for s_ in W_SITES},
      'activations only': {s_: 'f16' for s_ in A_SITES},
      'fwd activations only': {s_: 'f16' for s_ in FWD_PAIR if s_ not in W_SITES},
      'bwd activations only': {s_: 'f16' for s_ in BWD_PAIR if s_ not in W_SITES},
      'tanh proxy only': {},
      'all + tanh proxy': {s_: 'f16' for s_ in FWD_PAIR + BWD_PAIR},
    }
    if os.environ.get("TAIL"):
        exps = {
          'tail: weights fp16 only': {'t_w': 'f16'},
          'tail: act fp16 (fwd)': {'t_act': 'f16'},
          'tail: act + dgrad cot fp16': {'t_act': 'f16', 'tb_g': 'f16'},
          'tail: act + cot + wgrad ops fp16': {'t_act': 'f16', 'tb_g': 'f16', 'tw_act': 'f16', 'tw_g': 'f16'},
          'tail: all + weights fp16': {'t_act': 'f16', 'tb_g': 'f16', 'tw_act': 'f16', 'tw_g': 'f16', 't_w': 'f16'},
          'pair fp16 only': {s_: 'f16' for s_ in FWD_PAIR + BWD_PAIR},
          'pair + tail(act,cot,wgrad)': {**{s_: 'f16' for s_ in FWD_PAIR + BWD_PAIR}, 't_act': 'f16', 'tb_g': 'f16', 'tw_act': 'f16', 'tw_g': 'f16'},
        }
    seeds = [int(a_) for a_ in sys.argv[1:]] or [1, 4]
    for seed in seeds:
        kw, B, C, Z = _draw(seed)
        cfg = R.EnfConfig(**kw)
        case = make_case(cfg, B, C, Z, seed=100 + seed)
        TANH_APPROX[0] = False
        ref = run(cfg, case, {})
        print(f'--- fuzz seed {seed}: d={kw["num_hidden"]} H={kw["num_heads"]} {kw["invariant_type"]} B={B} C={C} Z={Z}')
        for en, sites in exps.items():
            TANH_APPROX[0] = 'tanh' in en
            got = run(cfg, case, sites)
            e = [rel_err(got[i], ref[i]) for i in range(4)]
            le = leaf_errs(got[4], ref[4])
            wk = max(le, key=le.get)
            print(f'{en:24s} out {e[0]:.2e} dp {e[1]:.2e} da {e[2]:.2e} ds {e[3]:.2e} dth(per leaf) {le[wk]:.2e} [{wk}]')
