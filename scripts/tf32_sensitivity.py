"""CPU experiment: which GEMM operands of the CDAE sweeps are sensitive to tf32 rounding?
Emulates the CUDA plan's operand rounding inside a copy of the oracle's 4-sweep math."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import numpy as np
import ardae_oracle as orc
from golden_util import rel_err, load_case, sub

def tf32(a):
    a32 = np.asarray(a, dtype=np.float32).copy()
    u = a32.view(np.uint32)
    u += 0x1000
    u &= 0xFFFFE000
    return a32.astype(np.float64)

class Q:
    def __init__(self, on): self.on = set(on)
    def __call__(self, tag, a): return tf32(a) if (tag in self.on or 'all' in self.on) else a

def run(cs, P, x, ctx, std, eps, q):
    B, S, d = x.shape; N, H = B*S, cs.H
    sig, sp = orc.sigmoid, orc.softplus
    xf, sf, ef = x.reshape(N, d), std.reshape(N, 1), eps.reshape(N, d)
    xt_full = xf + sf*ef; xt = q('xt_fwd', xt_full); xt_st = q('xt_st', xt_full)
    ik, ck, nk = cs.inp_keys, cs.ctx_keys, cs.nlp_keys
    W = lambda k: q('w_bwd', P[k + '.weight']); Wf = lambda k: q('w_fwd', P[k + '.weight'])
    U, A_ = [], []
    h = xt
    for k in ik:
        a = h @ Wf(k).T + P[k + '.bias']; hf = sp(a); h = q('act_fwd', hf); U.append(q('act_st', hf))
    cl = ctx.reshape(B, -1)
    for k in ck:
        cl = sp(cl @ P[k + '.weight'].T + P[k + '.bias'])
    W1 = W(nk[0]); W1f = Wf(nk[0])
    rowb = cl @ P[nk[0] + '.weight'][:, H:2*H].T + P[nk[0] + '.bias']
    V = []
    p = h @ W1f[:, :H].T + np.repeat(rowb, S, 0) + sf * P[nk[0] + '.weight'][:, 2*H][None, :]
    vf = sp(p); v = q('act_fwd', vf); V.append(q('act_st', vf))
    for k in nk[1:-1]:
        vf = sp(v @ Wf(k).T + P[k + '.bias']); v = q('act_fwd', vf); V.append(q('act_st', vf))
    wo = P[nk[-1] + '.weight']
    L = cs.L
    s_of = lambda u: 1 - np.exp(-u)
    DP = [None]*L; DA = [None]*L
    DP[L-1] = q('delta', -wo * s_of(V[L-1]))
    for l in range(L-1, 0, -1):
        DP[l-1] = q('delta', (DP[l] @ W(nk[l])) * s_of(V[l-1]))
    DA[L-1] = q('delta', (DP[0] @ W1[:, :H]) * s_of(U[L-1]))
    for l in range(L-1, 0, -1):
        DA[l-1] = q('delta', (DA[l] @ W(ik[l])) * s_of(U[l-1]))
    g = DA[0] @ W(ik[0])
    resid = sf*g + ef; loss = (resid**2).mean()
    r = q('r', (2.0/(N*d)) * sf * resid)
    UD, TA, VD, TP = [None]*L, [None]*L, [None]*L, [None]*L
    t = r
    for l in range(L):
        ad = t @ W(ik[l]).T; s = s_of(U[l])
        UD[l] = q('tan', ad*s); TA[l] = q('adj', DA[l]*ad*(1-s)); t = UD[l]
    for l in range(L):
        pd = t @ (W1[:, :H] if l == 0 else W(nk[l])).T; s = s_of(V[l])
        VD[l] = q('tan', pd*s); TP[l] = q('adj', DP[l]*pd*(1-s)); t = VD[l]
    G = {}
    G[nk[-1] + '.weight'] = -VD[L-1].sum(0, keepdims=True)
    for l in range(L-1, 0, -1):
        TP[l-1] = q('adj', (TP[l] @ W(nk[l])) * s_of(V[l-1]) + TP[l-1])
    TA[L-1] = q('adj', (TP[0] @ W1[:, :H]) * s_of(U[L-1]) + TA[L-1])
    for l in range(L-1, 0, -1):
        TA[l-1] = q('adj', (TA[l] @ W(ik[l])) * s_of(U[l-1]) + TA[l-1])
    for l in range(L-1, 0, -1):
        G[nk[l] + '.weight'] = TP[l].T @ V[l-1] + DP[l].T @ VD[l-1]
        G[ik[l] + '.weight'] = TA[l].T @ U[l-1] + DA[l].T @ UD[l-1]
    G[ik[0] + '.weight'] = TA[0].T @ xt_st + DA[0].T @ r
    G[nk[0] + '.weight.u'] = TP[0].T @ U[L-1] + DP[0].T @ UD[L-1]
    for l in range(L):
        G[ik[l] + '.bias'] = TA[l].sum(0); G[nk[l] + '.bias'] = TP[l].sum(0)
    return loss, g, G

name = sys.argv[1] if len(sys.argv) > 1 else 'toy_small'
z, meta = load_case(name); c = meta['cdae']; hp = meta['hp']
cs = orc.CdaeSpec(c['input_dim'], c['context_dim'], c['h_dim'], c['num_hidden_layers'])
P = sub(z, 'c0/')
x = np.repeat(hp['std_scale'] * (z['s0/z_cdae'] - z['s0/zbar']), hp['nstd'], axis=1)
ctx = z['s0/zbar']; std = z['s0/std'] * z['s0/noise/xi']; eps = z['s0/noise/eps_cdae']
if len(sys.argv) > 2:
    sc = float(sys.argv[2]); x = x * sc; std = std * sc
loss_o, g_o, G_o = orc.cdae_loss_and_grads(cs, P, x, ctx, std, eps)
H = cs.H
G_o[cs.nlp_keys[0] + '.weight.u'] = G_o[cs.nlp_keys[0] + '.weight'][:, :H]
for on in (['none'], ['all'], ['xt_fwd'], ['w_fwd'], ['act_fwd'], ['xt_st'], ['act_st'], ['w_bwd'], ['xt_st','act_st','w_bwd','r','delta','tan','adj']):
    loss, g, G = run(cs, P, x, ctx, std, eps, Q(on))
    errs = {k: rel_err(G[k], G_o[k]) for k in G}
    worst = max(errs.items(), key=lambda kv: kv[1])
    inp0 = errs[cs.inp_keys[0] + '.weight']
    print('%-12s loss rel %.2e score rel %.2e | worst grad %.2e (%s) | inp0.W %.2e' % ('+'.join(on), abs(loss-loss_o)/loss_o, rel_err(g, g_o.reshape(g.shape)), worst[1], worst[0], inp0))
