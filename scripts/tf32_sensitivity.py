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

def bf16(a):
    """round-to-nearest-even to bfloat16 (8 significant bits)"""
    a32 = np.asarray(a, dtype=np.float32).copy()
    u = a32.view(np.uint32)
    u += 0x7FFF + ((u >> 16) & 1)
    u &= 0xFFFF0000
    return a32.astype(np.float64)

def fp16_rowscaled(a):
    """fp16 with one power-of-two scale per row (row max -> [2^14, 2^15)): 11 significant bits, 2^-24 * rowmax floor"""
    a = np.asarray(a, dtype=np.float64)
    m = np.abs(a).max(axis=1, keepdims=True)
    e = np.where(m > 0, np.floor(np.log2(np.where(m > 0, m, 1.0))), 0.0)
    sc = 2.0 ** (14 - e)
    return (a * sc).astype(np.float16).astype(np.float64) / sc

def packed24(a):
    """fp32 with the low mantissa byte dropped (16 significant bits)"""
    a32 = np.asarray(a, dtype=np.float32).copy()
    u = a32.view(np.uint32)
    u += 0x80
    u &= 0xFFFFFF00
    return a32.astype(np.float64)

ROUND = {'tf32': tf32, 'bf16': bf16, 'fp16rs': fp16_rowscaled, 'p24': packed24}

class Q:
    """on: iterable of tags rounded to tf32, or dict tag -> rounding name"""
    def __init__(self, on):
        self.on = dict(on) if isinstance(on, dict) else {t: 'tf32' for t in on}
    def __call__(self, tag, a):
        r = self.on.get(tag, self.on.get('all'))
        if r is None and '@' in tag:   # per-layer forward tags: 'act_fwd@k' falls back to 'act_fwd' / 'fwd_from' (layers >= j)
            base, k = tag.split('@')
            r = self.on.get(base)
            j = self.on.get('fwd_from')
            if r is None and j is not None and int(k) >= j:
                r = 'tf32'
        return ROUND[r](a) if r else a

def run(cs, P, x, ctx, std, eps, q):
    B, S, d = x.shape; N, H = B*S, cs.H
    sig, sp = orc.sigmoid, orc.softplus
    xf, sf, ef = x.reshape(N, d), std.reshape(N, 1), eps.reshape(N, d)
    xt_full = xf + sf*ef; xt = q('xt_fwd', xt_full); xt_st = q('xt_st', xt_full)
    ik, ck, nk = cs.inp_keys, cs.ctx_keys, cs.nlp_keys
    W = lambda k: q('w_bwd', P[k + '.weight'])
    Wf = lambda k, li: q('w_fwd@%d' % li, P[k + '.weight'])   # forward layer li consumes (activation li-1, weight li)
    U, A_ = [], []
    h = xt
    for li, k in enumerate(ik):
        a = h @ Wf(k, li).T + P[k + '.bias']; hf = sp(a); h = q('act_fwd@%d' % (li + 1), hf); U.append(q('act_st', hf))
    cl = ctx.reshape(B, -1)
    for k in ck:
        cl = sp(cl @ P[k + '.weight'].T + P[k + '.bias'])
    W1 = W(nk[0]); W1f = Wf(nk[0], cs.L)
    rowb = cl @ P[nk[0] + '.weight'][:, H:2*H].T + P[nk[0] + '.bias']
    V = []
    p = h @ W1f[:, :H].T + np.repeat(rowb, S, 0) + sf * P[nk[0] + '.weight'][:, 2*H][None, :]
    vf = sp(p); v = q('act_fwd@%d' % (cs.L + 1), vf); V.append(q('act_st', vf))
    for li, k in enumerate(nk[1:-1]):
        vf = sp(v @ Wf(k, cs.L + 1 + li).T + P[k + '.bias']); v = q('act_fwd@%d' % (cs.L + 2 + li), vf); V.append(q('act_st', vf))
    wo = P[nk[-1] + '.weight']
    L = cs.L
    s_of = lambda u: 1 - np.exp(-u)
    DP = [None]*L; DA = [None]*L          # chain operands (stay on chip, tf32)
    DPs = [None]*L; DAs = [None]*L        # what later sweeps / contractions read back from the spill
    DP[L-1] = q('delta', -wo * s_of(V[L-1])); DPs[L-1] = q('delta_st', DP[L-1])
    for l in range(L-1, 0, -1):
        DP[l-1] = q('delta', (DP[l] @ W(nk[l])) * s_of(V[l-1])); DPs[l-1] = q('delta_st', DP[l-1])
    DA[L-1] = q('delta', (DP[0] @ W1[:, :H]) * s_of(U[L-1])); DAs[L-1] = q('delta_st', DA[L-1])
    for l in range(L-1, 0, -1):
        DA[l-1] = q('delta', (DA[l] @ W(ik[l])) * s_of(U[l-1])); DAs[l-1] = q('delta_st', DA[l-1])
    g = DA[0] @ W(ik[0])
    resid = sf*g + ef; loss = (resid**2).mean()
    r = q('r', (2.0/(N*d)) * sf * resid)
    UD, TA, VD, TP = [None]*L, [None]*L, [None]*L, [None]*L
    UDs, VDs = [None]*L, [None]*L
    t = r
    for l in range(L):
        ad = t @ W(ik[l]).T; s = s_of(U[l])
        UD[l] = q('tan', ad*s); UDs[l] = q('tan_st', UD[l]); TA[l] = q('t_st', q('adj', DAs[l]*ad*(1-s))); t = UD[l]
    for l in range(L):
        pd = t @ (W1[:, :H] if l == 0 else W(nk[l])).T; s = s_of(V[l])
        VD[l] = q('tan', pd*s); VDs[l] = q('tan_st', VD[l]); TP[l] = q('t_st', q('adj', DPs[l]*pd*(1-s))); t = VD[l]
    G = {}
    G[nk[-1] + '.weight'] = -VD[L-1].sum(0, keepdims=True)
    TPs, TAs = [None]*L, [None]*L
    TPs[L-1] = q('adj_st', TP[L-1])
    for l in range(L-1, 0, -1):
        TP[l-1] = q('adj', (TP[l] @ W(nk[l])) * s_of(V[l-1]) + TP[l-1]); TPs[l-1] = q('adj_st', TP[l-1])
    TA[L-1] = q('adj', (TP[0] @ W1[:, :H]) * s_of(U[L-1]) + TA[L-1]); TAs[L-1] = q('adj_st', TA[L-1])
    for l in range(L-1, 0, -1):
        TA[l-1] = q('adj', (TA[l] @ W(ik[l])) * s_of(U[l-1]) + TA[l-1]); TAs[l-1] = q('adj_st', TA[l-1])
    for l in range(L-1, 0, -1):
        G[nk[l] + '.weight'] = TPs[l].T @ V[l-1] + DPs[l].T @ VDs[l-1]
        G[ik[l] + '.weight'] = TAs[l].T @ U[l-1] + DAs[l].T @ UDs[l-1]
    G[ik[0] + '.weight'] = TAs[0].T @ xt_st + DAs[0].T @ r
    G[nk[0] + '.weight.u'] = TPs[0].T @ U[L-1] + DPs[0].T @ UDs[L-1]
    for l in range(L):
        G[ik[l] + '.bias'] = TA[l].sum(0); G[nk[l] + '.bias'] = TP[l].sum(0)
    return loss, g, G

def synthetic(H, L, d, B, S, seed=0):
    """config-2-like CDAE at the reference initialisation (default nn.Linear init) with inputs at the scale the
    step produces: x = std_scale (z - zbar) ~ 1e4, sigma = delta * rowstd * xi."""
    rng = np.random.RandomState(seed)
    cs = orc.CdaeSpec(d, d, H, L)
    P = {}
    def lin(o, i):
        k = 1.0 / np.sqrt(i)
        return rng.uniform(-k, k, (o, i)), rng.uniform(-k, k, (o,))
    for pre in ('ctx_encode', 'inp_encode'):
        for k in orc.mlp_keys(pre, L - 1):
            P[k + '.weight'], P[k + '.bias'] = lin(H, d if k.endswith('layers.0') else H)
    for k in orc.mlp_keys('neglogprob', L):
        o, i = (1, H) if k.endswith('.fc') else (H, 2 * H + 1 if k.endswith('layers.0') else H)
        P[k + '.weight'], P[k + '.bias'] = lin(o, i)
    x = 1e4 * 3.0 * rng.randn(B, S, d) * rng.uniform(0.2, 2.0, (B, 1, 1))
    ctx = rng.randn(B, 1, d)
    std = 0.1 * x.std(axis=1, ddof=1).mean(axis=-1)[:, None, None] * rng.randn(B, S, 1)
    eps = rng.randn(B, S, d)
    return cs, P, x, ctx, std, eps

name = sys.argv[1] if len(sys.argv) > 1 else 'toy_small'
if name.startswith('synth'):
    # synth:H,L,d,B,S
    H_, L_, d_, B_, S_ = [int(v) for v in name.split(':')[1].split(',')]
    cs, P, x, ctx, std, eps = synthetic(H_, L_, d_, B_, S_)
else:
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

def report(label, q):
    loss, g, G = run(cs, P, x, ctx, std, eps, q)
    errs = {k: rel_err(G[k], G_o[k]) for k in G}
    worst = max(errs.items(), key=lambda kv: kv[1])
    allg = rel_err(np.concatenate([G[k].ravel() for k in sorted(G)]), np.concatenate([G_o[k].ravel() for k in sorted(G)]))
    print('%-34s loss rel %.2e score rel %.2e | worst grad %.2e (%s) | all grads %.2e' % (
        label, abs(loss-loss_o)/loss_o, rel_err(g, g_o.reshape(g.shape)), worst[1], worst[0], allg))

PROD = ['xt_st', 'act_st', 'w_bwd', 'r', 'delta', 'tan', 'adj']   # what the CUDA plan rounds to tf32 today
print('== one operand class at a time, tf32 (fp64 elsewhere)')
for on in (['none'], ['all'], ['xt_fwd'], ['w_fwd'], ['act_fwd'], ['xt_st'], ['act_st'], ['w_bwd'], PROD):
    report('+'.join(on) if len(on) < 4 else 'PROD (tf32 backward plan)', Q(on))
print('== plain-tf32 forward (1 MMA instead of 3xTF32) from forward layer j on (layers 0..2L-1), on top of PROD')
for j in range(0, 2 * cs.L + 1):
    on = {t: 'tf32' for t in PROD}
    on['fwd_from'] = j
    report('PROD + tf32 forward layers >= %d' % j, Q(on))
print('== spill storage narrower than tf32, on top of PROD (chain operands stay tf32 on chip)')
SPILL = ['act_st', 'delta_st', 'tan_st', 't_st', 'adj_st']
for fmt in ('bf16', 'fp16rs', 'p24'):
    for cls in SPILL + ['ALL']:
        on = {t: 'tf32' for t in PROD}
        for t in (SPILL if cls == 'ALL' else [cls]):
            on[t] = fmt
        report('PROD + %s -> %s' % (cls, fmt), Q(on))

# == would the backward sweeps survive fp16 MMA operands (same 10-bit mantissa as tf32, 5-bit exponent)?
# chain operands (delta, tangents, adjoints, r) and backward weights rounded to fp16 after a STATIC power-of-two
# scale per class; prints the magnitude range of each class so the scale can be judged.
if os.environ.get('FP16_SWEEPS'):
    STATS = {}
    def make_f16(k):
        def f(a):
            a = np.asarray(a, dtype=np.float64)
            with np.errstate(over='ignore'):
                return (a * 2.0 ** k).astype(np.float16).astype(np.float64) / 2.0 ** k
        return f
    class QS(Q):
        def __call__(self, tag, a):
            base = tag.split('@')[0]
            if base in ('delta', 'tan', 'adj', 'r', 'w_bwd'):
                m = np.abs(np.asarray(a))
                st = STATS.setdefault(base, [0.0, np.inf, []])
                st[0] = max(st[0], m.max()); nz = m[m > 0]
                if nz.size:
                    st[1] = min(st[1], nz.min()); st[2].append(np.median(np.abs(a).max(axis=1)) if a.ndim == 2 else m.max())
            return Q.__call__(self, tag, a)
    print('== fp16 chain operands with static scales, on top of PROD + bf16 spill')
    base = {t: 'tf32' for t in PROD}
    for t in SPILL:
        base[t] = 'bf16'
    report('PROD + bf16 spill (today)', QS(base))
    for t, (mx, mn, med) in STATS.items():
        print('   class %-6s max %.3e  min nonzero %.3e  median row-max %.3e' % (t, mx, mn, float(np.median(med))))
    for kd, kt, ka, kr, kw in ((0, 0, 0, 0, 0), (8, 8, 8, 8, 4), (12, 12, 12, 12, 4), (4, 4, 4, 4, 4)):
        ROUND.update(f16d=make_f16(kd), f16t=make_f16(kt), f16a=make_f16(ka), f16r=make_f16(kr), f16w=make_f16(kw))
        on = dict(base)
        on.update(delta='f16d', tan='f16t', adj='f16a', r='f16r', w_bwd='f16w')
        report('fp16 sweeps, scales 2^(%d,%d,%d,%d;w %d)' % (kd, kt, ka, kr, kw), Q(on))
