"""Generates tests/golden/anchors.json: known-answer vectors for the hot path, computed with 50-digit mpmath from
the reference's formulas AS CODED (src/parameters/kernel.rs, src/gaussian_process/mod.rs, optimizer.rs), independent
of both the C oracle and the CUDA path.  Run:  python tests/golden/make_anchors.py

The first block reproduces the reference's own doc-test dataset (src/gaussian_process/mod.rs:7-8) with fixed
SquaredExp{ls:1, ampl:1}, zero prior and the builder's default noise (builder.rs:73) — the anchors listed in
SURVEY.md §8(c).  The reference's tests assert no numbers, so these are the pins ("parity unpinned" otherwise).
"""
import json
import os

import mpmath as mp

mp.mp.dps = 50
F = lambda v: float(v)


def dist2(a, b):
    return sum((mp.mpf(x) - mp.mpf(y)) ** 2 for x, y in zip(a, b))


def dot(a, b):
    return sum(mp.mpf(x) * mp.mpf(y) for x, y in zip(a, b))


def sgn(v):
    return mp.mpf(-1) if v < 0 else mp.mpf(1)


# kernels as coded (kernel.rs); returns (value, [gradients in get_parameters order])
def k_linear(p, a, b):
    return dot(a, b) + p[0], [mp.mpf(1)]


def k_polynomial(p, a, b):
    x = dot(a, b)
    inner = p[0] * x + p[1]
    gc = p[2] * inner ** (p[2] - 1)
    return inner ** p[2], [x * gc, gc, mp.log(inner) * inner ** p[2]]


def k_sqexp(p, a, b):
    d2 = dist2(a, b)
    e = mp.exp(-d2 / (2 * p[0] * p[0]))
    return abs(p[1]) * e, [d2 * abs(p[1]) * e / p[0] ** 3, sgn(p[1]) * e]


def k_exponential(p, a, b):
    r = mp.sqrt(dist2(a, b))
    e = mp.exp(-r / (2 * p[0] * p[0]))
    return abs(p[1]) * e, [r * abs(p[1]) * e / p[0] ** 3, sgn(p[1]) * e]


def k_matern1(p, a, b):
    r = mp.sqrt(dist2(a, b))
    l, ampl = abs(p[0]), abs(p[1])
    x = mp.sqrt(3) * r / l
    return ampl * (1 + x) * mp.exp(-x), [3 * ampl * r ** 2 * mp.exp(-x) / p[0] ** 3, sgn(p[1]) * (1 + x) * mp.exp(-x)]


def k_matern2(p, a, b):
    r = mp.sqrt(dist2(a, b))
    l, ampl = abs(p[0]), abs(p[1])
    x = mp.sqrt(5) * r / l
    val = ampl * (1 + x + 5 * r * r / (3 * l * l)) * mp.exp(-x)
    xs = mp.sqrt(5) * r / p[0]  # signed ls in the gradient (kernel.rs:891)
    gls = sgn(p[0]) * ampl * ((2 * l / 3 + 1) + r * mp.sqrt(5) * ((l ** 2 / 3 + l + 1) / l ** 2)) * mp.exp(-xs)
    gam = sgn(p[1]) * (1 + xs + 5 * r * r / (3 * l * l)) * mp.exp(-xs)
    return val, [gls, gam]


def k_hypertan(p, a, b):
    x = dot(a, b)
    gc = 1 / mp.cosh(p[0] * x + p[1]) ** 2
    return mp.tanh(p[0] * x + p[1]), [x * gc, gc]


def k_multiquadric(p, a, b):
    d2 = dist2(a, b)
    return mp.sqrt(d2 * d2 + p[0] * p[0]), [p[0] / mp.sqrt(d2 + p[0] * p[0])]


def k_rq(p, a, b):
    d2 = dist2(a, b)
    alpha, l = p[0], abs(p[1])
    val = (1 + d2 / (2 * p[0] * p[1] * p[1])) ** (-p[0])
    ga = ((d2 + 2 * l ** 2 * alpha) / (l ** 2 * alpha)) ** (-alpha) * (
        mp.mpf(2) ** alpha * (1 - mp.log((d2 + 2 * l ** 2 * alpha) / (2 * l ** 2 * alpha)))
        - (l ** 2 * mp.mpf(2) ** (alpha + 1) * alpha) / (d2 + 2 * l ** 2 * alpha))
    gl = d2 * (d2 / (2 * alpha * l * l) + 1) ** (-alpha - 1) / p[1] ** 3
    return val, [ga, gl]


LEAVES = {1: (k_linear, 1), 2: (k_polynomial, 3), 3: (k_sqexp, 2), 4: (k_exponential, 2), 5: (k_matern1, 2),
          6: (k_matern2, 2), 7: (k_hypertan, 2), 8: (k_multiquadric, 1), 9: (k_rq, 2)}


def eval_desc(ops, params, a, b):
    """postfix program -> (value, gradient list)"""
    st, po = [], 0
    params = [mp.mpf(p) for p in params]
    for op in ops:
        if op == 100:
            (v2, g2), (v1, g1) = st.pop(), st.pop()
            st.append((v1 + v2, g1 + g2))
        elif op == 101:
            (v2, g2), (v1, g1) = st.pop(), st.pop()
            st.append((v1 * v2, [g * v2 for g in g1] + [g * v1 for g in g2]))
        else:
            f, npar = LEAVES[op]
            st.append(f(params[po:po + npar], a, b))
            po += npar
    return st[0]


def gram(ops, params, X, noise):
    n = len(X)
    K = mp.zeros(n, n)
    for i in range(n):
        for j in range(n):
            K[i, j] = eval_desc(ops, params, X[j], X[i])[0]
        K[i, i] += mp.mpf(noise) ** 2
    return K


def gp_case(name, ops, params, X, y, noise, Xq, Xadd=None, yadd=None):
    n = len(X)
    K = gram(ops, params, X, noise)
    L = mp.cholesky(K)
    yv = mp.matrix(y)
    Kinv = mp.inverse(K)
    alpha = Kinv * yv
    Knq = mp.matrix(n, len(Xq))
    for i in range(n):
        for j in range(len(Xq)):
            Knq[i, j] = eval_desc(ops, params, X[i], Xq[j])[0]
    W = Kinv * Knq
    mean = [F(sum(W[i, j] * yv[i] for i in range(n))) for j in range(len(Xq))]
    var = [F(eval_desc(ops, params, Xq[j], Xq[j])[0] - sum(Knq[i, j] * W[i, j] for i in range(n)))
           for j in range(len(Xq))]
    cov = [[F(eval_desc(ops, params, Xq[i], Xq[j])[0] - sum(Knq[t, i] * W[t, j] for t in range(n)))
            for j in range(len(Xq))] for i in range(len(Xq))]
    # likelihood as coded (mod.rs:196-220): penalty = sum ln|k(x,x)+noise^2|, NOT log det
    data_fit = sum(yv[i] * alpha[i] for i in range(n))
    penalty = sum(mp.log(abs(eval_desc(ops, params, X[i], X[i])[0] + mp.mpf(noise) ** 2)) for i in range(n))
    lik = -(data_fit + penalty + n * mp.log(2 * mp.pi)) / 2
    # gradients as coded (optimizer.rs:24-60, :159-203)
    P = len(eval_desc(ops, params, X[0], X[0])[1])
    scale = data_fit / n
    g_unscaled, g_scaled = [], []
    for p in range(P):
        G = mp.matrix(n, n)
        for i in range(n):
            for j in range(n):
                G[i, j] = eval_desc(ops, params, X[min(i, j)], X[max(i, j)])[1][p]
        aGa = sum(alpha[i] * G[i, j] * alpha[j] for i in range(n) for j in range(n))
        tr = sum(Kinv[i, j] * G[j, i] for i in range(n) for j in range(n))
        g_unscaled.append(F((aGa - tr) / 2))
        g_scaled.append(F((aGa / scale - tr) / 2))
    g_noise = mp.mpf(noise) * (sum(a * a for a in alpha) - sum(Kinv[i, i] for i in range(n)))
    out = dict(name=name, ops=ops, params=[float(p) for p in params], X=X, y=y, noise=float(noise), Xq=Xq,
               L=[[F(L[i, j]) for j in range(n)] for i in range(n)], mean=mean, var=var, cov=cov,
               likelihood=F(lik), scale=F(scale), grad_unscaled=g_unscaled + [F(g_noise)], grad_scaled=g_scaled)
    if Xadd:
        Xall = X + Xadd
        L2 = mp.cholesky(gram(ops, params, Xall, noise))
        out["Xadd"], out["yadd"] = Xadd, yadd
        out["L_after_add"] = [[F(L2[i, j]) for j in range(len(Xall))] for i in range(len(Xall))]
    return out


def main():
    cases = []
    # --- reference doc-test dataset (mod.rs:7-8), SURVEY §8c anchors ---
    X = [[0.8], [1.2], [3.8], [4.2]]
    y = [3.0, 4.0, -2.0, -2.0]
    ymean = sum(y) / 4
    noise = F(mp.mpf("0.1") * mp.sqrt(sum((mp.mpf(v) - mp.mpf(ymean)) ** 2 for v in y) / 4))
    noise = 0.1 * (sum((v - ymean) ** 2 for v in y) / 4) ** 0.5  # f64 arithmetic as the builder does it
    cases.append(gp_case("doctest_sqexp", [3], [1.0, 1.0], X, y, noise, [[1.0], [2.0], [3.0]],
                         Xadd=[[0.0], [1.0], [2.0], [5.0]], yadd=[2.0, 3.0, -1.0, -2.0]))
    # --- 8 points in 2-d, every kernel + Sum/Prod ---
    X2 = [[0.1, 0.9], [0.4, 0.2], [0.7, 0.6], [0.3, 0.5], [0.95, 0.05], [0.55, 0.85], [0.15, 0.35], [0.8, 0.3]]
    y2 = [0.5, -0.3, 1.2, 0.1, -0.9, 0.8, 0.0, 0.4]
    Xq2 = [[0.2, 0.2], [0.6, 0.4], [0.9, 0.9]]
    Xadd2 = [[0.25, 0.75], [0.65, 0.15]]
    yadd2 = [0.3, -0.2]
    specs = [("linear", [1], [0.5]), ("polynomial", [2], [0.7, 1.1, 2.0]), ("sqexp", [3], [0.6, 1.3]),
             ("sqexp_neg", [3], [-0.6, -1.3]), ("exponential", [4], [0.8, 0.9]), ("matern1", [5], [0.7, 1.2]),
             ("matern2", [6], [0.9, 1.1]), ("matern2_neg", [6], [-0.9, -1.1]), ("hypertan", [7], [0.3, 0.1]),
             ("multiquadric", [8], [0.5]), ("rq", [9], [1.5, 0.8]),
             ("sum_sqexp_matern2", [3, 6, 100], [0.6, 1.3, 0.9, 0.4]),
             ("prod_sqexp_linear", [3, 1, 101], [0.6, 1.3, 0.5]),
             ("sum_prod_mix", [3, 5, 101, 9, 100], [0.6, 1.3, 0.7, 1.2, 1.5, 0.8])]
    for name, ops, params in specs:
        nz = 0.3 if name not in ("hypertan", "multiquadric") else 1.5  # keep K positive definite
        cases.append(gp_case("k8_" + name, ops, params, X2, y2, nz, Xq2, Xadd=Xadd2, yadd=yadd2))
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "anchors.json"), "w") as f:
        json.dump(dict(generator="tests/golden/make_anchors.py (mpmath, 50 digits)", cases=cases), f, indent=1)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
