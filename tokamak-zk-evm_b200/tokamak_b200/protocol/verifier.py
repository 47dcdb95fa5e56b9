"""The acceptance gate: Verifier::verify_snark and the per-argument checks of
verify-rust/src/lib.rs:98-330, on the host.  G1 linear combinations are done with exact affine arithmetic on Python
integers (about forty scalar multiplications per verification) and the final check is one product of ten pairings."""
import secrets

from ..transcript import TranscriptManager
from . import pairing
from .fr import Q_MOD, R_MOD, inv, root_of_unity


# ---------------------------------------------------------------- G1serde ops on the host (group_structures/mod.rs:888-947)
def g1_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0]:
        if (p[1] + q[1]) % Q_MOD == 0:
            return None
        lam = 3 * p[0] * p[0] * pow(2 * p[1], Q_MOD - 2, Q_MOD) % Q_MOD
    else:
        lam = (q[1] - p[1]) * pow(q[0] - p[0], Q_MOD - 2, Q_MOD) % Q_MOD
    x3 = (lam * lam - p[0] - q[0]) % Q_MOD
    return (x3, (lam * (p[0] - x3) - p[1]) % Q_MOD)


def g1_neg(p):
    return None if p is None else (p[0], (-p[1]) % Q_MOD)


def g1_mul(p, k):
    k %= R_MOD
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = g1_add(acc, acc)
        if bit == "1":
            acc = g1_add(acc, p)
    return acc


class _G:
    """Tiny operator wrapper so the verifier equations read like the reference's."""

    def __init__(self, p):
        self.p = p

    def __add__(self, o):
        return _G(g1_add(self.p, o.p))

    def __sub__(self, o):
        return _G(g1_add(self.p, g1_neg(o.p)))

    def __mul__(self, k):
        return _G(g1_mul(self.p, k))


def collect_challenges(points, scalars):
    mgr = TranscriptManager()
    mgr.add_proof0(points["U"], points["V"], points["W"], points["Q_AX"], points["Q_AY"], points["B"])
    thetas = mgr.get_thetas()
    mgr.add_proof1(points["R"])
    kappa0 = mgr.get_kappa0()
    mgr.add_proof2(points["Q_CX"], points["Q_CY"])
    chi, zeta = mgr.get_chi_zeta()
    mgr.add_proof3(scalars["V_eval"], scalars["R_eval"], scalars["R_omegaX_eval"], scalars["R_omegaX_omegaY_eval"])
    kappa1 = mgr.get_kappa1()
    return thetas, kappa0, chi, zeta, kappa1


def eval_a_pub(params, instance, chi):
    """a_pub_X.eval(chi, zeta): the interpolant of the l_free public values over the l_free-th roots of unity."""
    from .fr import lagrange_bases_at

    # the reference indexes a_pub_user[i] / a_pub_block[i] and panics on a short instance (verify-rust/src/lib.rs:150-170)
    if len(instance.a_pub_user) < params.l_user or len(instance.a_pub_block) < params.l_free - params.l_user:
        raise ValueError("instance has fewer public values than l_user / l_free - l_user")
    vals = list(instance.a_pub_user[:params.l_user]) + list(instance.a_pub_block[:params.l_free - params.l_user])
    lag = lagrange_bases_at(chi, params.l_free)
    return sum(v * b for v, b in zip(vals, lag)) % R_MOD


def check_proof_encoding(points, scalars):
    """Reject what the reference's deserialisation would never produce: coordinates >= q, points off y^2 = x^3 + 4
    (the identity is None), evaluations >= r.  An off-curve point would otherwise flow into g1_add and the pairing."""
    for name, p in points.items():
        if p is None:
            continue
        x, y = p
        if not (0 <= x < Q_MOD and 0 <= y < Q_MOD) or (y * y - x * x * x - 4) % Q_MOD != 0:
            raise ValueError(f"proof point {name} is not a canonical point of G1's curve")
    for name, v in scalars.items():
        if not 0 <= v < R_MOD:
            raise ValueError(f"proof scalar {name} is not a canonical element of Fr")


def verify_snark(params, sigma, preprocess, instance, points, scalars, kappa2=None):
    """Verifier::verify_snark (verify-rust/src/lib.rs:243-289).  sigma needs G, H, x, y, lagrange_KL and sigma2."""
    check_proof_encoding(points, scalars)
    check_proof_encoding(preprocess, {})
    P = {k: _G(v) for k, v in points.items()}
    pre = {k: _G(v) for k, v in preprocess.items()}
    thetas, kappa0, chi, zeta, kappa1 = collect_challenges(points, scalars)
    kappa2 = secrets.randbelow(R_MOD) if kappa2 is None else kappa2
    m_i, s_max = params.l_D - params.l, params.s_max
    w_mi_inv, w_s_inv = inv(root_of_unity(m_i)), inv(root_of_unity(s_max))
    t_n_eval = (pow(chi, params.n, R_MOD) - 1) % R_MOD
    t_mi_eval = (pow(chi, m_i, R_MOD) - 1) % R_MOD
    t_smax_eval = (pow(zeta, s_max, R_MOD) - 1) % R_MOD
    K0_eval = 1 if chi == 1 else t_mi_eval * inv(m_i) % R_MOD * inv(chi - 1) % R_MOD
    a_eval = eval_a_pub(params, instance, chi)
    G, sx, sy, KL = _G(sigma.G), _G(sigma.x), _G(sigma.y), _G(sigma.lagrange_KL)
    V_eval, R_eval = scalars["V_eval"], scalars["R_eval"]
    R_wX_eval, R_wXwY_eval = scalars["R_omegaX_eval"], scalars["R_omegaX_omegaY_eval"]
    k1 = lambda e: pow(kappa1, e, R_MOD)
    k2 = lambda e: pow(kappa2, e, R_MOD)
    # lhs_arith
    lhs_a = P["U"] * V_eval - P["W"] + (P["V"] - G * V_eval) * kappa1 - P["Q_AX"] * t_n_eval - P["Q_AY"] * t_smax_eval
    # lhs_copy
    F = P["B"] + pre["s0"] * thetas[0] + pre["s1"] * thetas[1] + G * thetas[2]
    Gp = P["B"] + sx * thetas[0] + sy * thetas[1] + G * thetas[2]
    term1 = (KL * ((R_eval - 1) % R_MOD)
             + (Gp * R_eval - F * R_wX_eval) * (kappa0 * (chi - 1) % R_MOD)
             + (Gp * R_eval - F * R_wXwY_eval) * (kappa0 * kappa0 % R_MOD * K0_eval % R_MOD)
             - P["Q_CX"] * t_mi_eval - P["Q_CY"] * t_smax_eval)
    lhs_c = (term1 * k1(2) + (P["R"] - G * R_eval) * k1(3) + (P["R"] - G * R_wX_eval) * kappa2 + (P["R"] - G * R_wXwY_eval) * k2(2))
    # lhs_binding
    lhs_b = P["A_free"] * ((1 + kappa2 * k1(4)) % R_MOD) - G * (kappa2 * k1(4) % R_MOD * a_eval % R_MOD)
    lhs = lhs_b + (lhs_a + lhs_c) * kappa2
    aux = (P["Pi_X"] * (kappa2 * chi % R_MOD) + P["Pi_Y"] * (kappa2 * zeta % R_MOD)
           + P["M_X"] * (k2(2) * w_mi_inv % R_MOD * chi % R_MOD) + P["M_Y"] * (k2(2) * zeta % R_MOD)
           + P["N_X"] * (k2(3) * w_mi_inv % R_MOD * chi % R_MOD) + P["N_Y"] * (k2(3) * w_s_inv % R_MOD * zeta % R_MOD))
    aux_x = P["Pi_X"] * kappa2 + P["M_X"] * k2(2) + P["N_X"] * k2(3)
    aux_y = P["Pi_Y"] * kappa2 + P["M_Y"] * k2(2) + P["N_Y"] * k2(3)
    s2 = sigma.sigma2
    left_g1 = [(lhs + aux).p, points["B"], points["U"], points["V"], points["W"]]
    left_g2 = [sigma.H, s2.alpha4, s2.alpha, s2.alpha2, s2.alpha3]
    right_g1 = [(pre["O_pub_fix"] + P["O_pub_free"]).p, points["O_mid"], points["O_prv"], aux_x.p, aux_y.p]
    right_g2 = [s2.gamma, s2.eta, s2.delta, s2.x, s2.y]
    return pairing.pairing_products_equal(left_g1, left_g2, right_g1, right_g2)


def verify_arith(params, sigma, points, scalars, proof4_test):
    """Verifier::verify_arith (:291-305): the arithmetic-constraint argument alone (testing-mode helper)."""
    P = {k: _G(v) for k, v in points.items()}
    T = {k: _G(v) for k, v in proof4_test.items()}
    thetas, kappa0, chi, zeta, kappa1 = collect_challenges(points, scalars)
    t_n_eval = (pow(chi, params.n, R_MOD) - 1) % R_MOD
    t_smax_eval = (pow(zeta, params.s_max, R_MOD) - 1) % R_MOD
    G = _G(sigma.G)
    V_eval = scalars["V_eval"]
    lhs_a = P["U"] * V_eval - P["W"] + (P["V"] - G * V_eval) * kappa1 - P["Q_AX"] * t_n_eval - P["Q_AY"] * t_smax_eval
    aux_a = T["Pi_AX"] * chi + T["Pi_AY"] * zeta
    s2 = sigma.sigma2
    return pairing.pairing_products_equal([(lhs_a + aux_a).p], [sigma.H], [proof4_test["Pi_AX"], proof4_test["Pi_AY"]], [s2.x, s2.y])
