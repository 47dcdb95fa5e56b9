"""Protocol driver: Prover.init and prove0..prove4, a restatement of prove/src/lib.rs:675-3206
over the backend interface.  Every polynomial stays on the backend (device-resident for GpuBackend); scalars are Python
integers mod r.  `prove()` runs the reference's main (prove/src/main.rs:39-62): init -> prove0 -> thetas -> prove1 ->
kappa0 -> prove2 -> (chi, zeta) -> prove3 -> kappa1 -> prove4, and returns the proof in the reference's layout."""
import secrets
import time
from dataclasses import dataclass
from typing import List


from .. import frs_from_ints, frs_sparse
from ..transcript import TranscriptManager
from . import qap
from .formats import format_proof
from .fr import R_MOD, inv, root_of_unity

ONE = 1
NEG_ONE = R_MOD - 1


@dataclass
class Mixer:
    """The 17 blinding scalars of Prover::init (prove/src/lib.rs:1048-1090); rW_X / rW_Y are padded to 4 with a zero."""
    rU_X: int
    rU_Y: int
    rV_X: int
    rV_Y: int
    rW_X: List[int]
    rW_Y: List[int]
    rB_X: List[int]
    rB_Y: List[int]
    rR_X: int
    rR_Y: int
    rO_mid: int

    @staticmethod
    def random():
        r = lambda: secrets.randbelow(R_MOD)
        return Mixer(r(), r(), r(), r(), [r(), r(), r(), 0], [r(), r(), r(), 0], [r(), r()], [r(), r()], r(), r(), r())

    @staticmethod
    def fixed(seed=0x5EED):
        """Deterministic blinding for byte-identical proof comparisons (BASELINE.json: 'under fixed blinding scalars')."""
        import hashlib

        def r(i):
            return int.from_bytes(hashlib.sha256(f"mixer-{seed}-{i}".encode()).digest(), "big") % R_MOD

        return Mixer(r(0), r(1), r(2), r(3), [r(4), r(5), r(6), 0], [r(7), r(8), r(9), 0], [r(10), r(11)], [r(12), r(13)], r(14), r(15), r(16))


def _lincomb(terms):
    """sum of c * X^sx * Y^sy * p over terms (c, p, sx, sy): one fused pass where the backend's polynomial type offers
    `lincomb` (tkm_poly_lincomb on the device), else the reference's chain of scalings, shifts and additions."""
    p0 = terms[0][1]
    if hasattr(type(p0), "lincomb"):
        return type(p0).lincomb([(c % R_MOD, p, sx, sy) for c, p, sx, sy in terms])
    acc = None
    for c, p, sx, sy in terms:
        t = (p.mul_monomial(sx, sy) if (sx or sy) else p) * (c % R_MOD)
        acc = t if acc is None else acc + t
    return acc


def comb(*terms):
    """poly_comb! (prove/src/lib.rs:30-38): sum_k c_k * p_k."""
    return _lincomb([(c, p, 0, 0) for c, p in terms])


def mul_by_x_minus_one(p):
    return _lincomb([(1, p, 1, 0), (R_MOD - 1, p, 0, 0)])


def mul_by_one_minus_x(p):
    return _lincomb([(1, p, 0, 0), (R_MOD - 1, p, 1, 0)])


def mul_by_linear_x(p, c):
    return _lincomb([(c[0], p, 0, 0), (c[1], p, 1, 0)])


def mul_by_linear_y(p, c):
    return _lincomb([(c[0], p, 0, 0), (c[1], p, 0, 1)])


def mul_by_term9(p, rB_X, rB_Y, t_mi_eval, t_smax_eval):
    const = (t_mi_eval * rB_X[0] + t_smax_eval * rB_Y[0]) % R_MOD
    return _lincomb([(const, p, 0, 0), (t_mi_eval * rB_X[1] % R_MOD, p, 1, 0), (t_smax_eval * rB_Y[1] % R_MOD, p, 0, 1)])


def _pow2(v):
    r = 1
    while r < v:
        r <<= 1
    return r


class Timings:
    def __init__(self):
        self.spans = {}

    def add(self, name, dt):
        self.spans[name] = self.spans.get(name, 0.0) + dt


class Prover:
    def __init__(self, backend, params, infos, r1cs_list, sigma, placements, permutation, instance, mixer=None, checks=False, library_csr=None):
        """Prover::init (prove/src/lib.rs:675-1206).  `checks` re-runs the reference's debug assertions (R1CS grid check,
        quotient identities at a random point)."""
        self.be, self.p, self.sigma, self.checks = backend, params, sigma, checks
        self.t = Timings()
        t0 = time.perf_counter()
        p = params
        p.validate()
        n, s_max, m_i, l_free = p.n, p.s_max, p.l_D - p.l, p.l_free
        self.m_i = m_i
        backend.init_ntt_domain(max(2 * n, 4 * m_i) * 2 * s_max)
        self.mixer = mixer or Mixer.random()
        # witness polynomials (gen_bXY, read_R1CS_gen_uvwXY)
        wt = qap.WitnessTable(p, placements, infos)
        self.t.add("init.witness_table", time.perf_counter() - t0)
        t1 = time.perf_counter()
        csr = library_csr or qap.LibraryCSR(r1cs_list)
        self.t.add("init.library_csr", time.perf_counter() - t1)
        t1 = time.perf_counter()
        self.uXY, self.vXY, self.wXY = backend.uvw_polys(p, csr, wt)
        self.t.add("init.build.witness.uvwXY", time.perf_counter() - t1)
        t1 = time.perf_counter()
        if hasattr(backend, "interface_poly"):  # scattered on the device from the resident witness table
            self.bXY = backend.interface_poly(p, wt)
        else:
            self.bXY = backend.from_rou_evals(qap.interface_evals_from_table(p, wt), m_i, s_max)
        self.t.add("init.build.witness.bXY", time.perf_counter() - t1)
        self.rXY = None
        t1 = time.perf_counter()
        # instance polynomials
        public_instance = list(instance.a_pub_user[:p.l_user]) + list(instance.a_pub_block[:l_free - p.l_user])
        if len(public_instance) != l_free:
            raise ValueError("instance length mismatch: expected l_free user+block values")
        self.a_free_X = backend.from_rou_evals(frs_from_ints(public_instance), l_free, 1)
        self.t_n = self._vanishing_x(n)
        self.t_mi = self._vanishing_x(m_i)
        self.t_smax = self._vanishing_y(s_max)
        self.omega_m_i, self.omega_s_max = root_of_unity(m_i), root_of_unity(s_max)
        if hasattr(backend, "permutation_polys"):  # built on the device: no m_I x s_max table on the host
            self.s0XY, self.s1XY = backend.permutation_polys(permutation, m_i, s_max, self.omega_m_i, self.omega_s_max)
        else:
            s0_ev, s1_ev = qap.permutation_evals(permutation, m_i, s_max, self.omega_m_i, self.omega_s_max)
            self.s0XY = backend.from_rou_evals(s0_ev, m_i, s_max)
            self.s1XY = backend.from_rou_evals(s1_ev, m_i, s_max)
        self.q = [None] * 4  # q0 (Q_AX part), q1, q2 (Q_CX part), q3
        self.cache = {}
        self.t.add("init.build.instance", time.perf_counter() - t1)
        self.t.add("init.build", time.perf_counter() - t0)
        t1 = time.perf_counter()
        self.binding = self._binding(placements, infos, wt)
        if hasattr(backend, "release_witness"):
            backend.release_witness(wt)
        self.t.add("init.binding", time.perf_counter() - t1)
        self.t.add("init", time.perf_counter() - t0)

    # ---- small fixed polynomials
    def _vanishing_x(self, k):
        return self.be.from_coeffs(frs_sparse(2 * k, {0: NEG_ONE, k: ONE}), 2 * k, 1)

    def _vanishing_y(self, k):
        return self.be.from_coeffs(frs_sparse(2 * k, {0: NEG_ONE, k: ONE}), 1, 2 * k)

    def _low_degree_x_times_vanishing(self, coeffs, exponent):
        size = _pow2(exponent + len(coeffs))
        out = {}
        for i, c in enumerate(coeffs):
            out[i] = (out.get(i, 0) - c) % R_MOD
            out[i + exponent] = (out.get(i + exponent, 0) + c) % R_MOD
        return frs_sparse(size, out), size

    def low_degree_x_times_vanishing(self, coeffs, exponent):
        out, size = self._low_degree_x_times_vanishing(coeffs, exponent)
        return self.be.from_coeffs(out, size, 1)

    def low_degree_y_times_vanishing(self, coeffs, exponent):
        out, size = self._low_degree_x_times_vanishing(coeffs, exponent)
        return self.be.from_coeffs(out, 1, size)

    def _mono(self, x):
        return self.be.from_coeffs(frs_from_ints([0, 1]), 2, 1) if x else self.be.from_coeffs(frs_from_ints([0, 1]), 1, 2)

    def _lagrange(self, size, idx, along_x):
        ev = frs_sparse(size, {idx: 1})
        return self.be.from_rou_evals(ev, size, 1) if along_x else self.be.from_rou_evals(ev, 1, size)

    def encode(self, poly, name):
        """Sigma1::encode_poly.  With a backend that can queue commitments (commit_async) this returns a pending handle:
        the stage collects its commitments and resolves them together (`_resolve`), so the serial tail of one MSM overlaps
        the next polynomial combination and accumulation instead of stalling the device."""
        t0 = time.perf_counter()
        if hasattr(self.be, "commit_async"):
            pt = self.be.commit_async(self.sigma.xy_powers, poly)
        else:
            pt = self.be.commit(self.sigma.xy_powers, poly)
        self.t.add("encode", time.perf_counter() - t0)
        self.t.add("encode." + name, time.perf_counter() - t0)
        return pt

    def _resolve(self, d):
        t0 = time.perf_counter()
        out = {k: (v.get() if hasattr(v, "get") else v) for k, v in d.items()}
        self.t.add("encode", time.perf_counter() - t0)
        self.t.add("encode.wait", time.perf_counter() - t0)
        return out

    # ---- binding (prove/src/lib.rs:1092-1176; sparse MSMs: group_structures/mod.rs:145-300)
    def _binding(self, placements, infos, wt):
        be, p, sg, mx = self.be, self.p, self.sigma, self.mixer
        A_free = self.encode(self.a_free_X, "A_free")
        # the four encodings and the two blinding sums are independent: a backend that can queue MSMs returns pending handles
        # and the serial tails overlap (one wait at the end instead of six)
        O_pub_free = sg.encode_O_pub_free(be, placements, infos, p, defer=True)
        O_mid_core = sg.encode_O_mid_no_zk(be, wt, p, defer=True)
        O_prv_core = sg.encode_O_prv_no_zk(be, wt, p, defer=True)
        msm_points = be.msm_points_async if hasattr(be, "msm_points_async") else be.msm_points
        # zero-knowledge terms (prove/src/lib.rs:1131-1160): the 17 scalar multiples as two small MSMs
        zk_mid = msm_points([sg.delta], [mx.rO_mid])
        terms = [(sg.eta, (-mx.rO_mid) % R_MOD), (sg.delta_inv_alphak_xh_tx[0][0], mx.rU_X), (sg.delta_inv_alphak_xh_tx[1][0], mx.rV_X)]
        terms += [(sg.delta_inv_alphak_xh_tx[2][h], mx.rW_X[h]) for h in range(3)]
        terms += [(sg.delta_inv_alpha4_xj_tx[j], mx.rB_X[j]) for j in range(2)]
        terms += [(sg.delta_inv_alphak_yi_ty[0][0], mx.rU_Y), (sg.delta_inv_alphak_yi_ty[1][0], mx.rV_Y)]
        terms += [(sg.delta_inv_alphak_yi_ty[2][i], mx.rW_Y[i]) for i in range(3)]
        terms += [(sg.delta_inv_alphak_yi_ty[3][i], mx.rB_Y[i]) for i in range(2)]
        zk_prv = msm_points([t[0] for t in terms], [t[1] for t in terms])
        r = self._resolve({"A_free": A_free, "O_pub_free": O_pub_free, "O_mid_core": O_mid_core, "O_prv_core": O_prv_core, "zk_mid": zk_mid, "zk_prv": zk_prv})
        return {"A_free": r["A_free"], "O_pub_free": r["O_pub_free"], "O_mid": be.g1_add(r["O_mid_core"], r["zk_mid"]),
                "O_prv": be.g1_add(r["O_prv_core"], r["zk_prv"])}

    # ---- prove0 (prove/src/lib.rs:1446-1782)
    def prove0(self):
        t0 = time.perf_counter()
        p, mx = self.p, self.mixer
        p0 = self.uXY * self.vXY - self.wXY
        self.q[0], self.q[1] = p0.div_by_vanishing_opt(p.n, p.s_max)
        if self.checks:
            self._check_quotient(p0, self.q[0], self.q[1], p.n, p.s_max)
        rW_X = self.be.from_coeffs(frs_from_ints(mx.rW_X), len(mx.rW_X), 1)
        rW_Y = self.be.from_coeffs(frs_from_ints(mx.rW_Y), 1, len(mx.rW_Y))
        U = self.encode(comb((ONE, self.uXY), (mx.rU_X, self.t_n), (mx.rU_Y, self.t_smax)), "U")
        V = self.encode(comb((ONE, self.vXY), (mx.rV_X, self.t_n), (mx.rV_Y, self.t_smax)), "V")
        W_zk = self.low_degree_x_times_vanishing(mx.rW_X, p.n) + self.low_degree_y_times_vanishing(mx.rW_Y, p.s_max)
        self.cache["w_zk"] = W_zk
        W = self.encode(self.wXY + W_zk, "W")
        Q_AX = self.encode(comb((ONE, self.q[0]), (mx.rU_X, self.vXY), (mx.rV_X, self.uXY), (NEG_ONE, rW_X),
                                (mx.rU_X * mx.rV_X, self.t_n), (mx.rU_Y * mx.rV_X, self.t_smax)), "Q_AX")
        Q_AY = self.encode(comb((ONE, self.q[1]), (mx.rU_Y, self.vXY), (mx.rV_Y, self.uXY), (NEG_ONE, rW_Y),
                                (mx.rU_X * mx.rV_Y, self.t_n), (mx.rU_Y * mx.rV_Y, self.t_smax)), "Q_AY")
        term_B_zk = self.low_degree_x_times_vanishing(mx.rB_X, self.m_i) + self.low_degree_y_times_vanishing(mx.rB_Y, p.s_max)
        self.cache["term_b_zk"] = term_B_zk
        B = self.encode(self.bXY + term_B_zk, "B")
        out = self._resolve({"U": U, "V": V, "W": W, "Q_AX": Q_AX, "Q_AY": Q_AY, "B": B})
        self.t.add("prove0", time.perf_counter() - t0)
        return out

    def _check_quotient(self, pXY, qx, qy, c, d):
        xe, ye = secrets.randbelow(R_MOD), secrets.randbelow(R_MOD)
        lhs = pXY.eval(xe, ye)
        rhs = (qx.eval(xe, ye) * (pow(xe, c, R_MOD) - 1) + qy.eval(xe, ye) * (pow(ye, d, R_MOD) - 1)) % R_MOD
        if lhs != rhs:
            raise AssertionError("quotient relation does not hold: the witness does not satisfy the constraints")

    def _fg(self, thetas):
        f = self.bXY + self.s0XY * thetas[0] + self.s1XY * thetas[1] + thetas[2]
        g = self.bXY + self._mono(True) * thetas[0] + self._mono(False) * thetas[1] + thetas[2]
        return f, g

    # ---- prove1 (:1784-1956)
    def prove1(self, thetas):
        t0 = time.perf_counter()
        p, mx = self.p, self.mixer
        f, g = self._fg(thetas)
        self.rXY = self.be.recursion_poly(f, g, self.m_i, p.s_max) if hasattr(self.be, "recursion_poly") else None
        if self.rXY is None:
            r_evals = self.be.recursion_evals(f.to_rou_evals(), g.to_rou_evals(), self.m_i, p.s_max)
            self.rXY = self.be.from_rou_evals(r_evals, self.m_i, p.s_max)
        RXY = self.rXY + (self.t_mi * mx.rR_X + self.t_smax * mx.rR_Y)
        out = self._resolve({"R": self.encode(RXY, "R")})
        self.t.add("prove1", time.perf_counter() - t0)
        return out

    # ---- prove2 (:1958-2270)
    def prove2(self, thetas, kappa0):
        t0 = time.perf_counter()
        p, mx, m_i, s_max = self.p, self.mixer, self.m_i, self.p.s_max
        k0sq = kappa0 * kappa0 % R_MOD
        r = self.rXY
        r_wX = r.scale_coeffs_x(inv(self.omega_m_i))
        r_wXwY = r_wX.scale_coeffs_y(inv(self.omega_s_max))
        f, g = self._fg(thetas)
        KL = self._lagrange(m_i, m_i - 1, True) * self._lagrange(s_max, s_max - 1, False)
        self.cache["lagrange_kl"] = KL
        K0 = self._lagrange(m_i, 0, True)
        # p_comb = (r - 1) KL + kappa0 (X - 1)(r g - r(X/w, Y) f) + kappa0^2 K0 (r g - r(X/w, Y/w) f)
        p_comb = self.be.p_comb(r, g, f, r_wX, r_wXwY, KL, K0, kappa0, 4 * m_i, 2 * s_max) if hasattr(self.be, "p_comb") else None
        if p_comb is None:
            rg = r * g
            p1 = (r - ONE) * KL
            p2 = mul_by_x_minus_one(rg - r_wX * f)
            p3 = K0 * (rg - r_wXwY * f)
            p_comb = comb((ONE, p1), (kappa0, p2), (k0sq, p3))
        self.q[2], self.q[3] = p_comb.div_by_vanishing_opt(m_i, s_max)
        if self.checks:
            self._check_quotient(p_comb, self.q[2], self.q[3], m_i, s_max)
        r_D1, r_D2, g_D = r - r_wX, r - r_wXwY, g - f
        out = {}
        for name, q, rB, rR, lin in (("Q_CX", self.q[2], mx.rB_X, mx.rR_X, mul_by_linear_x), ("Q_CY", self.q[3], mx.rB_Y, mx.rR_Y, mul_by_linear_y)):
            d1 = lin(r_D1, rB) + g_D * rR
            d2 = lin(r_D2, rB) + g_D * rR
            out[name] = self.encode(comb((ONE, q), (rR, KL), (kappa0, mul_by_x_minus_one(d1)), (k0sq, K0 * d2)), name)
        out = self._resolve(out)
        self.t.add("prove2", time.perf_counter() - t0)
        return out

    # ---- prove3 (:2272-2354)
    def prove3(self, chi, zeta):
        t0 = time.perf_counter()
        mx = self.mixer
        VXY = comb((ONE, self.vXY), (mx.rV_X, self.t_n), (mx.rV_Y, self.t_smax))
        V_eval = VXY.eval(chi, zeta)
        RXY = self.rXY + (self.t_mi * mx.rR_X + self.t_smax * mx.rR_Y)
        R_eval = RXY.eval(chi, zeta)
        R_wX = RXY.scale_coeffs_x(inv(self.omega_m_i))
        R_omegaX_eval = R_wX.eval(chi, zeta)
        R_omegaX_omegaY_eval = R_wX.scale_coeffs_y(inv(self.omega_s_max)).eval(chi, zeta)
        self.t.add("prove3", time.perf_counter() - t0)
        return {"V_eval": V_eval, "R_eval": R_eval, "R_omegaX_eval": R_omegaX_eval, "R_omegaX_omegaY_eval": R_omegaX_omegaY_eval}

    # ---- prove4 (:2356-3206)
    def prove4(self, proof3, thetas, kappa0, chi, zeta, kappa1):
        t0 = time.perf_counter()
        be, p, mx, m_i, s_max = self.be, self.p, self.mixer, self.m_i, self.p.s_max
        M = lambda *a: _prod(a)
        # Pi_A: arithmetic-constraint opening
        t_n_eval = (pow(chi, p.n, R_MOD) - 1) % R_MOD
        t_smax_eval = (pow(zeta, s_max, R_MOD) - 1) % R_MOD
        small_v_eval = self.vXY.eval(chi, zeta)
        rW_X = be.from_coeffs(frs_from_ints(mx.rW_X), len(mx.rW_X), 1)
        rW_Y = be.from_coeffs(frs_from_ints(mx.rW_Y), 1, len(mx.rW_Y))
        W_zk = self.cache.get("w_zk") or (self.low_degree_x_times_vanishing(mx.rW_X, p.n) + self.low_degree_y_times_vanishing(mx.rW_Y, s_max))
        VXY = comb((ONE, self.vXY), (mx.rV_X, self.t_n), (mx.rV_Y, self.t_smax))
        pA = comb((kappa1, VXY - proof3["V_eval"]),
                  (small_v_eval, self.uXY), (NEG_ONE, self.wXY),
                  (M(NEG_ONE, t_n_eval), self.q[0]), (M(NEG_ONE, t_smax_eval), self.q[1]),
                  (M(small_v_eval, mx.rU_X), self.t_n), (M(small_v_eval, mx.rU_Y), self.t_smax),
                  ((-(mx.rU_X * t_n_eval + mx.rU_Y * t_smax_eval)) % R_MOD, self.vXY),
                  (t_n_eval, rW_X), (t_smax_eval, rW_Y), (NEG_ONE, W_zk))
        Pi_AX_XY, Pi_AY_XY, _rem = pA.div_by_ruffini(chi, zeta)
        Pi_AX, Pi_AY = self.encode(Pi_AX_XY, "Pi_AX"), self.encode(Pi_AY_XY, "Pi_AY")
        # M, N: openings of R at (chi/w, zeta) and (chi/w, zeta/w)
        w_inv_x, w_inv_y = inv(self.omega_m_i), inv(self.omega_s_max)
        RXY = self.rXY + (self.t_mi * mx.rR_X + self.t_smax * mx.rR_Y)
        M_X_XY, M_Y_XY, rem2 = (RXY - proof3["R_omegaX_eval"]).div_by_ruffini(w_inv_x * chi % R_MOD, zeta)
        M_X, M_Y = self.encode(M_X_XY, "M_X"), self.encode(M_Y_XY, "M_Y")
        N_X_XY, N_Y_XY, rem3 = (RXY - proof3["R_omegaX_omegaY_eval"]).div_by_ruffini(w_inv_x * chi % R_MOD, w_inv_y * zeta % R_MOD)
        N_X, N_Y = self.encode(N_X_XY, "N_X"), self.encode(N_Y_XY, "N_Y")
        if self.checks:
            assert rem2 == 0 and rem3 == 0, "R opening remainders must vanish"
        # Pi_C: copy-constraint opening
        r = self.rXY
        r_wX = r.scale_coeffs_x(w_inv_x)
        r_wXwY = r_wX.scale_coeffs_y(w_inv_y)
        f, g = self._fg(thetas)
        t_mi_eval = (pow(chi, m_i, R_MOD) - 1) % R_MOD
        K0 = self._lagrange(m_i, 0, True)
        K0_eval = K0.eval(chi, zeta)
        small_r_eval, small_r_wX_eval, small_r_wXwY_eval = r.eval(chi, zeta), r_wX.eval(chi, zeta), r_wXwY.eval(chi, zeta)
        KL = self.cache.get("lagrange_kl") or (self._lagrange(m_i, m_i - 1, True) * self._lagrange(s_max, s_max - 1, False))
        term5 = comb((small_r_eval, g), ((-small_r_wX_eval) % R_MOD, f))
        term6 = comb((small_r_eval, g), ((-small_r_wXwY_eval) % R_MOD, f))
        k0sq = kappa0 * kappa0 % R_MOD
        pC = comb(((small_r_eval - 1) % R_MOD, KL), (M(kappa0, (chi - 1) % R_MOD), term5), (M(k0sq, K0_eval), term6),
                  ((-t_mi_eval) % R_MOD, self.q[2]), ((-t_smax_eval) % R_MOD, self.q[3]))
        r_D1, r_D2 = r - r_wX, r - r_wXwY
        r_D1_eval, r_D2_eval = r_D1.eval(chi, zeta), r_D2.eval(chi, zeta)
        term_B_zk = self.cache.get("term_b_zk") or (self.low_degree_x_times_vanishing(mx.rB_X, m_i) + self.low_degree_y_times_vanishing(mx.rB_Y, s_max))
        term10 = (g - f) * ((mx.rR_X * t_mi_eval + mx.rR_Y * t_smax_eval) % R_MOD)
        d1 = mul_by_term9(r_D1, mx.rB_X, mx.rB_Y, t_mi_eval, t_smax_eval) + term10
        LHS_zk1 = comb((M((chi - 1) % R_MOD, r_D1_eval), term_B_zk), (ONE, mul_by_one_minus_x(d1)), ((chi - 1) % R_MOD, term10))
        d2 = mul_by_term9(r_D2, mx.rB_X, mx.rB_Y, t_mi_eval, t_smax_eval) + term10
        LHS_zk2 = comb((M(K0_eval, r_D2_eval), term_B_zk), (K0_eval, term10), (NEG_ONE, K0 * d2))
        k1sq = kappa1 * kappa1 % R_MOD
        LHS_for_copy = comb((k1sq, pC), (M(k1sq, kappa0), LHS_zk1), (M(k1sq, k0sq), LHS_zk2), (M(k1sq, kappa1), RXY - proof3["R_eval"]))
        Pi_CX_XY, Pi_CY_XY, rem1 = LHS_for_copy.div_by_ruffini(chi, zeta)
        if self.checks:
            assert rem1 == 0, "copy-constraint opening remainder must vanish"
        Pi_CX, Pi_CY = self.encode(Pi_CX_XY, "Pi_CX"), self.encode(Pi_CY_XY, "Pi_CY")
        # Pi_B: opening of a_free
        A_eval = self.a_free_X.eval(chi, zeta)
        pi_B_XY, _piBy, _ = (self.a_free_X - A_eval).div_by_ruffini(chi, zeta)
        r = self._resolve({"Pi_AX": Pi_AX, "Pi_AY": Pi_AY, "M_X": M_X, "M_Y": M_Y, "N_X": N_X, "N_Y": N_Y, "Pi_CX": Pi_CX, "Pi_CY": Pi_CY,
                           "Pi_B": self.encode(pi_B_XY, "Pi_B")})
        Pi_AX, Pi_AY, M_X, M_Y, N_X, N_Y, Pi_CX, Pi_CY = (r[k] for k in ("Pi_AX", "Pi_AY", "M_X", "M_Y", "N_X", "N_Y", "Pi_CX", "Pi_CY"))
        Pi_B = be.g1_mul(r["Pi_B"], pow(kappa1, 4, R_MOD))
        Pi_X = be.g1_add(be.g1_add(Pi_AX, Pi_CX), Pi_B)
        Pi_Y = be.g1_add(Pi_AY, Pi_CY)
        self.t.add("prove4", time.perf_counter() - t0)
        proof4 = {"Pi_X": Pi_X, "Pi_Y": Pi_Y, "M_X": M_X, "M_Y": M_Y, "N_X": N_X, "N_Y": N_Y}
        proof4_test = {"Pi_CX": Pi_CX, "Pi_CY": Pi_CY, "Pi_AX": Pi_AX, "Pi_AY": Pi_AY, "Pi_B": Pi_B, "M_X": M_X, "M_Y": M_Y, "N_X": N_X, "N_Y": N_Y}
        return proof4, proof4_test


def _prod(vals):
    acc = 1
    for v in vals:
        acc = acc * (v % R_MOD) % R_MOD
    return acc


def prove(prover: Prover):
    """prove/src/main.rs:39-62.  Returns (points, scalars, formatted_proof, proof4_test)."""
    t0 = time.perf_counter()
    mgr = TranscriptManager()
    p0 = prover.prove0()
    mgr.add_proof0(p0["U"], p0["V"], p0["W"], p0["Q_AX"], p0["Q_AY"], p0["B"])
    thetas = mgr.get_thetas()
    p1 = prover.prove1(thetas)
    mgr.add_proof1(p1["R"])
    kappa0 = mgr.get_kappa0()
    p2 = prover.prove2(thetas, kappa0)
    mgr.add_proof2(p2["Q_CX"], p2["Q_CY"])
    chi, zeta = mgr.get_chi_zeta()
    p3 = prover.prove3(chi, zeta)
    mgr.add_proof3(p3["V_eval"], p3["R_eval"], p3["R_omegaX_eval"], p3["R_omegaX_omegaY_eval"])
    kappa1 = mgr.get_kappa1()
    p4, p4_test = prover.prove4(p3, thetas, kappa0, chi, zeta, kappa1)
    points = dict(prover.binding)
    for d in (p0, p1, p2, p4):
        points.update(d)
    prover.t.add("prove0-4", time.perf_counter() - t0)
    return points, p3, format_proof(points, p3), p4_test
