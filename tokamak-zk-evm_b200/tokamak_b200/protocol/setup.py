"""Trusted setup: Tau, Sigma1 / Sigma2 generation as `trusted-setup --fixed-tau` does it
(setup/trusted-setup/src/main.rs:71-190; Sigma::gen / Sigma1::gen / Sigma2::gen, libs/src/group_structures/mod.rs:312-551,
752-777).  The ~10.8 M generator multiples of Sigma1 are fixed-base multiplications on the device (backend.make_table);
Sigma2 is nine G2 scalar multiplications on the host."""
from dataclasses import dataclass

from . import pairing
from .fr import R_MOD, from_hex, inv, lagrange_bases_at, powers
from .qap import o_evaled


@dataclass
class Tau:
    x: int
    y: int
    alpha: int
    gamma: int
    delta: int
    eta: int

    @staticmethod
    def gen_fixed():
        """Tau::gen_fixed (libs/src/field_structures/mod.rs:43-65)."""
        return Tau(x=from_hex("0x7234cd9b97845e0125e84ae3ae81354e004558d8c82a83425652bc7b9ed49f7d"),
                   y=from_hex("0x6ed0eea55cbeeebdc7a41033ebd196ffecc1806fdbc13a8d41b8f1aa273a4037"),
                   alpha=from_hex("0x7234cd9b97845e0125e84ae3ae81354e004558d8c82a83425652bc7b9ed49f7d"),
                   gamma=from_hex("0x088dfe3d1b76775ec267d6d0e27b753ec904c76e0bc32ca8223dc2ae1a0ac6b4"),
                   delta=from_hex("0x04b8ce26374c547d8722ac51f5ed1e0f9cb891c332c69c865d96af150189a818"),
                   eta=from_hex("0x52eb2aeb35b72b94a19ea232e984850f2cda5542fdc10368955d8ac6274f8579"))


# the hard-coded generators of --fixed-tau (setup/trusted-setup/src/main.rs:71-78)
G1_FIXED = (0x0B001B4CC05FA01578BE7D4E821D6FF58F2A05C584FBA3CB31A37942DECE65EADEC9A878ADD2282F7C2513ABB8D4AB05,
            0x15E237775397ED22EEF43DD36CDCA277C9CF6FA7E4FFFF0A5BB4B20A82392CAACF0F63FB6CDB02BCCF2F5AF14970D6B9)
_G2X = "1116094a7c01d4fd8abcfea69c658c92c037765bee00556b8d4063c33540b316ac68a2d913d3adc3b43c7d7cc7505cfc17206c8ae661f247979b3f1daa7fb6d5f7ce9c17b5ed1d7e8b421a2508b3f09a603e6a5fab3fcde7364fd178d656ac36"
_G2Y = "15bf297a4b9842fb1a3a6f2dbf6b94de06997b11b2f72436c22efbb48d2f74b0de7239ea182a2ee50c23ae3d0be6fdee09459611409874fe4b04b1a7e42cb84eb4ae01728dc55dbd1343fda8d0fe94a299fc757acc1d2602a49a005b4ff90190"


def _g2_from_hex(h):
    """G2BaseField::from_hex of a 96-byte big-endian string: the high half is the imaginary part c1, the low half c0
    (little-endian limb order c0 || c1 read from a big-endian literal)."""
    return (int(h[96:], 16), int(h[:96], 16))


G2_FIXED = (_g2_from_hex(_G2X), _g2_from_hex(_G2Y))


@dataclass
class Sigma2:
    alpha: tuple
    alpha2: tuple
    alpha3: tuple
    alpha4: tuple
    gamma: tuple
    delta: tuple
    eta: tuple
    x: tuple
    y: tuple


class Sigma:
    """sigma = ([sigma_1]_1, [sigma_2]_2): G1 tables are backend handles (device-resident for the GPU backend)."""

    def __init__(self):
        self.G = self.H = None
        self.sigma2 = None
        self.lagrange_KL = None
        # Sigma1
        self.xy_powers = None                       # table [h_max][2 s_max]: x^h y^i
        self.x = self.y = self.delta = self.eta = None
        self.gamma_inv_o_inst = None                # table [l][1]
        self.eta_inv_li_o_inter_alpha4_kj = None    # table [m_I][s_max]
        self.delta_inv_li_o_prv = None              # table [m_D - l_D][s_max]
        self.delta_inv_alphak_xh_tx = None          # [3][3] points, k in 1..3, h in 0..2
        self.delta_inv_alpha4_xj_tx = None          # [2] points
        self.delta_inv_alphak_yi_ty = None          # [4][3] points, k in 1..4, i in 0..2

    # ---- Sigma1's encoders, with the reference's names (group_structures/mod.rs:59-119,145-300,584-700)
    def encode_poly(self, backend, poly):
        """[P(x, y)]_1 over xy_powers (impl_encode_poly!)."""
        return backend.commit(self.xy_powers, poly)

    @staticmethod
    def _msm_indexed(backend, defer):
        """The sparse-gather MSM of the backend; with defer=True and a backend that can queue it, a pending handle
        (.get()) so that independent encodings overlap their serial tails."""
        return backend.msm_indexed_async if defer and hasattr(backend, "msm_indexed_async") else backend.msm_indexed

    def _encode_statement(self, backend, table, witness_table, lo, hi, s_max, defer):
        """encode_statement_common (group_structures/mod.rs:266-300).  A backend that keeps the placement variables on the
        device (msm_indexed_witness) receives only the two index vectors; otherwise the values are collected on the host."""
        if hasattr(backend, "msm_indexed_witness"):
            idx, rows = witness_table.gather_indices(lo, hi, s_max)
            return backend.msm_indexed_witness(table, idx, witness_table, rows, defer)
        idx, vals = witness_table.gather(lo, hi, s_max)
        return self._msm_indexed(backend, defer)(table, idx, vals)

    def encode_O_pub_free(self, backend, placements, infos, params, defer=False):
        """encode_o_pub_free_common: the public sides of bufferPubOut (outputs), bufferPubIn and bufferBlockIn (inputs)
        against gamma_inv_o_inst; bufferEVMIn belongs to O_pub_fix."""
        import numpy as np

        from .. import frs_from_ints

        idx, sc = [], []
        for pl in placements:
            info = infos[pl.subcircuitId]
            if info.name == "bufferPubOut":
                s0, cnt = info.Out_idx
            elif info.name in ("bufferPubIn", "bufferBlockIn"):
                s0, cnt = info.In_idx
            else:
                continue
            for j in range(s0, s0 + cnt):
                idx.append(info.flattenMap[j])
                sc.append(pl.variables[j])
        return self._msm_indexed(backend, defer)(self.gamma_inv_o_inst, np.array(idx, dtype=np.uint32), frs_from_ints(sc))

    def encode_O_pub_fix(self, backend, a_pub_function, params):
        """encode_o_pub_fix_common: the function instance against the last m_function entries of gamma_inv_o_inst."""
        import numpy as np

        from .. import frs_from_ints

        m_function = params.l - params.l_free
        if m_function == 0:
            return None
        if len(a_pub_function) != m_function:
            raise ValueError(f"a_pub_function length mismatch: expected m_function={m_function}, got a_pub_function.len()={len(a_pub_function)}")
        start = params.l - m_function
        return backend.msm_indexed(self.gamma_inv_o_inst, np.arange(start, start + m_function, dtype=np.uint32), frs_from_ints(a_pub_function))

    def encode_O_mid_no_zk(self, backend, witness_table, params, defer=False):
        """encode_statement_common over the interface wires [l, l_D) against eta_inv_li_o_inter_alpha4_kj[wire][placement]."""
        return self._encode_statement(backend, self.eta_inv_li_o_inter_alpha4_kj, witness_table, params.l, params.l_D, params.s_max, defer)

    def encode_O_prv_no_zk(self, backend, witness_table, params, defer=False):
        """encode_statement_common over the private wires [l_D, m_D) against delta_inv_li_o_prv[wire][placement]."""
        return self._encode_statement(backend, self.delta_inv_li_o_prv, witness_table, params.l_D, params.m_D, params.s_max, defer)


def generate(backend, params, infos, r1cs_list, tau: Tau, g1_gen=G1_FIXED, g2_gen=G2_FIXED):
    p = params
    n, s_max, l, l_free, l_user = p.n, p.s_max, p.l, p.l_free, p.l_user
    m_i = p.l_D - l
    m_block, m_function = l_free - l_user, l - l_free
    k_vec = lagrange_bases_at(tau.x, m_i)
    l_vec = lagrange_bases_at(tau.y, s_max)
    m_vec = lagrange_bases_at(tau.x, l_free)
    o_vec = o_evaled(p, infos, r1cs_list, tau)

    sg = Sigma()
    sg.G, sg.H = g1_gen, g2_gen
    sg.lagrange_KL = backend.g1_mul(g1_gen, l_vec[s_max - 1] * k_vec[m_i - 1] % R_MOD)
    # Sigma1 (Sigma1::gen, group_structures/mod.rs:361-551)
    h_max = max(2 * n, 2 * m_i)
    sg.xy_powers = backend.make_table(powers(tau.x, h_max), powers(tau.y, 2 * s_max), g1_gen)
    sg.x, sg.y = backend.g1_mul(g1_gen, tau.x), backend.g1_mul(g1_gen, tau.y)
    sg.delta, sg.eta = backend.g1_mul(g1_gen, tau.delta), backend.g1_mul(g1_gen, tau.eta)
    # gamma^-1 (L_t(y) o_j(x) + M_j(x)): t = 0 user outputs, 1 user inputs, 2 block, 3 function (no M_j for the function part)
    user = [l_vec[0]] * p.l_user_out + [l_vec[1]] * (l_user - p.l_user_out) + [l_vec[2]] * m_block + [l_vec[3]] * m_function
    if len(user) != l:
        raise ValueError("user_vec length mismatch: expected l")
    ginv = inv(tau.gamma)
    col = [(user[j] * o_vec[j] + (m_vec[j] if j < l_free else 0)) % R_MOD * ginv % R_MOD for j in range(l)]
    sg.gamma_inv_o_inst = backend.make_table(col, [1], g1_gen)
    a4 = pow(tau.alpha, 4, R_MOD)
    einv, dinv = inv(tau.eta), inv(tau.delta)
    sg.eta_inv_li_o_inter_alpha4_kj = backend.make_table([(o_vec[l + j] + a4 * k_vec[j]) % R_MOD * einv % R_MOD for j in range(m_i)], l_vec, g1_gen)
    sg.delta_inv_li_o_prv = backend.make_table([o_vec[j] * dinv % R_MOD for j in range(l + m_i, p.m_D)], l_vec, g1_gen)
    t_n = (pow(tau.x, n, R_MOD) - 1) % R_MOD
    t_mi = (pow(tau.x, m_i, R_MOD) - 1) % R_MOD
    t_s = (pow(tau.y, s_max, R_MOD) - 1) % R_MOD
    sg.delta_inv_alphak_xh_tx = [[backend.g1_mul(g1_gen, dinv * pow(tau.alpha, k, R_MOD) % R_MOD * pow(tau.x, h, R_MOD) % R_MOD * t_n % R_MOD)
                                  for h in range(3)] for k in range(1, 4)]
    sg.delta_inv_alpha4_xj_tx = [backend.g1_mul(g1_gen, dinv * a4 % R_MOD * pow(tau.x, j, R_MOD) % R_MOD * t_mi % R_MOD) for j in range(2)]
    sg.delta_inv_alphak_yi_ty = [[backend.g1_mul(g1_gen, dinv * pow(tau.alpha, k, R_MOD) % R_MOD * pow(tau.y, i, R_MOD) % R_MOD * t_s % R_MOD)
                                  for i in range(3)] for k in range(1, 5)]
    # Sigma2 (Sigma2::gen, :752-777)
    g2 = lambda k: pairing.g2_mul(g2_gen, k % R_MOD)
    sg.sigma2 = Sigma2(alpha=g2(tau.alpha), alpha2=g2(pow(tau.alpha, 2, R_MOD)), alpha3=g2(pow(tau.alpha, 3, R_MOD)), alpha4=g2(a4),
                       gamma=g2(tau.gamma), delta=g2(tau.delta), eta=g2(tau.eta), x=g2(tau.x), y=g2(tau.y))
    return sg
