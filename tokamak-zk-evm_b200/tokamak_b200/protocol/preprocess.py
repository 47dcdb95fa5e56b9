"""Verifier preprocess (preprocess/src/lib.rs:31-82): s0, s1 commitments of the permutation polynomials and O_pub_fix, the
function-instance part of the public statement (encode_o_pub_fix_common, group_structures/mod.rs:145-182)."""
from . import qap
from .fr import root_of_unity


def preprocess(backend, params, sigma, permutation, instance):
    m_i, s_max = params.l_D - params.l, params.s_max
    backend.init_ntt_domain(max(2 * params.n, 4 * m_i) * 2 * s_max)
    s0_ev, s1_ev = qap.permutation_evals(permutation, m_i, s_max, root_of_unity(m_i), root_of_unity(s_max))
    s0 = backend.commit(sigma.xy_powers, backend.from_rou_evals(s0_ev, m_i, s_max))
    s1 = backend.commit(sigma.xy_powers, backend.from_rou_evals(s1_ev, m_i, s_max))
    O_pub_fix = sigma.encode_O_pub_fix(backend, instance.a_pub_function, params)
    return {"s0": s0, "s1": s1, "O_pub_fix": O_pub_fix}
