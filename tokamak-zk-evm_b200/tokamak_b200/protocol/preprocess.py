"""Verifier preprocess (preprocess/src/lib.rs:31-82): s0, s1 commitments of the permutation polynomials and O_pub_fix, the
function-instance part of the public statement (encode_o_pub_fix_common, group_structures/mod.rs:145-182)."""
import numpy as np

from .. import frs_from_ints
from . import qap
from .fr import root_of_unity


def preprocess(backend, params, sigma, permutation, instance):
    m_i, s_max = params.l_D - params.l, params.s_max
    backend.init_ntt_domain(max(2 * params.n, 4 * m_i) * 2 * s_max)
    s0_ev, s1_ev = qap.permutation_evals(permutation, m_i, s_max, root_of_unity(m_i), root_of_unity(s_max))
    s0 = backend.commit(sigma.xy_powers, backend.from_rou_evals(s0_ev, m_i, s_max))
    s1 = backend.commit(sigma.xy_powers, backend.from_rou_evals(s1_ev, m_i, s_max))
    m_function = params.l - params.l_free
    if m_function == 0:
        O_pub_fix = None
    else:
        if len(instance.a_pub_function) != m_function:
            raise ValueError(f"a_pub_function length mismatch: expected m_function={m_function}, got {len(instance.a_pub_function)}")
        start = params.l - m_function
        O_pub_fix = backend.msm_indexed(sigma.gamma_inv_o_inst, np.arange(start, start + m_function, dtype=np.uint32), frs_from_ints(instance.a_pub_function))
    return {"s0": s0, "s1": s1, "O_pub_fix": O_pub_fix}
