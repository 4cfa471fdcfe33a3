"""Single-site expectation values by FULL contraction of the block (src/algo/measurements.py:419-449, 499-519, 547-604, 607-655):

    <O>_site = <psi| O_site |psi> / <psi|psi>

Both are complete boundary-MPS contractions of the block with its six messages (`ContractionDepth.Full`, the scalar
`(mantissa, exp10)` branch of bubblecon, src/libs/bubblecon.py:3077-3088); in the numerator the site's ket is replaced by
the double-layer tensor with the operator sandwiched in.  One denominator, then three contractions (A, B, C of the centre
triangle) per operator.  The contractor is passed in: the product uses the device `bubblecon`, the oracle its numpy one.
"""
from __future__ import annotations

import numpy as np

from . import block_tn, contraction_order
from .lattice import SIDE_ANGLE, get_block


def sandwich(ket, mat):
    """double-layer tensor of one site with `mat` between ket and bra: [D^2, D^2, D^2, D^2] (measurements.py:419-433)."""
    ket = np.asarray(ket)
    D = ket.shape[1]
    ket_op = np.tensordot(ket, np.asarray(mat), axes=([0], [1]))            # [a, b, c, e, p']
    kob = np.tensordot(ket_op, np.conj(ket), axes=([4], [0]))                # [a, b, c, e, a*, b*, c*, e*]
    return np.transpose(kob, (0, 4, 1, 5, 2, 6, 3, 7)).reshape(D * D, D * D, D * D, D * D)


def _ratio(num, den, force_real):
    """(measurements.py:452-497) numerator / denominator in the common mantissa * 10^exp format"""
    if num[0] == 0 and den[0] == 0:
        raise FloatingPointError("Both numerator and denominator are zero.")
    m = num[0] / den[0]
    if force_real:
        m = float(np.real(m))
    return m * 10.0 ** (num[1] - den[1])


def calc_unit_cell_expectation_values_from_tn(bubblecon_fn, cell, messages: dict, N: int, operators, chi: int, direction: str = "U",
                                              force_real: bool = False) -> list:
    """-> [{'A': <O>, 'B': <O>, 'C': <O>} for O in operators]  (measurements.py:547-604).
    `cell`: (A, B, C) or one tensor per lattice site; `messages`: {side: [site arrays]}; `bubblecon_fn`: bubblecon(T, E, A,
    bubble_angle, order, D_trunc=, ket_tensors=, separate_exp=True) -> (mantissa, exp10).  The reference draws `direction` at
    random when it is not given; here it is an argument."""
    T, E, A, K, P = block_tn.assemble(N, cell, messages)
    T, E, A = block_tn.connect_corner(N, T, E, A, P, direction)
    order = list(contraction_order.kagome_order(N, direction, "Full"))
    den = bubblecon_fn(T, E, A, SIDE_ANGLE[direction], order, D_trunc=chi, ket_tensors=K, separate_exp=True)
    t = get_block(N).center_triangle
    out = []
    for op in operators:
        res = {}
        for c, key in enumerate("ABC"):
            idx = 3 * t + c
            T2, K2 = list(T), list(K)
            T2[idx] = sandwich(T[idx], op)
            K2[idx] = False
            num = bubblecon_fn(T2, E, A, SIDE_ANGLE[direction], order, D_trunc=chi, ket_tensors=K2, separate_exp=True)
            res[key] = _ratio(num, den, force_real)
        out.append(res)
    return out


def device_expectation_values(unit_cell, messages: dict, N: int, operators, chi: int, direction: str = "U", force_real: bool = False) -> list:
    """the same with the device contractor; `messages`: {side: Message} as returned by belief_propagation."""
    from .bubblecon import bubblecon
    msgs = {s: m.mps.A for s, m in messages.items()}
    return calc_unit_cell_expectation_values_from_tn(bubblecon, unit_cell.tensors(), msgs, N, operators, chi, direction, force_real)
