"""RCR (Windkessel) coupling around the hot path (BASELINE configs[0], 04-fluid/01-pipe3D_RCR): the host
mirror `svfsi_b200/cplbc.py` against the oracle's restatement of RCR_Integ_X / CALCDERCPLBC
(S/SETBC.f:1037-1123, 1292-1372), both against the closed-form solution of the Windkessel ODE, and the
semi-implicit sequence of S/MAIN.f:120-123, :280 / S/BAFINI.f:69-104."""
import numpy as np
import pytest

import common as cm
from svfsi_b200 import mesh

from oracle import oracle as ora
from svfsi_b200 import cplbc

FACES = [cplbc.RCR(121.0, 1.5e-4, 1212.0, Pd=0.0, Xo=0.0), cplbc.RCR(50.0, 4.0e-4, 700.0, Pd=30.0, Xo=200.0)]
DT = 5e-3


def _arrays(faces):
    return ([f.Rp for f in faces], [f.C for f in faces], [f.Rd for f in faces], [f.Pd for f in faces])


def test_rcr_integ_x_matches_the_oracle_and_the_closed_form():
    xo = np.array([10.0, 250.0]); Qo = np.array([3.0, -1.0]); Qn = np.array([3.5, -0.5])
    xn, y = cplbc.rcr_integ_x(xo, Qo, Qn, FACES, DT, time=3 * DT)
    xn_o, y_o = ora.rcr_integ_x(xo, Qo, Qn, *_arrays(FACES), DT, 3 * DT)
    assert np.abs(xn - xn_o).max() <= 1e-13 * np.abs(xn_o).max()
    assert np.abs(y - y_o).max() <= 1e-13 * np.abs(y_o).max()
    # C X' = Q(t) - (X - Pd)/Rd with Q linear in t: X = Pd + Rd (Q(t) - s tau) + K exp(-t/tau), tau = Rd C,
    # s = dQ/dt
    for k, f in enumerate(FACES):
        tau = f.Rd * f.C
        s = (Qn[k] - Qo[k]) / DT
        part0 = f.Pd + f.Rd * (Qo[k] - s * tau)
        K = xo[k] - part0
        exact = f.Pd + f.Rd * (Qn[k] - s * tau) + K * np.exp(-DT / tau)
        assert abs(xn[k] - exact) <= 1e-11 * abs(exact)
        assert abs(y[k] - (exact + Qn[k] * f.Rp)) <= 1e-11 * abs(y[k])


def test_resistance_is_the_derivative_of_the_face_pressure():
    flux = {"o": [2.0, 1.0], "n": [2.2, 0.9]}
    cpl = cplbc.CplBC(FACES, DT, "SI")
    cpl.init(lambda i, w: flux[w][i], time=DT)
    y_o, r_o = ora.calc_der_cplbc([f.Xo for f in FACES], flux["o"], flux["n"], *_arrays(FACES), DT, DT)
    assert np.allclose(cpl.r, r_o, rtol=1e-9, atol=0) and np.allclose(cpl.y, y_o, rtol=1e-13, atol=0)
    # dy/dQn in closed form: y = X(dt) + Qn Rp and dX/dQn = Rd (1 - tau/dt (1 - exp(-dt/tau)))
    for k, f in enumerate(FACES):
        tau = f.Rd * f.C
        exact = f.Rp + f.Rd * (1.0 - tau / DT * (1.0 - np.exp(-DT / tau)))
        assert abs(cpl.r[k] - exact) <= 1e-6 * exact


@pytest.mark.parametrize("scheme", ["SI", "I", "E"])
def test_coupling_sequence_of_the_time_loop(scheme):
    """S/MAIN.f: SETBCCPL at every Newton iteration integrates from xo (the state at the START of the
    step) with the current Qn; xo moves only at the end of the step"""
    cpl = cplbc.CplBC(FACES[:1], DT, scheme)
    Q = {"o": [0.0], "n": [0.0]}
    cpl.init(lambda i, w: Q[w][i])
    r0 = cpl.r.copy()
    assert (scheme == "E") == (r0[0] == 0.0)
    t = 0.0
    hist = []
    for step in range(3):
        t += DT
        for it in range(2):
            Q["n"] = [1.0 + step + 0.1 * it]
            g = cpl.setbccpl(lambda i, w: Q[w][i], t)
        hist.append((cpl.xo.copy(), cpl.xn.copy(), g.copy()))
        cpl.advance()
        assert np.array_equal(cpl.xo, hist[-1][1])
        Q["o"] = list(Q["n"])
    # the Newton iterations of one step restart from the same xo
    assert hist[1][0][0] == hist[0][1][0]
    # semi-implicit keeps the initial resistance, implicit recomputes it (same value here: the ODE is linear)
    assert np.allclose(cpl.r, r0, rtol=1e-6) or scheme == "E"
    # pressure rises with the flow that passes
    assert hist[2][2][0] > hist[1][2][0] > hist[0][2][0] > 0.0


def test_unknown_scheme_is_refused():
    with pytest.raises(ValueError):
        cplbc.CplBC(FACES, DT, "X")


def test_oracle_pipe_with_an_rcr_outlet():
    """the whole sequence on the small pipe with the oracle (the CPU reference of the device-resident RCR
    loop): Newton converges within each step, the resistance term is positive, the Windkessel pressure
    follows the outflow"""
    m, probs, _ = mesh.build_problem(8, 8, 20, nparts=1, L=4.0)
    out, (An, Yn), cpl = cm.oracle_rcr_time_loop(m, probs[0], [cplbc.RCR(121.0, 1.5e-4, 1212.0)])
    assert len(out) == 6
    for s in range(2):
        n0, n1, n2 = (out[3 * s + k][0] for k in range(3))
        assert n2 < n1 < n0 and n2 < 1e-2 * n0
    f = cplbc.RCR(121.0, 1.5e-4, 1212.0)
    tau = f.Rd * f.C
    assert abs(cpl.r[0] - (f.Rp + f.Rd * (1.0 - tau / cm.DT * (1.0 - np.exp(-cm.DT / tau))))) <= 1e-5 * cpl.r[0]
    assert all(o[3] > 0.0 for o in out)                  # outflow through the outlet
    assert out[-1][2] > out[2][2] > 0.0                    # g = X + Qn Rp grows while the capacitor charges
    assert np.isfinite(Yn).all() and np.isfinite(An).all()
