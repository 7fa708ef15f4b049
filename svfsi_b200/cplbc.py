"""Host-side mirror of svFSI's 0-D coupled boundary conditions for RCR (Windkessel) outlets -- the
`cplBC` object and SETBCCPL / CALCDERCPLBC / RCRINIT / RCR_Integ_X (S/SETBC.f:981-1123, 1263-1372), i.e.
what BASELINE configs[0] (04-fluid/01-pipe3D_RCR) adds around the hot path.  It stays on the HOST in the
reference too (rank 0 integrates nFa scalar ODEs and broadcasts, S/SETBC.f:1201-1255); the device side
only supplies the face fluxes (`gpu_face_integ_v_`) and consumes the result: the face pressure `g` goes
into the Neumann face assembly (h = g, S/SETBC.f:267-270), the resistance `r` into FSILS_SOLVE's
rank-one term (`res = gam*dt*r`, S/MAIN.f:186-192 -> ADDBCMUL).

    cpl = CplBC([RCR(Rp, C, Rd, Pd, Xo), ...], dt, scheme="SI")
    cpl.init(integ)                 # BAFINI.f:69-104: RCRINIT, then CALCDERCPLBC unless explicit
    per Newton iteration:   g = cpl.setbccpl(integ, time)        # S/MAIN.f:120-123
                            res_i = gam*dt*cpl.r[i]
    per time step:          cpl.advance()                        # cplBC%xo = cplBC%xn, S/MAIN.f:280

`integ(i, which)` returns the flux through coupled face i of Yo (which = "o") or Yn (which = "n"),
S/SETBC.f:1001-1002."""
from dataclasses import dataclass

import numpy as np

NTS = 100                      # sub-steps of RCR_Integ_X, S/SETBC.f:1298
ABS_TOL, REL_TOL = 1e-8, 1e-5  # CALCDERCPLBC, S/SETBC.f:1042-1043


@dataclass
class RCR:
    """lBc%RCR, S/READFILES.f:1713-1718: "RCR values" (Rp, C, Rd), "Distal pressure", "Initial pressure" """
    Rp: float
    C: float
    Rd: float
    Pd: float = 0.0
    Xo: float = 0.0


def rcr_integ_x(xo, Qo, Qn, faces, dt, time):
    """RCR_Integ_X (S/SETBC.f:1292-1372): nTS steps of Kutta's 3/8 rule on C dX/dt = Q - (X - Pd)/Rd with
    Q interpolated linearly from Qo to Qn over the time step; returns (xn, y) with y = X + Qn*Rp."""
    Rp = np.array([f.Rp for f in faces]); C = np.array([f.C for f in faces])
    Rd = np.array([f.Rd for f in faces]); Pd = np.array([f.Pd for f in faces])
    Qo = np.asarray(Qo, dtype=np.float64); Qn = np.asarray(Qn, dtype=np.float64)
    X = np.array(xo, dtype=np.float64)
    tt = max(time - dt, 0.0)
    dtt = dt / float(NTS)
    for n in range(1, NTS + 1):
        Qrk = []
        for i in range(1, 5):
            r = float(i - 1) / 3.0
            r = (float(n - 1) + r) / float(NTS)
            Qrk.append(Qo + (Qn - Qo) * r)
        f1 = (Qrk[0] - (X - Pd) / Rd) / C
        Xrk = X + dtt * f1 / 3.0
        f2 = (Qrk[1] - (Xrk - Pd) / Rd) / C
        Xrk = X - dtt * f1 / 3.0 + dtt * f2
        f3 = (Qrk[2] - (Xrk - Pd) / Rd) / C
        Xrk = X + dtt * f1 - dtt * f2 + dtt * f3
        f4 = (Qrk[3] - (Xrk - Pd) / Rd) / C
        X = X + (dtt / 8.0) * (f1 + 3.0 * (f2 + f3) + f4)
        tt = tt + dtt
        if np.isnan(X).any():
            raise FloatingPointError("RCR integration error detected")     # istat = -1 -> STOPSIM
    return X, X + Qn * Rp


class CplBC:
    """cplBC with RCR faces (all of the Neumann group, cplBC_Neu).  scheme: "SI" (svFSI's choice for RCR,
    S/READFILES.f:937: the resistance is computed once at initialisation), "I" (recomputed at every
    Newton iteration), "E" (no resistance term)."""

    def __init__(self, faces, dt, scheme="SI"):
        if scheme not in ("SI", "I", "E"):
            raise ValueError(f"Undefined cplBC%schm: {scheme}")
        self.fa = list(faces)
        self.dt = float(dt)
        self.schm = scheme
        n = len(self.fa)
        self.xo = np.zeros(n); self.xn = np.zeros(n)
        self.y = np.zeros(n)             # cplBC%fa%y -> bc%g
        self.r = np.zeros(n)             # bc%r
        self.Qo = np.zeros(n); self.Qn = np.zeros(n)

    # ------------------------------------------------------------------ BAFINI.f:69-104
    def init(self, integ, time=0.0):
        self.xo = np.array([f.Xo for f in self.fa], dtype=np.float64)     # RCRINIT, initRCR = F
        self.y[:] = 0.0
        if self.schm != "E":
            self.calcder(integ, time)

    def _fluxes(self, integ):
        for i in range(len(self.fa)):
            self.Qo[i] = integ(i, "o")
            self.Qn[i] = integ(i, "n")

    # ------------------------------------------------------------------ S/SETBC.f:1037-1123
    def calcder(self, integ, time):
        self._fluxes(integ)
        self.xn, self.y = rcr_integ_x(self.xo, self.Qo, self.Qn, self.fa, self.dt, time)
        diff = float(np.sqrt((self.Qo * self.Qo).sum() / float(len(self.fa))))
        diff = ABS_TOL if diff * REL_TOL < ABS_TOL else diff * REL_TOL
        orgY, orgQ = self.y.copy(), self.Qn.copy()
        for i in range(len(self.fa)):
            Qn = orgQ.copy(); Qn[i] = orgQ[i] + diff
            self.xn, y = rcr_integ_x(self.xo, self.Qo, Qn, self.fa, self.dt, time)
            self.r[i] = (y[i] - orgY[i]) / diff
        self.y, self.Qn = orgY, orgQ

    # ------------------------------------------------------------------ S/SETBC.f:981-1034
    def setbccpl(self, integ, time):
        if self.schm == "I":
            self.calcder(integ, time)
        else:
            self._fluxes(integ)
            self.xn, self.y = rcr_integ_x(self.xo, self.Qo, self.Qn, self.fa, self.dt, time)
        return self.y.copy()

    def advance(self):
        self.xo = self.xn.copy()        # S/MAIN.f:280
