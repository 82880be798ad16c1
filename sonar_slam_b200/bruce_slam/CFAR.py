"""`bruce_slam.CFAR.CFAR` -- detector front object of the reference
(bruce_slam/src/bruce_slam/CFAR.py:9-133), same constructor, attributes and methods.

The threshold factors (tau) are host-side scalar maths and stay in Python/scipy, as in
the reference (CFAR.py:71-121); detection runs on the GPU through `bruce_slam.cfar`.
False-alarm model: square-law detector in white Gaussian noise.

    CA    closed form            tau = N (Pfa^(-1/N) - 1)                       (CFAR.py:71-72)
    SOCA  root of  P_so(tau) - Pfa/2                                             (CFAR.py:98-110)
    GOCA  root of  (1 + tau/(N/2))^(-N/2) - P_so(tau) - Pfa/2                    (CFAR.py:112-115)
    OS    root of  N!/(N-k)! * Gamma(tau+N-k+1)/Gamma(tau+N+1) - Pfa             (CFAR.py:117-121)
  with P_so(t) = (2 + t/(N/2))^(-N/2) * sum_{j<N/2} C(N/2-1+j, j) (2 + t/(N/2))^(-j),
  each root searched with scipy.optimize.root from the CA value scaled by
  logspace(-2, 2, 10) in turn, first success wins (CFAR.py:74-96) -- the same starting
  points, so the same roots to the last bits (tests/golden/cfar_tau.json).
"""
import math

import numpy as np
from scipy.optimize import root

from . import cfar


class CFAR(object):
    """Constant False Alarm Rate detector: CA, SOCA, GOCA and OS variants."""

    def __init__(self, Ntc, Ngc, Pfa, rank=None):
        assert Ntc % 2 == 0
        assert Ngc % 2 == 0
        self.Ntc = Ntc  # training cells (both sides together)
        self.Ngc = Ngc  # guard cells (both sides together)
        self.Pfa = Pfa  # design false-alarm probability
        if rank is None:
            self.rank = self.Ntc / 2  # float, as in the reference (CFAR.py:24)
        else:
            self.rank = rank
            assert 0 <= self.rank < self.Ntc

        self.threshold_factor_CA = self.calc_WGN_threshold_factor_CA()
        self.threshold_factor_SOCA = self.calc_WGN_threshold_factor_SOCA()
        self.threshold_factor_GOCA = self.calc_WGN_threshold_factor_GOCA()
        self.threshold_factor_OS = self.calc_WGN_threshold_factor_OS()

        half_t, half_g = self.Ntc // 2, self.Ngc // 2
        self.params = {
            "CA": (half_t, half_g, self.threshold_factor_CA),
            "SOCA": (half_t, half_g, self.threshold_factor_SOCA),
            "GOCA": (half_t, half_g, self.threshold_factor_GOCA),
            "OS": (half_t, half_g, self.rank, self.threshold_factor_OS),
        }
        self.detector = {"CA": cfar.ca, "SOCA": cfar.soca, "GOCA": cfar.goca, "OS": cfar.os}
        self.detector2 = {"CA": cfar.ca2, "SOCA": cfar.soca2, "GOCA": cfar.goca2, "OS": cfar.os2}

    def __str__(self):
        rows = [
            "CFAR Detector Information\n",
            "=========================\n",
            "Number of training cells: {}\n".format(self.Ntc),
            "Number of guard cells: {}\n".format(self.Ngc),
            "Probability of false alarm: {}\n".format(self.Pfa),
            "Order statictics rank: {}\n".format(self.rank),
            "Threshold factors:\n",
            "      CA-CFAR: {:.3f}\n".format(self.threshold_factor_CA),
            "    SOCA-CFAR: {:.3f}\n".format(self.threshold_factor_SOCA),
            "    GOCA-CFAR: {:.3f}\n".format(self.threshold_factor_GOCA),
            "    OSCA-CFAR: {:.3f}\n".format(self.threshold_factor_OS),
        ]
        return "".join(rows)

    # ------------------------------------------------------------------ threshold factors
    def _solve(self, fun, name):
        start = self.calc_WGN_threshold_factor_CA()
        for scale in np.logspace(-2, 2, 10):
            sol = root(fun, start * scale)
            if sol.success:
                return sol.x[0]
        raise ValueError("Threshold factor of {} not found".format(name))

    def calc_WGN_threshold_factor_CA(self):
        return self.Ntc * (self.Pfa ** (-1.0 / self.Ntc) - 1)

    def calc_WGN_threshold_factor_SOCA(self):
        return self._solve(self.calc_WGN_pfa_SOCA, "SOCA")

    def calc_WGN_threshold_factor_GOCA(self):
        return self._solve(self.calc_WGN_pfa_GOCA, "GOCA")

    def calc_WGN_threshold_factor_OS(self):
        return self._solve(self.calc_WGN_pfa_OS, "OS")

    def calc_WGN_pfa_GOSOCA_core(self, x):
        x = float(np.ravel(x)[0])  # scipy hands the solver state over as a 1-element array
        half = self.Ntc / 2
        base = 2 + x / half
        acc = 0.0
        for j in range(int(half)):
            log_binom = math.lgamma(half + j) - math.lgamma(j + 1) - math.lgamma(half)
            acc += math.exp(log_binom) * base ** (-j)
        return acc * base ** (-half)

    def calc_WGN_pfa_SOCA(self, x):
        return self.calc_WGN_pfa_GOSOCA_core(x) - self.Pfa / 2

    def calc_WGN_pfa_GOCA(self, x):
        x = float(np.ravel(x)[0])
        half = self.Ntc / 2
        single = (1.0 + x / half) ** (-half)
        return single - self.calc_WGN_pfa_GOSOCA_core(x) - self.Pfa / 2

    def calc_WGN_pfa_OS(self, x):
        n, k = self.Ntc, self.rank
        x = float(np.ravel(x)[0])
        log_p = math.lgamma(n + 1) - math.lgamma(n - k + 1) + math.lgamma(x + n - k + 1) - math.lgamma(x + n + 1)
        return math.exp(log_p) - self.Pfa

    # ------------------------------------------------------------------ detection
    def detect(self, mat, alg="CA"):
        """Target mask (uint8 0/1, shape of `mat`)."""
        return self.detector[alg](mat, *self.params[alg])

    def detect2(self, mat, alg="CA"):
        """Target mask and the float32 threshold image."""
        return self.detector2[alg](mat, *self.params[alg])
