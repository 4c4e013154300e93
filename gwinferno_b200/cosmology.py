"""Host-side flat-LambdaCDM tables for the static dV_c/dz weight.

Mirrors the behaviour of the reference's ``Cosmology`` (gwinferno/cosmology.py:27-120): a
trapezoid table of the comoving distance on ``z = arange(0, 10, 1e-3)`` and
``dVc/dz = 4 pi Dc(z)^2 (c/H0)/E(z)`` with ``Dc`` linearly interpolated from the table.
Only the static pieces the population-likelihood path needs are provided (``dVcdz``,
``z2Dc``); the per-sample column itself is computed inside libgwi (csrc/plan.cpp) with the same
recipe -- this module is used by the Python host for the normalisation grids and by the
synthetic-catalog generator.
"""

import numpy as np

C_SI = 299792458.0  # m/s
# gwinferno/cosmology.py:19-22 (the LVK Planck-2015 variant is the one the models use,
# models/bsplines/single.py:8, models/parametric/parametric.py:4)
PLANCK_2015_LVK_Ho = 67.90 / 1e-3
PLANCK_2015_LVK_OmegaMatter = 0.3065
PLANCK_2015_LVK_OmegaLambda = 1.0 - PLANCK_2015_LVK_OmegaMatter
DEFAULT_DZ = 1e-3


class FlatLambdaCDM:
    def __init__(self, Ho, omega_matter, omega_lambda, max_z=10.0, dz=DEFAULT_DZ):
        self.Ho = Ho
        self.c_over_Ho = C_SI / Ho
        self.OmegaMatter = omega_matter
        self.OmegaLambda = omega_lambda
        self._extend(max_z, dz)

    def z2E(self, z):
        opz = 1.0 + np.asarray(z, dtype=np.float64)
        return np.sqrt(self.OmegaLambda + self.OmegaMatter * opz**3)

    def dDcdz(self, z):
        return self.c_over_Ho / self.z2E(z)

    def _extend(self, max_z, dz):
        # sequential trapezoid, Dc[i+1] = Dc[i] + 0.5 (f(z_i) + f(z_i + dz)) dz
        # (gwinferno/cosmology.py:48-77); cumsum adds in the same order as the reference loop.
        self.z = np.arange(0.0, max_z, dz)
        zl = self.z[:-1]
        inc = 0.5 * (self.dDcdz(zl) + self.dDcdz(zl + dz)) * dz
        self.Dc = np.concatenate([[0.0], np.cumsum(inc)])

    def z2Dc(self, z):
        z = np.asarray(z, dtype=np.float64)
        if z.size and np.max(z) > self.z[-1]:
            raise ValueError("redshift beyond the tabulated range (z < 10)")
        return np.interp(z, self.z, self.Dc)

    def dVcdz(self, z):
        z = np.asarray(z, dtype=np.float64)
        Dc = self.z2Dc(z)
        return 4.0 * np.pi * Dc**2 * self.dDcdz(z)


Planck15 = FlatLambdaCDM(PLANCK_2015_LVK_Ho, PLANCK_2015_LVK_OmegaMatter, PLANCK_2015_LVK_OmegaLambda)
