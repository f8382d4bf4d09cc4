"""Physical and ion-specific constants with the attribute names of the reference's ``constants.Constants``
(constants.py:7-97 of arparaschiv/solar-coronal-inversion), table-driven.

``blos_proc`` receives this *module* and instantiates ``Constants(line)`` per line (CLEDB_PROC.py:451).
"""

# ion key -> (log10 ion temperature [K], atomic mass [amu], reference wavelength [nm], F factor, g_u, g_l, J_u, J_l)
_LINES = {
    'fe-xiii_1074': (6.22, 55.847, 1074.6153, 0.0, 1.5, 1, 1, 0),
    'fe-xiii_1079': (6.22, 55.847, 1079.7803, 0.0, 1.5, 1.5, 2, 1),
    'si-x_1430': (6.15, 28.0855, 1430.2231, 0.5, 1.334, 0.665, 1.5, 0.5),
    'si-ix_3934': (6.05, 28.0855, 3926.6551, 0.0, 1.5, 1, 1, 0),
}
_AMU_KG = 1.672621E-27


class Constants:
    """General constants plus the constants of one spectral line (``ion`` is one of the keys above)."""

    def __init__(self, ion):
        self.l_speed = 2.9979e+8                        # speed of light [m/s]
        self.kb = 1.3806488e-23                         # Boltzmann constant [J/K]
        self.e_mass = 9.10938356e-31                    # electron mass [kg]
        self.e_charge = 1.602176634e-19                 # electron charge [C]
        self.bohrmagneton = 9.274009994e-24 * 1.e-4     # Bohr magneton, per Gauss
        self.planckconst = 6.62607004e-34               # Planck constant [J s]
        if ion not in _LINES:
            print("Not supported ion or wrong string. Ion not Fe fe-xiii_1074, fe-xiii_1079, si-x_1430 or "
                  "si-ix_3934.\nIon specific constants not returned!")
            return
        logt, amu, ref, ffac, gu, gl, ju, jl = _LINES[ion]
        self.ion_temp = logt
        self.ion_mass = amu * _AMU_KG
        self.line_ref = ref
        self.width_th = self.line_ref / self.l_speed * (4. * 0.69314718 * self.kb * (10. ** self.ion_temp) / self.ion_mass) ** 0.5
        self.F_factor = ffac
        self.gu, self.gl, self.ju, self.jl = gu, gl, ju, jl
        # LS-coupling effective Lande factor
        self.g_eff = 0.5 * (self.gu + self.gl) + 0.25 * (self.gu - self.gl) * (self.ju * (self.ju + 1) - self.jl * (self.jl + 1))
