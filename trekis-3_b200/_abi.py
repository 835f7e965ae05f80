"""ctypes mirror of include/trekis3_gpu.h and include/trekis3_host.h (plumbing only).

The structs must match the C headers field by field; tests/test_abi.py checks the sizes
against the values the C side reports.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)

MAX_ATOMS, MAX_SHELLS, NR, NTHETA, MAX_NT = 8, 32, 50, 180, 256

P_D = C.POINTER(C.c_double)
P_I64 = C.POINTER(C.c_int64)


class Config(C.Structure):
    _fields_ = [
        ("shi_E", C.c_double), ("shi_mass", C.c_double), ("shi_fixed_Zeff", C.c_double),
        ("shi_Z", C.c_int32), ("shi_kind_Zeff", C.c_int32),
        ("Tim", C.c_double), ("dt", C.c_double), ("dt_flag", C.c_int32), ("include_photons", C.c_int32),
        ("cut_off", C.c_double), ("layer", C.c_double), ("hole_mass", C.c_double),
        ("work_function", C.c_double), ("bar_length", C.c_double), ("bar_height", C.c_double),
        ("kind_of_EMFP", C.c_int32), ("reserved0", C.c_int32),
        ("seed", C.c_uint64),
    ]


class Tables(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int32), ("n_shells", C.c_int32), ("vb_shell", C.c_int32), ("nshl_atom1", C.c_int32),
        ("atom_Z", C.c_int32 * MAX_ATOMS), ("atom_nshl", C.c_int32 * MAX_ATOMS), ("atom_first", C.c_int32 * MAX_ATOMS),
        ("atom_mass", C.c_double * MAX_ATOMS), ("atom_pers", C.c_double * MAX_ATOMS),
        ("shell_atom", C.c_int32 * MAX_SHELLS), ("shell_num", C.c_int32 * MAX_SHELLS),
        ("shell_Ip", C.c_double * MAX_SHELLS), ("shell_Nel", C.c_double * MAX_SHELLS),
        ("shell_auger", C.c_double * MAX_SHELLS), ("shell_radiat", C.c_double * MAX_SHELLS),
        ("n_ei", C.c_int32), ("ei_E", P_D), ("ei_L", P_D),
        ("n_ee", C.c_int32), ("ee_E", P_D), ("ee_L", P_D),
        ("n_hi", C.c_int32), ("hi_E", P_D), ("hi_L", P_D),
        ("n_he", C.c_int32), ("he_E", P_D), ("he_L", P_D),
        ("n_ph", C.c_int32), ("ph_E", P_D), ("ph_L", P_D),
        ("n_shi", C.c_int32), ("shi_E", P_D), ("shi_L", P_D), ("shi_dEdx", P_D),
        ("dshi_off", P_I64), ("dshi_E", P_D), ("dshi_L", P_D),
        ("eid_off", P_I64), ("eid_hw", P_D), ("eid_L", P_D),
        ("eed_off", P_I64), ("eed_hw", P_D), ("eed_L", P_D),
        ("hid_off", P_I64), ("hid_hw", P_D), ("hid_L", P_D),
        ("hed_off", P_I64), ("hed_hw", P_D), ("hed_L", P_D),
        ("n_dos", C.c_int32), ("dos_E", P_D), ("dos_DOS", P_D), ("dos_int", P_D), ("dos_effm", P_D),
        ("n_r", C.c_int32), ("out_R", P_D), ("out_V", P_D),
        ("shell_kocs", C.c_int32 * MAX_SHELLS), ("shell_Ek", C.c_double * MAX_SHELLS), ("at_dens", C.c_double),
        ("delta_cdf", C.c_int32), ("osc_off", C.c_int32 * (MAX_SHELLS + 1)), ("osc_E0", P_D), ("osc_alpha", P_D),
        ("n_dsf_e", C.c_int32), ("dsf_e_dE", P_D), ("dsf_e_emit", P_D), ("dsf_e_absorb", P_D), ("ee_emit", P_D), ("ee_absorb", P_D),
        ("n_dsf_h", C.c_int32), ("dsf_h_dE", P_D), ("dsf_h_emit", P_D), ("dsf_h_absorb", P_D), ("he_emit", P_D), ("he_absorb", P_D),
    ]


TALLY_NAMES = [
    "Out_ne", "Out_Ee", "Out_nphot", "Out_Ephot", "Out_Ee_vs_E", "Out_Eh_vs_E", "Out_Elat", "Out_nh", "Out_Eh",
    "Out_Ehkin", "Out_tot_Ne", "Out_tot_Nphot", "Out_tot_E", "Out_E_e", "Out_E_phot", "Out_E_at", "Out_E_h",
    "Out_Eat_dens", "Out_theta", "Out_theta_h", "Out_Ne_Em", "Out_E_Em", "Out_Ee_vs_E_Em", "Out_field_all",
    "Out_E_field", "Out_diff_coeff",
]
N_TALLIES = len(TALLY_NAMES)


class TallyLayout(C.Structure):
    _fields_ = [
        ("Nt", C.c_int32), ("n_r", C.c_int32), ("n_atoms", C.c_int32), ("nshl1", C.c_int32), ("n_dos", C.c_int32),
        ("reserved", C.c_int32),
        ("off", C.c_int64 * N_TALLIES), ("len", C.c_int64 * N_TALLIES), ("total", C.c_int64),
        ("time_grid", C.c_double * (MAX_NT + 1)),
    ]

    def shape(self, name):
        Nt, NRr, NA, NS, ND = self.Nt, self.n_r, self.n_atoms, self.nshl1, self.n_dos
        return {
            "Out_ne": (Nt, NRr), "Out_Ee": (Nt, NRr), "Out_nphot": (Nt, NRr), "Out_Ephot": (Nt, NRr),
            "Out_Ee_vs_E": (Nt, NRr), "Out_Eh_vs_E": (Nt, ND), "Out_Elat": (Nt, NRr),
            "Out_nh": (Nt, NRr, NA, NS), "Out_Eh": (Nt, NRr, NA, NS), "Out_Ehkin": (Nt, NRr, NA, NS),
            "Out_tot_Ne": (Nt,), "Out_tot_Nphot": (Nt,), "Out_tot_E": (Nt,), "Out_E_e": (Nt,), "Out_E_phot": (Nt,),
            "Out_E_at": (Nt,), "Out_E_h": (Nt, NA, NS), "Out_Eat_dens": (Nt, NRr),
            "Out_theta": (Nt + 1, NTHETA), "Out_theta_h": (Nt + 1, NTHETA), "Out_Ne_Em": (Nt,), "Out_E_Em": (Nt,),
            "Out_Ee_vs_E_Em": (Nt, NRr), "Out_field_all": (Nt, NRr), "Out_E_field": (Nt,), "Out_diff_coeff": (Nt,),
        }[name]


EVENT_NAMES = ["shi", "el_inelastic", "el_elastic", "vbh_inelastic", "vbh_elastic", "auger", "radiative",
               "auger_frozen", "photon"]
ERROR_NAMES = ["err10", "err20", "err21", "err22", "err23", "err25", "err30", "err40", "err41", "err50", "err51",
               "err52", "queue_overflow", "nan", "auger_balance"]
# algorithmic bytes per event: compulsory particle-state traffic, tables excluded
EVENT_BYTES = {"shi": 176, "el_inelastic": 320, "el_elastic": 144, "vbh_inelastic": 384, "vbh_elastic": 208,
               "auger": 384, "radiative": 280, "auger_frozen": 208, "photon": 248}


class Stats(C.Structure):
    _fields_ = [
        ("events", C.c_uint64 * len(EVENT_NAMES)), ("errors", C.c_uint64 * len(ERROR_NAMES)),
        ("n_electrons", C.c_uint64), ("n_photons", C.c_uint64), ("n_waves", C.c_uint64),
        ("kernel_launches", C.c_uint64), ("device_ms", C.c_double), ("algorithmic_bytes", C.c_double),
        ("max_energy_drift", C.c_double), ("cold_events", C.c_uint64 * 2), ("warm_events", C.c_uint64 * 2),
    ]

    def as_dict(self):
        d = {"events": {n: int(self.events[i]) for i, n in enumerate(EVENT_NAMES)},
             "errors": {n: int(self.errors[i]) for i, n in enumerate(ERROR_NAMES) if self.errors[i]},
             "n_electrons": int(self.n_electrons), "n_photons": int(self.n_photons), "n_waves": int(self.n_waves),
             "kernel_launches": int(self.kernel_launches), "device_ms": float(self.device_ms),
             "algorithmic_bytes": float(self.algorithmic_bytes), "max_energy_drift": float(self.max_energy_drift),
             "cold_events": {"electron": int(self.cold_events[0]), "vbhole": int(self.cold_events[1])},
             "warm_events": {"electron": int(self.warm_events[0]), "vbhole": int(self.warm_events[1])}}
        d["total_events"] = sum(d["events"].values())
        return d


def lib_path(name):
    return {"host": os.path.join(HERE, "libtrekis3_host.so"),
            "gpu": os.path.join(HERE, "libtrekis3_gpu.so"),
            "oracle": os.path.join(REPO, "oracle", "_build", "libtrk3_oracle.so")}[name]
