"""Host half of the drop-in (libtrekis3_host.so): input readers, table builder, output writer."""
import ctypes as C
import hashlib
import os
import shutil

import numpy as np

from . import _abi
from ._abi import Config, Tables, TallyLayout

_lib = None


def _host():
    global _lib
    if _lib is None:
        path = _abi.lib_path("host")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` or "
                               "`make -C trekis-3_b200/csrc host`")
        lib = C.CDLL(path)
        lib.trk3h_load.restype = C.c_void_p
        lib.trk3h_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        lib.trk3h_free.argtypes = [C.c_void_p]
        lib.trk3h_build_tables.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
        lib.trk3h_set_dcs_evaluator.argtypes = [C.c_void_p]
        lib.trk3h_set_dcs_evaluator.restype = None
        lib.trk3h_save_tables.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
        lib.trk3h_load_tables.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
        lib.trk3h_write_reference_cache.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.c_char_p, C.c_int]
        lib.trk3h_read_reference_cache.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        lib.trk3h_reference_cache_name.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
        lib.trk3h_config.restype = C.POINTER(Config)
        lib.trk3h_config.argtypes = [C.c_void_p]
        lib.trk3h_tables.restype = C.POINTER(Tables)
        lib.trk3h_tables.argtypes = [C.c_void_p]
        lib.trk3h_set.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        lib.trk3h_get.restype = C.c_double
        lib.trk3h_get.argtypes = [C.c_void_p, C.c_char_p]
        lib.trk3h_get_string.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
        lib.trk3h_num_warnings.argtypes = [C.c_void_p]
        lib.trk3h_warning.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        PD = C.POINTER(C.c_double)
        lib.trk3h_eval_TotIMFP.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, PD, PD]
        lib.trk3h_eval_EMFP.argtypes = [C.c_void_p, C.c_double, C.c_int, PD, PD]
        lib.trk3h_eval_dcs_phonon.argtypes = [C.c_void_p, C.c_double, C.c_double, PD, PD]
        lib.trk3h_eval_SHI.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, PD, PD, PD]
        lib.trk3h_eval_photon.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, PD]
        lib.trk3h_sumrules.argtypes = [C.c_void_p, C.c_int, C.c_int, PD, PD]
        lib.trk3h_grid.argtypes = [C.c_void_p, C.c_double, C.c_double, PD, C.c_int]
        lib.trk3h_save_output.argtypes = [C.c_void_p, C.POINTER(TallyLayout), PD, C.c_int, C.c_char_p, C.c_char_p,
                                          C.c_int, C.c_char_p, C.c_int]
        lib.trk3_tally_layout_init.argtypes = [C.POINTER(Config), C.POINTER(Tables), C.POINTER(TallyLayout)]
        lib.trk3_host_version.restype = C.c_char_p
        _lib = lib
    return _lib


# The five BASELINE.json configurations as edits of the shipped default INPUT_PARAMETERS.txt
# (SURVEY.md 8d): (material, SHI Z, E [MeV], photons, NMC)
CONFIGS = {
    "C1": ("Al2O3", 54, 167.0, 0, 100),
    "C2": ("SiO2_cryst", 79, 2187.0, 1, 1000),
    "C3": ("Diamond", 54, 167.0, 0, 100),
    "C4": ("Au", 92, 2600.0, 0, 1000),
    "C5": ("Al2O3", 54, 167.0, 0, 100000),
}


def make_run_dir(path, config="C1", data_dir=None, nmc=None, extra_lines=("gnuplot no", "grid 1"), edits=None):
    """Create a TREKIS run directory: INPUT_PARAMETERS.txt (shipped default with the config's edits,
    SURVEY.md 8d) plus INPUT_CDF / INPUT_DOS / INPUT_EADL links to the data directory."""
    data_dir = data_dir or os.path.join(_abi.REPO, "data")
    os.makedirs(path, exist_ok=True)
    material, z, e_mev, photons, n = CONFIGS[config] if isinstance(config, str) else config
    if nmc is not None:
        n = nmc
    with open(os.path.join(data_dir, "INPUT_PARAMETERS.default.txt")) as f:
        lines = f.read().splitlines()
    lines[0] = f"{material}         ! material name"
    lines[1] = f"{z}         ! SHI atomic number"
    lines[2] = f"{e_mev}        ! [MeV] total SHI energy"
    lines[15] = f"{photons}           ! include radiative decay of deep holes? (0=no, 1=yes)"
    lines[17] = f"{n}           ! number of MC iterations to be performed"
    for k, v in (edits or {}).items():      # 1-based line number -> full line text
        lines[k - 1] = v
    lines = lines[:19] + list(extra_lines)
    with open(os.path.join(path, "INPUT_PARAMETERS.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    for d in ("INPUT_CDF", "INPUT_DOS", "INPUT_EADL", "INPUT_DSF"):
        dst = os.path.join(path, d)
        if os.path.islink(dst) or os.path.exists(dst):
            continue
        if d == "INPUT_DSF" and not os.path.isdir(os.path.join(data_dir, d)):
            continue                        # DSF cross sections (kind_of_EMFP = 2) are optional inputs; the reference ships none
        os.symlink(os.path.join(data_dir, d), dst)
    return path


class Case:
    """Parsed inputs + host-built tables of one TREKIS-3 run (everything do_Monte_Carlo receives)."""

    def __init__(self, handle, directory):
        self._h = handle
        self.dir = directory

    @classmethod
    def load(cls, directory):
        lib = _host()
        err = C.create_string_buffer(1024)
        h = lib.trk3h_load(os.fsencode(directory), err, 1024)
        if not h:
            raise RuntimeError("TREKIS input error: " + err.value.decode(errors="replace"))
        return cls(h, directory)

    def __del__(self):
        try:
            if self._h:
                _host().trk3h_free(self._h)
                self._h = None
        except Exception:
            pass

    # -- scalars -----------------------------------------------------------------------------
    def get(self, key):
        return _host().trk3h_get(self._h, key.encode())

    def set(self, key, value):
        rc = _host().trk3h_set(self._h, key.encode(), float(value))
        if rc != 0:
            raise KeyError(key)

    def get_string(self, key):
        buf = C.create_string_buffer(512)
        if _host().trk3h_get_string(self._h, key.encode(), buf, 512) != 0:
            raise KeyError(key)
        return buf.value.decode()

    @property
    def warnings(self):
        lib = _host()
        out = []
        for i in range(lib.trk3h_num_warnings(self._h)):
            buf = C.create_string_buffer(1024)
            lib.trk3h_warning(self._h, i, buf, 1024)
            out.append(buf.value.decode())
        return out

    # -- tables ------------------------------------------------------------------------------
    def _input_digest(self, shi_window_only):
        hsh = hashlib.sha256()
        for rel in ("INPUT_PARAMETERS.txt",):
            with open(os.path.join(self.dir, rel), "rb") as f:
                # only the lines that influence the tables (not NMC/threads/time grid)
                lines = f.read().splitlines()
                keep = [lines[i] for i in (0, 1, 2, 3, 8, 9, 10, 11, 12, 13, 14, 15) if i < len(lines)] + lines[19:]
                hsh.update(b"\n".join(keep))
        # every file the reader opened: the .cdf / .dos actually used (the optional 'CDF <file>' / 'DOS <file>' flags redirect
        # them) and the atomic data bases that supply what a .cdf leaves out (masses, Ip, Nel, Ek, decay times, form factors)
        rels = [self.get_string("cdf_file"), self.get_string("dos_file")]
        mat = self.get_string("material")
        if os.path.isdir(os.path.join(self.dir, "INPUT_DSF", mat)):
            rels += [f"INPUT_DSF/{mat}/{n}" for n in sorted(os.listdir(os.path.join(self.dir, "INPUT_DSF", mat)))]
        rels += [f"INPUT_EADL/{n}" for n in ("INPUT_atomic_data.dat", "radiative_widths.dat", "EADL2023.ALL", "EPDL2023.ALL",
                                              "Atomic_form_factors.dat")]
        for rel in rels:
            p = os.path.join(self.dir, rel)
            if rel and os.path.isfile(p):
                hsh.update(rel.encode())
                with open(p, "rb") as f:
                    hsh.update(f.read())
        hsh.update(_host().trk3_host_version())
        hsh.update(b"window" if shi_window_only else b"full")
        return hsh.hexdigest()[:20]

    def build_tables(self, threads=0, shi_window_only=False, verbose=False, cache_dir=None, evaluator=None):
        """Build (or load from the binary cache) all MFP and differential cross-section tables.

        evaluator: who evaluates the q-integrals of the loss function (>99 % of the work): None = the host threads,
        point by point, as the reference does; "gpu" = all of them at once by trk3_dcs_eval of libtrekis3_gpu.so
        (raises without the CUDA library); "host" = the same record / evaluate / replay path with
        the host threads as the evaluator (tests).  The tables are identical in all three cases."""
        lib = _host()
        err = C.create_string_buffer(1024)
        cache = None
        if cache_dir:
            os.makedirs(cache_dir, exist_ok=True)
            cache = os.path.join(cache_dir, f"{self.get_string('material')}_{self.get_string('ion')}_"
                                            f"{self._input_digest(shi_window_only)}.trk3tab")
            if os.path.exists(cache):
                if lib.trk3h_load_tables(self._h, os.fsencode(cache), err, 1024) == 0:
                    return self
        fn = None
        if evaluator == "gpu":
            from .engine import _gpu
            fn = C.cast(_gpu().trk3_dcs_eval, C.c_void_p)
        elif evaluator == "host":
            fn = C.cast(lib.trk3h_dcs_eval_host, C.c_void_p)
        elif evaluator is not None:
            raise ValueError("evaluator must be None, 'host' or 'gpu'")
        lib.trk3h_set_dcs_evaluator(fn)
        try:
            rc = lib.trk3h_build_tables(self._h, threads, int(shi_window_only), int(verbose), err, 1024)
        finally:
            lib.trk3h_set_dcs_evaluator(None)
        if rc != 0:
            raise RuntimeError("table build failed: " + err.value.decode(errors="replace"))
        if cache:
            tmp = cache + f".tmp{os.getpid()}"
            if lib.trk3h_save_tables(self._h, os.fsencode(tmp), err, 1024) == 0:
                os.replace(tmp, cache)
        return self

    def write_reference_cache(self, out_root):
        """Write the built tables as the reference's own on-disk cache under <out_root>/OUTPUT_<material>/ (names, rows
        and number formats of Analytical_IMFPs.f90:262-786, 917-1606, 2306-2707): a TREKIS-3 working directory that
        holds these files skips its table integration.  Returns the number of files written."""
        n = C.c_int(0)
        err = C.create_string_buffer(1024)
        if _host().trk3h_write_reference_cache(self._h, os.fsencode(out_root), C.byref(n), err, 1024) != 0:
            raise RuntimeError("write_reference_cache failed: " + err.value.decode(errors="replace"))
        return n.value

    def read_reference_cache(self, out_root, threads=0):
        """Take the tables from a reference-format cache under <out_root>/OUTPUT_<material>/ instead of building them
        (the differential ion table of the given ion energy is always computed, as MAIN.f90:231-238 does).  Raises if a
        file is missing or its row count differs from the energy grid -- the conditions under which the reference
        recomputes (Analytical_IMFPs.f90:307-327)."""
        err = C.create_string_buffer(1024)
        if _host().trk3h_read_reference_cache(self._h, os.fsencode(out_root), int(threads), err, 1024) != 0:
            raise RuntimeError("read_reference_cache failed: " + err.value.decode(errors="replace"))
        return self

    def reference_cache_name(self, which):
        out = C.create_string_buffer(512)
        if _host().trk3h_reference_cache_name(self._h, which.encode(), out, 512) != 0:
            raise KeyError(which)
        return out.value.decode()

    @property
    def config(self):
        p = _host().trk3h_config(self._h)
        if not p:
            raise RuntimeError("tables are not built yet (call build_tables)")
        return p.contents

    @property
    def tables(self):
        p = _host().trk3h_tables(self._h)
        if not p:
            raise RuntimeError("tables are not built yet (call build_tables)")
        return p.contents

    def layout(self):
        lay = TallyLayout()
        rc = _host().trk3_tally_layout_init(C.byref(self.config), C.byref(self.tables), C.byref(lay))
        if rc != 0:
            raise RuntimeError(f"trk3_tally_layout_init failed ({rc})")
        return lay

    def table_arrays(self):
        """numpy views of the flattened tables (copies), for inspection and tests."""
        t = self.tables
        ns = t.n_shells

        def arr(ptr, n):
            return np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n > 0 and ptr else np.zeros(0)

        def off(ptr, n):
            return np.ctypeslib.as_array(ptr, shape=(n,)).copy()

        out = {
            "ei_E": arr(t.ei_E, t.n_ei), "ei_L": arr(t.ei_L, ns * t.n_ei).reshape(ns, -1),
            "ee_E": arr(t.ee_E, t.n_ee), "ee_L": arr(t.ee_L, t.n_ee),
            "hi_E": arr(t.hi_E, t.n_hi), "hi_L": arr(t.hi_L, ns * t.n_hi).reshape(ns, -1),
            "he_E": arr(t.he_E, t.n_he), "he_L": arr(t.he_L, t.n_he),
            "shi_E": arr(t.shi_E, t.n_shi), "shi_L": arr(t.shi_L, ns * t.n_shi).reshape(ns, -1),
            "shi_dEdx": arr(t.shi_dEdx, ns * t.n_shi).reshape(ns, -1),
            "dos_E": arr(t.dos_E, t.n_dos), "dos_DOS": arr(t.dos_DOS, t.n_dos), "dos_int": arr(t.dos_int, t.n_dos),
            "dos_effm": arr(t.dos_effm, t.n_dos), "out_R": arr(t.out_R, t.n_r), "out_V": arr(t.out_V, t.n_r),
            "shell_Ip": np.array(t.shell_Ip[:ns]), "shell_Nel": np.array(t.shell_Nel[:ns]),
            "shell_auger": np.array(t.shell_auger[:ns]), "shell_radiat": np.array(t.shell_radiat[:ns]),
        }
        if t.n_ph > 0:
            out["ph_E"] = arr(t.ph_E, t.n_ph)
            out["ph_L"] = arr(t.ph_L, ns * t.n_ph).reshape(ns, -1)
        for name, n_rows in (("dshi", ns), ("eid", ns * t.n_ei), ("eed", t.n_ee), ("hid", t.n_hi), ("hed", t.n_he)):
            o = off(getattr(t, name + "_off"), n_rows + 1)
            out[name + "_off"] = o
            a, b = ("E", "L") if name == "dshi" else ("hw", "L")
            out[f"{name}_{a}"] = arr(getattr(t, f"{name}_{a}"), int(o[-1]))
            out[f"{name}_{b}"] = arr(getattr(t, f"{name}_{b}"), int(o[-1]))
        return out

    # -- single-point evaluations (table-builder parity tests) ---------------------------------
    def eval_TotIMFP(self, E, atom, shell, kind=0):
        L, d = C.c_double(), C.c_double()
        _host().trk3h_eval_TotIMFP(self._h, E, atom, shell, kind, C.byref(L), C.byref(d))
        return L.value, d.value

    def eval_EMFP(self, E, kind=0):
        L, d = C.c_double(), C.c_double()
        _host().trk3h_eval_EMFP(self._h, E, kind, C.byref(L), C.byref(d))
        return L.value, d.value

    def eval_dcs_phonon(self, Ee, hw):
        """(value of Diff_cross_section_phonon for an electron, amplitude of the first phonon oscillator as used)"""
        v, a0 = C.c_double(), C.c_double()
        if _host().trk3h_eval_dcs_phonon(self._h, Ee, hw, C.byref(v), C.byref(a0)) != 0:
            raise RuntimeError("trk3h_eval_dcs_phonon failed")
        return v.value, a0.value

    def eval_SHI(self, E, atom, shell):
        s, d, z = C.c_double(), C.c_double(), C.c_double()
        _host().trk3h_eval_SHI(self._h, E, atom, shell, C.byref(s), C.byref(d), C.byref(z))
        return s.value, d.value, z.value

    def eval_photon(self, E, atom, shell):
        L = C.c_double()
        _host().trk3h_eval_photon(self._h, E, atom, shell, C.byref(L))
        return L.value

    def sumrules(self, atom, shell):
        k, f = C.c_double(), C.c_double()
        _host().trk3h_sumrules(self._h, atom, shell, C.byref(k), C.byref(f))
        return k.value, f.value

    def grid(self, Emin, Emax):
        n = _host().trk3h_grid(self._h, Emin, Emax, None, 0)
        buf = (C.c_double * n)()
        _host().trk3h_grid(self._h, Emin, Emax, buf, n)
        return np.array(buf)

    # -- output --------------------------------------------------------------------------------
    def save_output(self, tallies, nmc, out_root):
        lay = self.layout()
        t = np.ascontiguousarray(tallies, dtype=np.float64)
        err = C.create_string_buffer(1024)
        od = C.create_string_buffer(2048)
        rc = _host().trk3h_save_output(self._h, C.byref(lay), t.ctypes.data_as(C.POINTER(C.c_double)), int(nmc),
                                       os.fsencode(out_root), od, 2048, err, 1024)
        if rc != 0:
            raise RuntimeError("save_output failed: " + err.value.decode(errors="replace"))
        return od.value.decode()


def split_tallies(lay, buf):
    """dict name -> numpy array (Fortran order, reference shapes) of a packed tally buffer."""
    out = {}
    for i, name in enumerate(_abi.TALLY_NAMES):
        a = np.asarray(buf[lay.off[i]: lay.off[i] + lay.len[i]])
        out[name] = a.reshape(lay.shape(name), order="F")
    return out
