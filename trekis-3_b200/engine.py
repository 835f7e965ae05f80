"""CUDA engine binding (libtrekis3_gpu.so).  There is no CPU fallback: every entry point raises
if the CUDA library cannot be loaded or no GPU is present."""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import Config, Tables, TallyLayout, Stats

_lib = None


def _gpu():
    global _lib
    if _lib is None:
        path = _abi.lib_path("gpu")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: the CUDA engine has no CPU fallback. Build it with "
                               "`python -c 'import __graft_entry__ as g; g.build()'`.")
        lib = C.CDLL(path)
        PD = C.POINTER(C.c_double)
        lib.trk3_mc_create.argtypes = [C.POINTER(Config), C.POINTER(Tables), C.c_int, C.POINTER(C.c_void_p)]
        lib.trk3_mc_run.argtypes = [C.c_void_p, C.c_int64, C.c_int64, PD, C.POINTER(Stats)]
        lib.trk3_mc_run_device.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(Stats)]
        lib.trk3_mc_device_tallies.restype = C.c_void_p
        lib.trk3_mc_device_tallies.argtypes = [C.c_void_p]
        lib.trk3_mc_zero_device_tallies.argtypes = [C.c_void_p]
        lib.trk3_mc_download_tallies.argtypes = [C.c_void_p, PD]
        lib.trk3_mc_iteration_energies.argtypes = [C.c_void_p, PD, C.c_int64, C.POINTER(C.c_int64)]
        lib.trk3_mc_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        lib.trk3_mc_layout.restype = C.POINTER(TallyLayout)
        lib.trk3_mc_layout.argtypes = [C.c_void_p]
        lib.trk3_mc_last_error.restype = C.c_char_p
        lib.trk3_mc_last_error.argtypes = [C.c_void_p]
        lib.trk3_mc_destroy.argtypes = [C.c_void_p]
        lib.trk3_mc_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        lib.trk3_mc_set_device_tallies.argtypes = [C.c_void_p, C.c_void_p]
        lib.trk3_mc_kernel_times.argtypes = [C.c_void_p, PD, C.POINTER(C.c_uint64), C.c_int]
        lib.trk3_mc_reload_tables.argtypes = [C.c_void_p, C.POINTER(Config), C.POINTER(Tables)]
        lib.trk3_mc_table_bytes.restype = C.c_uint64
        lib.trk3_mc_table_bytes.argtypes = [C.c_void_p]
        lib.trk3_nccl_unique_id.argtypes = [C.c_void_p]
        lib.trk3_mc_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.trk3_mc_set_comm.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.trk3_mc_comm_size.argtypes = [C.c_void_p]
        lib.trk3_mc_reset.argtypes = [C.c_void_p]
        lib.trk3_dcs_stats.argtypes = [PD, C.POINTER(C.c_int64), C.c_int]
        lib.trk3_gpu_version.restype = C.c_char_p
        _lib = lib
    return _lib


def dcs_stats(reset=False):
    """Device time [ms] and number of q-integrals evaluated by trk3_dcs_eval (the GPU table builder) so far."""
    ms, n = C.c_double(0.0), C.c_int64(0)
    _gpu().trk3_dcs_stats(C.byref(ms), C.byref(n), int(reset))
    return ms.value, n.value


NCCL_UNIQUE_ID_BYTES = 128


def nccl_unique_id():
    """128-byte NCCL id (rank 0 creates it, every rank passes it to Engine.comm_init)."""
    buf = C.create_string_buffer(NCCL_UNIQUE_ID_BYTES)
    rc = _gpu().trk3_nccl_unique_id(buf)
    if rc != 0:
        raise RuntimeError(f"trk3_nccl_unique_id failed ({rc}): NCCL not available")
    return buf.raw


def gpu_library_loaded():
    """True once libtrekis3_gpu.so is mapped into this process (used by tests to prove no fallback ran)."""
    return _lib is not None


class Engine:
    """One engine per GPU: tables uploaded once, iterations run in batches of wavefronts."""

    def __init__(self, case, device=-1, seed=None, **options):
        lib = _gpu()
        self.case = case
        cfg = Config.from_buffer_copy(case.config)
        if seed is not None:
            cfg.seed = int(seed)
        self._cfg = cfg
        h = C.c_void_p()
        rc = lib.trk3_mc_create(C.byref(cfg), C.byref(case.tables), int(device), C.byref(h))
        if rc != 0 or not h:
            msg = lib.trk3_mc_last_error(h).decode() if h else "no engine"
            raise RuntimeError(f"trk3_mc_create failed ({rc}): {msg}")
        self._h = h
        self.layout = lib.trk3_mc_layout(h).contents
        for k, v in options.items():
            self.set_option(k, v)

    def reload_tables(self, case, seed=None):
        """Copy the case's configuration and tables to the device again (same shapes): the per-call input transfer."""
        cfg = Config.from_buffer_copy(case.config)
        if seed is not None:
            cfg.seed = int(seed)
        self._cfg = cfg
        self.case = case
        self._check(_gpu().trk3_mc_reload_tables(self._h, C.byref(cfg), C.byref(case.tables)), "trk3_mc_reload_tables")

    def table_bytes(self):
        """Bytes copied host->device by the last table binding."""
        return int(_gpu().trk3_mc_table_bytes(self._h))

    def set_option(self, name, value):
        rc = _gpu().trk3_mc_set_option(self._h, name.encode(), float(value))
        if rc != 0:
            raise KeyError(name)

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {_gpu().trk3_mc_last_error(self._h).decode()}")

    def run(self, it_begin, it_end, tallies=None):
        """Run iterations [it_begin, it_end); host buffers in, host buffers out (copies inside)."""
        if tallies is None:
            tallies = np.zeros(self.layout.total, dtype=np.float64)
        st = Stats()
        self._check(_gpu().trk3_mc_run(self._h, int(it_begin), int(it_end),
                                       tallies.ctypes.data_as(C.POINTER(C.c_double)), C.byref(st)), "trk3_mc_run")
        return tallies, st.as_dict()

    def run_device(self, it_begin, it_end):
        st = Stats()
        self._check(_gpu().trk3_mc_run_device(self._h, int(it_begin), int(it_end), C.byref(st)), "trk3_mc_run_device")
        return st.as_dict()

    def comm_init(self, nranks, rank, unique_id):
        """Attach this engine to an NCCL communicator of `nranks` engines (one per GPU/process): from now on run() and
        run_device() are collective and end with ONE all-reduce of the tally buffer, issued by the library on its own
        stream (replaces the 26 MPI_Reduce of Monte_Carlo.f90:131-389)."""
        buf = C.create_string_buffer(bytes(unique_id), NCCL_UNIQUE_ID_BYTES)
        self._check(_gpu().trk3_mc_comm_init(self._h, int(nranks), int(rank), buf), "trk3_mc_comm_init")

    def comm_init_torch(self, device=None):
        """comm_init with the id distributed through an initialised torch.distributed process group (plumbing only)."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        on_gpu = dist.get_backend() == "nccl"
        t = torch.zeros(NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            t = torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8).clone()
        if on_gpu:
            t = t.cuda(device)
        dist.broadcast(t, src=0)
        self.comm_init(world, rank, bytes(t.cpu().numpy().tobytes()))

    def comm_size(self):
        return int(_gpu().trk3_mc_comm_size(self._h))

    def reset(self):
        """Back to a defined state after a failed run (see trk3_mc_reset)."""
        self._check(_gpu().trk3_mc_reset(self._h), "trk3_mc_reset")

    def set_stream(self, cuda_stream_handle):
        self._check(_gpu().trk3_mc_set_stream(self._h, C.c_void_p(int(cuda_stream_handle))), "set_stream")

    def set_device_tallies(self, device_ptr):
        """Accumulate into a caller-owned device buffer (layout.total float64), e.g. tensor.data_ptr()."""
        self._check(_gpu().trk3_mc_set_device_tallies(self._h, C.c_void_p(int(device_ptr))), "set_device_tallies")

    KERNEL_CLASSES = ("k_wave<electron,hot>", "k_wave<vbhole,hot>", "k_wave<corehole>", "k_wave<photon>", "k_shi", "finalize",
                      "k_wave<electron,cold>", "k_wave<vbhole,cold>", "k_wave<electron,warm>", "k_wave<vbhole,warm>")

    def kernel_times(self):
        ms = (C.c_double * 10)()
        n = (C.c_uint64 * 10)()
        _gpu().trk3_mc_kernel_times(self._h, ms, n, 10)
        return {k: {"ms": ms[i], "launches": int(n[i])} for i, k in enumerate(self.KERNEL_CLASSES)}

    def device_tallies_ptr(self):
        return _gpu().trk3_mc_device_tallies(self._h)

    def zero_device_tallies(self):
        self._check(_gpu().trk3_mc_zero_device_tallies(self._h), "zero")

    def download_tallies(self, into=None):
        if into is None:
            into = np.zeros(self.layout.total, dtype=np.float64)
        self._check(_gpu().trk3_mc_download_tallies(self._h, into.ctypes.data_as(C.POINTER(C.c_double))), "download")
        return into

    def iteration_energies(self, n_max):
        buf = np.zeros((n_max, self.layout.Nt), dtype=np.float64)
        n = C.c_int64()
        self._check(_gpu().trk3_mc_iteration_energies(self._h, buf.ctypes.data_as(C.POINTER(C.c_double)), buf.size,
                                                      C.byref(n)), "iteration_energies")
        return buf[: n.value]

    def close(self):
        if getattr(self, "_h", None):
            _gpu().trk3_mc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_handles = {}        # persistent plugin handles: (device, table shapes) -> Engine


def _shape_key(case, device):
    t = case.tables
    lay = case.layout()
    return (int(device), int(lay.total), int(lay.Nt), int(t.n_shells), int(t.n_ei), int(t.n_ee), int(t.n_hi), int(t.n_he),
            int(t.n_ph), int(t.n_shi), int(t.n_dos), int(t.n_r), int(t.eid_off[t.n_shells * t.n_ei]),
            int(t.eed_off[t.n_ee]), int(t.hid_off[t.n_hi]), int(t.hed_off[t.n_he]), int(t.dshi_off[t.n_shells]),
            bool(case.config.work_function > 0), bool(case.config.include_photons))


def release_handles():
    """Destroy the persistent engines kept by do_Monte_Carlo."""
    for e in _handles.values():
        e.close()
    _handles.clear()


def do_Monte_Carlo(case, NMC=None, device=-1, seed=None, it_begin=0, persistent=True, comm=None, **options):
    """Replacement of `call do_Monte_Carlo(NMC, SHI, ...)` (Monte_Carlo.f90:39): returns the summed
    Out_* tallies (not yet divided by NMC, as in the reference) and the run statistics.

    Every call copies its inputs (configuration + tables) host->device and the tallies device->host.  With
    `persistent` (default) the engine handle -- device queues and scratch -- is kept between calls with the same table
    shapes, as a plugin linked into the Fortran host would do; `persistent=False` creates and destroys an engine.

    comm = (nranks, rank, nccl_unique_id) makes the call collective over `nranks` processes, one GPU each, every rank passing its
    own share of the global iterations [it_begin, it_begin + NMC): the library sums the tallies of all ranks with one NCCL
    all-reduce before they are copied back, and every rank returns the reduced arrays (the MPI build of the reference:
    Monte_Carlo.f90:111-129 + :131-389).  The communicator is created once per persistent handle."""
    n = int(NMC if NMC is not None else case.get("NMC"))
    if not persistent:
        eng = Engine(case, device=device, seed=seed, **options)
        try:
            if comm is not None:
                eng.comm_init(*comm)
            return eng.run(it_begin, it_begin + n)
        finally:
            eng.close()
    key = _shape_key(case, device)
    eng = _handles.get(key)
    if eng is None:
        eng = _handles[key] = Engine(case, device=device, seed=seed)
        if comm is not None:
            eng.comm_init(*comm)
    else:
        eng.reload_tables(case, seed=seed)
        if comm is not None and eng.comm_size() != comm[0]:
            eng.comm_init(*comm)
    for k, v in options.items():
        eng.set_option(k, v)
    return eng.run(it_begin, it_begin + n)
