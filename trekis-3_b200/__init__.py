"""trekis3_b200 -- B200-native Monte-Carlo cascade engine of TREKIS-3 (hot path of Monte_Carlo.f90).

Host side mirrors the reference's driver surface:
    Case.load(dir)          <- Read_input_file            (Reading_files_and_parameters.f90:162)
    Case.build_tables()     <- Analytical_*_dEdx          (Universal_MC_for_SHI_MAIN.f90:146-247)
    do_Monte_Carlo(case)    <- do_Monte_Carlo             (Monte_Carlo.f90:39)   [CUDA, sm_100a]
    Case.save_output(...)   <- Save_output                (Sorting_output_data.f90:340)
The CUDA engine has no CPU fallback: do_Monte_Carlo raises if libtrekis3_gpu.so or a GPU is missing.
"""
from ._abi import Config, Tables, TallyLayout, Stats, TALLY_NAMES, EVENT_NAMES, EVENT_BYTES, lib_path  # noqa: F401
from .host import Case, make_run_dir, CONFIGS  # noqa: F401
from .engine import Engine, do_Monte_Carlo, gpu_library_loaded, release_handles, nccl_unique_id  # noqa: F401

__version__ = "0.1.0"
