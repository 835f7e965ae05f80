#!/usr/bin/env python
"""bench.py -- headline benchmark of the TREKIS-3 Monte-Carlo hot path on B200.

Metric (BASELINE.json): MC iterations/s (whole job, all GPUs); events/s reported beside it.
Workload at N=1: BASELINE.json configs[1] -- Au 2187 MeV in SiO2_cryst, photon transport and Auger/radiative
decays on, 1000 MC iterations per step.  A step is one do_Monte_Carlo call (1000 ion impacts + full cascades
to 100 fs).  N>1, two modes:
  --scaling weak   (default for C1-C4) every rank runs its own NMC iterations per step (disjoint global ranges);
  --scaling strong (default for --config C5, BASELINE.json configs[4]) the NMC = 1e5 iterations of a step are split
                   contiguously over the ranks, value = NMC / max-rank time.
Either way a step ends with ONE ncclAllReduce of the packed tally buffer issued by the library itself on its own stream
(trk3_mc_comm_init: the all-reduce sits behind the C ABI, replacing the 26 MPI_Reduce of Monte_Carlo.f90:131-389);
torch.distributed only carries the 128-byte NCCL id and the barrier / max-over-ranks of the timing.

  python bench.py --gpus N --steps K --warmup W            our CUDA engine
  python bench.py --impl reference ...                      the CPU restatement of the reference on the host cores

One JSON line on stdout (rank 0).  See the task contract for the keys.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "MC iterations/s (Monte_Carlo.f90 cascade engine, whole job)"
UNIT = "iterations/s"


def load_case(config, nmc):
    import trekis3_b200 as tk
    run_dir = os.path.join(ROOT, ".bench_run", f"{config}_{os.getpid()}")
    tk.make_run_dir(run_dir, config, nmc=nmc)
    case = tk.Case.load(run_dir)
    # the tables the reference main builds: the ion over its whole energy grid (Analytical_IMFPs.f90:2242-2510).  Built once
    # (host threads: a few seconds; excluded from every timing, as the reference caches them too) and kept in .table_cache
    case.build_tables(shi_window_only=False, cache_dir=os.path.join(ROOT, ".table_cache"))
    return case


def workload_name(config, nmc):
    import trekis3_b200 as tk
    mat, z, e, ph, _ = tk.CONFIGS[config]
    ion = {54: "Xe", 79: "Au", 92: "U"}.get(z, str(z))
    return f"{config}: {ion} {e:g} MeV in {mat}, photons {'on' if ph else 'off'}, {nmc} MC iterations per step, T=100 fs"


def bench_config(args, world, tally_doubles):
    """The `config` object of the JSON line: the workload and how the GPU arm runs it.  Both arms print the SAME object (the reference
    arm runs "on your arm's config"), what is specific to the CPU run is in its `cpu_baseline.sample`."""
    nmc, strong = args.nmc, args.scaling == "strong"
    return {"workload": workload_name(args.config, nmc), "iterations_in_flight": args.batch,
            "step": (f"one step = one call of {nmc} iterations = ONE batch: latency-bound by the chain of hot generations; "
                     "`throughput` holds the same configuration with 4096 iterations in flight") if nmc <= args.batch else
                    f"one step = one call of {nmc} iterations, run as batches of {args.batch} iterations in flight",
            "l2": "256 MiB buffer written between timed steps (L2 flush); queues stream through HBM",
            "parallelism": (f"{nmc} iterations per step split contiguously over {world} GPU(s) (strong scaling)" if strong else
                            f"every one of {world} GPU(s) runs its own {nmc} iterations per step (weak scaling)") +
                           f"; one ncclAllReduce of {tally_doubles} doubles per step, issued by libtrekis3_gpu.so on its own stream",
            "inputs": "shipped INPUT_CDF/INPUT_DOS files; radiative widths from data/INPUT_EADL/radiative_widths.dat "
                      "(approximate, EADL2023.ALL is not redistributable)"}


_JSON_FD = None


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its
    version banner there): from here on fd 1 goes to stderr, and the JSON line alone goes to the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled in-process every few
    milliseconds (a `nvidia-smi -lms` child needs longer to start than a short timed region lasts); nvidia-smi is the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.sm, self.mx, self.reasons = [], [], set()
        self.nvml, self.handle, self.stop_flag, self.thread = None, None, threading.Event(), None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:        # CUDA ordinal -> NVML device through the PCI address (CUDA_VISIBLE_DEVICES may renumber)
            import torch
            pr = torch.cuda.get_device_properties(self.gpu)
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
        return pynvml, h

    def _sample(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
            mask = int(get(self.handle))
            for bit, name in self.BITS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _poll(self):
        while not self.stop_flag.is_set():
            try:
                self._sample()
            except Exception:
                break
            self.stop_flag.wait(0.004)

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.mx.append(float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)))
            self._sample()                      # the first NVML queries of a process can take tens of milliseconds: not in the timed region
            self.sm.clear(); self.reasons.clear()
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml:
            self.stop_flag.set()
            if self.thread:
                self.thread.join(timeout=1.0)
            if not self.sm:                 # a timed region shorter than one polling period: one sample right at its end
                try:
                    self._sample()
                except Exception:
                    pass
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml"}
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def cpu_reference_rate(case, seconds, threads=0):
    """Oracle (CPU restatement of the reference algorithm, all host threads) on a bounded sample."""
    import oracle_api
    cores = os.cpu_count() or 1
    n_done, t0 = 0, time.perf_counter()
    chunk = max(cores, 1)
    events = 0
    while True:
        _, st, _, _ = oracle_api.run(case, n_done, n_done + chunk, rng_mode=0, threads=threads)
        n_done += chunk
        events += st["total_events"]
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
        # grow the chunk so that the thread pool start-up does not dominate
        chunk = min(max(chunk, int(chunk * 2)), max(cores, int((seconds - dt) / max(dt / n_done, 1e-9))))
        if chunk < 1:
            break
    dt = time.perf_counter() - t0
    return n_done / dt, events / dt, n_done, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    case = load_case(args.config, args.nmc)
    import oracle_api
    flags = oracle_api.use_native_build()       # compiled on THIS machine: -O3 -march=native
    per_step = max(2.0, min(12.0, 150.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_reference_rate(case, min(per_step, 2.0))
    rates, ev_rates, n_tot, t_tot = [], [], 0, 0.0
    for _ in range(args.steps):
        r, e, n, dt, cores = cpu_reference_rate(case, per_step)
        rates.append(r); ev_rates.append(e); n_tot += n; t_tot += dt
    value = n_tot / t_tot
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(1, args.steps), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, max(1, args.gpus), case.layout().total),
        "sample": f"{n_tot} iterations in {t_tot:.1f} s",
        "events_per_s": sum(ev_rates) / len(ev_rates),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                         "sample": f"oracle/trk3_oracle.cpp (reference algorithm incl. O(N) next-event search, {flags}), "
                                   f"{n_tot} iterations of the same workload in {t_tot:.1f} s on {os.cpu_count()} threads; "
                                   "the Fortran reference cannot be built: no Fortran compiler in the image or on the box "
                                   "(profiles/r2_probe_fortran.txt; recipe: oracle/build_ref.sh)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def table_bytes(case):
    return int(sum(a.nbytes for a in case.table_arrays().values()))


def run_ours(args):
    import numpy as np
    import torch
    import trekis3_b200 as tk

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # rank 0 builds the tables (or finds them in .table_cache), the other ranks then load them from the cache
    if dist is not None and rank != 0:
        dist.barrier()
    case = load_case(args.config, args.nmc)
    if dist is not None and rank == 0:
        dist.barrier()
    nmc = args.nmc
    lay = case.layout()
    stream = torch.cuda.Stream()
    tally = torch.zeros(lay.total, dtype=torch.float64, device="cuda")
    eng = tk.Engine(case, device=local_rank, batch=args.batch)
    eng.set_stream(stream.cuda_stream)
    eng.set_device_tallies(tally.data_ptr())
    if dist is not None:
        eng.comm_init_torch(local_rank)         # from here on run_device() is collective: the library all-reduces the tallies
    strong = args.scaling == "strong"

    def my_range(k):
        """global iteration range of this rank in step k"""
        if strong:          # the NMC iterations of the step split contiguously over the ranks
            return k * nmc + rank * nmc // world, k * nmc + (rank + 1) * nmc // world
        first = ((k * world) + rank) * nmc      # weak: every rank its own NMC iterations
        return first, first + nmc
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")     # > 126 MB L2

    def step(k):
        # iterations of this rank for step k: disjoint global ranges -> independent Philox streams
        lo, hi = my_range(k)
        with torch.cuda.stream(stream):
            tally.zero_()
            st = eng.run_device(lo, hi)         # ends with the single collective of the path, issued by the library (N > 1)
        return st

    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    eng.set_option("profile", 1)
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    events_total, launches, drift, errors = 0, 0, 0.0, {}
    ev_by_class = {}
    cold_ev = {"electron": 0, "vbhole": 0}
    warm_ev = {"electron": 0, "vbhole": 0}
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                            # L2 flush between timed iterations (untimed)
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            ev[k][0].record(stream)
        st = step(args.warmup + k)
        with torch.cuda.stream(stream):
            ev[k][1].record(stream)
        events_total += st["total_events"]; launches += st["kernel_launches"]
        drift = max(drift, st["max_energy_drift"])
        for n, v in st["events"].items():
            ev_by_class[n] = ev_by_class.get(n, 0) + v
        for n, v in st["errors"].items():
            errors[n] = errors.get(n, 0) + v
        for n, v in st["cold_events"].items():
            cold_ev[n] += v
        for n, v in st["warm_events"].items():
            warm_ev[n] += v
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms, float(events_total)], dtype=torch.float64, device="cuda")
    if dist is not None:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, events_all = float(tmax[0]), float(tsum[1])
    else:
        events_all = float(events_total)
    job_iterations = nmc if strong else world * nmc          # iterations of one step, all ranks
    value = job_iterations * args.steps / (ms * 1e-3)
    ktimes = eng.kernel_times()

    # ---- end-to-end through the public API with HOST buffers (tables H2D + tallies D2H inside the timed region)
    e2e_steps = max(1, args.steps)
    host_tally = np.zeros(lay.total)
    e2e_comm = None
    if dist is not None:
        id_t = torch.zeros(tk.engine.NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            id_t = torch.frombuffer(bytearray(tk.nccl_unique_id()), dtype=torch.uint8).clone().cuda()
        dist.broadcast(id_t, src=0)
        e2e_comm = (world, rank, bytes(id_t.cpu().numpy().tobytes()))
    lo, hi = my_range(0)
    # warm-up of the public path: the first call creates the plugin handle, the following ones re-bind the inputs as every timed call does
    for k in range(max(2, min(args.warmup, 3))):
        tk.do_Monte_Carlo(case, NMC=hi - lo, device=local_rank, batch=args.batch, comm=e2e_comm)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e2e_calls = []
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        lo, hi = my_range(k)
        tc = time.perf_counter()
        tl, st_e = tk.do_Monte_Carlo(case, NMC=hi - lo, device=local_rank, it_begin=lo, batch=args.batch, comm=e2e_comm)   # collective for N > 1
        host_tally += tl
        e2e_calls.append((round((time.perf_counter() - tc) * 1e3, 3), round(st_e["device_ms"], 3)))
    torch.cuda.synchronize()
    e2e_t = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([e2e_t], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); e2e_t = float(tt[0])
    e2e_value = job_iterations * e2e_steps / e2e_t
    h2d = int(tk.engine._handles[tk.engine._shape_key(case, local_rank)].table_bytes()) + 4096     # tables + configuration block
    d2h = int(lay.total * 8 + (my_range(0)[1] - my_range(0)[0]) * lay.Nt * 20)
    tk.release_handles()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (per-class CUDA-event times measured live above)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    B = tk.EVENT_BYTES
    class_bytes = {
        "k_wave<electron,hot>": ev_by_class.get("el_inelastic", 0) * B["el_inelastic"] + (ev_by_class.get("el_elastic", 0) - cold_ev["electron"] - warm_ev["electron"]) * B["el_elastic"],
        "k_wave<electron,warm>": warm_ev["electron"] * B["el_elastic"],
        "k_wave<vbhole,warm>": warm_ev["vbhole"] * B["vbh_elastic"],
        "k_wave<vbhole,hot>": ev_by_class.get("vbh_inelastic", 0) * B["vbh_inelastic"] + (ev_by_class.get("vbh_elastic", 0) - cold_ev["vbhole"] - warm_ev["vbhole"]) * B["vbh_elastic"],
        "k_wave<electron,cold>": cold_ev["electron"] * B["el_elastic"],
        "k_wave<vbhole,cold>": cold_ev["vbhole"] * B["vbh_elastic"],
        "k_wave<corehole>": ev_by_class.get("auger", 0) * B["auger"] + ev_by_class.get("radiative", 0) * B["radiative"] + ev_by_class.get("auger_frozen", 0) * B["auger_frozen"],
        "k_wave<photon>": ev_by_class.get("photon", 0) * B["photon"],
        "k_shi": ev_by_class.get("shi", 0) * B["shi"],
    }
    dom = max((k for k in ktimes if k in class_bytes), key=lambda k: ktimes[k]["ms"])
    dom_ms, dom_n = ktimes[dom]["ms"], max(1, ktimes[dom]["launches"])
    achieved = class_bytes[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    # ---- what ncu measured for the SAME launches (scripts/ncu_round2.sh: `ncu --set full` of every launch of one batch of this
    # workload -> profiles/ncu_classes_<config>.json): DRAM traffic per launch of the class, warp instructions per collision, L2 hit
    # rate, FP64 pipe and issue-slot utilisation
    ncu = {}
    try:
        with open(os.path.join(ROOT, "profiles", f"ncu_classes_{args.config}.json")) as f:
            ncu = json.load(f).get("classes", {})
    except Exception:
        pass
    ncls = ncu.get(dom, {})
    traffic = ncls.get("dram_bytes_per_launch")
    # the bound that does hold these kernels: instruction issue.  Warp instructions per collision come from the ncu capture of the
    # same workload, collisions and time are measured live; peak = SMs x 4 schedulers x SM clock
    sm_clock_hz = float((clk or {}).get("sm_mhz") or 1965.0) * 1e6
    issue_peak = torch.cuda.get_device_properties(local_rank).multi_processor_count * 4 * sm_clock_hz
    n_coll = {
        "k_wave<electron,hot>": ev_by_class.get("el_inelastic", 0) + ev_by_class.get("el_elastic", 0) - cold_ev["electron"] - warm_ev["electron"],
        "k_wave<vbhole,hot>": ev_by_class.get("vbh_inelastic", 0) + ev_by_class.get("vbh_elastic", 0) - cold_ev["vbhole"] - warm_ev["vbhole"],
        "k_wave<electron,cold>": cold_ev["electron"], "k_wave<vbhole,cold>": cold_ev["vbhole"],
        "k_wave<electron,warm>": warm_ev["electron"], "k_wave<vbhole,warm>": warm_ev["vbhole"],
        "k_shi": ev_by_class.get("shi", 0),
    }
    issue = {}
    for k, n in n_coll.items():
        c = ncu.get(k, {})
        if c.get("warp_inst_per_collision") and k in ktimes and ktimes[k]["ms"] > 0 and n > 0:
            rate = n * c["warp_inst_per_collision"] / (ktimes[k]["ms"] * 1e-3)
            issue[k] = {"warp_inst_per_collision": c["warp_inst_per_collision"], "achieved_Gwarp_inst_s": rate / 1e9, "frac": rate / issue_peak,
                        "threads_per_inst_ncu": c.get("threads_per_inst"), "issue_slots_pct_ncu": c.get("issue_slots_pct"),
                        "fp64_pipe_pct_ncu": c.get("fp64_pipe_pct"), "l2_hit_pct_ncu": c.get("l2_hit_pct"), "l1_hit_pct_ncu": c.get("l1_hit_pct"),
                        "stall_no_instruction_cycles_per_issue_ncu": c.get("stall_no_instruction")}
    # algorithmic FP64 work (secondary figure): ~600 flop-equivalents per event against the FP64 peak of the device
    fp64_peak = 148 * 64 * 2 * sm_clock_hz            # 64 DFMA lanes per SM (B200), 2 flop each
    fp64 = {"flop_equivalents_per_event": 600, "achieved_TFLOP_s": 600.0 * events_all / world / (ms * 1e-3) / 1e12,
            "peak_TFLOP_s": fp64_peak / 1e12, "frac": 600.0 * events_all / world / (ms * 1e-3) / fp64_peak}
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "launches": dom_n, "avg_launch_ms": dom_ms / dom_n,
                "algorithmic_bytes_per_launch": class_bytes[dom] / dom_n,
                "traffic_note": "traffic = dram__bytes_read.sum + dram__bytes_write.sum summed over ALL launches of this kernel class in one batch of the "
                                "same workload (ncu --set full), divided by their number: the same launches algorithmic_bytes_per_launch averages over",
                "kernel_share_of_step": dom_ms / ms if ms > 0 else None,
                "primary": "instruction_issue",
                "instruction_issue": {"peak_Gwarp_inst_s": issue_peak / 1e9, "kernels": issue,
                                      "note": "the roofline that does bound these kernels: warp instructions issued / (SMs x 4 schedulers x SM clock); "
                                              "instructions per collision from the ncu capture of every launch of one batch (profiles/ncu_classes_*.json), "
                                              "collisions (device counters) and CUDA-event times measured live"},
                "fp64": fp64,
                "per_kernel": {k: {"GB/s": (class_bytes[k] / (ktimes[k]["ms"] * 1e-3) / 1e9 if ktimes[k]["ms"] > 0 else 0.0),
                                   "ms": ktimes[k]["ms"], "launches": ktimes[k]["launches"]} for k in class_bytes},
                "note": "algorithmic bytes = collisions x compulsory particle-state bytes (SURVEY.md 8d).  Histories stay in registers between "
                        "collisions and the tables are L1/L2-resident, so the HBM fraction is small by construction; the kernels are bound by "
                        "instruction issue and fetch (fp64 libm sequences; loop bodies of 64-200 KB of SASS against a 32 KB instruction cache: "
                        "stall no_instruction 2-3 cycles per issued instruction), the hot cascade also by the latency of its longest history; "
                        "the hot/warm/core-hole/photon kernels of one generation run on concurrent streams, so their class times overlap"}

    # ---- the same workload with the GPU filled: a step of `nmc` iterations is one latency-bound batch (a chain of ~10 hot
    # generations whatever the batch size); 4096 iterations in flight give the throughput figure of the engine (state both)
    throughput = None
    if world == 1 and args.config != "C5" and not args.no_throughput:
        try:
            eng_t = tk.Engine(case, device=local_rank, batch=4096)
            eng_t.run_device(10_000_000, 10_000_000 + 4096)                       # warm-up: allocates the queues of 4096 iterations
            best = min(eng_t.run_device(10_000_000 + 4096 * (i + 1), 10_000_000 + 4096 * (i + 2))["device_ms"] for i in range(2))
            throughput = {"iterations_in_flight": 4096, "ms_per_4096_iterations": best, "value": 4096.0 / (best * 1e-3), "unit": UNIT,
                          "note": "device time of one 4096-iteration batch of the same configuration, tables resident (min of 2 after a warm-up)"}
            eng_t.close()
        except Exception as e:                       # never lose the headline over the side figure
            throughput = {"error": str(e)}

    # ---- CPU baseline: oracle on the host cores, bounded sample of the same workload
    cpu = None
    if not args.no_cpu_baseline:
        r, e, n, dt, cores = cpu_reference_rate(case, args.cpu_seconds)
        cpu = {"value": r, "unit": UNIT, "cores": cores, "kind": "port", "events_per_s": e,
               "sample": f"oracle/trk3_oracle.cpp (restated reference, incl. its O(N) next-event search), {n} iterations of the "
                         f"same workload in {dt:.1f} s on {cores} host threads"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, world, lay.total),
        "events_per_s": events_all / (ms * 1e-3), "events_per_s_per_gpu": events_all / (ms * 1e-3) / world,
        "events_by_class": ev_by_class, "cold_events": cold_ev, "warm_events": warm_ev, "wall_s": t_wall,
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_call": [c for c, _ in e2e_calls], "device_ms_per_call": [d for _, d in e2e_calls],
                "note": "trekis3_b200.do_Monte_Carlo(case) with host buffers: configuration + tables host->device, MC, tallies and "
                        "per-iteration energies device->host in every call; the plugin handle (device queues) persists between calls"},
        "gpu_launches": launches,
        "throughput": throughput,
        "kernel_times_ms": ktimes,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "max_energy_drift": drift, "errors": errors,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--nmc", type=int, default=None, help="iterations per step (default: the config's NMC)")
    ap.add_argument("--batch", type=int, default=None,
                    help="iterations in flight on the GPU (default: 4096 for the high-statistics configuration C5, else 1024: a C2 step "
                         "of 1000 iterations is one batch)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-throughput", action="store_true", help="skip the side figure with 4096 iterations in flight")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="N > 1: weak = NMC iterations per rank and step (default), strong = NMC iterations per step split over the "
                         "ranks (default for --config C5, the configuration BASELINE.json names for 1/2/4/8 GPUs)")
    args = ap.parse_args()
    if args.scaling is None:
        args.scaling = "strong" if args.config == "C5" else "weak"
    if args.batch is None:
        args.batch = 4096 if args.config == "C5" else 1024
    protect_stdout()
    import trekis3_b200 as tk
    if args.nmc is None:
        args.nmc = tk.CONFIGS[args.config][4]
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 0)
    if not all(os.path.exists(tk.lib_path(n)) for n in ("host", "gpu", "oracle")):
        import __graft_entry__ as g
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            g.build()
        else:
            for _ in range(600):
                if all(os.path.exists(tk.lib_path(n)) for n in ("host", "gpu", "oracle")):
                    break
                time.sleep(0.5)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
