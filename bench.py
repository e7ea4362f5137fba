#!/usr/bin/env python
"""bench.py -- grid-cell-updates/s of the mp + advect model step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (time_step.f90:512-529: mp(halo) -> halo_send ->
mp(subset) -> halo_retrieve -> advect) over the synthetic 512x512x40 ideal hill case
(BASELINE.json configs[1]: MPDATA order 2 + FCT, mp_simple).  One cell-update = one (i,k,j)
cell advanced one step through the microphysics and the MPDATA of all 5 advected scalars.

  value    device-resident throughput (fields already in HBM), CUDA events on the library's
           compute stream, max over ranks.
  e2e      the same step through icar_physics_step_host: HOST arrays in (pinned), host arrays
           out, H2D/D2H inside the timed region.
  roofline ALGORITHMIC bytes (128 B per cell-update for mp_simple +
           MPDATA x 5 scalars) / measured time, against the measured HBM copy peak.
  cpu_baseline  the oracle (CPU restatement of the reference Fortran, "port") timed on this
           box's host cores on a bounded sample (rank 0, N=1 only).

--impl reference times that CPU restatement with all host threads instead (the reference
Fortran itself cannot be built: no Fortran compiler in the image).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "grid_cell_updates_per_s"
UNIT = "cell-updates/s"
# bytes per cell-update, SURVEY.md section 8d: mp_simple 56 B + MPDATA S=5 72 B; Thompson 84 B + MPDATA S=9 104 B
# Algorithmic bytes per cell and launch of each kernel family (distinct fields read + written
# once, 4 B each; S = advected scalars) and the DRAM traffic one launch moved at 512x512x40 in the committed
# `ncu --set full` captures (profiles/r1_v11_top_kernels.txt, profiles/r1_thompson_v3.txt), in bytes.
KERNEL_BYTES = {
    "k_mp_simple": lambda S: 56.0, "k_mp_thompson": lambda S: 84.0, "k_wind_normalise": lambda S: 44.0,
    "k_face_coefficients": lambda S: 64.0, "k_upwind_all": lambda S: 20.0 + 8.0 * S,
    "k_corrector_fused": lambda S: 44.0 + 12.0 * S,
}
NCU_TRAFFIC_512 = {("k_corrector_fused", 5): 1142.5e6, ("k_mp_simple", 5): 620.0e6, ("k_upwind_all", 5): 620.0e6,
                   # Thompson: 0.68 GB of field reads + 11.4 GB of writes, almost all of it register spills around
                   # the out-of-line math calls (profiles/r1_thompson_v3.txt)
                   ("k_mp_thompson", 9): 12056.0e6}
SCHEMES = {
    "simple": dict(mp=2, b_alg=128.0, scalars=("th", "qv", "qc", "qr", "qs"), acc=("precipitation", "snowfall"),
                   workload="ideal hill {nx}x{ny}x{nz}, MPDATA order2+FCT, mp_simple (BASELINE configs[1])"),
    "thompson": dict(mp=1, b_alg=188.0, scalars=("th", "qv", "qc", "qr", "qs", "qi", "qg", "ni", "nr"),
                     acc=("precipitation", "snowfall", "graupel"),
                     workload="ideal hill {nx}x{ny}x{nz}, MPDATA order2+FCT, Thompson microphysics (BASELINE configs[2])"),
}


def make_state(args):
    from icar_b200 import ideal
    st = ideal.make_ideal_case(nx=args.nx, nz=args.nz, ny=args.ny, perturb=True)
    if args.mp == "thompson":
        ideal.add_thompson_fields(st)
    for k in SCHEMES[args.mp]["acc"]:
        st[k] = np.zeros((args.ny, args.nx), np.float64)
    return st


_THOMPSON = None


def host_cpu():
    """(model name, logical cores) of the box the CPU baseline runs on."""
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    model = ln.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return model, os.cpu_count() or 1


def oracle_stepper(args):
    """step(state, dt, scratch) -> scratch for the scheme's CPU restatement."""
    from oracle import oracle as O
    O.build()
    if args.mp == "thompson":
        global _THOMPSON
        if _THOMPSON is None:
            _THOMPSON = O.Thompson(big_tables=True)      # the table build takes seconds: once per process
        T = _THOMPSON
        return O, lambda w, dt, scr: T.step(w, np.float32(dt), np.float32(dt), 2, True, scr)
    return O, lambda w, dt, scr: O.step_simple(w, np.float32(dt), np.float32(dt), 2, True, scr)


def measured_peak():
    """HBM copy peak from MEASURED_PEAKS.json (driver-written), else the recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def _reasons(self):
        nv = self.nv
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            try:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            except Exception:
                return
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        for n, bit in names.items():
            if mask & bit:
                self.reasons.add(n)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self._reasons()
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        if self.ok:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread:
            self._stop.set()
            self._thread.join()
        return {
            "sm_mhz": int(statistics.median(self.samples)) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


def physical_device(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# --------------------------------------------------------------------------------------------
def cpu_oracle_rate(args, st, dt, min_seconds, max_steps, threads=None):
    """Times the oracle step on the full workload; returns (cells/s, steps, seconds, threads)."""
    O, step = oracle_stepper(args)
    if threads:
        O.set_threads(threads)
    thr = O.max_threads()
    work = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in st.items()}
    cells = work["th"].size
    ny, nz, nx = work["th"].shape
    scratch = (np.empty((ny, nz, nx - 1), np.float32), np.empty((ny - 1, nz, nx), np.float32),
               np.empty((ny, nz, nx), np.float32))
    steps, t0 = 0, time.perf_counter()
    while steps < max_steps:
        step(work, dt, scratch)
        steps += 1
        if time.perf_counter() - t0 >= min_seconds:
            break
    el = time.perf_counter() - t0
    return cells * steps / el, steps, el, thr


def cpu_oracle_rate_one_thread(args, st, dt, rows=48):
    """One host thread (one coarray image, where the reference's advection is serial) on a
    y-band of the workload; returns (cells/s, rows, seconds)."""
    O, step = oracle_stepper(args)
    all_threads = O.max_threads()
    rows = min(rows, st["th"].shape[0])
    band = {k: (np.ascontiguousarray(v[:rows + (1 if k in ("v", "jacobian_v") else 0)])
                if isinstance(v, np.ndarray) and v.ndim >= 2 else v) for k, v in st.items()}
    _, nz, nx = band["th"].shape
    scratch = (np.empty((rows, nz, nx - 1), np.float32), np.empty((rows - 1, nz, nx), np.float32),
               np.empty((rows, nz, nx), np.float32))
    O.set_threads(1)
    try:
        step(band, dt, scratch)                      # page in
        t0 = time.perf_counter()
        step(band, dt, scratch)
        el = time.perf_counter() - t0
    finally:
        O.set_threads(all_threads)
    return band["th"].size / el, rows, el


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference Fortran on all host threads."""
    if rank != 0:
        return
    nx, nz, ny = args.nx, args.nz, args.ny
    # bounded sample: a y-band of the workload sized so that K+W steps end within minutes
    st = make_state(args)
    dt = float(st["dt"])
    O, step = oracle_stepper(args)
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its ranks, which would
    # time the reference on one core
    try:
        O.set_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        O.set_threads(os.cpu_count() or 1)
    thr = O.max_threads()

    def scratch_for(b):
        n = b["th"].shape[0]
        return (np.empty((n, nz, nx - 1), np.float32), np.empty((n - 1, nz, nx), np.float32),
                np.empty((n, nz, nx), np.float32))

    probe_ny = min(ny, 32)
    band = {k: (np.ascontiguousarray(v[:probe_ny + (1 if k in ("v", "jacobian_v") else 0)])
                if isinstance(v, np.ndarray) and v.ndim >= 2 else v) for k, v in st.items()}
    t0 = time.perf_counter()
    step(band, dt, scratch_for(band))
    per_row = (time.perf_counter() - t0) / probe_ny
    budget = 150.0 / max(1, args.steps + args.warmup)          # seconds per step
    sample_ny = int(max(8, min(ny, budget / max(per_row, 1e-9))))
    band = {k: (np.ascontiguousarray(v[:sample_ny + (1 if k in ("v", "jacobian_v") else 0)])
                if isinstance(v, np.ndarray) and v.ndim >= 2 else v) for k, v in st.items()}
    cells = nx * nz * sample_ny
    scratch = scratch_for(band)
    for _ in range(args.warmup):
        step(band, dt, scratch)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(band, dt, scratch)
    el = time.perf_counter() - t0
    value = cells * args.steps / el
    sample = f"{args.steps} steps of a {nx}x{nz}x{sample_ny} y-band of the workload"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": SCHEMES[args.mp]["workload"].format(nx=nx, ny=ny, nz=nz), "grid": [nx, nz, ny], "sample": sample,
                   "note": "CPU restatement (oracle port) of the reference Fortran; the Fortran itself "
                           "cannot be compiled in this image (no gfortran)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": thr, "kind": "port", "sample": sample,
                         "cpu_model": host_cpu()[0], "host_logical_cores": host_cpu()[1]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import icar_b200
    from icar_b200 import ideal

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    nx, nz, ny = args.nx, args.nz, args.ny
    halo = args.halo
    scheme = SCHEMES[args.mp]
    st = make_state(args)
    dt = float(st["dt"])
    if distributed:
        grid = (icar_b200.Grid.xslab if args.slabs == "x" else icar_b200.Grid.yslab)(nx, nz, ny, world, rank + 1, halo)
        local = ideal.slab_of(st, grid)
    else:
        grid = icar_b200.Grid.single_image(nx, nz, ny, halo_size=1)
        local = st
    dom = icar_b200.Domain(grid, icar_b200.Options(dx=st["dx"], microphysics=scheme["mp"]), device=local_rank)
    dom.upload_state(local)
    if distributed:
        ids = [icar_b200.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        dom.comm_init(rank + 1, world, ids[0])

    def barrier():
        dom.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cells_global = nx * nz * ny

    # ---------------- optional: what update_winds does once per forcing interval (BASELINE configs[4]: "full step incl.
    # linear_winds"): linear-theory perturbation of u, v for the background flow, then balance_uvw; timed on its own
    linear_winds = None
    if args.linear_winds:
        zi = np.empty_like(st["dz_interface"])             # global z_interface = terrain + the layers below
        zi[:, 0, :] = st["terrain"]
        for k in range(1, nz):
            zi[:, k, :] = zi[:, k - 1, :] + st["dz_interface"][:, k - 1, :]
        dom.linear_winds_init(st["terrain"], buffer=50)     # lt_parameters default (options_obj.f90:1452)
        dom.synchronize()
        entries = None
        if args.linear_winds == "spatial":
            dom.linear_winds_set_z(local["z"])
        dom.timer_start()
        if args.linear_winds == "spatial":                  # the reference's per-cell table semantics, entries on demand
            entries = dom.linear_perturb_spatial(icar_b200.LtOptions(linear_update_fraction=1.0), global_z_interface=zi,
                                                 global_dz_interface=st["dz_interface"])
        else:                                               # one background state, evaluated directly
            dom.linear_perturb(10.0, 5.0, 3e-5, global_z_interface=zi, global_dz_interface=st["dz_interface"],
                               linear_update_fraction=1.0)
        dom.balance_uvw()
        lw_ms = max_over_ranks(dom.timer_stop())
        dt = float(dom.compute_dt(st["dz_levels"]))          # the winds changed: update_dt (time_step.f90:375-423)
        fnx, fny, _ = dom.linear_winds_dims()
        linear_winds = {"ms": lw_ms, "fft_grid": [fnx, fny], "levels": nz, "dt_s_after": dt,
                        "table_entries_evaluated": entries,
                        "api": ("icar_linear_perturb_spatial (spatial_winds: variable N, table look-up per cell, entries on demand)"
                                if args.linear_winds == "spatial" else
                                "icar_linear_perturb (space_varying_dz, N^2 = 3e-5, direct evaluation)") +
                               ", cuFFT Z2Z, + icar_balance_uvw; once per forcing interval, not inside the timed steps"}

    # ---------------- device-resident timing
    for _ in range(max(args.warmup, 3)):
        dom.physics_step(dt)
    if args.graph:
        dom.physics_steps(dt, 2, graph=True)               # capture outside the timed region
    barrier()
    sampler = ClockSampler(physical_device(local_rank))
    launches0 = dom.kernel_launches
    sampler.start()
    dom.timer_start()
    if args.graph:
        dom.physics_steps(dt, args.steps, graph=True)      # pairs of steps replayed from a CUDA graph
    else:
        for _ in range(args.steps):
            dom.physics_step(dt)
    ms = dom.timer_stop()
    barrier()
    clocks = sampler.stop()
    launches = dom.kernel_launches - launches0
    ms = max_over_ranks(ms)
    ms_per_step = ms / args.steps
    value = cells_global / (ms_per_step * 1e-3)

    # ---------------- per-kernel timing (CUDA events around every launch, on the launching stream) for the
    # roofline of the dominant kernel; a separate pass of the same steps so that the events do not sit
    # inside the timed region above
    dom.profile_enable(True)
    for _ in range(args.steps):
        dom.physics_step(dt)
    prof = dom.profile_read()
    dom.profile_enable(False)
    barrier()

    # ---------------- end-to-end timing: host arrays in, host arrays out
    host = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in local.items()}
    for k in scheme["scalars"] + scheme["acc"]:
        host[k] = dom.download(k)
    if args.derive_diagnostics:
        # exner and density are taken on the device from pressure and potential temperature (icar_step_host)
        host = {k: v for k, v in host.items() if k not in ("pii", "rho")}
    hd = dom.host_domain(host)
    dom.host_register(hd)
    host_step = dom.step_host if args.derive_diagnostics else dom.physics_step_host
    in_keys = ("u", "v", "w", "pressure") + (() if args.derive_diagnostics else ("pii", "rho")) + scheme["scalars"] + scheme["acc"]
    out_keys = scheme["scalars"] + scheme["acc"]
    h2d = int(sum(host[k].nbytes for k in in_keys))
    d2h = int(sum(host[k].nbytes for k in out_keys))
    e2e_steps = max(3, min(args.steps, 10))
    e2e_api = None
    if args.e2e_bands > 1 and not distributed:
        # the same step pipelined over bands of rows (icar_b200/pipeline.py): upload of the next band, kernels
        # of this one and download of the previous one overlap.  Inputs and outputs are different host arrays.
        from icar_b200.pipeline import HostPipeline
        pipe = HostPipeline(local, icar_b200.Options(dx=st["dx"], microphysics=scheme["mp"]),
                            bands=args.e2e_bands, device=local_rank, derive_diagnostics=args.derive_diagnostics)
        other = {k: np.ascontiguousarray(host[k].copy()) for k in out_keys}
        hd_other = dom.host_domain(other)
        dom.host_register(hd_other)
        sets = [host, dict(host, **other)]
        n_bands = len(pipe.bounds)
        held = sum(v1 - v0 + 1 for (_, _, v0, v1) in pipe.bounds)
        h2d = int(h2d * held / ny)              # ghost rows are uploaded by both neighbours

        def host_step(dt_, hd_, _n=[0]):
            src, dst = sets[_n[0] % 2], sets[(_n[0] + 1) % 2]
            pipe.step(dt_, src, {k: dst[k] for k in out_keys})
            _n[0] += 1
        e2e_api = (f"icar_physics_step_host_io over {n_bands} bands of rows, 4 ghost rows, one context and host thread "
                   "per band (icar_b200.pipeline.HostPipeline; pinned host arrays, inputs and outputs swapped every step)")
    for _ in range(2):
        host_step(dt, hd)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_step(dt, hd)                      # returns after the D2H copies completed
    e2e_s = time.perf_counter() - t0
    barrier()
    dom.host_unregister(hd)
    if e2e_api:
        dom.host_unregister(hd_other)
        pipe.close()
    e2e_s = max_over_ranks(e2e_s)
    e2e_value = cells_global / (e2e_s / e2e_steps)
    if distributed:
        t = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        h2d, d2h = int(t[0].item()), int(t[1].item())
    diag = dom.mp_diagnostics()

    # ---------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, csteps, csec, thr = cpu_oracle_rate(args, st, dt, min_seconds=10.0, max_steps=5)
        cpu = {"value": rate, "unit": UNIT, "cores": thr, "kind": "port",
               "sample": f"{csteps} full step(s) of the same {nx}x{nz}x{ny} workload, {csec:.1f} s, "
                         f"OpenMP over j with {thr} threads (oracle/: C++ restatement of the Fortran)"}
        cpu["cpu_model"], cpu["host_logical_cores"] = host_cpu()
        r1, rows1, sec1 = cpu_oracle_rate_one_thread(args, st, dt)
        cpu["one_thread"] = {"value": r1, "unit": UNIT, "cores": 1,
                             "sample": f"one step of a {nx}x{nz}x{rows1} y-band of the workload, {sec1:.1f} s "
                                       "(one coarray image of the reference)"}

    peak, peak_src = measured_peak()
    B_ALG = scheme["b_alg"]
    achieved = value * B_ALG / 1e9 / world          # GB/s per GPU on algorithmic bytes
    S = len(scheme["scalars"])
    cells_local = grid.nx * grid.nz * grid.ny
    kernels = {}
    for name, (kms, kn) in prof.items():
        e = {"ms_per_step": kms / args.steps, "launches_per_step": kn / args.steps, "avg_us": kms / max(kn, 1) * 1e3}
        if name in KERNEL_BYTES:
            # one launch of a family covers the tile once (mp strips + interior together do)
            per_step_bytes = KERNEL_BYTES[name](S) * cells_local
            e["algorithmic_bytes_per_cell"] = KERNEL_BYTES[name](S)
            e["achieved_gbs"] = per_step_bytes / (kms / args.steps * 1e-3) / 1e9
        kernels[name] = e
    dominant = max(kernels, key=lambda k: kernels[k]["ms_per_step"]) if kernels else None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": scheme["workload"].format(nx=nx, ny=ny, nz=nz), "grid": [nx, nz, ny], "dt_s": dt,
                       "advected_scalars": len(scheme["scalars"]),
                       "decomposition": (f"{args.slabs}-slabs over {world} GPUs, halo {halo}, NCCL send/recv" if distributed
                                         else "single image"),
                       "l2": "resident working set ~1.1 GB per step > 126 MB L2, no explicit flush",
                       "mp_diagnostics": list(diag),
                       "stepping": "icar_physics_steps (CUDA graph replay)" if args.graph else "icar_physics_step per step"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
                    "api": e2e_api or ("icar_step_host (pinned host arrays; exner / density derived on the device)"
                                       if args.derive_diagnostics else "icar_physics_step_host (pinned host arrays)")},
            "gpu_launches": int(launches),
            # the dominant kernel, timed live with CUDA events (rank 0's tile); `traffic` = DRAM bytes of one
            # launch from the committed ncu --set full capture (only known for the default 512x512x40 tile)
            "roofline": ({"bound": "hbm", "kernel": dominant, "achieved": kernels[dominant].get("achieved_gbs"),
                          "peak": peak, "unit": "GB/s",
                          "frac": (kernels[dominant]["achieved_gbs"] / peak) if "achieved_gbs" in kernels[dominant] else None,
                          "traffic": (NCU_TRAFFIC_512.get((dominant, S)) if (grid.nx, grid.nz, grid.ny) == (512, 40, 512) else None),
                          "algorithmic_bytes_per_launch": kernels[dominant].get("algorithmic_bytes_per_cell", 0) * cells_local,
                          "avg_launch_us": kernels[dominant]["ms_per_step"] * 1e3,
                          "share_of_step": kernels[dominant]["ms_per_step"] / sum(k["ms_per_step"] for k in kernels.values()),
                          "peak_source": peak_src,
                          "note": "issue-bound, not HBM-bound: see DESIGN.md section 3 (instruction counts per cell)"}
                         if dominant else None),
            # the BASELINE metric's own figure: whole step on the algorithmic bytes of SURVEY.md section 8d
            "step_roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                              "kernel": "whole step (all kernels of mp + advect), per GPU",
                              "algorithmic_bytes_per_cell_update": B_ALG},
            "kernels": kernels,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if linear_winds:
            line["linear_winds"] = linear_winds
        print(json.dumps(line), flush=True)
    dom.close()
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=512)
    ap.add_argument("--nz", type=int, default=40)
    ap.add_argument("--ny", type=int, default=512)
    ap.add_argument("--halo", type=int, default=3, help="ghost width for N>1 (3 = decomposition-invariant)")
    ap.add_argument("--slabs", default="y", choices=["x", "y"],
                    help="decomposition axis for N>1 (y keeps full-width rows and contiguous halo planes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--derive-diagnostics", action="store_true",
                    help="e2e through icar_step_host: exner and density are not uploaded but derived on the device "
                         "(diagnostic_update), as step()'s loop does before the physics")
    ap.add_argument("--e2e-bands", type=int, default=1,
                    help="e2e pipelined over this many bands of rows (icar_b200/pipeline.py); 1 = the plain host step")
    ap.add_argument("--linear-winds", nargs="?", const="direct", default=None, choices=["direct", "spatial"],
                    help="before the steps, run the once-per-forcing-interval wind update (icar_linear_perturb + "
                         "icar_balance_uvw) and report its time under `linear_winds` (not yet run on hardware)")
    ap.add_argument("--graph", action="store_true",
                    help="run the timed steps through icar_physics_steps (CUDA-graph replay) instead of one "
                         "icar_physics_step call per step")
    ap.add_argument("--mp", default="simple", choices=["simple", "thompson"],
                    help="microphysics scheme: simple = BASELINE configs[1] (the headline), thompson = configs[2]")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
