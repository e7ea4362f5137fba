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
  cpu_baseline  the oracle (CPU restatement of the reference Fortran, "port"), rebuilt on this box with the
           reference's own compiler flags, timed on the host cores on a bounded sample (rank 0, N=1 only).
  parity   (N=1) the product against the strict oracle, in this process, outside the timed region: BASELINE config 1
           (100x100x20, 20 steps) and two steps of the benchmark state itself, per field.
  decomposition_check  (N>1) every rank re-runs two steps on the undivided domain and compares the cells it owns.

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

# torchrun exports OMP_NUM_THREADS=1; the CPU legs set their own thread count.  Read by CUDA when it initialises: the
# e2e pipeline drives several contexts from one process.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "grid_cell_updates_per_s"
UNIT = "cell-updates/s"
# bytes per cell-update, SURVEY.md section 8d: mp_simple 56 B + MPDATA S=5 72 B; Thompson 84 B + MPDATA S=9 104 B
# Algorithmic bytes per cell and launch of each kernel family (distinct fields read + written
# once, 4 B each; S = advected scalars).
KERNEL_BYTES = {
    "k_mp_simple": lambda S: 56.0, "k_mp_thompson": lambda S: 84.0, "k_wind_normalise": lambda S: 44.0,
    "k_face_coefficients": lambda S: 56.0, "k_upwind_all": lambda S: 20.0 + 8.0 * S,
    "k_corrector_fused": lambda S: 48.0 + 12.0 * S,
}
# DRAM bytes and executed instructions of one launch of each kernel SYMBOL, from the committed `ncu --set full`
# captures (scripts/ncu_kernels_json.py); used only when the symbol and the tile are the ones timed here
NCU_KERNELS = os.path.join(ROOT, "profiles", "ncu_kernels.json")
SCHEMES = {
    "simple": dict(mp=2, b_alg=128.0, scalars=("th", "qv", "qc", "qr", "qs"), acc=("precipitation", "snowfall"),
                   workload="ideal hill {nx}x{ny}x{nz}, MPDATA order2+FCT, mp_simple (BASELINE configs[1])"),
    "thompson": dict(mp=1, b_alg=188.0, scalars=("th", "qv", "qc", "qr", "qs", "qi", "qg", "ni", "nr"),
                     acc=("precipitation", "snowfall", "graupel"),
                     workload="ideal hill {nx}x{ny}x{nz}, MPDATA order2+FCT, Thompson microphysics (BASELINE {cfg})"),
}
BASELINE_CONFIGS = {(512, 512, 40): "configs[2]", (1024, 1024, 50): "configs[3]", (2048, 2048, 60): "configs[4]"}


def make_state(args, rows=None):
    """The synthetic state of the workload; rows=(j0, j1): only that band of rows (a rank of a large y-slab run)."""
    from icar_b200 import ideal
    st = ideal.make_ideal_case(nx=args.nx, nz=args.nz, ny=args.ny, perturb=True, rows=rows)
    if args.mp == "thompson":
        ideal.add_thompson_fields(st)
    for k in SCHEMES[args.mp]["acc"]:
        st[k] = np.zeros((st["ny"], args.nx), np.float64)
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
    from oracle import oracle as O0
    flags = O0.use_fast_build()                 # the restatement rebuilt on this host with the reference's own flags
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
        "config": {"workload": SCHEMES[args.mp]["workload"].format(nx=nx, ny=ny, nz=nz, cfg=BASELINE_CONFIGS.get((nx, ny, nz), "grid, not one of its configs")), "grid": [nx, nz, ny], "sample": sample,
                   "note": "CPU restatement (oracle port) of the reference Fortran; the Fortran itself "
                           "cannot be compiled in this image (no gfortran)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": thr, "kind": "port", "sample": sample, "flags": flags,
                         "cpu_model": host_cpu()[0], "host_logical_cores": host_cpu()[1]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def rel_to_max(a, b):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(float(np.abs(b).max()), 1e-30))


def parity_record(args, icar_b200, ideal, st, dt, local_rank):
    """The product against the STRICT oracle (the checker), in this process, outside the timed region."""
    from oracle import oracle as O
    O.build()
    scheme = SCHEMES[args.mp]
    rec = {"oracle": "oracle/ (C++ restatement of the reference Fortran, " + O.STRICT_FLAGS + "); pinned bit for bit to the "
                     "reference's own Fortran source executed by oracle/f90run (tests/test_reference_source.py)"}
    if args.mp == "simple":
        # BASELINE config 1: the ideal hill case 100x100x20, 20 coupled steps
        c1 = ideal.make_ideal_case(nx=100, nz=20, ny=100)
        ref = ideal.copy_state(c1)
        d1 = icar_b200.Domain(icar_b200.Grid.single_image(100, 20, 100), icar_b200.Options(dx=c1["dx"]), device=local_rank)
        d1.upload_state(c1)
        for _ in range(20):
            d1.physics_step(float(c1["dt"]))
            O.step_simple(ref, np.float32(c1["dt"]), np.float32(c1["dt"]), 2, True)
        err = {k: rel_to_max(d1.download(k), ref[k]) for k in scheme["scalars"]}
        same = {k: bool(np.array_equal(d1.download(k), ref[k])) for k in scheme["scalars"]}
        d1.close()
        rec["config1_100x100x20_20_steps"] = {"max_rel_err": err, "bitwise_equal": same, "tolerance": 1e-6,
                                              "within_tolerance": bool(max(err.values()) <= 1e-6)}
    # the benchmark state itself: two steps from the initial state
    _, step = oracle_stepper(args)
    ref = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in st.items()}
    d2 = icar_b200.Domain(icar_b200.Grid.single_image(args.nx, args.nz, args.ny),
                          icar_b200.Options(dx=st["dx"], microphysics=scheme["mp"]), device=local_rank)
    d2.upload_state(st)
    ny, nz, nx = ref["th"].shape
    scratch = (np.empty((ny, nz, nx - 1), np.float32), np.empty((ny - 1, nz, nx), np.float32), np.empty((ny, nz, nx), np.float32))
    for _ in range(2):
        d2.physics_step(dt)
        step(ref, dt, scratch)
    err = {k: rel_to_max(d2.download(k), ref[k]) for k in scheme["scalars"]}
    same = {k: bool(np.array_equal(d2.download(k), ref[k])) for k in scheme["scalars"]}
    d2.close()
    rec[f"workload_{nx}x{nz}x{ny}_2_steps"] = {"max_rel_err": err, "bitwise_equal": same}
    return rec


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
    band = bool(args.band_state and distributed and args.slabs == "y")
    if band:
        # a rank builds only the rows it holds (ghost rows included): the global state of BASELINE configs[4] would be
        # 25 GB of host memory per rank.  No undivided domain exists then, so no decomposition check and no other axis.
        g0 = icar_b200.Grid.yslab(nx, nz, ny, world, rank + 1, halo)
        st = make_state(args, rows=(g0.jms - 1, g0.jme))
        args.no_other_axis = True
    else:
        st = make_state(args)
    dt = float(st["dt"])
    options = icar_b200.Options(dx=st["dx"], microphysics=scheme["mp"])

    def make_domain(slabs):
        if distributed:
            grid = (icar_b200.Grid.xslab if slabs == "x" else icar_b200.Grid.yslab)(nx, nz, ny, world, rank + 1, halo)
            local = st if band else ideal.slab_of(st, grid)
        else:
            grid = icar_b200.Grid.single_image(nx, nz, ny, halo_size=1)
            local = st
        dom = icar_b200.Domain(grid, options, device=local_rank)
        dom.upload_state(local)
        if distributed:
            ids = [icar_b200.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            dom.comm_init(rank + 1, world, ids[0])
        return dom, grid, local

    dom, grid, local = make_domain(args.slabs)

    def barrier(d=None):
        (d or dom).synchronize()
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

    def timed_steps(d, steps, graph):
        """W untimed steps, then `steps` timed ones: CUDA events on the library's compute stream, max over ranks."""
        for _ in range(max(args.warmup, 3)):
            d.physics_step(dt)
        if graph:
            d.physics_steps(dt, 2, graph=True)             # capture outside the timed region
        barrier(d)
        l0 = d.kernel_launches
        d.timer_start()
        if graph:
            d.physics_steps(dt, steps, graph=True)         # pairs of steps replayed from a CUDA graph
        else:
            for _ in range(steps):
                d.physics_step(dt)
        ms = d.timer_stop()
        barrier(d)
        return max_over_ranks(ms) / steps, d.kernel_launches - l0

    def decomposition_check(d, g, slabs, local_state):
        """Two steps from the initial state on this rank's slab (with its neighbours) and on the undivided domain (alone,
        on this rank's GPU): the cells the rank owns must be equal bit for bit."""
        one = icar_b200.Domain(icar_b200.Grid.single_image(nx, nz, ny), options, device=local_rank)
        one.upload_state(st)
        d.upload_state(local_state)
        for _ in range(2):
            one.physics_step(dt)
            d.physics_step(dt)
        d.synchronize()
        if slabs == "x":
            lo = 0 if rank == 0 else g.its - g.ims
            hi = g.nx if rank == world - 1 else g.ite - g.ims + 1
            g0 = g.ims - 1 + lo
            own, ref = (lambda a: a[..., lo:hi]), (lambda a: a[..., g0:g0 + (hi - lo)])
        else:
            lo = 0 if rank == 0 else g.jts - g.jms
            hi = g.ny if rank == world - 1 else g.jte - g.jms + 1
            g0 = g.jms - 1 + lo
            own, ref = (lambda a: a[lo:hi]), (lambda a: a[g0:g0 + (hi - lo)])
        fields = scheme["scalars"] + scheme["acc"]
        bad = [k for k in fields if not np.array_equal(own(d.download(k)), ref(one.download(k)))]
        one.close()
        t = torch.tensor([len(bad)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return {"bitwise_equal": bool(t.item() == 0), "fields": len(fields), "steps": 2, "ranks": world,
                "mismatching_field_comparisons": int(t.item()),
                "what": "cells each rank owns vs. the same steps on the undivided domain"}

    # ---------------- optional: what update_winds does once per forcing interval (BASELINE configs[4]: "full step incl.
    # linear_winds"): linear-theory perturbation of u, v for the background flow, then balance_uvw; timed on its own
    linear_winds = None
    if args.linear_winds:
        # global terrain and layer geometry (the case is uniform in y: a band's first row stands for every row)
        g_terrain = np.broadcast_to(st["terrain"][:1], (ny, nx)) if band else st["terrain"]
        g_dzi = np.broadcast_to(st["dz_interface"][:1], (ny, nz, nx)) if band else st["dz_interface"]
        zi = np.empty((ny, nz, nx), np.float32)            # global z_interface = terrain + the layers below
        zi[:, 0, :] = g_terrain
        for k in range(1, nz):
            zi[:, k, :] = zi[:, k - 1, :] + g_dzi[:, k - 1, :]
        dom.linear_winds_init(np.ascontiguousarray(g_terrain), buffer=50)     # lt_parameters default (options_obj.f90:1452)
        dom.synchronize()
        entries = None
        if args.linear_winds == "spatial":
            dom.linear_winds_set_z(local["z"])
        dom.timer_start()
        if args.linear_winds == "spatial":                  # the reference's per-cell table semantics, entries on demand
            entries = dom.linear_perturb_spatial(icar_b200.LtOptions(linear_update_fraction=1.0), global_z_interface=zi,
                                                 global_dz_interface=g_dzi)
        else:                                               # one background state, evaluated directly
            dom.linear_perturb(10.0, 5.0, 3e-5, global_z_interface=zi, global_dz_interface=g_dzi,
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
    sampler = ClockSampler(physical_device(local_rank))
    sampler.start()
    ms_per_step, launches = timed_steps(dom, args.steps, args.graph)
    clocks = sampler.stop()
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
    fused_symbol = dom.fused_kernel_name

    # ---------------- end-to-end timing: host arrays in, host arrays out
    derive = not args.upload_diagnostics
    host = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in local.items()}
    for k in scheme["scalars"] + scheme["acc"]:
        host[k] = dom.download(k)
    if derive:
        # exner and density are taken on the device from pressure and potential temperature (icar_step_host)
        host = {k: v for k, v in host.items() if k not in ("pii", "rho")}
    hd = dom.host_domain(host)
    dom.host_register(hd)
    host_step = dom.step_host if derive else dom.physics_step_host
    in_keys = ("u", "v", "w", "pressure") + (() if derive else ("pii", "rho")) + scheme["scalars"] + scheme["acc"]
    out_keys = scheme["scalars"] + scheme["acc"]
    h2d = int(sum(host[k].nbytes for k in in_keys))
    d2h = int(sum(host[k].nbytes for k in out_keys))
    e2e_steps = max(3, min(args.steps, 10))
    e2e_api = None
    if args.e2e_bands > 1 and not distributed:
        # the same step pipelined over bands of rows (icar_b200/pipeline.py): upload of the next band, kernels
        # of this one and download of the previous one overlap.  Inputs and outputs are different host arrays.
        from icar_b200.pipeline import HostPipeline
        pipe = HostPipeline(local, options, bands=args.e2e_bands, device=local_rank, derive_diagnostics=derive)
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
        e2e_api = (f"icar_b200.pipeline.HostPipeline: icar_physics_step_host_io_async over {n_bands} bands of rows, 4 ghost "
                   "rows, one context per band, queued in band order from one host thread and awaited (pinned host arrays, "
                   "inputs and outputs swapped every step"
                   + ("; exner / density derived on the device)" if derive else ")"))
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

    # ---------------- correctness inside the driver-run artefact
    parity = decomp = other_axis = None
    if distributed and band:
        decomp = {"skipped": "--band-state: no rank holds the undivided domain at this size; the same kernels and schedule "
                             "are checked bit for bit at 512x512x40 and 1024x1024x50"}
    elif distributed:
        decomp = decomposition_check(dom, grid, args.slabs, local)
        if not args.no_other_axis:
            # the other decomposition axis once (north_star prescribes x; y keeps full-width rows and contiguous planes)
            oax = "x" if args.slabs == "y" else "y"
            dom.close()
            dom, grid, local = make_domain(oax)
            oms, _ = timed_steps(dom, max(10, args.steps // 2), args.graph)
            other_axis = {"slabs": oax, "ms_per_step": oms, "value": cells_global / (oms * 1e-3),
                          "decomposition_check": decomposition_check(dom, grid, oax, local)}
    elif rank == 0 and not args.no_parity:
        parity = parity_record(args, icar_b200, ideal, st, dt, local_rank)

    # ---------------- CPU baseline (rank 0, N=1 only): the restatement built with the reference's own flags
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        flags = O.use_fast_build()
        rate, csteps, csec, thr = cpu_oracle_rate(args, st, dt, min_seconds=10.0, max_steps=5)
        cpu = {"value": rate, "unit": UNIT, "cores": thr, "kind": "port", "flags": flags,
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

    # ---------------- what the committed ncu captures say about the kernels timed here: DRAM traffic of the dominant
    # kernel (only if the capture is of the SAME symbol on the same tile) and the issue ceiling of the step
    try:
        with open(NCU_KERNELS) as f:
            ncu = json.load(f)
    except Exception:
        ncu = {}

    def captured(family):
        want = fused_symbol if family == "k_corrector_fused" else None
        for sym, rec in ncu.items():
            if rec.get("tile_cells") != cells_local or not sym.startswith(family):
                continue
            if want is None or sym == want:
                return sym, rec
        return None, None

    traffic = traffic_src = None
    if dominant and not distributed:
        sym, rec = captured(dominant)
        if rec:
            traffic = rec["dram_bytes_per_launch"]
            traffic_src = {"capture": rec["capture"], "summary": rec.get("summary"), "symbol": sym,
                           "timed_symbol": fused_symbol if dominant == "k_corrector_fused" else dominant}
    alu = None
    if not distributed:
        per_cell, missing = {}, []
        for fam in kernels:
            sym, rec = captured(fam)
            if rec:
                per_cell[sym] = rec["lane_instructions_per_cell"]
            else:
                missing.append(fam)
        if per_cell:
            sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
            mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965
            rate = sms * 128 * mhz * 1e6
            total = sum(per_cell.values())
            alu = {"lane_instructions_per_cell": per_cell, "total_per_cell_update": total,
                   "issue_rate_lane_instructions_per_s": rate, "floor_ms_per_step": total * cells_local / rate * 1e3,
                   "frac_of_issue_ceiling": (total * cells_local / rate * 1e3) / ms_per_step,
                   "not_captured": missing,
                   "note": "executed warp instructions x 32 / cells from the committed ncu captures; 128 lanes x SMs x "
                           "SM clock.  f32x2 instructions count once here but occupy the FMA pipe for two cycles."}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": scheme["workload"].format(nx=nx, ny=ny, nz=nz, cfg=BASELINE_CONFIGS.get((nx, ny, nz), "grid, not one of its configs")), "grid": [nx, nz, ny], "dt_s": dt,
                       "advected_scalars": len(scheme["scalars"]),
                       "arithmetic": "the reference's IEEE binary32 operations in the reference's order (advection and "
                                     "mp_simple bit-identical to the oracle); there is no other mode",
                       "decomposition": (f"{args.slabs}-slabs over {world} GPUs, halo {halo}, boundary cells stored into the "
                                         "neighbours' ghost cells over NVLink (CUDA IPC), no message layer" if distributed
                                         else "single image"),
                       "l2": "resident working set ~1.1 GB per step > 126 MB L2, no explicit flush",
                       "mp_diagnostics": list(diag),
                       "stepping": "icar_physics_steps (CUDA graph replay)" if args.graph else "icar_physics_step per step"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
                    "api": e2e_api or ("icar_step_host (pinned host arrays; exner / density derived on the device)"
                                       if derive else "icar_physics_step_host (pinned host arrays)")},
            "gpu_launches": int(launches),
            # the dominant kernel, timed live with CUDA events (rank 0's tile); `traffic` = DRAM bytes of one launch
            # from the committed ncu --set full capture of the same symbol on the same tile, else null
            "roofline": ({"bound": "hbm", "kernel": dominant, "symbol": fused_symbol if dominant == "k_corrector_fused" else dominant,
                          "achieved": kernels[dominant].get("achieved_gbs"),
                          "peak": peak, "unit": "GB/s",
                          "frac": (kernels[dominant]["achieved_gbs"] / peak) if "achieved_gbs" in kernels[dominant] else None,
                          "traffic": traffic, "traffic_source": traffic_src,
                          "algorithmic_bytes_per_launch": kernels[dominant].get("algorithmic_bytes_per_cell", 0) * cells_local,
                          "avg_launch_us": kernels[dominant]["ms_per_step"] * 1e3,
                          "share_of_step": kernels[dominant]["ms_per_step"] / sum(k["ms_per_step"] for k in kernels.values()),
                          "peak_source": peak_src,
                          "note": "issue-bound, not HBM-bound: see alu_ceiling and DESIGN.md section 3"}
                         if dominant else None),
            # the BASELINE metric's own figure: whole step on the algorithmic bytes of SURVEY.md section 8d
            "step_roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                              "kernel": "whole step (all kernels of mp + advect), per GPU",
                              "algorithmic_bytes_per_cell_update": B_ALG},
            "kernels": kernels,
        }
        if alu:
            line["alu_ceiling"] = alu
        if parity:
            line["parity"] = parity
        if decomp:
            line["decomposition_check"] = decomp
        if other_axis:
            line["xslab" if other_axis["slabs"] == "x" else "yslab"] = other_axis
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
    ap.add_argument("--upload-diagnostics", action="store_true",
                    help="e2e: upload exner and density with the other host arrays (icar_physics_step_host) instead of "
                         "deriving them on the device from pressure and potential temperature (diagnostic_update, what "
                         "step()'s loop does before the physics; the default)")
    ap.add_argument("--e2e-bands", type=int, default=4,
                    help="e2e pipelined over this many bands of rows (icar_b200/pipeline.py; N=1 only); 1 = the plain host step")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison (N=1) after the timed region")
    ap.add_argument("--no-other-axis", action="store_true", help="N>1: do not also time the other decomposition axis")
    ap.add_argument("--band-state", action="store_true",
                    help="N>1, y-slabs: every rank builds only its own rows of the synthetic state (BASELINE configs[4], "
                         "2048x2048x60: the global state would be 25 GB of host memory per rank); skips the decomposition check")
    ap.add_argument("--linear-winds", nargs="?", const="direct", default=None, choices=["direct", "spatial"],
                    help="before the steps, run the once-per-forcing-interval wind update (icar_linear_perturb + "
                         "icar_balance_uvw) and report its time under `linear_winds`")
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
