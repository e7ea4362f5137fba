"""One model step on HOST arrays, pipelined over bands of rows on one GPU.

`Domain.physics_step_host` moves 680 MB over PCIe per step at 512x512x40 and does it strictly in sequence:
upload, kernels, download.  PCIe is full duplex and the copy engines run next to the SMs, so the step can
tend to max(upload, download) instead of their sum if the work is cut along y:

* the domain is split into `bands` contiguous bands of rows; band b is held by its own context (its own
  buffers and stream) as a stand-alone tile of the rows [first - halo, last + halo];
* microphysics is columnar, so running it on the ghost rows too gives them exactly the values the
  neighbouring band computes; MPDATA order 2 + FCT reaches three rows, and the tile's outermost row is a
  boundary-ring row the microphysics skips (jts = jms + 1, grid_obj.f90:241-250), hence halo = 4: the rows a
  band owns come out bit-identical to the undivided domain's (the same argument, and the same kernels, as the
  halo-3 slab decomposition over several GPUs, tests/test_gpu_multi.py);
* every band uploads its rows, steps, and downloads only the rows it owns (icar_physics_step_host_io).  With
  pinned host arrays (`ordered=True`, the default) one host thread QUEUES the bands one after the other
  (icar_physics_step_host_io_async) and then waits for all of them: the copy engine serves the uploads in band order,
  so band 0 computes and downloads while bands 1.. are still arriving.  (Queued from one host thread per band instead,
  the bands' copies interleave, every band's upload finishes late and little overlaps: 10.3 ms against 8.9 at
  512x512x40.)  With pageable arrays (`ordered=False`) the copies block the calling thread, so the bands run on host
  threads (ctypes releases the GIL).

Inputs and outputs must be DIFFERENT host arrays: a band reads its neighbours' rows as ghost input while they
may already be writing their results.  Callers keep two sets and swap them every step.

Reference: the loop body of step(), src/main/time_step.f90:464-549, with the halo exchange of :515-526
replaced by recomputation on the ghost rows.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import Domain, Grid, Options

# members of a state dict (icar_b200.ideal naming) by the number of y rows they have relative to ny
_PLUS_ONE_ROW = ("v", "jacobian_v")
STATIC = ("jacobian", "jacobian_u", "jacobian_v", "jacobian_w", "advection_dz", "dz_mass")


def band_bounds(ny, bands, halo):
    """[(first_owned, last_owned, first_held, last_held)] for `bands` contiguous bands of 0..ny-1."""
    bands = max(1, min(bands, ny // max(2 * halo, 1) or 1))
    base, rem = divmod(ny, bands)
    out, a = [], 0
    for b in range(bands):
        n = base + (1 if b < rem else 0)
        last = a + n - 1
        out.append((a, last, max(0, a - halo), min(ny - 1, last + halo)))
        a = last + 1
    return out


def _rows(name, array, v0, v1):
    """The rows v0..v1 of a (ny | ny+1, ...) array as a contiguous view."""
    extra = 1 if name in _PLUS_ONE_ROW else 0
    view = array[v0:v1 + 1 + extra]
    assert view.flags["C_CONTIGUOUS"]
    return view


class HostPipeline:
    """`bands` stand-alone tiles of one (nx, nz, ny) domain on one GPU, stepped concurrently from host arrays."""

    HALO = 4

    def __init__(self, state, options: Options | None = None, bands: int = 8, device: int = 0,
                 derive_diagnostics: bool = False, ordered: bool = True):
        """state: an icar_b200.ideal state dict of the WHOLE domain; its static members (jacobians, dz) are
        uploaded once, band by band."""
        self.nx, self.nz, self.ny = int(state["nx"]), int(state["nz"]), int(state["ny"])
        self.options = options or Options(dx=state["dx"])
        self.derive = bool(derive_diagnostics)
        self.ordered = bool(ordered)
        self.bounds = band_bounds(self.ny, bands, self.HALO)
        self.domains = []
        for (a, b, v0, v1) in self.bounds:
            dom = Domain(Grid.single_image(self.nx, self.nz, v1 - v0 + 1), self.options, device=device)
            for k in STATIC:
                if k in state:
                    dom.upload(k, _rows(k, state[k], v0, v1))
            self.domains.append(dom)
        self.pool = ThreadPoolExecutor(max_workers=len(self.domains))
        self._views = {}

    def _host_domains(self, host_in, host_out):
        key = (tuple((k, v.ctypes.data) for k, v in host_in.items() if isinstance(v, np.ndarray)),
               tuple((k, v.ctypes.data) for k, v in host_out.items() if isinstance(v, np.ndarray)))
        if key not in self._views:
            for k, v in host_out.items():               # checked once per pair of array sets, not per step
                if isinstance(v, np.ndarray) and isinstance(host_in.get(k), np.ndarray) and np.may_share_memory(v, host_in[k]):
                    raise ValueError(f"{k}: input and output arrays of a pipelined step must be different arrays")
            per_band = []
            for dom, (a, b, v0, v1) in zip(self.domains, self.bounds):
                vin = {k: _rows(k, v, v0, v1) for k, v in host_in.items()
                       if isinstance(v, np.ndarray) and v.ndim >= 2 and k not in STATIC
                       and not (self.derive and k in ("pii", "rho"))}
                vout = {k: _rows(k, v, v0, v1) for k, v in host_out.items() if isinstance(v, np.ndarray) and v.ndim >= 2}
                per_band.append((dom.host_domain(vin), dom.host_domain(vout), a - v0, b - v0, vin, vout))
            self._views[key] = per_band
        return self._views[key]

    def step(self, dt, host_in, host_out):
        """One step: host_in (u, v, w, pressure, [pii, rho,] scalars, accumulators) -> host_out (scalars,
        accumulators).  The two must not share memory."""
        work = self._host_domains(host_in, host_out)
        if self.ordered:
            for dom, (hin, hout, r0, r1, _, _) in zip(self.domains, work):
                dom.physics_step_host_io_async(dt, hin, hout, r0, r1, self.derive)
            for dom in self.domains:
                dom.synchronize()
            return
        futures = [self.pool.submit(dom.physics_step_host_io, dt, hin, hout, r0, r1, self.derive)
                   for dom, (hin, hout, r0, r1, _, _) in zip(self.domains, work)]
        for f in futures:
            f.result()

    @property
    def kernel_launches(self):
        return sum(d.kernel_launches for d in self.domains)

    def close(self):
        self.pool.shutdown(wait=True)
        for d in self.domains:
            d.close()
        self.domains = []
