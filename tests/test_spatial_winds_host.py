"""spatial_winds (linear_winds.f90:840-1127) without a GPU: the NumPy restatement oracle/spatial_winds_np.py on its
own terms, and the kernels of icar_b200/csrc/lw_kernels.cuh -- run on the CPU by tests/cpp/lw_kernels_host.cpp, with
the host logic of lw_spatial_winds (icar_b200/csrc/linear_winds.cu) restated in Python and numpy.fft in cuFFT's
place -- against it.  The device run is tests/test_gpu_linear_winds.py::test_spatial_winds_*."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from icar_b200 import ideal
from lw_expected import z_interface_of
from oracle import linear_winds_np as LW
from oracle import spatial_winds_np as SW
from test_linear_winds_host import EmulatedLinearWinds, cptr, emu, fptr          # noqa: F401  (emu is a fixture)

f = np.float32


def iptr(a): return a.ctypes.data_as(C.POINTER(C.c_int))


def varied_state(nx=40, nz=7, ny=24, moist=True):
    """The hill case with a background wind that crosses several nodes of the speed and direction axes, and some
    cloud water aloft so that calc_stability takes both of its branches."""
    st = ideal.make_ideal_case(nx=nx, nz=nz, ny=ny, perturb=True)
    i = np.arange(nx + 1, dtype=np.float64)[None, None, :]
    j = np.arange(ny + 1, dtype=np.float64)[:, None, None]
    st["u"] = (st["u"] * (1 + 0.5 * np.sin(2 * np.pi * i / nx)) * (1 + 0.1 * np.cos(2 * np.pi * j[:ny] / ny))).astype(f)
    st["v"] = ((5.0 * (1 + 0.8 * np.cos(2 * np.pi * j / ny))) * np.ones((1, nz, nx)) * (1 - 1.5 * (i[:, :, :nx] > 30))).astype(f)
    if moist:
        st["qc"] = np.zeros_like(st["qv"])
        st["qc"][:, 3:5, 10:25] = f(2e-4)
    return st


def run_oracle(st, o, buffer_terrain_dx, varying, calls=1, tile=None):
    terrain_xy = np.ascontiguousarray(st["terrain"].T)
    zi = z_interface_of(st)
    lh = (np.cumsum(st["dz_levels"], dtype=f) - st["dz_levels"] / f(2)).astype(f)

    def bounds(z):
        if varying:
            zb = (zi[:, z, :] - st["terrain"]).T
            return zb, (zi[:, z, :] - st["terrain"] + st["dz_interface"][:, z, :]).T
        half = st["dz_levels"][z] / f(2)
        return lh[z] - half, lh[z] + half
    cache = SW.EntryCache(terrain_xy, st["dx"], o, bounds)
    if tile is None:
        loc = {k: st[k].copy() for k in ("u", "v", "th", "pii", "z", "qv", "qc", "qr", "qs") if k in st}
        origin = (0, 0)
    else:
        i0, j0, nx, ny = tile
        sl = lambda a, ex=0, ey=0: np.ascontiguousarray(a[j0:j0 + ny + ey, :, i0:i0 + nx + ex])
        loc = {k: sl(st[k]) for k in ("th", "pii", "z", "qv", "qc", "qr", "qs") if k in st}
        loc["u"], loc["v"] = sl(st["u"], ex=1), sl(st["v"], ey=1)
        origin = (i0, j0)
    up, vp = np.zeros_like(loc["u"]), np.zeros_like(loc["v"])
    for _ in range(calls):
        nsq = SW.spatial_winds(loc, cache, o, up, vp, tile=origin)
    return loc["u"], loc["v"], nsq, cache, lh, zi


# ------------------------------------------------------------------------------------------------ the restatement

def test_table_axes_and_positions():
    o = SW.LTOptions()
    dirs, nsqs, spds = SW.table_axes(o)
    assert len(dirs) == 24 and len(nsqs) == 5 and len(spds) == 6
    np.testing.assert_allclose(spds, [0, 6, 12, 18, 24, 30])
    np.testing.assert_allclose(np.exp(nsqs[[0, -1]]), [1e-7, 6e-4], rtol=1e-5)
    pos, nxt, w = SW.positions(spds, np.array([0.0, 3.0, 6.0, 29.0, 30.0, 45.0, -1.0], f))
    assert pos.tolist() == [0, 0, 0, 4, 4, 5, 0] and nxt.tolist() == [1, 1, 1, 5, 5, 5, 0]
    np.testing.assert_allclose(w, [1, 0.5, 0, 1 / 6, 0, 1, 1], atol=1e-6)
    # directions: wind FROM the south-west quadrant etc., measured clockwise from +v
    d = SW.calc_direction(np.array([0, 1, 1, 0, -1, -1], f), np.array([1, 1, 0, -1, -1, 1], f))
    np.testing.assert_allclose(d, np.array([0, 0.25, 0.5, 1.0, 1.25, 1.75]) * np.pi, atol=1e-6)


def test_stability_dry_and_moist():
    th = np.array([280.0, 300.0], f)
    n2 = SW.calc_stability(th[:1], th[1:], f(1), f(1), f(1000), f(3000), f(0.003), f(0.002), np.array([0.0], f))
    np.testing.assert_allclose(n2, 9.81 * np.log(300 / 280) / 2000, rtol=1e-5)
    moist = SW.calc_stability(th[:1], th[1:], f(0.95), f(0.85), f(1000), f(3000), f(0.003), f(0.002), np.array([1e-4], f))
    assert 0 < moist[0] < n2[0]                                             # latent heating weakens the stability


def test_box_smoothing_repeats_the_edges():
    a = np.zeros((6, 1, 15), f)
    a[:, 0, :] = np.arange(15, dtype=f)[None, :]
    s = SW.smooth_array_ydim3(a, 2)
    np.testing.assert_allclose(s[3, 0, 2:13], np.arange(2, 13), rtol=1e-6)  # linear ramp: unchanged in the interior
    np.testing.assert_allclose(s[3, 0, 0], (0 + 0 + 0 + 1 + 2) / 5, rtol=1e-6)
    np.testing.assert_allclose(s[3, 0, 14], (12 + 13 + 14 + 14 + 14) / 5, rtol=1e-6)


def test_on_a_table_node_the_interpolation_returns_the_entry():
    """A uniform flow that sits exactly on a node of all three axes: one entry carries all the weight, so the
    perturbation is that of the direct evaluation (icar_linear_perturb) for the node's (U, V, N^2)."""
    st = ideal.make_ideal_case(nx=40, nz=7, ny=24)
    o = SW.LTOptions(variable_N=False, smooth_nsq=False, buffer=7, linear_update_fraction=f(1), N_squared=f(np.exp(SW.table_axes(SW.LTOptions())[1][2])))
    dirs, nsqs, spds = SW.table_axes(o)
    U, V = (np.sin(dirs[3]) * spds[2]).astype(f), (np.cos(dirs[3]) * spds[2]).astype(f)
    st["u"][:] = U; st["v"][:] = V
    u, v, nsq, cache, lh, zi = run_oracle(st, o, None, varying=False)
    from lw_expected import expected_winds
    want_u, want_v, _, _ = expected_winds(st, 7, U, V, np.exp(nsqs[2]).astype(f), 1.0, 1.0, False, calls=1)
    # the column-mean wind is the node's up to REAL rounding, so a neighbour can carry ~1e-6 of the weight
    assert np.abs(u - want_u).max() <= 1e-4 and np.abs(v[:-1] - want_v[:-1]).max() <= 1e-4
    assert np.abs(u - st["u"]).max() > 0.05
    # the extra v row (k = ny + 1) is looked up with the column means of the ALREADY UPDATED u of row ny
    # (u1d is taken inside `do k = 1, nyv`, :961-973): it is off the node, so it mixes in neighbouring entries
    assert np.abs(v[-1] - want_v[-1]).max() > 1e-2


# ------------------------------------------------------------------------------------- the kernels, emulated

class EmulatedSpatialWinds(EmulatedLinearWinds):
    """lw_spatial_winds with emulated kernels; buffers are C-order images of the device arrays."""

    def spatial(self, o, g, fields, u_dev, v_dev, bounds, varying):
        emu = self.emu
        nx, nz, ny, ld, ldu, i0, j0 = g
        gnx, gny = self.nx, self.ny
        if self.upert is None:
            self.upert, self.vpert = np.zeros_like(u_dev), np.zeros_like(v_dev)
        dirs, nsqs, spds = SW.table_axes(o)
        n = np.array([len(dirs), len(nsqs), len(spds)], np.int32)
        oi = np.array([o.variable_N, o.smooth_nsq, o.vert_smooth, o.stability_window_size], np.int32)
        of = np.array([o.min_stability, o.max_stability, o.N_squared, o.linear_contribution, o.linear_update_fraction], f)
        nsq = np.zeros((ny, nz, ld), f)
        P = lambda k: fptr(fields[k]) if fields.get(k) is not None else None
        emu.emu_nsq_column(nx, nz, ny, ld, P("th"), P("pii"), P("z"), P("qv"), P("qc"), P("qi"), P("qr"), P("qs"), iptr(oi), fptr(of), fptr(nsq))
        if o.smooth_nsq:
            rowmeans = np.zeros((ny, nz, ld), np.float64)
            emu.emu_smooth(fptr(nsq), cptr(rowmeans), nx, nz, ny, ld, o.stability_window_size)
        self.nsq = nsq
        n_entries = int(n.prod())
        words = (n_entries + 31) // 32
        u1d = np.zeros((ny + 1, nx + 1), f); v1d = np.zeros_like(u1d)
        stride_u, stride_v = (gnx + 1) * gny, gnx * (gny + 1)
        self.evaluated = 0
        for phase in range(2):
            row0, nrows = (0, ny) if phase == 0 else (ny, 1)
            emu.emu_colmean(fptr(u_dev), fptr(v_dev), nx, nz, ny, ld, ldu, row0, nrows, fptr(u1d), fptr(v1d))
            for j in range(nz):
                bitmap = np.zeros(words, np.uint32)
                emu.emu_mark_entries(iptr(n), fptr(dirs), fptr(nsqs), fptr(spds), fptr(u1d), fptr(v1d), fptr(nsq), nx, nz, ny, ld, j,
                                     o.stability_window_size, row0, nrows, cptr(bitmap))
                used = [e for e in range(n_entries) if (bitmap[e >> 5] >> (e & 31)) & 1]
                slot_of = np.zeros(n_entries, np.int32)
                planes_u = np.zeros((len(used), stride_u), f); planes_v = np.zeros((len(used), stride_v), f)
                for s, e in enumerate(used):
                    slot_of[e] = s
                    i_n, i_s, i_d = e % n[1], (e // n[1]) % n[2], e // (n[1] * n[2])
                    U, V = (np.sin(dirs[i_d]) * spds[i_s]).astype(f), (np.cos(dirs[i_d]) * spds[i_s]).astype(f)
                    N2 = np.exp(nsqs[i_n]).astype(f)
                    zb, zt = bounds(j)
                    if varying:
                        self.varyingz(U, V, N2, zb, zt, o.minimum_layer_size)
                    else:
                        self.constz(U, V, N2, zb, zt, o.minimum_layer_size)
                    if U == 0 and V == 0:
                        self.up = np.zeros((self.fny, self.fnx), np.complex128); self.vp = self.up.copy()
                    emu.emu_stagger(cptr(self.up), cptr(self.vp), self.fnx, self.buffer, gnx, gny, fptr(planes_u[s]), fptr(planes_v[s]))
                    self.evaluated += 1
                args = (iptr(n), fptr(dirs), fptr(nsqs), fptr(spds), iptr(oi), fptr(of), fptr(u1d), fptr(v1d), fptr(nsq), nx, nz, ny, ld)
                if phase == 0:
                    emu.emu_interp(*args, ldu, j, 0, ny, 0, iptr(slot_of), fptr(planes_u), C.c_longlong(stride_u), gnx, i0, j0,
                                   fptr(u_dev), fptr(self.upert))
                emu.emu_interp(*args, ld, j, row0, nrows, 1, iptr(slot_of), fptr(planes_v), C.c_longlong(stride_v), gnx, i0, j0,
                               fptr(v_dev), fptr(self.vpert))


def padded(a, pitch, fill=1e30):
    out = np.full(a.shape[:-1] + (pitch,), f(fill))
    out[..., :a.shape[-1]] = a
    return out


@pytest.mark.parametrize("varying,tile,variable_N", [(False, None, True), (True, None, True), (False, (9, 5, 21, 13), True),
                                                      (False, None, False)])
def test_spatial_winds_with_emulated_kernels(emu, varying, tile, variable_N):
    st = varied_state()
    o = SW.LTOptions(buffer=7, variable_N=variable_N)
    calls = 1 if varying else 2            # (terrain-following bounds: many heights per entry; the relaxation memory is exercised by the others)
    want_u, want_v, want_nsq, cache, lh, zi = run_oracle(st, o, None, varying, calls=calls, tile=tile)
    nxg, nyg, nz = st["nx"], st["ny"], st["nz"]
    i0, j0, nx, ny = tile or (0, 0, nxg, nyg)
    ld, ldu = -(-nx // 32) * 32, -(-(nx + 1) // 32) * 32
    sl = lambda a, ex=0, ey=0: np.ascontiguousarray(a[j0:j0 + ny + ey, :, i0:i0 + nx + ex])
    fields = {k: padded(sl(st[k]), ld, fill=0.0) for k in ("th", "pii", "z", "qv", "qc", "qr", "qs")}
    u0, v0 = sl(st["u"], ex=1), sl(st["v"], ey=1)
    u_dev, v_dev = padded(u0, ldu), padded(v0, ld)

    def bounds(z):
        if varying:
            zb = zi[:, z, :] - st["terrain"]
            return zb, zb + st["dz_interface"][:, z, :]
        half = st["dz_levels"][z] / f(2)
        return lh[z] - half, lh[z] + half
    lw = EmulatedSpatialWinds(emu, st["terrain"], o.buffer, st["dx"])
    for _ in range(calls):
        lw.spatial(o, (nx, nz, ny, ld, ldu, i0, j0), fields, u_dev, v_dev, bounds, varying)
    assert (u_dev[..., nx + 1:] == f(1e30)).all() and (v_dev[..., nx:] == f(1e30)).all()
    np.testing.assert_allclose(lw.nsq[..., :nx], want_nsq, rtol=2e-6)            # log N^2 after both smoothings
    du, dv = u_dev[..., :nx + 1] - u0, v_dev[..., :nx] - v0
    wu, wv = want_u - u0, want_v - v0
    assert max(np.abs(wu).max(), np.abs(wv).max()) > 0.05                        # the hill perturbs the flow
    # m/s, on winds of ~10 m/s.  The restatement is bit-equal to the reference's spatial_winds (tests/test_reference_source.py);
    # the kernels take the column means and the smoothing sums in another order, and a last-bit difference in log N^2 moves
    # the interpolation weights: a few 1e-5 m/s
    assert np.abs(du - wu).max() <= 5e-5 and np.abs(dv - wv).max() <= 5e-5
    # several table entries per level were needed, far fewer than the 720 of the full table
    assert 8 * nz <= lw.evaluated <= 720 * nz * 2 * 0.4       # (per call; the counter restarts with every call)
