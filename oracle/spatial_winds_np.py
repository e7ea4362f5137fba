"""TEST INFRASTRUCTURE -- NumPy restatement of spatial_winds (src/physics/linear_winds.f90:840-1127), the routine
linear_perturb runs every forcing interval: per-cell stability, its smoothing, the column-mean background wind, the
position of every cell in the (speed, direction, log N^2) look-up table, the trilinear interpolation between the
eight surrounding table entries, the relaxation of the stored perturbation and its addition to u, v.

The table entries are what initialize_spatial_winds (:596-830) computes -- linear_perturbation for the node's
(U, V, N^2) and the level's bounds, staggered (:756-762) -- but they are evaluated ON DEMAND here (a dictionary
keyed by node and level) instead of being tabulated in full: the full table is 720 x nx x nz x ny x 2 REALs with
the namelist defaults (options_obj.f90:1469-1477).  The values are the same.

PARITY: pinned to the reference's own source.  oracle/f90run executes spatial_winds itself (with calc_stability,
calc_direction, calc_speed, calc_weight, smooth_array, linear_space of atm_utilities.f90 / array_utilities.f90) on a mock
domain; given the same look-up table this file's spatial_winds returns the same u, v and stored perturbations in every
bit (tests/test_reference_source.py::test_live_spatial_winds; REAL log / exp / atan / sin / cos are glibc's, like the
compiled Fortran's).  The table ENTRIES: initialize_spatial_winds is executed too (one image, its own loop over the
nodes, the coarray PUT of copy_data_to_remote) and every entry of its table equals EntryCache.get
(::test_live_spatial_winds_table).

Also restated: calc_stability / calc_dry_stability / calc_moist_stability / calc_sat_lapse_rate / calc_direction /
calc_speed / calc_u / calc_v (src/utilities/atm_utilities.f90:334-467), calc_weight / linear_space / smooth_array_3d
with ydim = 3 (src/utilities/array_utilities.f90:215-237, 263-288, 290-362).

Arrays: 3-D fields are (ny, nz, nx) C-order like everywhere in icar_b200 (u: nx + 1, v: ny + 1)."""
import numpy as np

from oracle import linear_winds_np as LW

f = np.float32
PI = f(3.1415927)
GRAVITY, RD, RW, CP, LH_VAP = f(9.81), f(287.058), f(461.5), f(1012.0), f(2260000.0)


def _libm1(name):
    """A REAL elemental intrinsic as the compiled Fortran evaluates it: glibc's single-precision function, element by element
    (NumPy's own vectorised float32 log / exp / arctan differ from glibc's in the last bit now and then)."""
    import ctypes
    import ctypes.util
    fn = getattr(ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6"), name)
    fn.argtypes, fn.restype = [ctypes.c_float], ctypes.c_float

    def apply(x):
        a = np.asarray(x, f)
        if a.ndim == 0:
            return f(fn(float(a)))
        a = np.ascontiguousarray(a)
        out = np.empty_like(a)
        fo, fx = out.reshape(-1), a.reshape(-1)
        for n in range(fx.size):
            fo[n] = fn(float(fx[n]))
        return out
    return apply


logf, expf, atanf, sinf, cosf = (_libm1(n) for n in ("logf", "expf", "atanf", "sinf", "cosf"))


class LTOptions:
    """lt_parameters with the namelist defaults (options_obj.f90:1448-1479)."""

    def __init__(self, **kw):
        self.variable_N = True
        self.smooth_nsq = True
        self.buffer = 50
        self.stability_window_size = 10
        self.vert_smooth = 10
        self.max_stability = f(6e-4)
        self.min_stability = f(1e-7)
        self.N_squared = f(3e-5)
        self.linear_contribution = f(1)
        self.linear_update_fraction = f(0.2)
        self.dirmax, self.dirmin = f(2) * PI, f(0)
        self.spdmax, self.spdmin = f(30), f(0)
        self.n_dir_values, self.n_nsq_values, self.n_spd_values = 24, 5, 6
        self.minimum_layer_size = f(100)
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError(k)
            setattr(self, k, v)
        self.nsqmax = logf(f(self.max_stability))                  # log(max_stability), REAL
        self.nsqmin = logf(f(self.min_stability))


def linear_space(lo, hi, n):
    """array_utilities.f90:233-235: (i-1.0)/real(n-1.0) * (max - min) + min, REAL."""
    i = np.arange(n, dtype=f)
    return (i / f(n - 1) * (f(hi) - f(lo)) + f(lo)).astype(f)


def table_axes(o):
    """dir_values, nsq_values, spd_values (linear_winds.f90:657-661)."""
    return (linear_space(o.dirmin, o.dirmax, o.n_dir_values), linear_space(o.nsqmin, o.nsqmax, o.n_nsq_values),
            linear_space(o.spdmin, o.spdmax, o.n_spd_values))


def calc_direction(u, v):
    """atm_utilities.f90:334-355 on arrays."""
    u, v = np.asarray(u, f), np.asarray(v, f)
    with np.errstate(divide="ignore", invalid="ignore"):
        a = atanf((u / v).astype(f))
    out = np.where(v < 0, a + PI, np.where(u >= 0, a, a + f(2) * PI)).astype(f)
    zero = v == 0
    out = np.where(zero & (u > 0), PI / f(2), out)
    out = np.where(zero & ~(u > 0), PI * f(1.5), out)
    return out.astype(f)


def calc_speed(u, v):
    return np.sqrt((np.asarray(u, f) ** 2 + np.asarray(v, f) ** 2).astype(f)).astype(f)


def calc_stability(th_top, th_bot, pii_top, pii_bot, z_top, z_bot, qv_top, qv_bot, qc):
    """atm_utilities.f90:448-467 with that module's variable_N = .True. (the branch spatial_winds reaches)."""
    dry = (GRAVITY * (logf(th_top) - logf(th_bot)).astype(f) / (z_top - z_bot)).astype(f)
    t_top, t_bot = (th_top * pii_top).astype(f), (th_bot * pii_bot).astype(f)
    t = ((t_top + t_bot) / f(2)).astype(f)
    qv = ((qv_top + qv_bot) / f(2)).astype(f)
    dz = (z_top - z_bot).astype(f)
    L = LH_VAP
    sat_lapse = (GRAVITY * ((f(1) + (L * qv) / (RD * t)) / (CP + (L * L * qv * (RD / RW)) / (RD * t * t)))).astype(f)
    moist = ((GRAVITY / t) * ((t_top - t_bot) / dz + sat_lapse) * (f(1) + (L * qv) / (RD * t))
             - (GRAVITY / (f(1) + qv + qc) * (qv_top - qv_bot) / dz)).astype(f)
    return np.where(qc < f(1e-7), dry, moist).astype(f)


def smooth_array_ydim3(a, w):
    """smooth_array_3d(wind, windowsize, ydim=3) (array_utilities.f90:290-362) on an (ny, nz, nx) array: a box mean
    of 2w+1 along y, then along x, both with the edge value repeated outside the array, REAL*8 sums.  The Fortran
    keeps running sums (whose first term is rounded to REAL); here every window is summed afresh, so results can
    differ in the last bit of the REAL result."""
    ny, nz, nx = a.shape
    n = 2 * w + 1
    d = a.astype(np.float64)
    idx = np.clip(np.arange(ny)[:, None] + np.arange(-w, w + 1)[None, :], 0, ny - 1)
    rowmeans = d[idx].sum(axis=1) / n                                   # (ny, nz, nx)
    idx = np.clip(np.arange(nx)[:, None] + np.arange(-w, w + 1)[None, :], 0, nx - 1)
    out = rowmeans[:, :, idx].sum(axis=3) / n
    return out.astype(f)


def nsquared_field(st, o):
    """linear_winds.f90:898-959: log N^2 per cell, vertically smoothed in place (sequentially in j), then
    smooth_array over (x, y).  st: th, pii, z, qv (+ qc, qi, qr, qs when present), all (ny, nz, nx)."""
    th = st["th"]
    ny, nz, nx = th.shape
    nsq = np.empty((ny, nz, nx), f)
    vs = o.vert_smooth
    hyd = np.zeros_like(th)
    for k in ("qc", "qi", "qr", "qs"):                                  # :905-913, this order
        if st.get(k) is not None:
            hyd = (hyd + st[k]).astype(f)
    for j in range(1, nz + 1):
        if o.variable_N:
            top = min(j + vs, nz)
            bottom = max(1, j - (vs - (top - j)))
            b, t = bottom - 1, top - 1
            val = calc_stability(th[:, b], th[:, t], st["pii"][:, b], st["pii"][:, t], st["z"][:, b], st["z"][:, t],
                                 st["qv"][:, b], st["qv"][:, t], hyd[:, j - 1])
            val = np.maximum(o.min_stability, np.minimum(o.max_stability, val)).astype(f)
        else:
            val = np.full((ny, nx), o.N_squared, f)
        nsq[:, j - 1] = logf(val)
    if o.smooth_nsq:
        for j in range(1, nz + 1):                                       # in place: levels below j are already smoothed
            top = min(j + vs, nz)
            bottom = max(1, j - (vs - (top - j)))
            acc = nsq[:, j - 1].copy()
            for s in range(bottom, j):
                acc = (acc + nsq[:, s - 1]).astype(f)
            for s in range(j + 1, top + 1):
                acc = (acc + nsq[:, s - 1]).astype(f)
            nsq[:, j - 1] = (acc / f(top - bottom + 1)).astype(f)
        nsq = smooth_array_ydim3(nsq, o.stability_window_size)
    return nsq


def positions(values, cur):
    """`pos = 1; do step: if (cur > values(step)) pos = step` (:1045-1071) and calc_weight (array_utilities.f90:263-288).
    Returns 0-based pos, next and the REAL weight of pos."""
    values = np.asarray(values, f)
    n = len(values)
    pos = np.zeros(cur.shape, np.int64)
    for step in range(n):
        pos = np.where(cur > values[step], step, pos)
    below = cur < values[0]
    last = pos == n - 1
    nxt = np.where(below, 0, np.where(last, n - 1, np.minimum(pos + 1, n - 1)))
    with np.errstate(divide="ignore", invalid="ignore"):
        w = ((values[nxt] - cur) / (values[nxt] - values[pos])).astype(f)
    w = np.where(below | last, f(1), w).astype(f)
    return pos, nxt, w


class EntryCache:
    """hi_u_LUT(spd, dir, nsq, :, z, :) / hi_v_LUT(...) computed when first asked for (linear_winds.f90:700-762)."""

    def __init__(self, terrain_xy, dx, o, layer_bounds):
        """layer_bounds(z) -> (z_bottom, z_top): scalars (constz) or (nx, ny) arrays (varyingz), 0-based level."""
        self.o = o
        self.tf, self.total = LW.terrain_frequency(terrain_xy, o.buffer)
        self.lt = LW.LinearTheoryData(self.tf.shape[0], self.tf.shape[1], dx)
        self.nx, self.ny = terrain_xy.shape
        self.bounds = layer_bounds
        self.dirs, self.nsqs, self.spds = table_axes(o)
        self.store = {}
        self.computed = 0

    def get(self, spd, dr, nsq, z):
        key = (int(spd), int(dr), int(nsq), int(z))
        if key not in self.store:
            U = (sinf(self.dirs[dr]) * self.spds[spd]).astype(f)      # calc_u / calc_v (atm_utilities.f90:373-391)
            V = (cosf(self.dirs[dr]) * self.spds[spd]).astype(f)
            N2 = expf(self.nsqs[nsq])
            zb, zt = self.bounds(z)
            if np.ndim(zb) == 0:
                p = LW.linear_perturbation_constz(U, V, N2, zb, zt, self.o.minimum_layer_size, self.tf, self.lt)
            else:
                p = LW.linear_perturbation_varyingz(U, V, N2, zb, zt, self.o.minimum_layer_size, self.tf, self.lt, self.total)
            b, nx, ny = self.total, self.nx, self.ny
            tu = (p[0][b - 1:b + nx, b:b + ny] + p[0][b:b + nx + 1, b:b + ny]).real.astype(f) / f(2)
            tv = (p[1][b:b + nx, b - 1:b + ny] + p[1][b:b + nx, b:b + ny + 1]).real.astype(f) / f(2)
            self.store[key] = (np.ascontiguousarray(tu.T), np.ascontiguousarray(tv.T))    # (ny, nx+1), (ny+1, nx)
            self.computed += 1
        return self.store[key]


def interpolate(cache, which, z, rows, cols, spos, nexts, dpos, nextd, npos, nextn, sw, dw, nw):
    """:1084-1091 for one level: the eight entries gathered cell by cell, combined in the Fortran's order."""
    def lut(s, d, n):
        out = np.empty(s.shape, f)
        keys = np.stack([s, d, n], axis=-1).reshape(-1, 3)
        for ks, kd, kn in np.unique(keys, axis=0):
            m = (s == ks) & (d == kd) & (n == kn)
            out[m] = cache.get(ks, kd, kn, z)[which][rows[m], cols[m]]
        return out
    one = f(1)
    first = (nw * (dw * lut(spos, dpos, npos) + (one - dw) * lut(spos, nextd, npos))
             + (one - nw) * (dw * lut(spos, dpos, nextn) + (one - dw) * lut(spos, nextd, nextn))).astype(f)
    second = (nw * (dw * lut(nexts, dpos, npos) + (one - dw) * lut(nexts, nextd, npos))
              + (one - nw) * (dw * lut(nexts, dpos, nextn) + (one - dw) * lut(nexts, nextd, nextn))).astype(f)
    return (sw * first + (one - sw) * second).astype(f)


def spatial_winds(st, cache, o, u_pert, v_pert, tile=(0, 0)):
    """spatial_winds(domain, reverse=.False., vsmooth, winsz, update) on st["u"] (ny, nz, nx+1), st["v"] (ny+1, nz, nx),
    in place; u_pert / v_pert are hi_u_perturbation / hi_v_perturbation (same shapes, updated in place).
    tile = (i0, j0): where this image's tile sits in the global domain the table entries cover."""
    u3d, v3d = st["u"], st["v"]
    ny, nz, nxu = u3d.shape
    nx, nyv = nxu - 1, ny + 1
    i0, j0 = tile
    nsq = nsquared_field(st, o)                                          # log N^2, (ny, nz, nx)
    winsz = o.stability_window_size
    frac, contrib = f(o.linear_update_fraction), f(o.linear_contribution)
    dirs, nsqs, spds = cache.dirs, cache.nsqs, cache.spds
    ii = np.arange(nxu)
    vi = np.minimum(ii, nx - 1)
    for k in range(nyv):                                                 # rows in order: row ny sees the updated u of row ny-1
        uk = min(k, ny - 1)
        u1d = np.zeros(nxu, f)
        v1d = np.zeros(nxu, f)
        for j in range(nz):                                              # sum(u3d(i,:,uk)) / nz, REAL, in level order
            u1d = (u1d + u3d[uk, j, :]).astype(f)
            v1d = (v1d + v3d[k, j, vi]).astype(f)
        u1d, v1d = (u1d / f(nz)).astype(f), (v1d / f(nz)).astype(f)
        curdir, curspd = calc_direction(u1d, v1d), calc_speed(u1d, v1d)
        dpos, nextd, dw = positions(dirs, curdir)
        spos, nexts, sw = positions(spds, curspd)
        for j in range(nz):
            bottom, top = max(j - winsz, 0), min(j + winsz, nz - 1)
            acc = np.zeros(nxu, f)
            for s in range(bottom, top + 1):
                acc = (acc + nsq[uk, s, vi]).astype(f)
            curnsq = (acc / f(top - bottom + 1)).astype(f)
            npos, nextn, nw = positions(nsqs, curnsq)
            if k < ny:
                rows = np.full(nxu, j0 + k)
                new = interpolate(cache, 0, j, rows, i0 + ii, spos, nexts, dpos, nextd, npos, nextn, sw, dw, nw)
                u_pert[k, j, :] = (u_pert[k, j, :] * (f(1) - frac) + frac * new).astype(f)
                u3d[k, j, :] = (u3d[k, j, :] + u_pert[k, j, :] * contrib).astype(f)
            m = slice(0, nx)
            rows = np.full(nx, j0 + k)
            new = interpolate(cache, 1, j, rows, i0 + ii[m], spos[m], nexts[m], dpos[m], nextd[m], npos[m], nextn[m],
                              sw[m], dw[m], nw[m])
            v_pert[k, j, :] = (v_pert[k, j, :] * (f(1) - frac) + frac * new).astype(f)
            v3d[k, j, :] = (v3d[k, j, :] + v_pert[k, j, :] * contrib).astype(f)
    return nsq
