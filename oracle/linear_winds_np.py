"""TEST INFRASTRUCTURE -- NumPy restatement of the core of ICAR's linear-theory wind perturbation
(Barstad & Gronas 2006, appendix A), the piece the cuFFT implementation
(icar_b200/csrc/linear_winds.cu) is checked against.
PARITY: the Fortran cannot be compiled here and holds no expected values, but the reference ships a Python
implementation of the same solution in its x-z form, helpers/ideal_linear.py:solve(); it is run in this container
by tests/golden/make_linear_theory_golden.py and linear_perturbation_at_height reproduces its u_hat to 4.5e-8 of the
field maximum (tests/test_oracle_linear_winds.py) -- so the per-height solution is PINNED; and (round 2) the Fortran itself
is EXECUTED by oracle/f90run (FFTW's calls bound to numpy.fft): initialize_linear_theory_data, the per-height solution
with l /= 0, both layer means, the two shifts of fftshift.f90 and add_buffer_topo equal this file's functions bit for
bit (add_buffer_topo's smoothing sums to 1e-15), tests/test_reference_source.py::test_live_linear_winds.

Follows, function by function:
    src/utilities/fftshift.f90:87-109, 217-239      fftshift2cc / ifftshift2cc (through a SINGLE precision temporary)
    src/physics/linear_winds.f90:351-418            add_buffer_topo
    src/physics/linear_winds.f90:426-470            initialize_linear_theory_data (k, l, kl)
    src/physics/linear_winds.f90:1201-1222          setup_linwinds: buffered terrain -> terrain_frequency
    src/physics/linear_winds.f90:181-237            linear_perturbation_at_height
    src/physics/linear_winds.f90:239-277            linear_perturbation_constz
    src/physics/linear_winds.f90:280-344            linear_perturbation_varyingz

Not restated (the reference's look-up table over speed / direction / stability and its interpolation,
initialize_spatial_winds :596-830 and spatial_winds :840-1127): at 2048 x 2048 x 60 that table would be 7 TB
(SURVEY.md), so the B200 path evaluates the perturbation directly for the background wind at hand.

Arrays are indexed [i, j] = (x, y) like the Fortran (first index fastest there; here plain C order)."""
import numpy as np

SMALL_VALUE = np.float32(1e-15)          # linear_winds.f90:131
PI = np.float32(3.1415927)               # icar_constants: pi (REAL)


def _shift_index(n):
    """ii(i) of fftshift.f90: ii = mod(i + (n+1)/2, n), 0 -> n   (1-based); returned 0-based."""
    i = np.arange(1, n + 1)
    ii = (i + (n + 1) // 2) % n
    ii[ii == 0] = n
    return ii - 1


def fftshift2cc(a):
    """fftshift.f90:87-109: tmp(ii, jj) = a(i, j) with a COMPLEX (single precision) temporary."""
    nx, ny = a.shape
    tmp = np.empty((nx, ny), np.complex64)
    tmp[np.ix_(_shift_index(nx), _shift_index(ny))] = a.astype(np.complex64)
    return tmp.astype(np.complex128)


def ifftshift2cc(a):
    """fftshift.f90:217-239: tmp(i, j) = a(ii, jj), again through single precision."""
    nx, ny = a.shape
    return a[np.ix_(_shift_index(nx), _shift_index(ny))].astype(np.complex64).astype(np.complex128)


def add_buffer_topo(terrain, smooth_window, buffer):
    """linear_winds.f90:351-418.  terrain (nx, ny) REAL -> (nx + 2 buffer, ny + 2 buffer) complex double whose
    borders blend the opposite edges into each other (the FFT sees a periodic field) and are then smoothed."""
    f = np.float32
    terrain = np.asarray(terrain, f)
    nx, ny = terrain.shape[0] + 2 * buffer, terrain.shape[1] + 2 * buffer
    b = np.full((nx, ny), terrain.min(), np.complex128)
    b[buffer:nx - buffer, buffer:ny - buffer] = terrain
    for i in range(1, buffer + 1):                                             # :367-372
        weight = f(i) / (f(buffer) * f(2))
        pos = buffer - i
        first, last = terrain[0, :], terrain[nx - 2 * buffer - 1, :]
        b[pos, buffer:ny - buffer] = first * (f(1) - weight) + last * weight
        b[nx - pos - 1, buffer:ny - buffer] = first * weight + last * (f(1) - weight)
    for i in range(1, buffer + 1):                                             # :373-378
        weight = f(i) / (f(buffer) * f(2))
        w, omw = np.float64(weight), np.float64(f(1) - weight)                 # (1-weight) is a REAL expression
        pos = buffer - i
        lo, hi = b[:, buffer].copy(), b[:, ny - buffer - 1].copy()
        b[:, pos] = lo * omw + hi * w
        b[:, ny - pos - 1] = lo * w + hi * omw
    if smooth_window > 0:                                                      # :382-415 (in place, in this order)
        for j in range(1, buffer + 1):
            window = min(j, smooth_window)
            for i in range(1, nx + 1):
                xs, xe = max(1, i - window), min(nx, i + window)
                ys, ye = max(1, buffer - j + 1 - window), min(ny, buffer - j + 1 + window)
                b[i - 1, buffer - j] = b[xs - 1:xe, ys - 1:ye].sum() / ((xe - xs + 1) * (ye - ys + 1))
                ys, ye = max(1, ny - (buffer - j) - window), min(ny, ny - (buffer - j) + window)
                b[i - 1, ny - (buffer - j) - 1] = b[xs - 1:xe, ys - 1:ye].sum() / ((xe - xs + 1) * (ye - ys + 1))
            for i in range(1, ny + 1):
                xs, xe = max(1, buffer - j + 1 - window), min(nx, buffer - j + 1 + window)
                ys, ye = max(1, i - window), min(ny, i + window)
                b[buffer - j, i - 1] = b[xs - 1:xe, ys - 1:ye].sum() / ((xe - xs + 1) * (ye - ys + 1))
                xs, xe = max(1, nx - (buffer - j) - window), min(nx, nx - (buffer - j) + window)
                b[nx - (buffer - j) - 1, i - 1] = b[xs - 1:xe, ys - 1:ye].sum() / ((xe - xs + 1) * (ye - ys + 1))
    return b


class LinearTheoryData:
    """linear_theory_type (data_structures.f90:184-195) as initialize_linear_theory_data fills it (:426-470):
    k and l run from -pi/dx to +pi/dx INCLUSIVE over nx (ny) points -- the reference's own spacing, not the
    FFT's 2 pi / (n dx) -- and kl = k^2 + l^2 with zeros replaced by 1e-15.  All REAL."""

    def __init__(self, nx, ny, dx):
        f = np.float32
        offset = PI / f(dx)
        gain = f(2) * offset / f(nx - 1)
        k1 = (np.arange(nx, dtype=f) * gain - offset).astype(f)
        gain = f(2) * offset / f(ny - 1)
        l1 = (np.arange(ny, dtype=f) * gain - offset).astype(f)
        self.k = np.repeat(k1[:, None], ny, axis=1)
        self.l = np.repeat(l1[None, :], nx, axis=0)
        self.kl = (self.k ** 2 + self.l ** 2).astype(f)
        self.kl[self.kl == 0] = SMALL_VALUE
        self.nx, self.ny = nx, ny


def terrain_frequency(terrain, buffer):
    """setup_linwinds :1201-1222: two buffering passes (smoothing window 5, then a flat 2-cell rim), forward FFT
    normalised by the number of cells, fftshift.  Returns (terrain_frequency, total buffer)."""
    first = add_buffer_topo(terrain, 5, buffer)
    second = add_buffer_topo(first.real.astype(np.float32), 0, 2)
    nx, ny = second.shape
    tf = np.fft.fft2(second) / (nx * ny)                  # FFTW_FORWARD: exp(-i ...), unnormalised
    return fftshift2cc(tf), buffer + 2


def linear_perturbation_at_height(U, V, Nsq, z, fourier_terrain, lt):
    """linear_winds.f90:181-237.  U, V, Nsq, z REAL scalars; returns (u_perturb, v_perturb), complex double
    (nx, ny) fields on the buffered grid: the inverse transforms of uhat and vhat (FFTW_BACKWARD, unnormalised)."""
    f = np.float32
    U, V, Nsq, z = f(U), f(V), f(Nsq), f(z)
    if U == 0 and V == 0:
        zero = np.zeros(fourier_terrain.shape, np.complex128)
        return zero, zero.copy()
    sig = (U * lt.k + V * lt.l).astype(f)
    sig[sig == 0] = SMALL_VALUE
    denom = (sig ** 2).astype(f).astype(np.complex128)                # sig**2 is REAL, stored in a complex double
    msq = (Nsq / denom) * lt.kl                                        # REAL / complex double * REAL
    with np.errstate(invalid="ignore"):
        mimag = np.sqrt(-msq).real * 1j                                # (0,0) + real(sqrt(-msq)) * i
        m = np.sqrt(msq)                                               # vertical wave number, hydrostatic
    m = np.where(sig < 0, -m, m)
    m = np.where(msq.real < 0, mimag, m)
    ineta = 1j * fourier_terrain * np.exp(1j * m * np.float64(z))
    ineta = ineta / (lt.kl / ((0 - m) * sig))
    uhat = ifftshift2cc(lt.k * ineta)
    vhat = ifftshift2cc(lt.l * ineta)
    n = uhat.size
    return np.fft.ifft2(uhat) * n, np.fft.ifft2(vhat) * n              # FFTW does not normalise


def linear_perturbation_constz(U, V, Nsq, z_bottom, z_top, minimum_step, fourier_terrain, lt):
    """linear_winds.f90:239-277: the mean over n = max(1, ceiling((z_top - z_bottom) / minimum_step)) heights
    at the middles of equal sub-layers."""
    f = np.float32
    if f(U) == 0 and f(V) == 0:
        zero = np.zeros(fourier_terrain.shape, np.complex128)
        return zero, zero.copy()
    n_steps = max(1, int(np.ceil((f(z_top) - f(z_bottom)) / f(minimum_step))))
    step_size = (f(z_top) - f(z_bottom)) / f(n_steps)
    current_z = f(z_bottom) + step_size / f(2)
    ua = np.zeros(fourier_terrain.shape, np.complex128)
    va = np.zeros(fourier_terrain.shape, np.complex128)
    for _ in range(n_steps):
        u, v = linear_perturbation_at_height(U, V, Nsq, current_z, fourier_terrain, lt)
        ua += u
        va += v
        current_z = f(current_z + step_size)
    return ua / n_steps, va / n_steps


def linear_perturbation_varyingz(U, V, Nsq, z_bottom, z_top, minimum_step, fourier_terrain, lt, buffer):
    """linear_winds.f90:280-344: z_bottom / z_top are (nx, ny) fields of the un-buffered domain; every sampled
    height contributes to a column with the fraction of its sub-layer that lies inside that column's layer."""
    f = np.float32
    shape = fourier_terrain.shape
    if f(U) == 0 and f(V) == 0:
        zero = np.zeros(shape, np.complex128)
        return zero, zero.copy()
    z_bottom, z_top = np.asarray(z_bottom, f), np.asarray(z_top, f)
    start_z, end_z = z_bottom.min(), z_top.max()
    itop = np.full(shape, z_top.max(), f)
    ibot = np.full(shape, z_bottom.min(), f)
    nx0, ny0 = z_top.shape
    itop[buffer - 1:buffer - 1 + nx0, buffer - 1:buffer - 1 + ny0] = z_top       # (buffer:buffer+n-1) 1-based, as written
    ibot[buffer - 1:buffer - 1 + nx0, buffer - 1:buffer - 1 + ny0] = z_bottom
    layer_count = np.zeros(shape, f)
    step_size = min(f(minimum_step), (z_top - z_bottom).min())
    current_z = f(start_z + step_size / f(2))
    ua = np.zeros(shape, np.complex128)
    va = np.zeros(shape, np.complex128)
    half = step_size / f(2)
    while current_z < end_z:
        u, v = linear_perturbation_at_height(U, V, Nsq, current_z, fourier_terrain, lt)
        frac = np.maximum(f(0), np.minimum(half, current_z - ibot) + np.minimum(f(0), itop - current_z)
                          + np.minimum(half, itop - current_z) + np.minimum(f(0), current_z - ibot)) / step_size
        frac = frac.astype(f)
        layer_count += frac
        ua += u * frac
        va += v * frac
        current_z = f(current_z + step_size)
    with np.errstate(divide="ignore", invalid="ignore"):
        return ua / layer_count, va / layer_count
