"""What icar_linear_perturb must leave in u, v according to the NumPy restatement (oracle/linear_winds_np.py):
the LUT builder's per-level perturbation (linear_winds.f90:716-737), its stagger (:756-762) and spatial_winds'
update (:1094-1111).  Shared by the GPU test and by the CPU emulation of the kernels.  TEST INFRASTRUCTURE."""
import numpy as np

from oracle import linear_winds_np as LW


def z_interface_of(st):
    """global z_interface = terrain + the layers below; (ny, nz, nx) like every 3-D array of icar_b200.ideal."""
    zi = np.empty_like(st["dz_interface"])
    zi[:, 0, :] = st["terrain"]
    for k in range(1, st["nz"]):
        zi[:, k, :] = zi[:, k - 1, :] + st["dz_interface"][:, k - 1, :]
    return zi


def stagger(field, buffer, nx, ny):
    """initialize_spatial_winds :756-762 (0-based): u faces (nx+1, ny), v faces (nx, ny+1), REAL."""
    f = np.float32
    b = buffer
    tu = (field[0][b - 1:b + nx, b:b + ny] + field[0][b:b + nx + 1, b:b + ny]).real.astype(f) / f(2)
    tv = (field[1][b:b + nx, b - 1:b + ny] + field[1][b:b + nx, b:b + ny + 1]).real.astype(f) / f(2)
    return tu, tv


def expected_winds(st, buffer, U, V, Nsq, frac, contrib, varying, calls, tile=None, minimum_layer_size=100.0):
    """u, v after `calls` calls of linear_perturb on the state `st`; tile = (i0, j0, nx, ny) of a sub-domain whose
    u, v (shapes (ny, nz, nx+1), (ny+1, nz, nx)) are st["u_tile"], st["v_tile"]."""
    f = np.float32
    nxg, nyg, nz = st["nx"], st["ny"], st["nz"]
    terrain = np.ascontiguousarray(st["terrain"].T)
    tf, total = LW.terrain_frequency(terrain, buffer)
    lt = LW.LinearTheoryData(tf.shape[0], tf.shape[1], st["dx"])
    zi = z_interface_of(st)
    lh = (np.cumsum(st["dz_levels"], dtype=f) - st["dz_levels"] / f(2)).astype(f)
    if tile is None:
        i0, j0, nx, ny = 0, 0, nxg, nyg
        u, v = st["u"].copy(), st["v"].copy()
    else:
        i0, j0, nx, ny = tile
        u, v = st["u_tile"].copy(), st["v_tile"].copy()
    up, vp = np.zeros_like(u), np.zeros_like(v)
    frac, contrib = f(frac), f(contrib)
    for _ in range(calls):
        for k in range(nz):
            if varying:
                zb = (zi[:, k, :] - st["terrain"]).T
                zt = (zi[:, k, :] - st["terrain"] + st["dz_interface"][:, k, :]).T
                pert = LW.linear_perturbation_varyingz(U, V, Nsq, zb, zt, minimum_layer_size, tf, lt, total)
            else:
                half = st["dz_levels"][k] / f(2)
                pert = LW.linear_perturbation_constz(U, V, Nsq, lh[k] - half, lh[k] + half, minimum_layer_size, tf, lt)
            tu, tv = stagger(pert, total, nxg, nyg)
            tu, tv = tu[i0:i0 + nx + 1, j0:j0 + ny], tv[i0:i0 + nx, j0:j0 + ny + 1]
            up[:, k, :] = up[:, k, :] * (f(1) - frac) + frac * tu.T
            vp[:, k, :] = vp[:, k, :] * (f(1) - frac) + frac * tv.T
            u[:, k, :] += up[:, k, :] * contrib
            v[:, k, :] += vp[:, k, :] * contrib
    return u, v, lh, zi
