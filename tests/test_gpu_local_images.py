"""The multi-image path on ONE GPU: the images of a slab decomposition as contexts of one process on one device,
connected with icar_comm_init_local, so that the halo exchange kernels (k_halo_begin / k_halo_push / k_halo_wait: stores
into the neighbour's ghost cells + flags, csrc/halo.cu) and the multi-image schedule of icar_physics_step run -- and are
checked -- on every box, not only on multi-GPU ones (tests/test_gpu_multi.py does the same over CUDA IPC, one process
per GPU).

Bar: with halo 3 the cells an image owns equal the single-image run BIT FOR BIT (the dependency radius of MPDATA order 2
+ FCT is 3 cells; the reference exchanges 1 cell and is therefore decomposition-dependent, tests/test_decomp_gloo.py)."""
import numpy as np
import pytest

from icar_b200 import ideal

pytestmark = pytest.mark.gpu


def owned(grid, axis, first, last):
    """(slice of the tile, slice of the global domain) of the cells an image owns, along `axis`."""
    if axis == "x":
        lo = 0 if first else grid.its - grid.ims
        hi = grid.ime - grid.ims + 1 if last else grid.ite - grid.ims + 1
        g0 = grid.ims - 1 + lo
        return (Ellipsis, slice(lo, hi)), (Ellipsis, slice(g0, g0 + hi - lo))
    lo = 0 if first else grid.jts - grid.jms
    hi = grid.jme - grid.jms + 1 if last else grid.jte - grid.jms + 1
    g0 = grid.jms - 1 + lo
    return (slice(lo, hi),), (slice(g0, g0 + hi - lo),)


def run(gpu_lib, dims, nimages, axis, mp, steps, graph):
    nx, nz, ny = dims
    st = ideal.make_ideal_case(nx=nx, nz=nz, ny=ny, perturb=True)
    st["qv"] *= np.float32(3.0)
    if mp == 1:
        ideal.add_thompson_fields(st)
    dt = float(st["dt"])
    opt = dict(dx=st["dx"], microphysics=mp)
    one = gpu_lib.Domain(gpu_lib.Grid.single_image(nx, nz, ny), gpu_lib.Options(**opt))
    one.upload_state(st)
    make = gpu_lib.Grid.xslab if axis == "x" else gpu_lib.Grid.yslab
    grids = [make(nx, nz, ny, nimages, i + 1, 3) for i in range(nimages)]
    doms = [gpu_lib.Domain(g, gpu_lib.Options(**opt)) for g in grids]
    for d, g in zip(doms, grids):
        d.upload_state(ideal.slab_of(st, g))
    gpu_lib.comm_init_local(doms)
    # update_dt: each image's CFL time step; their minimum (co_min) is the single-image value
    assert min(d.compute_dt(st["dz_levels"]) for d in doms) == one.compute_dt(st["dz_levels"])
    if graph:
        for _ in range(steps // 2):                       # every image queues the same two steps before anyone waits
            for d in doms:
                d.physics_steps(dt, 2, graph=True)
    else:
        for _ in range(steps):
            for d in doms:
                d.physics_step(dt)
    for _ in range(steps):
        one.physics_step(dt)
    fields = ("th", "qv", "qc", "qr", "qs", "precipitation") + (("qi", "qg", "ni", "nr", "snowfall", "graupel") if mp == 1 else ())
    launches = 0
    for i, (d, g) in enumerate(zip(doms, grids)):
        mine, theirs = owned(g, axis, i == 0, i == nimages - 1)
        for k in fields:
            a, b = d.download(k)[mine], one.download(k)[theirs]
            assert np.array_equal(a, b), (axis, nimages, i, k, float(np.abs(a - b).max()))
        launches += d.kernel_launches
    for d in doms:
        d.close()
    one.close()
    return launches


@pytest.mark.parametrize("axis", ["x", "y"])
@pytest.mark.parametrize("graph", [False, True])
def test_two_images_on_one_gpu_equal_the_single_image_bit_for_bit(gpu_lib, axis, graph):
    assert run(gpu_lib, (96, 12, 96), 2, axis, 2, 4, graph) > 0


def test_three_images_the_middle_one_has_two_neighbours(gpu_lib):
    run(gpu_lib, (130, 10, 48), 3, "x", 2, 4, False)
    run(gpu_lib, (48, 10, 131), 3, "y", 2, 4, True)      # uneven slabs: 44 + 44 + 43 rows


@pytest.mark.parametrize("axis", ["x", "y"])
def test_two_images_thompson(gpu_lib, axis):
    run(gpu_lib, (64, 12, 64), 2, axis, 1, 4, axis == "y")


def test_exchange_calls_are_checked(gpu_lib):
    """icar_halo_retrieve without a send, a second send before the retrieve, and upload / download while the neighbours
    may be writing the ghost cells are refused; an unconnected slab says so."""
    nx, nz, ny = 64, 8, 32
    st = ideal.make_ideal_case(nx=nx, nz=nz, ny=ny)
    grids = [gpu_lib.Grid.yslab(nx, nz, ny, 2, i + 1, 3) for i in range(2)]
    doms = [gpu_lib.Domain(g, gpu_lib.Options(dx=st["dx"])) for g in grids]
    for d, g in zip(doms, grids):
        d.upload_state(ideal.slab_of(st, g))
    gpu_lib.comm_init_local(doms)
    with pytest.raises(gpu_lib.IcarError, match="without icar_halo_send"):
        doms[0].halo_retrieve()
    for d in doms:
        d.halo_send()
    with pytest.raises(gpu_lib.IcarError, match="has not been retrieved"):
        doms[0].halo_send()
    with pytest.raises(gpu_lib.IcarError, match="ghost cells"):
        doms[0].download("qv")
    for d in doms:
        d.halo_retrieve()
    # the exchange moved image 2's first owned rows into image 1's ghost rows
    a = doms[0].download("th")[grids[0].jte - grids[0].jms + 1:]
    b = doms[1].download("th")[3:6]
    assert np.array_equal(a, b)
    for d in doms:
        d.close()
