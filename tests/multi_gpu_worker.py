"""Worker of tests/test_gpu_multi.py: one rank per GPU (torch.distributed.run), x-slab decomposition
with halo 3, NCCL halo exchange inside libicar_b200.  Every rank also runs the whole domain on its
own GPU as a single image; the owned cells of the slab must equal it BIT FOR BIT."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import icar_b200
    from icar_b200 import ideal

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    nx, nz, ny, steps, halo = (int(a) for a in sys.argv[1:6])
    axis = sys.argv[6] if len(sys.argv) > 6 else "x"
    mp = int(sys.argv[7]) if len(sys.argv) > 7 else 2      # 2 = mp_simple (5 scalars), 1 = Thompson (9 scalars)
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")          # plumbing only: id broadcast + barrier
    st = ideal.make_ideal_case(nx=nx, nz=nz, ny=ny, perturb=True)
    st["qv"] *= np.float32(3.0)
    if mp == 1:
        ideal.add_thompson_fields(st)
    dt = float(st["dt"])
    one = icar_b200.Domain(icar_b200.Grid.single_image(nx, nz, ny), icar_b200.Options(dx=st["dx"], microphysics=mp), device=local)
    one.upload_state(st)
    grid = (icar_b200.Grid.xslab if axis == "x" else icar_b200.Grid.yslab)(nx, nz, ny, world, rank + 1, halo)
    dom = icar_b200.Domain(grid, icar_b200.Options(dx=st["dx"], microphysics=mp), device=local)
    dom.upload_state(ideal.slab_of(st, grid))
    ids = [icar_b200.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    dom.comm_init(rank + 1, world, ids[0])
    # update_dt: the CFL time step of the slab, `co_min` over the images (ncclAllReduce), must equal the
    # single-image value (the per-cell maxima partition exactly)
    dt_one = one.compute_dt(st["dz_levels"])
    dt_all = dom.compute_dt(st["dz_levels"])
    assert dt_one == dt_all == min(dt, 120.0), (dt_one, dt_all, dt)
    if axis == "x":
        lo = 0 if rank == 0 else grid.its - grid.ims
        hi = grid.nx if rank == world - 1 else grid.ite - grid.ims + 1
        g0 = grid.ims - 1 + lo
        own = lambda a: a[..., lo:hi]
        ref = lambda a: a[..., g0:g0 + (hi - lo)]
    else:
        lo = 0 if rank == 0 else grid.jts - grid.jms
        hi = grid.ny if rank == world - 1 else grid.jte - grid.jms + 1
        g0 = grid.jms - 1 + lo
        own = lambda a: a[lo:hi]
        ref = lambda a: a[g0:g0 + (hi - lo)]
    fields = ("th", "qv", "qc", "qr", "qs", "precipitation") + (("qi", "qg", "ni", "nr", "snowfall", "graupel") if mp == 1 else ())
    bad = []
    # `steps` eager steps, then `steps` more replayed from the CUDA graph of a pair of steps (the exchange kernels are
    # part of the captured pair); the owned cells are compared with the single image after each phase
    for phase in ("eager", "graph"):
        for _ in range(steps):
            one.physics_step(dt)
        if phase == "eager":
            for _ in range(steps):
                dom.physics_step(dt)
        else:
            dom.physics_steps(dt, steps, graph=True)
        dom.synchronize()
        for k in fields:
            a, b = own(dom.download(k)), ref(one.download(k))
            if not np.array_equal(a, b):
                bad.append((phase, k, float(np.abs(a - b).max())))
        dist.barrier()
    print(f"rank {rank}: mp {mp}, axis {axis}, owned {g0}..{g0 + hi - lo - 1}, mismatching fields: {bad}", flush=True)
    dist.barrier()
    dom.close(); one.close()
    dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
