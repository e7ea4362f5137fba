"""Worker of tests/test_gpu_multi.py: one rank per GPU (torch.distributed.run), x-slab decomposition
with halo 3, NCCL halo exchange inside libicar_b200.  Every rank also runs the whole domain on its
own GPU as a single image; the owned cells of the slab must equal it BIT FOR BIT."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def slab(st, g):
    lo, hi = g.ims - 1, g.ime
    out = {}
    for k, v in st.items():
        if not isinstance(v, np.ndarray):
            out[k] = v
        elif k in ("u", "jacobian_u"):
            out[k] = np.ascontiguousarray(v[:, :, lo:hi + 1])
        elif v.ndim in (2, 3):
            out[k] = np.ascontiguousarray(v[..., lo:hi])
        else:
            out[k] = v
    return out


def main():
    import torch
    import torch.distributed as dist
    import icar_b200
    from icar_b200 import ideal

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    nx, nz, ny, steps, halo = (int(a) for a in sys.argv[1:6])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")          # plumbing only: id broadcast + barrier
    st = ideal.make_ideal_case(nx=nx, nz=nz, ny=ny, perturb=True)
    st["qv"] *= np.float32(3.0)
    dt = float(st["dt"])
    one = icar_b200.Domain(icar_b200.Grid.single_image(nx, nz, ny), icar_b200.Options(dx=st["dx"]), device=local)
    one.upload_state(st)
    grid = icar_b200.Grid.xslab(nx, nz, ny, world, rank + 1, halo)
    dom = icar_b200.Domain(grid, icar_b200.Options(dx=st["dx"]), device=local)
    dom.upload_state(slab(st, grid))
    ids = [icar_b200.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    dom.comm_init(rank + 1, world, ids[0])
    for _ in range(steps):
        one.physics_step(dt)
        dom.physics_step(dt)
    dom.synchronize()
    lo = 0 if rank == 0 else grid.its - grid.ims
    hi = grid.nx if rank == world - 1 else grid.ite - grid.ims + 1
    g0 = grid.ims - 1 + lo
    bad = []
    for k in ("th", "qv", "qc", "qr", "qs"):
        a = dom.download(k)[:, :, lo:hi]
        b = one.download(k)[:, :, g0:g0 + (hi - lo)]
        if not np.array_equal(a, b):
            bad.append((k, float(np.abs(a - b).max())))
    pa = dom.download("precipitation")[:, lo:hi]
    pb = one.download("precipitation")[:, g0:g0 + (hi - lo)]
    if not np.array_equal(pa, pb):
        bad.append(("precipitation", float(np.abs(pa - pb).max())))
    print(f"rank {rank}: owned columns {g0}..{g0 + hi - lo - 1}, mismatching fields: {bad}", flush=True)
    dist.barrier()
    dom.close(); one.close()
    dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
