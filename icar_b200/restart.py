"""The restart / output file of libicar_b200 (icar_write_restart, include/icar_b200.h) read and written with
NumPy -- the converter a user runs next to the model, since there is no NetCDF library in this image (the
reference writes one NetCDF file per image, src/io/output_obj.f90, and reads it back in src/io/restart.f90:11-90).

Layout (little-endian):
    char magic[8] = "ICARB2RS"; int32 version = 1; int32 nvars; int32 grid[19] (icar_grid member order);
    int32 microphysics; float64 model_time; float64 last_model_time;
    nvars x { int32 var; int32 elem_size; int32 dims[3] = (nx', ny', nz') in file order; int64 nbytes; data }
3-D data are in the reference's FILE order (x, y, z) (output_obj.f90:423-426): as a C-order NumPy array that is
(nz', ny', nx'); `memory_order=True` gives the in-memory (x, z, y) order, i.e. NumPy (ny', nz', nx'), the shape
Domain.upload / download use."""
import struct

import numpy as np

MAGIC = b"ICARB2RS"
GRID_FIELDS = ("ims", "ime", "kms", "kme", "jms", "jme", "its", "ite", "kts", "kte", "jts", "jte",
               "ids", "ide", "kds", "kde", "jds", "jde", "halo_size")


def read_restart_file(path, memory_order=False):
    """-> dict(grid=dict, microphysics=int, model_time=float, last_model_time=float, vars={var id: array})."""
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:8] != MAGIC:
        raise ValueError(f"{path}: not an icar_b200 restart file")
    if len(raw) < 112:
        raise ValueError(f"{path}: truncated header")
    version, nvars = struct.unpack_from("<ii", raw, 8)
    if version != 1:
        raise ValueError(f"{path}: version {version}")
    grid = dict(zip(GRID_FIELDS, struct.unpack_from("<19i", raw, 16)))
    microphysics, = struct.unpack_from("<i", raw, 92)
    model_time, last_model_time = struct.unpack_from("<dd", raw, 96)
    pos, out = 112, {}
    for _ in range(nvars):
        if pos + 28 > len(raw):
            raise ValueError(f"{path}: truncated (variable record)")
        var, es, nx, ny, nz, nbytes = struct.unpack_from("<iiiiiq", raw, pos)
        pos += 28
        if nbytes != es * nx * ny * nz or pos + nbytes > len(raw):
            raise ValueError(f"{path}: truncated or corrupt at variable {var}")
        a = np.frombuffer(raw, dtype=np.float32 if es == 4 else np.float64, count=nx * ny * nz, offset=pos)
        if es == 8:                                        # the 2-D REAL*8 accumulators
            a = a.reshape(ny, nx)
        else:
            a = a.reshape(nz, ny, nx)
            if memory_order:
                a = np.ascontiguousarray(a.transpose(1, 0, 2))
        out[var] = a.copy()
        pos += nbytes
    return dict(grid=grid, microphysics=microphysics, model_time=model_time, last_model_time=last_model_time, vars=out)


def write_restart_file(path, grid, microphysics, model_time, variables, last_model_time=-999.0):
    """variables: {var id: array}; 3-D REAL arrays in memory order (ny', nz', nx'), 2-D REAL*8 arrays (ny', nx')."""
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<ii", 1, len(variables)))
        f.write(struct.pack("<19i", *[int(grid[k]) for k in GRID_FIELDS]))
        f.write(struct.pack("<i", int(microphysics)))
        f.write(struct.pack("<dd", float(model_time), float(last_model_time)))
        for var in sorted(variables):
            a = np.asarray(variables[var])
            if a.ndim == 3:
                a = np.ascontiguousarray(a.astype(np.float32).transpose(1, 0, 2))          # (nz', ny', nx')
                nz, ny, nx = a.shape
            else:
                a = np.ascontiguousarray(a.astype(np.float64))
                (ny, nx), nz = a.shape, 1
            f.write(struct.pack("<iiiiiq", int(var), a.itemsize, nx, ny, nz, a.nbytes))
            f.write(a.tobytes())
