"""Restart / output file without a GPU: the header parser of the library
(icar_restart_info, a pure host function) against files written by the NumPy mirror of the format
(icar_b200/restart.py), and the mirror's own round trip.  The device side (icar_write_restart /
icar_read_restart and the bit-identical continuation) is in tests/test_gpu_restart.py."""
import numpy as np
import pytest

import icar_b200
from icar_b200 import restart as R


def sample(tmp_path, nx=7, nz=3, ny=5):
    g = icar_b200.Grid.single_image(nx, nz, ny)
    grid = {k: getattr(g, k) for k in R.GRID_FIELDS}
    rng = np.random.default_rng(0)
    variables = {icar_b200.VAR["u"]: rng.random((ny, nz, nx + 1), dtype=np.float32),
                 icar_b200.VAR["v"]: rng.random((ny + 1, nz, nx), dtype=np.float32),
                 icar_b200.VAR["water_vapor"]: rng.random((ny, nz, nx), dtype=np.float32),
                 icar_b200.VAR["precipitation"]: rng.random((ny, nx))}
    path = str(tmp_path / "restart_000001.bin")
    R.write_restart_file(path, grid, 2, 3600.5, variables, last_model_time=3590.25)
    return path, g, grid, variables


def test_numpy_round_trip_and_file_order(tmp_path):
    path, g, grid, variables = sample(tmp_path)
    back = R.read_restart_file(path, memory_order=True)
    assert back["grid"] == grid and back["microphysics"] == 2
    assert back["model_time"] == 3600.5 and back["last_model_time"] == 3590.25
    assert sorted(back["vars"]) == sorted(variables)
    for k, a in variables.items():
        assert np.array_equal(back["vars"][k], a) and back["vars"][k].dtype == a.dtype
    # file order is the reference's (x, y, z): output_obj.f90:423-426
    qv = R.read_restart_file(path)["vars"][icar_b200.VAR["water_vapor"]]
    assert qv.shape == (3, 5, 7)
    assert np.array_equal(qv, variables[icar_b200.VAR["water_vapor"]].transpose(1, 0, 2))


def test_library_reads_the_header(tmp_path):
    path, g, grid, variables = sample(tmp_path)
    got, mp, t, ids = icar_b200.restart_info(path)
    assert {k: getattr(got, k) for k in R.GRID_FIELDS} == grid
    assert mp == 2 and t == 3600.5 and ids == sorted(variables)


def test_library_rejects_foreign_and_truncated_files(tmp_path):
    path, *_ = sample(tmp_path)
    raw = open(path, "rb").read()
    bad = str(tmp_path / "bad.bin")
    for content, what in ((b"NETCDF.." + raw[8:], "not an icar_b200"), (raw[:60], "truncated"), (raw[:-9], "truncated"),
                          (raw[:8] + (7).to_bytes(4, "little") + raw[12:], "version")):
        open(bad, "wb").write(content)
        with pytest.raises(icar_b200.IcarError, match=what):
            icar_b200.restart_info(bad)
        with pytest.raises(ValueError):
            R.read_restart_file(bad)
    with pytest.raises(icar_b200.IcarError, match="cannot open"):
        icar_b200.restart_info(str(tmp_path / "missing.bin"))
