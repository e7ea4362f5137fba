"""icar_write_restart / icar_read_restart on the device (src/io/output_obj.f90,
src/io/restart.f90:11-90): a run continued from a file is bit-identical to one that never stopped, and the file
holds what icar_download returns, in the reference's file order."""
import os

import numpy as np
import pytest

from icar_b200 import ideal
from icar_b200 import restart as R

pytestmark = [pytest.mark.gpu]


def make(gpu_lib, st, mp, **opt):
    dom = gpu_lib.Domain(gpu_lib.Grid.single_image(st["nx"], st["nz"], st["ny"]),
                         gpu_lib.Options(dx=st["dx"], microphysics=mp, **opt))
    dom.upload_state(st)
    return dom


@pytest.mark.parametrize("mp,update_interval", [(2, 0), (2, 90), (1, 0)])
def test_continuing_from_a_file_is_bit_identical(gpu_lib, tmp_path, mp, update_interval):
    st = ideal.make_ideal_case(nx=70, nz=12, ny=22, perturb=True)      # 70: rows are padded on the device
    st["qv"] *= np.float32(3.0)
    if mp == 1:
        st = ideal.add_thompson_fields(st)
    dt = float(st["dt"])
    names = ("th", "qv", "qc", "qr", "qs") + (("qi", "ni", "nr", "qg") if mp == 1 else ())
    a = make(gpu_lib, st, mp, update_interval=update_interval)
    for _ in range(3):
        a.physics_step(dt)
    path = str(tmp_path / "restart.bin")
    a.write_restart(path)
    for _ in range(4):
        a.physics_step(dt)
    # what the file holds is what download returns (taken before the four extra steps: read it back into b)
    b = make(gpu_lib, st, mp, update_interval=update_interval)
    for k in names:                                                    # scribble, so that only the file can restore it
        b.upload(k, np.full_like(b.download(k), 9.0))
    b.read_restart(path)
    held = R.read_restart_file(path, memory_order=True)
    assert held["model_time"] == pytest.approx(3 * dt) and held["microphysics"] == mp
    for k in names:
        assert np.array_equal(held["vars"][gpu_lib.VAR[gpu_lib.ALIASES.get(k, k)]], b.download(k)), k
    for _ in range(4):
        b.physics_step(dt)
    for k in names + ("precipitation",):
        assert np.array_equal(a.download(k), b.download(k)), k


def test_a_file_of_another_grid_or_scheme_is_refused(gpu_lib, tmp_path):
    st = ideal.make_ideal_case(nx=40, nz=8, ny=12)
    a = make(gpu_lib, st, 2)
    path = str(tmp_path / "restart.bin")
    a.write_restart(path)
    other = ideal.make_ideal_case(nx=40, nz=8, ny=13)
    with pytest.raises(gpu_lib.IcarError, match="different grid"):
        make(gpu_lib, other, 2).read_restart(path)
    with pytest.raises(gpu_lib.IcarError, match="different microphysics"):
        make(gpu_lib, st, 0).read_restart(path)
