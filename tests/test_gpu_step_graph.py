"""icar_physics_steps: the `do while` loop of step() (time_step.f90:464-549) at a fixed dt, with pairs of steps
replayed from a CUDA graph.  The bar is bit identity with the same number of icar_physics_step calls -- the
graph only changes how the launches reach the GPU."""
import numpy as np
import pytest

from icar_b200 import ideal

pytestmark = pytest.mark.gpu

SIMPLE = ("th", "qv", "qc", "qr", "qs")
THOMPSON = SIMPLE + ("qi", "qg", "ni", "nr")
ACC = ("precipitation", "snowfall")


def make(gpu_lib, st, **opt):
    dom = gpu_lib.Domain(gpu_lib.Grid.single_image(st["nx"], st["nz"], st["ny"]), gpu_lib.Options(dx=st["dx"], **opt))
    dom.upload_state(st)
    return dom


def assert_same(a, b, fields):
    for k in fields:
        np.testing.assert_array_equal(a.download(k), b.download(k), err_msg=k)
    assert a.model_time == b.model_time
    assert a.kernel_launches == b.kernel_launches


@pytest.mark.parametrize("nsteps", [2, 7, 12])
def test_graph_replay_is_bit_identical_to_eager_steps(gpu_lib, nsteps):
    st = ideal.make_ideal_case(nx=96, nz=24, ny=40, perturb=True)
    st["qv"] *= np.float32(3.0)
    dt = float(st["dt"])
    eager, graph = make(gpu_lib, st), make(gpu_lib, st)
    for _ in range(nsteps):
        eager.physics_step(dt)
    graph.physics_steps(dt, nsteps, graph=True)
    assert_same(eager, graph, SIMPLE + ACC)
    # the captured pair is reused by a second call, and an odd tail runs eagerly
    for _ in range(5):
        eager.physics_step(dt)
    graph.physics_steps(dt, 5, graph=True)
    assert_same(eager, graph, SIMPLE + ACC)
    assert float(graph.download("precipitation").sum()) > 0


def test_graph_is_recaptured_when_dt_changes_or_a_field_is_replaced(gpu_lib):
    st = ideal.make_ideal_case(nx=64, nz=20, ny=24, perturb=True)
    st["qv"] *= np.float32(3.0)
    dt = float(st["dt"])
    eager, graph = make(gpu_lib, st), make(gpu_lib, st)
    for dom_dt, n in ((dt, 6), (0.5 * dt, 4), (dt, 3), (0.5 * dt, 6)):
        for _ in range(n):
            eager.physics_step(dom_dt)
        graph.physics_steps(dom_dt, n, graph=True)
        assert_same(eager, graph, SIMPLE + ACC)
    # an odd number of eager steps in between leaves every field in its other buffer
    eager.physics_step(dt)
    graph.physics_step(dt)
    for dom in (eager, graph):
        dom.upload("qv", st["qv"])
    for _ in range(4):
        eager.physics_step(dt)
    graph.physics_steps(dt, 4, graph=True)
    assert_same(eager, graph, SIMPLE + ACC)


def test_graph_with_update_interval_and_exact_arithmetic(gpu_lib):
    """mp() due every third step (mp_driver.f90:698-713) changes which launches a pair holds; the staged
    (exact arithmetic) advection rotates its scratch buffers with a period that need not be two steps."""
    st = ideal.make_ideal_case(nx=48, nz=16, ny=20, perturb=True)
    st["qv"] *= np.float32(3.0)
    dt = float(st["dt"])
    for opt in (dict(update_interval=int(2.5 * dt)), dict(exact_arithmetic=1), dict(microphysics=0)):
        eager, graph = make(gpu_lib, st, **opt), make(gpu_lib, st, **opt)
        for _ in range(9):
            eager.physics_step(dt)
        graph.physics_steps(dt, 9, graph=True)
        assert_same(eager, graph, SIMPLE + ACC)


def test_graph_replay_thompson(gpu_lib):
    nx, nz, ny = 48, 20, 20
    st = ideal.add_thompson_fields(ideal.make_ideal_case(nx=nx, nz=nz, ny=ny), moisten=3.0)
    for k in ACC + ("graupel",):
        st[k] = np.zeros((ny, nx), np.float64)
    dt = float(st["dt"])
    eager, graph = make(gpu_lib, st, microphysics=1), make(gpu_lib, st, microphysics=1)
    for _ in range(8):
        eager.physics_step(dt)
    graph.physics_steps(dt, 8, graph=True)
    assert_same(eager, graph, THOMPSON + ACC + ("graupel",))


def test_steps_without_graph_and_bad_arguments(gpu_lib):
    st = ideal.make_ideal_case(nx=32, nz=10, ny=16, perturb=True)
    dt = float(st["dt"])
    eager, loop = make(gpu_lib, st), make(gpu_lib, st)
    for _ in range(3):
        eager.physics_step(dt)
    loop.physics_steps(dt, 3, graph=False)
    assert_same(eager, loop, SIMPLE)
    loop.physics_steps(dt, 0)
    assert loop.model_time == eager.model_time
    with pytest.raises(gpu_lib.IcarError):
        loop.physics_steps(dt, -1)
