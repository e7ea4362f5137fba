"""Host-side logic of icar_b200.pipeline.HostPipeline without a GPU: the bands, the row views handed to each
band's context and the rows each band writes back, with the contexts replaced by a recorder."""
import numpy as np
import pytest

import icar_b200.pipeline as P
from icar_b200 import ideal

PER_STEP = ("u", "v", "w", "pressure", "pii", "rho", "th", "qv", "qc", "qr", "qs", "precipitation", "snowfall")
OUT = ("th", "qv", "qc", "qr", "qs", "precipitation", "snowfall")


class RecordingDomain:
    """Stands in for icar_b200.Domain: 'steps' by adding one to the rows it is asked to return."""
    calls = []

    def __init__(self, grid, options, device=0):
        self.grid, self.static = grid, {}

    def upload(self, name, array):
        self.static[name] = array.shape

    def host_domain(self, views):
        return dict(views)

    def physics_step_host_io(self, dt, hin, hout, row0, row1, derive):
        RecordingDomain.calls.append((self.grid.ny, tuple(sorted(hin)), row0, row1, derive))
        for k, v in hout.items():
            v[row0:row1 + 1] = hin[k][row0:row1 + 1] + 1

    physics_step_host_io_async = physics_step_host_io       # the ordered mode queues the bands from one thread ...

    def synchronize(self):                                  # ... and waits for each
        pass

    kernel_launches = 0

    def close(self):
        pass


@pytest.mark.parametrize("ordered", [True, False])
@pytest.mark.parametrize("bands,derive", [(3, True), (4, False), (1, False)])
def test_bands_return_exactly_their_own_rows(monkeypatch, bands, derive, ordered):
    monkeypatch.setattr(P, "Domain", RecordingDomain)
    RecordingDomain.calls = []
    st = ideal.make_ideal_case(nx=16, nz=8, ny=40, perturb=True)
    for k in ("precipitation", "snowfall"):
        st[k] = np.zeros((40, 16))
    pipe = P.HostPipeline(st, None, bands=bands, derive_diagnostics=derive, ordered=ordered)
    assert len(pipe.domains) == bands
    for dom, (a, b, v0, v1) in zip(pipe.domains, pipe.bounds):
        rows = v1 - v0 + 1
        assert dom.grid.ny == rows
        assert dom.static["jacobian_v"][0] == rows + 1 and dom.static["jacobian_u"] == (rows, 8, 17)
    src = {k: np.ascontiguousarray(st[k].copy()) for k in PER_STEP}
    dst = {k: np.full_like(st[k], -7) for k in OUT}
    pipe.step(1.0, src, dst)
    for k in OUT:                                   # every row written exactly once, by its owner
        np.testing.assert_array_equal(dst[k], src[k] + 1)
    assert len(RecordingDomain.calls) == bands
    for (rows, keys, r0, r1, d), (a, b, v0, v1) in zip(sorted(RecordingDomain.calls, key=lambda c: c[0:1]),
                                                       sorted(pipe.bounds, key=lambda bb: bb[3] - bb[2] + 1)):
        assert d == derive
        assert ("pii" in keys) == (not derive) and "jacobian" not in keys
    with pytest.raises(ValueError):                 # in-place is not allowed: neighbours read these rows as ghosts
        pipe.step(1.0, src, {k: src[k] for k in OUT})
    pipe.close()
