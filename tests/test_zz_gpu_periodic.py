"""GPU tier, translational periodic boundaries (the device code is also checked on the CPU
tier against the oracle and the reference's own BC code): one generation
in a box with periodic x planes, periodic oblique planes and reflective z planes."""
import numpy as np
import pytest

import reference_cases as rc
from oracle.binding import Oracle

pytestmark = pytest.mark.gpu


def test_periodic_box_generation_matches_oracle(gpu_engine, monkeypatch):
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", "2048")
    m, model, box = rc.build("periodic_box")
    o = Oracle(model)
    n = 20000
    src = o.sample_initial_source(n, box, False)
    gpu_engine.init(model, 0, n)
    gi = gpu_engine.gen_in(n, 0, 0, 1.0, False)
    fb, out = gpu_engine.transport_generation(gi, src)
    fb_o, out_o = o.transport_generation(gi, src)
    assert out.n_lost == 0 and out.global_tally[3] == 0.0      # nothing leaks from a periodic / reflective box
    assert abs(out.n_fission - out_o.n_fission) <= 5
    assert abs(out.n_events - out_o.n_events) <= 2e-4 * out_o.n_events
    np.testing.assert_allclose(out.global_tally[:3], out_o.global_tally[:3], rtol=2e-4)
    gpu_engine.finalize()
