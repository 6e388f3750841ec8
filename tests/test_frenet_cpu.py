"""Host-side logic of dcarl_b200.frenet that needs no device (the sample grids, the route ingest) and the
test-side numpy restatement of the ego projection, against the reference's golden run."""
import numpy as np
import pytest

from dcarl_b200 import frenet as fr
from oracle import cport
from helpers_frenet import frenet_state, params_from_golden


@pytest.fixture(scope="module")
def g(golden):
    return golden("frenet")


def test_params_are_the_reference_grids_and_match_the_oracle_struct(g):
    p, q = params_from_golden(g, fr.frenet_params), params_from_golden(g, cport.frenet_params)
    assert bytes(p) == bytes(q)                                   # same layout, same values as the oracle's mirror
    assert list(p.d_samples[:p.n_d]) == [-4.0, -2.0, 0.0, 2.0, 4.0]
    assert list(p.tv_samples[:p.n_v]) == np.arange(30 / 3.6 - 15 / 3.6, 30 / 3.6 + 15 / 3.6, 15 / 3.6).tolist()
    assert (p.n_t[0], p.n_pred) == (14, 15)
    with pytest.raises(ValueError):
        fr.frenet_params(D_ROAD_W=0.25, MAX_LEFT_WIDTH=-8, MAX_RIGHT_WIDTH=8)


def test_route_ingest_and_start_state_match_the_reference(g):
    for i in range(int(g["n_cases"])):
        nl, nk = int(g["n_lane"][i]), int(g["n_knots"][i])
        dense = fr.dense_polyline2d(g["lane_xy"][i, :nl], 2)
        assert dense.shape == (nk, 2)
        assert np.array_equal(dense[:, 0], g["knots_x"][i, :nk]) and np.array_equal(dense[:, 1], g["knots_y"][i, :nk])
        s, d, vd = frenet_state(g["ego_xy"][i, 0], g["ego_xy"][i, 1], g["ego_v"][i, 0], g["ego_v"][i, 1], dense)
        assert np.allclose([s, -d, vd], g["start"][i, :3], rtol=0, atol=1e-12)
