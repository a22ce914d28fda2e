"""Pin the CPU oracle (oracle/ttrl_oracle.c) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O
from topotrafficrl_b200 import abi
from tests import common as T


@pytest.fixture(scope="module")
def kat():
    return T.golden("kat_functions.npz")


@pytest.fixture(scope="module")
def inter():
    net, table, cfg, routes = T.intersection_scene()
    return O.Oracle(cfg, table, routes), table


def test_scalar_helpers(kat):
    for x, w, nz in zip(kat["wrap_in"], kat["wrap_out"], kat["not_zero_out"]):
        assert abs(O.wrap_to_pi(float(x)) - w) <= 1e-15
        assert O.not_zero(float(x)) == nz


def test_lane_geometry(kat, inter):
    orc, table = inter
    for li in range(table.n_lanes):
        for k, (p, h, s, r) in enumerate(zip(kat["lane_pts"], kat["lane_h"], kat["lane_s"], kat["lane_r"])):
            np.testing.assert_allclose(orc.lane_local(li, *p), kat["lane_local"][li, k], rtol=0, atol=1e-12)
            np.testing.assert_allclose(orc.lane_position(li, s, r), kat["lane_position"][li, k], rtol=0, atol=1e-12)
            assert abs(orc.lane_heading_at(li, s) - kat["lane_heading"][li, k]) <= 1e-14
            assert abs(orc.lane_distance_with_heading(li, p[0], p[1], h) - kat["lane_dwh"][li, k]) <= 1e-12
    got = [orc.closest_lane(p[0], p[1], h) for p, h in zip(kat["lane_pts"], kat["lane_h"])]
    assert (np.array(got) == kat["closest"]).all()


def test_sine_lane(kat):
    from topotrafficrl_b200.road import RoadNetwork, SineLane
    from topotrafficrl_b200 import scenes
    p = kat["sine_params"]
    net = RoadNetwork()
    net.add_lane("a", "b", SineLane(p[0:2], p[2:4], p[4], p[5], p[6], speed_limit=15))
    table = net.to_table()
    cfg = scenes.build_config(table, scenes.merged_config(scenes.HIGHWAY_CONFIG, None), "highway")
    orc = O.Oracle(cfg, table)
    for k, (pt, s, r) in enumerate(zip(kat["lane_pts"], kat["lane_s"], kat["lane_r"])):
        np.testing.assert_allclose(orc.lane_local(0, *pt), kat["sine_local"][k], rtol=0, atol=1e-12)
        np.testing.assert_allclose(orc.lane_position(0, s, r), kat["sine_position"][k], rtol=0, atol=1e-12)
        assert abs(orc.lane_heading_at(0, s) - kat["sine_heading"][k]) <= 1e-14


def test_controllers(kat, inter):
    orc, _ = inter
    for row, want in zip(kat["steer_in"], kat["steer_out"]):
        assert abs(orc.steering_control(row[0], row[1], row[2], row[3], int(row[4])) - want) <= 1e-12
    got = [orc.speed_to_index(float(s)) for s in kat["s2i_in"]]
    assert (np.array(got) == kat["s2i_out"]).all()
    for row, want in zip(kat["idm_in"], kat["idm_out"]):
        got = orc.idm_acceleration(row[0], row[1:7], row[7:13] if row[13] > 0.5 else None)
        assert abs(got - want) <= 1e-9 * max(1.0, abs(want))


def test_collision_and_regulation_geometry(kat, inter):
    orc, _ = inter
    for row, want in zip(kat["sat_in"], kat["sat_out"]):
        inter_, will, t = orc.polygons_intersecting(row[0:4], row[4:8], 1 / 15)
        assert inter_ == bool(want[0]) and will == bool(want[1])
        np.testing.assert_allclose(t, want[2:4], rtol=0, atol=1e-12)
    for row, want in zip(kat["rect_in"], kat["rect_out"]):
        assert orc.rotated_rectangles_intersect(row[0:5], row[5:10]) == bool(want)


@pytest.mark.parametrize("name,scene", [
    ("intersection_substeps.npz", "intersection"),
    ("highway_n8_substeps.npz", "highway8"),
    ("highway_n50_substeps.npz", "highway50"),
    ("highway_n200_substeps.npz", "highway200"),
    ("highway_grid_n40_substeps.npz", "highway40"),
])
def test_substep_resynced(name, scene):
    """Inject every golden 'before' state, run ONE oracle sub-step, compare with the reference's 'after'."""
    g = T.golden(name)
    if scene == "intersection":
        _, table, cfg, routes = T.intersection_scene()
        orc = O.Oracle(cfg, table, routes)
    else:
        _, table, cfg, _ = T.highway_scene(int(scene[7:]))
        orc = O.Oracle(cfg, table)
    st = T.batch_state(g, "before")
    want = T.batch_state(g, "after")
    orc.substep(st, g["action"].astype(np.int32))
    T.compare_states(st, want, T.TOL_SUBSTEP, name)


@pytest.mark.parametrize("name,over", [
    ("intersection_steps_kin.npz", None),
    ("intersection_steps_grid_dense.npz", T.GRID_DENSE),
    ("intersection_steps_grid_road.npz", T.GRID_ROAD),
])
def test_intersection_step(name, over):
    """Full env.step(): 15 sub-steps + obs + reward + flags + clear + spawn (reference draws injected)."""
    g = T.golden(name)
    _, table, cfg, routes = T.intersection_scene(over)
    orc = O.Oracle(cfg, table, routes)
    st = T.batch_state(g, "before")
    want = T.batch_state(g, "after")
    obs, reward, term, trunc, _ = orc.step(st, g["action"].astype(np.int32), T.draws_array(g["draw"]))
    T.compare_states(st, want, T.TOL_STEP, name)
    np.testing.assert_allclose(obs.reshape(g["obs"].shape), g["obs"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(reward, g["reward"], rtol=0, atol=1e-6)
    assert (term.astype(bool) == g["terminated"]).all() and (trunc.astype(bool) == g["truncated"]).all()
    np.testing.assert_allclose(st.env_d[abi.ED_TIME], want.env_d[abi.ED_TIME])


@pytest.mark.parametrize("name,n,over", [
    ("highway_n8_steps.npz", 8, None),
    ("highway_n50_steps.npz", 50, None),
    ("highway_grid_n40_steps.npz", 40, T.HIGHWAY_GRID),
])
def test_highway_step(name, n, over):
    g = T.golden(name)
    _, table, cfg, _ = T.highway_scene(n, overrides=over)
    orc = O.Oracle(cfg, table)
    st = T.batch_state(g, "before")
    want = T.batch_state(g, "after")
    obs, reward, term, trunc, _ = orc.step(st, g["action"].astype(np.int32))
    T.compare_states(st, want, T.TOL_STEP, name)
    np.testing.assert_allclose(obs.reshape(g["obs"].shape), g["obs"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(reward, g["reward"], rtol=0, atol=1e-6)
    assert (term.astype(bool) == g["terminated"]).all() and (trunc.astype(bool) == g["truncated"]).all()
