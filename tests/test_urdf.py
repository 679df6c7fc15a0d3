"""URDF / SRDF ingestion (simple-mpc_b200/python/smpc_urdf.py): what a maintainer needs to run this path on the real Go2 /
Talos of example-robot-data (reference tests/test_utils.cpp:14-97, examples/utils.py:10-31). The real files are absent here
(SURVEY.md §8c), so the reader is pinned by round trips of the synthetic robots and by hand-computed known answers."""
import numpy as np
import pytest
import smpc_urdf as U
from smpc_ctypes import go2_like_model, talos_like_model, fk_numpy, total_mass, com_numpy


@pytest.mark.parametrize("which", ["go2", "talos"])
def test_model_desc_round_trips_through_urdf(which):
    m, lim = go2_like_model() if which == "go2" else talos_like_model()
    feet = ["FL_foot", "FR_foot", "RL_foot", "RR_foot"] if which == "go2" else ["left_sole_link", "right_sole_link"]
    urdf, srdf = U.model_desc_to_urdf(m, feet, limits=lim)
    d = U.describe(urdf, srdf)
    m2, lim2 = U.to_model_desc(d, feet, "base", "reference")
    assert (m2.njoints, m2.nq, m2.nv, m2.nfeet) == (m.njoints, m.nq, m.nv, m.nfeet)
    rng = np.random.default_rng(0)
    for _ in range(3):
        q = np.array(m.q_ref[: m.nq]) + np.concatenate([rng.normal(size=3) * 0.1, np.zeros(4), rng.normal(size=m.nq - 7) * 0.3])
        quat = q[3:7] + rng.normal(size=4) * 0.1
        q[3:7] = quat / np.linalg.norm(quat)
        a, b = fk_numpy(m, q), fk_numpy(m2, q)
        for (Ra, pa), (Rb, pb) in zip(a["joints"] + a["feet"], b["joints"] + b["feet"]):
            assert np.abs(Ra - Rb).max() < 1e-12 and np.abs(pa - pb).max() < 1e-12
        assert np.abs(com_numpy(m, q) - com_numpy(m2, q)).max() < 1e-12
    assert abs(total_mass(m) - total_mass(m2)) < 1e-12
    for i in range(m.njoints):
        assert np.abs(np.array(m.inertia[i][:]) - np.array(m2.inertia[i][:])).max() < 1e-12
    assert np.allclose(lim["qmin"], lim2["qmin"]) and np.allclose(lim["qmax"], lim2["qmax"])
    # "<foot>_ref" frames sit under the base where the feet are at the reference configuration
    for f in range(m.nfeet):
        assert np.abs(np.array(m.footref_p[f][:]) - np.array(m2.footref_p[f][:])).max() < 1e-12


URDF = """<?xml version="1.0"?>
<robot name="toy">
  <link name="trunk"><inertial><origin xyz="0.1 0 0" rpy="0 0 0"/><mass value="2"/><inertia ixx="1" ixy="0" ixz="0" iyy="2" iyz="0" izz="3"/></inertial></link>
  <link name="imu"><inertial><origin xyz="0 0 0" rpy="0 0 0"/><mass value="1"/><inertia ixx="0.1" ixy="0" ixz="0" iyy="0.1" iyz="0" izz="0.1"/></inertial></link>
  <link name="b_leg"><inertial><origin xyz="0 0 -0.5" rpy="0 0 0"/><mass value="1"/><inertia ixx="0.2" ixy="0" ixz="0" iyy="0.2" iyz="0" izz="0.01"/></inertial></link>
  <link name="a_leg"/>
  <link name="a_foot"/>
  <joint name="imu_joint" type="fixed"><parent link="trunk"/><child link="imu"/><origin xyz="-0.2 0 0" rpy="0 0 0"/></joint>
  <joint name="zz_hip" type="revolute"><parent link="trunk"/><child link="b_leg"/><origin xyz="0 0.1 0" rpy="0 0 1.5707963267948966"/><axis xyz="0 1 0"/><limit lower="-1" upper="2" effort="30" velocity="1"/></joint>
  <joint name="aa_hip" type="revolute"><parent link="trunk"/><child link="a_leg"/><origin xyz="0 -0.1 0" rpy="0 0 0"/><axis xyz="1 0 0"/><limit lower="-0.5" upper="0.5" effort="10" velocity="1"/></joint>
  <joint name="a_foot_fixed" type="fixed"><parent link="a_leg"/><child link="a_foot"/><origin xyz="0 0 -1" rpy="0 0 0"/></joint>
</robot>"""
SRDF = """<robot name="toy"><group_state name="standing" group="all">
  <joint name="root_joint" value="0 0 1 0 0 0 1"/><joint name="aa_hip" value="0.3"/><joint name="zz_hip" value="-0.2"/></group_state></robot>"""


def test_known_answers_of_a_hand_written_urdf():
    d = U.describe(URDF, SRDF)
    J = d["joints"]
    # joint order: children visited in alphabetical order of their joint names (urdfdom's std::map), fixed joints folded away
    assert [j["name"] for j in J] == ["root_joint", "aa_hip", "zz_hip"]
    # the fixed imu link is merged into the trunk body: mass 3, centre of mass (2 * 0.1 + 1 * (-0.2)) / 3 = 0
    assert abs(J[0]["mass"] - 3.0) < 1e-15 and np.abs(J[0]["lever"]).max() < 1e-15
    # parallel-axis: Izz = 3 + 2 * 0.1^2 + 0.1 + 1 * 0.2^2
    assert abs(J[0]["inertia"][2, 2] - (3 + 0.02 + 0.1 + 0.04)) < 1e-14
    assert np.allclose(J[2]["R"], [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-15) and np.allclose(J[2]["p"], [0, 0.1, 0])
    assert (J[1]["lower"], J[1]["upper"], J[1]["effort"]) == (-0.5, 0.5, 10.0)
    jidx, R, p = d["frames"]["a_foot"]
    assert jidx == 1 and np.allclose(p, [0, 0, -1])
    assert d["frames"]["imu"][0] == 0 and np.allclose(d["frames"]["imu"][2], [-0.2, 0, 0])
    q = d["reference_configurations"]["standing"]
    assert np.allclose(q, [0, 0, 1, 0, 0, 0, 1, 0.3, -0.2])
    # locking a joint at its reference value folds it away like buildReducedRobot (reference examples/utils.py:14-27)
    dl = U.describe(URDF, SRDF, lock=["zz_hip"], lock_state="standing")
    assert [j["name"] for j in dl["joints"]] == ["root_joint", "aa_hip"]
    assert abs(dl["joints"][0]["mass"] - 4.0) < 1e-15
    # the locked leg's centre of mass: R_origin * Ry(-0.2) * (0, 0, -0.5) + origin
    c = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]]) @ U.axis_angle_R([0, 1, 0], -0.2) @ np.array([0, 0, -0.5]) + np.array([0, 0.1, 0])
    assert np.allclose(dl["joints"][0]["lever"] * 4.0, 1.0 * c, atol=1e-14)
    with pytest.raises(ValueError):
        U.describe(URDF.replace('type="revolute"><parent link="trunk"/><child link="a_leg"', 'type="prismatic"><parent link="trunk"/><child link="a_leg"'))


def test_urdf_model_drives_the_host_api():
    """the parsed robot builds a simple_mpc.Model that RobotModelHandler accepts (the path a maintainer with the real
    example-robot-data files takes)"""
    import host_api_utils as hu
    sm = hu.import_simple_mpc()
    m, lim = go2_like_model()
    feet = ["FL_foot", "FR_foot", "RL_foot", "RR_foot"]
    urdf, srdf = U.model_desc_to_urdf(m, feet, limits=lim)
    model = U.to_model(sm, U.describe(urdf, srdf))
    assert model.nq == 19 and model.nv == 18
    mh = sm.RobotModelHandler(model, "reference", "base")
    for n in feet:
        mh.addPointFoot(n, "base")
    ref = sm.RobotModelHandler(sm.Model.BuildGo2Like(), "standing", "base")
    assert abs(mh.getMass() - ref.getMass()) < 1e-12
    assert np.abs(np.array(mh.getReferenceState()) - np.array(ref.getReferenceState())).max() < 1e-12
