"""pik_urdf_chain (host-only C-ABI): URDF -> chain table with urdfdom / MoveIt semantics."""
import math

import numpy as np
import pytest

from pick_ik_b200 import capi, robots


@pytest.mark.parametrize("name", sorted(robots.ROBOTS))
def test_generated_urdf_round_trips_bit_exact(name):
    chain = robots.ROBOTS[name]()
    desc, names = capi.urdf_chain(robots.to_urdf(chain), chain.base_link, chain.tip_link)
    ref = chain.joint_desc()
    assert names == [j.name for j in chain.joints]
    assert len(desc) == len(ref)
    moving = ref["type"] != 0
    for field in ref.dtype.names:
        if field == "axis":  # a fixed joint has none
            np.testing.assert_array_equal(desc[field][moving], ref[field][moving], err_msg=f"{name}: {field}")
        else:
            np.testing.assert_array_equal(desc[field], ref[field], err_msg=f"{name}: {field}")
    # and the table pik_robot_create derives from it is the same
    a, b = capi.Robot(desc), capi.Robot(chain)
    assert a.n == b.n
    for i in range(a.n):
        va, vb = a.variable(i), b.variable(i)
        for f, _ in capi.Variable._fields_:
            assert getattr(va, f) == getattr(vb, f)


URDF = """<?xml version='1.0' encoding="utf-8"?>
<!-- a comment with <joint name="ghost" type="revolute"> inside -->
<robot name="toy" xmlns:xacro="http://www.ros.org/wiki/xacro">
  <material name="grey"><color rgba="0.5 0.5 0.5 1"/></material>
  <link name="world"/>
  <link name="base"><visual><origin xyz="9 9 9"/><geometry><box size="1 1 1"/></geometry></visual></link>
  <link name="l1"/><link name="l2"/><link name="l3"/><link name="tool"/><link name="finger"/>
  <joint name="mount" type="fixed"><parent link="world"/><child link="base"/><origin xyz="0 0 0.5"/></joint>
  <joint name="shoulder" type="revolute">
    <origin rpy="0 0 1.5" xyz="0.1 0 0.2"/>
    <parent link="base"/> <child link="l1"/>
    <axis xyz="0 0 2"/>
    <limit lower="-2.0" upper="2.0" velocity="-1.5" effort="3"/>
    <safety_controller soft_lower_limit="-1.75" soft_upper_limit="2.5" k_position="1" k_velocity="1"/>
    <dynamics damping="0.1"/>
  </joint>
  <joint name='slide' type='prismatic'><parent link='l1'/><child link='l2'/>
    <limit lower='0' upper='0.3' velocity='0.2' effort='1'/></joint>
  <joint name="wrist" type="continuous"><parent link="l2"/><child link="l3"/><axis xyz="0 1 0"/>
    <limit velocity="4" effort="1"/></joint>
  <joint name="flange" type="fixed"><parent link="l3"/><child link="tool"/><origin rpy="0.1 0.2 0.3"/></joint>
  <joint name="finger_joint" type="prismatic"><parent link="tool"/><child link="finger"/>
    <limit lower="0" upper="0.04" velocity="0.1" effort="1"/><mimic joint="slide"/></joint>
  <transmission name="t"><joint name="shoulder"><hardwareInterface>x</hardwareInterface></joint></transmission>
  <gazebo reference="l1"><joint name="nested" type="revolute"/></gazebo>
</robot>
"""


def test_hand_written_urdf_semantics():
    desc, names = capi.urdf_chain(URDF, "world", "tool")
    assert names == ["mount", "shoulder", "slide", "wrist", "flange"]
    assert list(desc["type"]) == [0, 1, 2, 1, 0]
    sh = desc[1]
    np.testing.assert_array_equal(sh["axis"], [0.0, 0.0, 1.0])  # normalised
    assert (sh["min_position"], sh["max_position"]) == (-1.75, 2.0)  # soft limits intersected with <limit>
    assert sh["max_velocity"] == 1.5 and sh["bounded"] == 1
    np.testing.assert_array_equal(sh["origin_R"], robots.rpy_to_matrix(0, 0, 1.5).reshape(9))
    np.testing.assert_array_equal(sh["origin_t"], [0.1, 0.0, 0.2])
    sl = desc[2]
    np.testing.assert_array_equal(sl["axis"], [1.0, 0.0, 0.0])  # URDF default axis
    np.testing.assert_array_equal(sl["origin_R"], np.eye(3).reshape(9))
    wr = desc[3]
    assert wr["bounded"] == 0 and (wr["min_position"], wr["max_position"]) == (-math.pi, math.pi)
    np.testing.assert_array_equal(desc[4]["origin_R"], robots.rpy_to_matrix(0.1, 0.2, 0.3).reshape(9))
    # a sub-chain, and the Robot built from it
    sub, sub_names = capi.urdf_chain(URDF, "base", "l3")
    assert sub_names == ["shoulder", "slide", "wrist"]
    assert capi.Robot(sub).n == 3


def test_errors():
    with pytest.raises(capi.PikError) as e:
        capi.urdf_chain(URDF, "world", "finger")  # mimic joint on the chain
    assert e.value.status == -7
    with pytest.raises(capi.PikError) as e:
        capi.urdf_chain(URDF, "l2", "l1")  # tip is not below base
    assert e.value.status == -2
    with pytest.raises(capi.PikError) as e:
        capi.urdf_chain(URDF.replace('type="continuous"', 'type="floating"'), "world", "tool")
    assert e.value.status == -7
    with pytest.raises(capi.PikError) as e:
        capi.urdf_chain(URDF.replace("</robot>", ""), "world", "tool")  # unbalanced document
    assert e.value.status == -2
    with pytest.raises(capi.PikError) as e:
        capi.urdf_chain(URDF.replace('lower="-2.0"', 'lower="abc"'), "world", "tool")
    assert e.value.status == -2
    desc, _ = capi.urdf_chain(URDF, "tool", "tool")  # empty chain
    assert len(desc) == 0
