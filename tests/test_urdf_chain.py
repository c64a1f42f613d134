"""pik_urdf_chain (host-only C-ABI): URDF -> chain table with urdfdom / MoveIt semantics."""
import ctypes as C
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


def test_urdf_tree_round_trips_the_tree_fixtures():
    """pik_urdf_tree: several tips, mimic joints, parents-first order -- the arguments of pik_robot_create_tree."""
    for name in ("two_arm", "three_tip", "mimic_arm"):
        tree = robots.TREES[name]()
        xml, tips = robots.tree_to_urdf(tree)
        got = capi.urdf_tree(xml, "base", tips)
        desc, parent, tip_joint, mimic_of, factor, offset = tree.tree_arrays()
        # the reader walks the tree depth-first (children in document order): map its joints back by name
        index = {j.name: k for k, j in enumerate(tree.joints)}
        order = [index[n] for n in got["joint_names"]]
        assert sorted(order) == list(range(len(tree.joints))), name
        back = {old: new for new, old in enumerate(order)}
        for new, old in enumerate(order):
            for f in desc.dtype.names:
                if f == "axis" and desc[old]["type"] == 0:
                    continue  # a fixed joint has no <axis>
                np.testing.assert_array_equal(got["desc"][new][f], desc[old][f], err_msg=f"{name} {got['joint_names'][new]} {f}")
            assert got["parent"][new] == (back[parent[old]] if parent[old] >= 0 else -1)
            assert got["mimic_of"][new] == (back[mimic_of[old]] if mimic_of[old] >= 0 else -1)
            if mimic_of[old] >= 0:
                assert (got["mimic_factor"][new], got["mimic_offset"][new]) == (factor[old], offset[old])
        assert [order[t] for t in got["tip_joint"]] == list(tip_joint)
        # and the robot built from it has the variables of the fixture (joint order within a depth may differ from the
        # fixture's only where branches interleave, which these fixtures do not do for their moving joints)
        h = C.c_void_p()
        vp = C.c_void_p
        rc = capi.lib().pik_robot_create_tree(got["desc"].ctypes.data_as(vp), len(got["desc"]), got["parent"].ctypes.data_as(vp),
                                              got["tip_joint"].ctypes.data_as(vp), len(tips), got["mimic_of"].ctypes.data_as(vp),
                                              got["mimic_factor"].ctypes.data_as(vp), got["mimic_offset"].ctypes.data_as(vp), C.byref(h))
        assert rc == 0
        assert capi.lib().pik_robot_num_variables(h) == tree.num_variables
        assert capi.lib().pik_robot_num_tips(h) == tree.num_tips
        capi.lib().pik_robot_destroy(h)


def test_urdf_tree_floating_planar_and_errors():
    for name in ("floating_arm", "planar_arm"):
        tree = robots.TREES[name]()
        xml, tips = robots.tree_to_urdf(tree)
        got = capi.urdf_tree(xml, "base", tips)
        assert got["desc"][0]["type"] == (3 if name == "floating_arm" else 4) and got["desc"][0]["bounded"] == 0
        # the single-tip chain reader has no multi-variable joints
        with pytest.raises(capi.PikError) as e:
            capi.urdf_chain(xml, "base", tips[0])
        assert e.value.status == -7
    xml, tips = robots.tree_to_urdf(robots.mimic_arm())
    with pytest.raises(capi.PikError) as e:
        capi.urdf_chain(xml, "base", tips[0])  # a mimic joint on the chain: use pik_urdf_tree
    assert e.value.status == -7
    with pytest.raises(capi.PikError) as e:
        capi.urdf_tree(xml, "base", ["no_such_link"])
    assert e.value.status == -2
    # a name that does not fit PIK_URDF_NAME_BYTES is an error, not a truncation
    long_name = "j" * 80
    doc = (f'<robot name="r"><link name="a"/><link name="b"/><joint name="{long_name}" type="revolute"><parent link="a"/>'
           '<child link="b"/><axis xyz="0 0 1"/><limit lower="-1" upper="1" velocity="1" effort="1"/></joint></robot>')
    with pytest.raises(capi.PikError) as e:
        capi.urdf_chain(doc, "a", "b")
    assert e.value.status == -1


def test_srdf_group_chains():
    srdf = """<?xml version="1.0"?>
    <robot name="two_arm">
      <!-- planning groups -->
      <group name="left_arm"><chain base_link="base" tip_link="l_tool_link"/></group>
      <group name="right_arm"><chain base_link="base" tip_link="r_tool_link"/></group>
      <group name="both_arms"><group name="left_arm"/><group name="right_arm"/></group>
      <group name="by_joints"><joint name="torso"/></group>
      <end_effector name="l" parent_link="l_tool_link" group="left_arm"/>
    </robot>"""
    assert capi.srdf_group(srdf, "left_arm") == ("base", ["l_tool_link"])
    base, tips = capi.srdf_group(srdf, "both_arms")
    assert base == "base" and sorted(tips) == ["l_tool_link", "r_tool_link"]
    for group, status in (("by_joints", -7), ("no_such_group", -2)):
        with pytest.raises(capi.PikError) as e:
            capi.srdf_group(srdf, group)
        assert e.value.status == status
    # SRDF group -> URDF tree -> robot: the stand-alone replacement of RobotModel + JointModelGroup
    tree = robots.two_arm()
    xml, _ = robots.tree_to_urdf(tree)
    got = capi.urdf_tree(xml, base, tips)
    assert len(got["desc"]) == len(tree.joints) and len(got["tip_joint"]) == 2
