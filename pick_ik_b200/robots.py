"""Robot chain fixtures for the batched IK engine.

Each robot is a serial chain of joints from the model root to the tip link, in the layout of
``pik_joint_desc`` (include/pik.h).  The reference gets the same information from a MoveIt
``RobotModel`` (``Robot::from``, /root/reference/src/robot.cpp:44-85 and
``get_active_variable_indices``, robot.cpp:122-160); there is no MoveIt/URDF here, so the
chains of the robots the reference's tests and BASELINE.json's configs name are tabulated
(SURVEY.md Appendix C).

rpy -> rotation follows urdfdom ``Rotation::setFromRPY`` (quaternion, normalised) and Eigen
``Quaterniond::toRotationMatrix`` as MoveIt does when it builds joint origin transforms
(SURVEY.md Appendix B.3).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

JOINT_FIXED = 0
JOINT_REVOLUTE = 1
JOINT_PRISMATIC = 2
JOINT_FLOATING = 3  # 7 variables: x y z, quaternion x y z w (/root/reference/src/forward_kinematics.cpp:64-70)
JOINT_PLANAR = 4    # 3 variables: x y theta (forward_kinematics.cpp:71-79)

# numpy mirror of pik_joint_desc / orc_joint_desc (natural C alignment, 176 bytes)
JOINT_DESC_DTYPE = np.dtype(
    [
        ("type", np.int32),
        ("bounded", np.int32),
        ("origin_R", np.float64, (9,)),
        ("origin_t", np.float64, (3,)),
        ("axis", np.float64, (3,)),
        ("min_position", np.float64),
        ("max_position", np.float64),
        ("max_velocity", np.float64),
    ],
    align=True,
)


def rpy_to_matrix(r: float, p: float, y: float) -> np.ndarray:
    """urdfdom setFromRPY -> quaternion -> Eigen toRotationMatrix (row-major 3x3)."""
    phi, the, psi = r / 2.0, p / 2.0, y / 2.0
    qx = math.sin(phi) * math.cos(the) * math.cos(psi) - math.cos(phi) * math.sin(the) * math.sin(psi)
    qy = math.cos(phi) * math.sin(the) * math.cos(psi) + math.sin(phi) * math.cos(the) * math.sin(psi)
    qz = math.cos(phi) * math.cos(the) * math.sin(psi) - math.sin(phi) * math.sin(the) * math.cos(psi)
    qw = math.cos(phi) * math.cos(the) * math.cos(psi) + math.sin(phi) * math.sin(the) * math.sin(psi)
    nrm = math.sqrt(qx * qx + qy * qy + qz * qz + qw * qw)
    qx, qy, qz, qw = qx / nrm, qy / nrm, qz / nrm, qw / nrm
    return quat_to_matrix(qw, qx, qy, qz)


def quat_to_matrix(w: float, x: float, y: float, z: float) -> np.ndarray:
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array(
        [
            [1.0 - (tyy + tzz), txy - twz, txz + twy],
            [txy + twz, 1.0 - (txx + tzz), tyz - twx],
            [txz - twy, tyz + twx, 1.0 - (txx + tyy)],
        ],
        dtype=np.float64,
    )


@dataclass
class Joint:
    name: str
    type: int
    xyz: Sequence[float] = (0.0, 0.0, 0.0)
    rpy: Sequence[float] = (0.0, 0.0, 0.0)
    axis: Sequence[float] = (0.0, 0.0, 1.0)
    lower: float = 0.0
    upper: float = 0.0
    velocity: float = 0.0
    continuous: bool = False  # URDF "continuous": position_bounded_ = false, bounds -pi..pi
    # kinematic trees (RobotTree): index of the joint whose child link this joint hangs on (-1: the model root)
    parent: int = -1
    # mimic joints: index of the joint followed, value = mimic_factor * master + mimic_offset
    mimic_of: int = -1
    mimic_factor: float = 1.0
    mimic_offset: float = 0.0


@dataclass
class RobotChain:
    name: str
    base_link: str
    tip_link: str
    joints: List[Joint] = field(default_factory=list)

    @property
    def variable_names(self) -> List[str]:
        return [j.name for j in self.joints if j.type != JOINT_FIXED]

    @property
    def num_variables(self) -> int:
        return len(self.variable_names)

    def joint_desc(self) -> np.ndarray:
        return joints_to_desc(self.joints)


def joints_to_desc(joints: Sequence[Joint]) -> np.ndarray:
    out = np.zeros(len(joints), dtype=JOINT_DESC_DTYPE)
    for i, j in enumerate(joints):
        out[i]["type"] = j.type
        out[i]["origin_R"] = rpy_to_matrix(*j.rpy).reshape(9)
        out[i]["origin_t"] = np.asarray(j.xyz, dtype=np.float64)
        ax = np.asarray(j.axis, dtype=np.float64)
        if j.type in (JOINT_REVOLUTE, JOINT_PRISMATIC):
            ax = ax / math.sqrt(float(ax @ ax))  # RevoluteJointModel::setAxis normalises
        out[i]["axis"] = ax
        if j.type == JOINT_FIXED:
            continue
        if j.continuous:
            out[i]["bounded"] = 0
            out[i]["min_position"] = -math.pi
            out[i]["max_position"] = math.pi
        else:
            out[i]["bounded"] = 1
            out[i]["min_position"] = j.lower
            out[i]["max_position"] = j.upper
        out[i]["max_velocity"] = abs(j.velocity)
    return out


VARIABLES_OF = {JOINT_FIXED: 0, JOINT_REVOLUTE: 1, JOINT_PRISMATIC: 1, JOINT_FLOATING: 7, JOINT_PLANAR: 3}


@dataclass
class RobotTree:
    """A kinematic tree with several tip links (one goal pose per tip, /root/reference/src/goal.cpp:80-89,163-175),
    floating / planar joints and mimic joints: the arguments of pik_robot_create_tree / orc_robot_build_tree.
    Joints are listed parents first; Joint.parent is the index of the joint above (-1: the model root)."""
    name: str
    joints: List[Joint] = field(default_factory=list)
    tip_joints: List[int] = field(default_factory=list)

    @property
    def num_variables(self) -> int:
        return sum(VARIABLES_OF[j.type] for j in self.joints if j.mimic_of < 0)

    @property
    def num_tips(self) -> int:
        return len(self.tip_joints)

    def joint_desc(self) -> np.ndarray:
        return joints_to_desc(self.joints)

    def tree_arrays(self):
        parent = np.array([j.parent for j in self.joints], dtype=np.int32)
        tips = np.array(self.tip_joints, dtype=np.int32)
        mimic_of = np.array([j.mimic_of for j in self.joints], dtype=np.int32)
        factor = np.array([j.mimic_factor for j in self.joints], dtype=np.float64)
        offset = np.array([j.mimic_offset for j in self.joints], dtype=np.float64)
        return self.joint_desc(), parent, tips, mimic_of, factor, offset


def to_urdf(chain: RobotChain) -> str:
    """The chain as a URDF document (links named after their parent joints; the decimal literals are
    repr() round-trips, so pik_urdf_chain reproduces joint_desc() bit for bit).  Test fixture generator
    for the C-ABI URDF reader -- the reference's robots live in URDFs this container does not have."""
    def f(v):
        return " ".join(repr(float(x)) for x in v)

    lines = ['<?xml version="1.0"?>', f'<robot name="{chain.name}">', f'  <link name="{chain.base_link}"/>']
    parent = chain.base_link
    for k, j in enumerate(chain.joints):
        child = chain.tip_link if k == len(chain.joints) - 1 else f"{j.name}_link"
        kind = {JOINT_FIXED: "fixed", JOINT_PRISMATIC: "prismatic"}.get(j.type, "continuous" if j.continuous else "revolute")
        lines.append(f'  <link name="{child}"/>')
        lines.append(f'  <joint name="{j.name}" type="{kind}">')
        lines.append(f'    <parent link="{parent}"/>')
        lines.append(f'    <child link="{child}"/>')
        lines.append(f'    <origin xyz="{f(j.xyz)}" rpy="{f(j.rpy)}"/>')
        if j.type != JOINT_FIXED:
            lines.append(f'    <axis xyz="{f(j.axis)}"/>')
            if j.continuous:
                lines.append(f'    <limit effort="10" velocity="{float(j.velocity)!r}"/>')
            else:
                lines.append(f'    <limit effort="10" lower="{float(j.lower)!r}" upper="{float(j.upper)!r}" '
                             f'velocity="{float(j.velocity)!r}"/>')
        lines.append("  </joint>")
        parent = child
    lines.append("</robot>")
    return "\n".join(lines) + "\n"


H = 1.57079632679  # the literal the URDFs carry


def panda(tip: str = "panda_hand") -> RobotChain:
    """Franka Panda, group panda_arm (moveit_resources_panda_description/urdf/panda.urdf)."""
    R, F = JOINT_REVOLUTE, JOINT_FIXED
    joints = [
        Joint("panda_joint1", R, (0, 0, 0.333), (0, 0, 0), (0, 0, 1), -2.8973, 2.8973, 2.1750),
        Joint("panda_joint2", R, (0, 0, 0), (-H, 0, 0), (0, 0, 1), -1.7628, 1.7628, 2.1750),
        Joint("panda_joint3", R, (0, -0.316, 0), (H, 0, 0), (0, 0, 1), -2.8973, 2.8973, 2.1750),
        Joint("panda_joint4", R, (0.0825, 0, 0), (H, 0, 0), (0, 0, 1), -3.0718, -0.0698, 2.1750),
        Joint("panda_joint5", R, (-0.0825, 0.384, 0), (-H, 0, 0), (0, 0, 1), -2.8973, 2.8973, 2.6100),
        Joint("panda_joint6", R, (0, 0, 0), (H, 0, 0), (0, 0, 1), -0.0175, 3.7525, 2.6100),
        Joint("panda_joint7", R, (0.088, 0, 0), (H, 0, 0), (0, 0, 1), -2.8973, 2.8973, 2.6100),
        Joint("panda_joint8", F, (0, 0, 0.107)),
    ]
    if tip == "panda_hand":
        joints.append(Joint("panda_hand_joint", F, (0, 0, 0), (0, 0, -0.785398163397)))
    elif tip != "panda_link8":
        raise ValueError(f"link not found: {tip}")
    return RobotChain("panda", "panda_link0", tip, joints)


PANDA_HOME = (0.0, -math.pi / 4, 0.0, -3.0 * math.pi / 4, 0.0, math.pi / 2, math.pi / 4)


def ur5() -> RobotChain:
    """UR5 (ur_description ur5.urdf.xacro, joint_limited=false), base_link -> ee_link."""
    R, F = JOINT_REVOLUTE, JOINT_FIXED
    P2 = math.pi / 2
    tp = 2.0 * math.pi
    joints = [
        Joint("shoulder_pan_joint", R, (0, 0, 0.089159), (0, 0, 0), (0, 0, 1), -tp, tp, 3.15),
        Joint("shoulder_lift_joint", R, (0, 0.13585, 0), (0, P2, 0), (0, 1, 0), -tp, tp, 3.15),
        Joint("elbow_joint", R, (0, -0.1197, 0.425), (0, 0, 0), (0, 1, 0), -tp, tp, 3.15),
        Joint("wrist_1_joint", R, (0, 0, 0.39225), (0, P2, 0), (0, 1, 0), -tp, tp, 3.2),
        Joint("wrist_2_joint", R, (0, 0.093, 0), (0, 0, 0), (0, 0, 1), -tp, tp, 3.2),
        Joint("wrist_3_joint", R, (0, 0, 0.09465), (0, 0, 0), (0, 1, 0), -tp, tp, 3.2),
        Joint("ee_fixed_joint", F, (0, 0.0823, 0), (0, 0, P2)),
    ]
    return RobotChain("ur5", "base_link", "ee_link", joints)


def fetch() -> RobotChain:
    """Fetch, group arm_with_torso (fetch_description/robots/fetch.urdf), tip gripper_link."""
    R, F, P = JOINT_REVOLUTE, JOINT_FIXED, JOINT_PRISMATIC
    joints = [
        Joint("torso_lift_joint", P, (-0.086875, 0, 0.37743), (-6.123e-17, 0, 0), (0, 0, 1), 0.0, 0.38615, 0.1),
        Joint("shoulder_pan_joint", R, (0.119525, 0, 0.34858), (0, 0, 0), (0, 0, 1), -1.6056, 1.6056, 1.256),
        Joint("shoulder_lift_joint", R, (0.117, 0, 0.06), (0, 0, 0), (0, 1, 0), -1.221, 1.518, 1.454),
        Joint("upperarm_roll_joint", R, (0.219, 0, 0), (0, 0, 0), (1, 0, 0), velocity=1.571, continuous=True),
        Joint("elbow_flex_joint", R, (0.133, 0, 0), (0, 0, 0), (0, 1, 0), -2.251, 2.251, 1.521),
        Joint("forearm_roll_joint", R, (0.197, 0, 0), (0, 0, 0), (1, 0, 0), velocity=1.571, continuous=True),
        Joint("wrist_flex_joint", R, (0.1245, 0, 0), (0, 0, 0), (0, 1, 0), -2.16, 2.16, 2.268),
        Joint("wrist_roll_joint", R, (0.1385, 0, 0), (0, 0, 0), (1, 0, 0), velocity=2.268, continuous=True),
        Joint("gripper_axis", F, (0.16645, 0, 0)),
    ]
    return RobotChain("fetch", "base_link", "gripper_link", joints)


def rr(link1_length: float = 2.0) -> RobotChain:
    """The in-code 2-R planar robot of the reference's tests
    (/root/reference/tests/ik_tests.cpp:15-48 uses a->b at x=2; tests/robot_tests.cpp:9-38 x=1).
    RobotModelBuilder::addChain gives revolute joints limits +-pi and no velocity limit."""
    R, F = JOINT_REVOLUTE, JOINT_FIXED
    joints = [
        Joint("base-a-joint", R, (0, 0, 0), (0, 0, 0), (0, 0, 1), -math.pi, math.pi, 0.0),
        Joint("a-b-joint", R, (link1_length, 0, 0), (0, 0, 0), (0, 0, 1), -math.pi, math.pi, 0.0),
        Joint("b-ee-joint", F, (1.0, 0, 0)),
    ]
    return RobotChain("rr", "base", "ee", joints)


def skew6() -> RobotChain:
    """Synthetic 6-variable chain exercising general (non axis-aligned) revolute axes, a
    general-axis prismatic joint, negative axes and interleaved fixed joints (no reference
    counterpart; parity-test coverage for the REV_GENERAL / PRISMATIC device paths)."""
    R, F, P = JOINT_REVOLUTE, JOINT_FIXED, JOINT_PRISMATIC
    joints = [
        Joint("f0", F, (0.1, 0.0, 0.2), (0.1, -0.2, 0.3)),
        Joint("j0", R, (0, 0, 0.1), (0.3, 0, 0), (0, 0, -1), -2.5, 2.5, 2.0),
        Joint("j1", R, (0.2, 0, 0), (0, 0.4, 0), (1, 1, 0), -2.0, 2.0, 1.5),
        Joint("f1", F, (0.0, 0.05, 0.0), (0, 0, 0.5)),
        Joint("j2", P, (0.1, 0, 0.1), (0, 0, 0), (1, 2, 2), -0.2, 0.3, 0.5),
        Joint("j3", R, (0.25, 0, 0), (-0.7, 0.1, 0), (0, -1, 0), velocity=3.0, continuous=True),
        Joint("j4", R, (0.15, 0.02, 0), (0, 0, 1.1), (1, 0, 0), -3.0, 3.0, 0.0),
        Joint("j5", R, (0.1, 0, 0.05), (0.2, 0.2, 0.2), (0.3, -0.5, 0.8), -2.8, 2.8, 2.5),
        Joint("f2", F, (0.05, 0, 0.08), (0, 1.0, 0)),
        Joint("f3", F, (0.0, 0.01, 0.0), (0.5, 0, 0)),
    ]
    return RobotChain("skew6", "root", "tool", joints)


def snake16() -> RobotChain:
    """Synthetic 16-variable chain at the variable-table limit (kMaxVars = 16): axis-aligned and general
    revolute axes, two prismatic joints, two continuous joints, interleaved fixed joints and no tool frame
    (no reference counterpart; maximum-size coverage for the device tables and loops)."""
    R, F, P = JOINT_REVOLUTE, JOINT_FIXED, JOINT_PRISMATIC
    axes = [(0, 0, 1), (0, 1, 0), (1, 0, 0), (0, 0, -1), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 0),
            (0, -1, 0), (0, 0, 1), (0.2, 0.3, 0.9), (0, 1, 0), (1, 0, 0), (0, 0, 1), (0, 1, 0), (-1, 0, 0)]
    joints = [Joint("base", F, (0.0, 0.0, 0.1), (0.0, 0.0, 0.2))]
    for k, ax in enumerate(axes):
        xyz = (0.09 if k % 3 == 0 else 0.02, 0.03 if k % 4 == 1 else 0.0, 0.07 if k % 2 else 0.11)
        rpy = (0.3 * ((k % 5) - 2), 0.25 * ((k % 3) - 1), 0.4 * ((k % 7) - 3))
        if k in (4, 11):
            joints.append(Joint(f"s{k}", P, xyz, rpy, ax, -0.08, 0.12, 0.4))
        elif k in (6, 13):
            joints.append(Joint(f"s{k}", R, xyz, rpy, ax, velocity=2.0, continuous=True))
        else:
            joints.append(Joint(f"s{k}", R, xyz, rpy, ax, -2.2 + 0.05 * k, 2.4 - 0.04 * k, 1.0 + 0.1 * k))
        if k == 7:
            joints.append(Joint("mid", F, (0.01, 0.0, 0.02), (0.1, 0.0, -0.1)))
    return RobotChain("snake16", "root", "s15_link", joints)


def single(kind: str = "revolute") -> RobotChain:
    """One moving joint (the smallest chain the engine accepts): a revolute joint about a general axis or a
    prismatic joint, between two fixed joints (minimum-size coverage: n = 1, 1/n mutation probability = 1)."""
    if kind == "revolute":
        moving = Joint("j", JOINT_REVOLUTE, (0.1, 0.0, 0.2), (0.2, -0.1, 0.3), (0.0, 0.6, 0.8), -2.0, 2.5, 1.0)
    else:
        moving = Joint("j", JOINT_PRISMATIC, (0.1, 0.0, 0.2), (0.2, -0.1, 0.3), (0.0, 0.0, 1.0), -0.3, 0.4, 0.5)
    return RobotChain("single_" + kind, "root", "tool",
                      [Joint("pre", JOINT_FIXED, (0.0, 0.1, 0.0), (0.0, 0.3, 0.0)), moving,
                       Joint("post", JOINT_FIXED, (0.3, 0.0, 0.1), (0.1, 0.0, 0.0))])


def single_prismatic() -> RobotChain:
    return single("prismatic")


def two_arm() -> RobotTree:
    """Two 3-joint arms on a common revolute torso, one tool tip each (a fixed joint between the torso and the
    branch point, fixed tool frames): the smallest tree with a shared chain prefix and two goal poses."""
    R, F = JOINT_REVOLUTE, JOINT_FIXED
    j = [
        Joint("torso", R, (0, 0, 0.4), (0, 0, 0), (0, 0, 1), -1.5, 1.5, 1.0, parent=-1),
        Joint("chest", F, (0, 0, 0.3), (0.1, 0, 0), parent=0),
        Joint("l_shoulder", R, (0, 0.2, 0), (0, 0, 0.3), (0, 1, 0), -2.0, 2.0, 1.5, parent=1),
        Joint("l_elbow", R, (0.3, 0, 0), (0, 0, 0), (0, 1, 0), -2.2, 2.2, 1.5, parent=2),
        Joint("l_wrist", R, (0.25, 0, 0), (0, 0, 0), (1, 0, 0), velocity=2.0, continuous=True, parent=3),
        Joint("l_tool", F, (0.1, 0, 0), (0, 0.2, 0), parent=4),
        Joint("r_shoulder", R, (0, -0.2, 0), (0, 0, -0.3), (0, 1, 0), -2.0, 2.0, 1.5, parent=1),
        Joint("r_elbow", R, (0.3, 0, 0), (0, 0, 0), (0, 1, 0), -2.2, 2.2, 1.5, parent=6),
        Joint("r_wrist", R, (0.25, 0, 0), (0, 0, 0), (0.6, 0, 0.8), -2.5, 2.5, 2.0, parent=7),
        Joint("r_tool", F, (0.1, 0, 0), (0, -0.2, 0), parent=8),
    ]
    return RobotTree("two_arm", j, [5, 9])


def three_tip() -> RobotTree:
    """Three tips: two branches straight off the model root, a tip on an intermediate link (a moving joint's own
    child link, no tool frame) and a prismatic joint."""
    R, F, P = JOINT_REVOLUTE, JOINT_FIXED, JOINT_PRISMATIC
    j = [
        Joint("a0", R, (0.1, 0, 0), (0, 0, 0), (0, 0, 1), -2.0, 2.0, 1.0, parent=-1),
        Joint("a1", R, (0.3, 0, 0), (0.2, 0, 0), (0, 1, 0), -1.8, 1.8, 1.0, parent=0),   # tip 0: a1's child link
        Joint("a2", P, (0.2, 0, 0), (0, 0, 0), (1, 0, 0), -0.1, 0.3, 0.4, parent=1),
        Joint("a_tool", F, (0.05, 0, 0.02), (0, 0, 0.1), parent=2),                        # tip 1
        Joint("base_b", F, (-0.2, 0.1, 0), (0, 0, 1.0), parent=-1),
        Joint("b0", R, (0, 0, 0.2), (0, 0, 0), (0, 0, 1), -3.0, 3.0, 2.0, parent=4),
        Joint("b1", R, (0.25, 0, 0), (0, 0.3, 0), (0, -1, 0), -1.5, 1.5, 2.0, parent=5),   # tip 2
    ]
    return RobotTree("three_tip", j, [1, 3, 6])


def floating_arm() -> RobotTree:
    """A floating base (7 variables, src/forward_kinematics.cpp:64-70) carrying a 2-joint arm."""
    R, F = JOINT_REVOLUTE, JOINT_FIXED
    j = [
        Joint("world_joint", JOINT_FLOATING, (0, 0, 0.1), (0, 0, 0), (0, 0, 0), -1.0, 1.0, 0.0, parent=-1),
        Joint("j1", R, (0, 0, 0.2), (0, 0, 0), (0, 1, 0), -2.0, 2.0, 1.0, parent=0),
        Joint("j2", R, (0.3, 0, 0), (0, 0, 0), (0, 0, 1), -2.0, 2.0, 1.0, parent=1),
        Joint("tool", F, (0.2, 0, 0), (0, 0, 0), parent=2),
    ]
    return RobotTree("floating_arm", j, [3])


def planar_arm() -> RobotTree:
    """A planar base (x, y, theta; src/forward_kinematics.cpp:71-79) with unbounded translation (MoveIt's default)
    carrying a 3-joint arm."""
    R, F = JOINT_REVOLUTE, JOINT_FIXED
    j = [
        Joint("base_joint", JOINT_PLANAR, (0, 0, 0.05), (0, 0, 0), (0, 0, 0), velocity=0.5, continuous=True, parent=-1),
        Joint("j1", R, (0.1, 0, 0.3), (0, 0, 0), (0, 0, 1), -2.5, 2.5, 1.0, parent=0),
        Joint("j2", R, (0.3, 0, 0), (0, 0, 0), (0, 1, 0), -2.0, 2.0, 1.0, parent=1),
        Joint("j3", R, (0.25, 0, 0), (0, 0, 0), (0, 1, 0), -2.0, 2.0, 1.0, parent=2),
        Joint("tool", F, (0.1, 0, 0), (0, 0, 0), parent=3),
    ]
    return RobotTree("planar_arm", j, [4])


def mimic_arm() -> RobotTree:
    """A serial arm whose third joint mimics the second (multiplier -0.5, offset 0.1): 3 variables for 4 moving
    joints (/root/reference/src/robot.cpp:145-147 skips mimic joints; MoveIt's FK drives them from their master)."""
    R, F = JOINT_REVOLUTE, JOINT_FIXED
    j = [
        Joint("j0", R, (0, 0, 0.2), (0, 0, 0), (0, 0, 1), -2.5, 2.5, 1.0, parent=-1),
        Joint("j1", R, (0.1, 0, 0.1), (0, 0, 0), (0, 1, 0), -1.5, 1.5, 1.0, parent=0),
        Joint("j1_mimic", R, (0.3, 0, 0), (0, 0, 0), (0, 1, 0), -1.5, 1.5, 1.0, parent=1, mimic_of=1, mimic_factor=-0.5,
              mimic_offset=0.1),
        Joint("j2", R, (0.3, 0, 0), (0, 0, 0), (1, 0, 0), -3.0, 3.0, 1.0, parent=2),
        Joint("tool", F, (0.15, 0, 0), (0, 0, 0), parent=3),
    ]
    return RobotTree("mimic_arm", j, [4])


def tree_to_urdf(tree: RobotTree, base_link: str = "base"):
    """The tree as a URDF document and the names of its tip links (links are named after their parent joints).  The
    decimal literals are repr() round-trips, so pik_urdf_tree reproduces tree_arrays() bit for bit."""
    def f(v):
        return " ".join(repr(float(x)) for x in v)

    kinds = {JOINT_FIXED: "fixed", JOINT_PRISMATIC: "prismatic", JOINT_FLOATING: "floating", JOINT_PLANAR: "planar"}
    lines = ['<?xml version="1.0"?>', f'<robot name="{tree.name}">', f'  <link name="{base_link}"/>']
    link_of = [f"{j.name}_link" for j in tree.joints]
    for k, j in enumerate(tree.joints):
        parent = base_link if j.parent < 0 else link_of[j.parent]
        kind = kinds.get(j.type, "continuous" if j.continuous else "revolute")
        lines.append(f'  <link name="{link_of[k]}"/>')
        lines.append(f'  <joint name="{j.name}" type="{kind}">')
        lines.append(f'    <parent link="{parent}"/>')
        lines.append(f'    <child link="{link_of[k]}"/>')
        lines.append(f'    <origin xyz="{f(j.xyz)}" rpy="{f(j.rpy)}"/>')
        if j.type in (JOINT_REVOLUTE, JOINT_PRISMATIC):
            lines.append(f'    <axis xyz="{f(j.axis)}"/>')
            if j.continuous:
                lines.append(f'    <limit effort="10" velocity="{float(j.velocity)!r}"/>')
            else:
                lines.append(f'    <limit effort="10" lower="{float(j.lower)!r}" upper="{float(j.upper)!r}" '
                             f'velocity="{float(j.velocity)!r}"/>')
        elif j.type in (JOINT_FLOATING, JOINT_PLANAR):
            lines.append(f'    <limit effort="10" velocity="{float(j.velocity)!r}"/>')
        if j.mimic_of >= 0:
            lines.append(f'    <mimic joint="{tree.joints[j.mimic_of].name}" multiplier="{float(j.mimic_factor)!r}" '
                         f'offset="{float(j.mimic_offset)!r}"/>')
        lines.append("  </joint>")
    lines.append("</robot>")
    return "\n".join(lines) + "\n", [link_of[t] for t in tree.tip_joints]


TREES = {"two_arm": two_arm, "three_tip": three_tip, "floating_arm": floating_arm, "planar_arm": planar_arm,
         "mimic_arm": mimic_arm}

ROBOTS = {"single": single, "single_prismatic": single_prismatic, "panda": panda, "ur5": ur5, "fetch": fetch, "rr": rr, "skew6": skew6, "snake16": snake16}
