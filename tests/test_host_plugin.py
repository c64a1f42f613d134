"""Builds and runs tests/cpp/test_plugin.cpp: the C++ host-side mirror of the reference's
PickIKPlugin (pick_ik_b200/host) over the C-ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_plugin")


def build():
    from pick_ik_b200 import build as pik_build

    lib = pik_build.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    srcs = [os.path.join(ROOT, "tests", "cpp", "test_plugin.cpp"), os.path.join(ROOT, "pick_ik_b200", "host", "pick_ik_plugin.cpp")]
    import hashlib

    h = hashlib.sha256()
    for d in srcs + [os.path.join(ROOT, "pick_ik_b200", "host", "pick_ik_plugin.hpp"), os.path.join(ROOT, "include", "pik.h")]:
        with open(d, "rb") as fh:
            h.update(fh.read())
    stamp = EXE + ".hash"
    if os.path.exists(EXE) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        return
    libdir = os.path.dirname(lib)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-o", EXE] + srcs +
                          ["-L" + libdir, "-lpik_b200", "-Wl,-rpath," + libdir, "-lpthread"])
    with open(stamp, "w") as fh:
        fh.write(h.hexdigest())


def run(*args):
    build()
    yaml = os.path.join(ROOT, "pick_ik_b200", "host", "pick_ik_parameters.yaml")
    return subprocess.run([EXE, yaml] + list(args), capture_output=True, text=True, timeout=600)


def test_plugin_host_logic_cpu():
    r = run("--cpu-only")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout


@pytest.mark.gpu
def test_plugin_solves_on_gpu():
    r = run()
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout


MOVEIT_EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_moveit_plugin")


def build_moveit():
    from pick_ik_b200 import build as pik_build

    lib = pik_build.build()
    os.makedirs(os.path.dirname(MOVEIT_EXE), exist_ok=True)
    srcs = [os.path.join(ROOT, "tests", "cpp", "test_moveit_plugin.cpp"),
            os.path.join(ROOT, "pick_ik_b200", "host", "moveit", "pick_ik_b200_plugin.cpp"),
            os.path.join(ROOT, "pick_ik_b200", "host", "pick_ik_plugin.cpp")]
    libdir = os.path.dirname(lib)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "tests", "cpp", "mock_moveit"),
                           "-o", MOVEIT_EXE] + srcs + ["-L" + libdir, "-lpik_b200", "-Wl,-rpath," + libdir, "-lpthread"])


def test_real_moveit_translation_unit_compiles_against_the_api_mocks_and_registers():
    """pick_ik_b200/host/moveit/pick_ik_b200_plugin.cpp: every `override` matches kinematics::KinematicsBase, the class
    registers through PLUGINLIB_EXPORT_CLASS, initialize flattens a RobotModel; without MoveIt headers the translation
    unit is empty."""
    build_moveit()
    r = subprocess.run([MOVEIT_EXE, "--cpu-only"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only",
                           os.path.join(ROOT, "pick_ik_b200", "host", "moveit", "pick_ik_b200_plugin.cpp")])
    xml = open(os.path.join(ROOT, "pick_ik_b200", "host", "moveit", "pick_ik_b200_kinematics_description.xml")).read()
    assert 'type="pick_ik_b200::MoveItPickIKPlugin"' in xml and 'base_class_type="kinematics::KinematicsBase"' in xml


@pytest.mark.gpu
def test_moveit_plugin_solves_two_tip_goal_on_gpu():
    build_moveit()
    r = subprocess.run([MOVEIT_EXE], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout
