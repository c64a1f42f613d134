// test_plugin.cpp -- the host-side PickIKPlugin mirror, exercised the way the reference exercises its
// solvers (tests/ik_tests.cpp:137-293, tests/ik_memetic_tests.cpp:110-206) but through the plugin surface
// (which the reference itself does not test).  `--cpu-only` runs the sections that need no GPU.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../pick_ik_b200/host/pick_ik_plugin.hpp"

using namespace pick_ik_b200;
using compat::ChainModel;
using compat::KinematicsQueryOptions;
using compat::MoveItErrorCodes;
using compat::Pose;

static int g_failures = 0, g_checks = 0;
#define CHECK(cond)                                                                 \
    do {                                                                            \
        ++g_checks;                                                                 \
        if (!(cond)) {                                                              \
            ++g_failures;                                                           \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);           \
        }                                                                           \
    } while (0)
static bool approx(double a, double b, double margin) { return std::fabs(a - b) <= margin; }

// urdfdom setFromRPY -> quaternion -> rotation matrix
static void rpy_to_matrix(double r, double p, double y, double* R) {
    double const phi = r / 2, the = p / 2, psi = y / 2;
    double qx = std::sin(phi) * std::cos(the) * std::cos(psi) - std::cos(phi) * std::sin(the) * std::sin(psi);
    double qy = std::cos(phi) * std::sin(the) * std::cos(psi) + std::sin(phi) * std::cos(the) * std::sin(psi);
    double qz = std::cos(phi) * std::cos(the) * std::sin(psi) - std::sin(phi) * std::sin(the) * std::cos(psi);
    double qw = std::cos(phi) * std::cos(the) * std::cos(psi) + std::sin(phi) * std::sin(the) * std::sin(psi);
    double const nrm = std::sqrt(qx * qx + qy * qy + qz * qz + qw * qw);
    qx /= nrm; qy /= nrm; qz /= nrm; qw /= nrm;
    double const tx = 2 * qx, ty = 2 * qy, tz = 2 * qz, twx = tx * qw, twy = ty * qw, twz = tz * qw;
    double const txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

static void add_joint(ChainModel& m, char const* joint, char const* link, int type, double x, double y, double z,
                      double rr, double rp, double ry, double lower, double upper, double vel) {
    pik_joint_desc jd;
    std::memset(&jd, 0, sizeof(jd));
    jd.type = type;
    jd.bounded = type != PIK_JOINT_FIXED;
    rpy_to_matrix(rr, rp, ry, jd.origin_R);
    jd.origin_t[0] = x; jd.origin_t[1] = y; jd.origin_t[2] = z;
    jd.axis[2] = 1.0;
    jd.min_position = lower; jd.max_position = upper; jd.max_velocity = vel;
    m.joints.push_back(jd);
    m.joint_names.push_back(joint);
    m.link_names.push_back(link);
}

// tests/ik_tests.cpp:15-48: base -> a (revolute z), a -> b (revolute z, x = 2), b -> ee (fixed, x = 1)
static ChainModel make_rr_model_for_ik() {
    ChainModel m;
    m.group_name = "group";
    m.model_frame = "base";
    add_joint(m, "base-a-joint", "a", PIK_JOINT_REVOLUTE, 0, 0, 0, 0, 0, 0, -M_PI, M_PI, 0);
    add_joint(m, "a-b-joint", "b", PIK_JOINT_REVOLUTE, 2, 0, 0, 0, 0, 0, -M_PI, M_PI, 0);
    add_joint(m, "b-ee-joint", "ee", PIK_JOINT_FIXED, 1, 0, 0, 0, 0, 0, 0, 0, 0);
    return m;
}

// moveit_resources panda, group panda_arm + panda_hand
static ChainModel make_panda_model() {
    double const H = 1.57079632679;
    ChainModel m;
    m.group_name = "panda_arm";
    m.model_frame = "panda_link0";
    add_joint(m, "panda_joint1", "panda_link1", PIK_JOINT_REVOLUTE, 0, 0, 0.333, 0, 0, 0, -2.8973, 2.8973, 2.1750);
    add_joint(m, "panda_joint2", "panda_link2", PIK_JOINT_REVOLUTE, 0, 0, 0, -H, 0, 0, -1.7628, 1.7628, 2.1750);
    add_joint(m, "panda_joint3", "panda_link3", PIK_JOINT_REVOLUTE, 0, -0.316, 0, H, 0, 0, -2.8973, 2.8973, 2.1750);
    add_joint(m, "panda_joint4", "panda_link4", PIK_JOINT_REVOLUTE, 0.0825, 0, 0, H, 0, 0, -3.0718, -0.0698, 2.1750);
    add_joint(m, "panda_joint5", "panda_link5", PIK_JOINT_REVOLUTE, -0.0825, 0.384, 0, -H, 0, 0, -2.8973, 2.8973, 2.6100);
    add_joint(m, "panda_joint6", "panda_link6", PIK_JOINT_REVOLUTE, 0, 0, 0, H, 0, 0, -0.0175, 3.7525, 2.6100);
    add_joint(m, "panda_joint7", "panda_link7", PIK_JOINT_REVOLUTE, 0.088, 0, 0, H, 0, 0, -2.8973, 2.8973, 2.6100);
    add_joint(m, "panda_joint8", "panda_link8", PIK_JOINT_FIXED, 0, 0, 0.107, 0, 0, 0, 0, 0, 0);
    add_joint(m, "panda_hand_joint", "panda_hand", PIK_JOINT_FIXED, 0, 0, 0, 0, 0, -0.785398163397, 0, 0, 0);
    return m;
}

static Pose make_pose(double x, double y, double z, double qw, double qx, double qy, double qz) {
    Pose p;
    p.position.x = x; p.position.y = y; p.position.z = z;
    p.orientation.w = qw; p.orientation.x = qx; p.orientation.y = qy; p.orientation.z = qz;
    return p;
}

// ---------------------------------------------------------------------------------------------
static void test_params(std::string const& yaml_path) {
    Params p;
    CHECK(set_param(p, "mode", "local") && p.mode == "local");
    CHECK(set_param(p, "memetic_population_size", "128") && p.memetic_population_size == 128);
    CHECK(set_param(p, "stop_optimization_on_valid_solution", "false") && !p.stop_optimization_on_valid_solution);
    CHECK(!set_param(p, "no_such_parameter", "1"));
    CHECK(!set_param(p, "gd_max_iters", "many"));
    Params q;
    CHECK(load_params(q, "mode: local\n# comment\nrotation_scale: 0.25\n") == 2);
    CHECK(q.mode == "local" && q.rotation_scale == 0.25);
    CHECK(load_params(q, "this is not yaml") == -1);
    // the parameter contract file parses back to the built-in defaults
    std::ifstream in(yaml_path);
    CHECK(in.good());
    std::stringstream ss;
    ss << in.rdbuf();
    Params y;
    y.mode = "";
    y.memetic_gd_max_iters = -1;
    CHECK(load_params(y, ss.str()) == 25);
    Params def;
    CHECK(y.mode == def.mode && y.memetic_gd_max_iters == def.memetic_gd_max_iters && y.rotation_scale == def.rotation_scale &&
          y.memetic_population_size == def.memetic_population_size && y.gd_min_cost_delta == def.gd_min_cost_delta);
    pik_params pp;
    CHECK(to_pik_params(def, false, pp) && pp.mode == PIK_MODE_GLOBAL && pik_params_validate(&pp) == PIK_OK);
    def.mode = "sideways";
    CHECK(!to_pik_params(def, false, pp));
}

static void test_initialize_errors(bool have_gpu) {
    PickIKPlugin plugin;
    auto const model = make_rr_model_for_ik();
    bool threw = false;
    try {
        plugin.initialize(model, "group", "base", {"no_such_link"}, 0.0);
    } catch (std::invalid_argument const&) {
        threw = true;  // src/pick_ik_plugin.cpp:65-67
    }
    CHECK(threw);
    CHECK(!plugin.initialize(model, "other_group", "base", {"ee"}, 0.0));
    if (!have_gpu) CHECK(!plugin.initialize(model, "group", "base", {"ee"}, 0.0));  // no device: fails, no CPU fallback
    std::vector<Pose> out;
    CHECK(!plugin.getPositionFK({}, {}, out));  // src/pick_ik_plugin.cpp:300-305
}

// the stand-alone path: URDF text -> ChainModel (pik_urdf_chain) -> the same table the hand-built model gives
static void test_chain_from_urdf() {
    char const* urdf =
        "<?xml version='1.0'?><robot name='rr'><link name='base'/><link name='l1'/><link name='l2'/><link name='ee'/>"
        "<!-- two revolute joints about z and a fixed tool frame -->"
        "<joint name='j1' type='revolute'><parent link='base'/><child link='l1'/><axis xyz='0 0 1'/>"
        "<limit lower='-3.0' upper='3.0' velocity='1' effort='1'/></joint>"
        "<joint name='j2' type='revolute'><parent link='l1'/><child link='l2'/><origin xyz='2 0 0'/><axis xyz='0 0 1'/>"
        "<limit lower='-3.0' upper='3.0' velocity='1' effort='1'/></joint>"
        "<joint name='tool' type='fixed'><parent link='l2'/><child link='ee'/><origin xyz='1 0 0' rpy='0 0 0.25'/></joint>"
        "</robot>";
    auto const m = pick_ik_b200::compat::chain_from_urdf(urdf, "group", "base", "ee");
    CHECK(m.joints.size() == 3 && m.joint_names.size() == 3 && m.link_names.size() == 3);
    CHECK(m.joint_names[0] == "j1" && m.joint_names[2] == "tool" && m.link_names[2] == "ee" && m.model_frame == "base");
    CHECK(m.joints[0].type == PIK_JOINT_REVOLUTE && m.joints[2].type == PIK_JOINT_FIXED);
    CHECK(m.joints[1].origin_t[0] == 2.0 && m.joints[1].min_position == -3.0 && m.joints[1].max_velocity == 1.0);
    double R[9];
    rpy_to_matrix(0, 0, 0.25, R);
    for (int i = 0; i < 9; ++i) CHECK(m.joints[2].origin_R[i] == R[i]);
    pik_robot* robot = nullptr;
    CHECK(pik_robot_create(m.joints.data(), (int32_t)m.joints.size(), &robot) == PIK_OK);
    CHECK(pik_robot_num_variables(robot) == 2);
    pik_robot_destroy(robot);
    bool threw = false;
    try {
        pick_ik_b200::compat::chain_from_urdf(urdf, "group", "ee", "base");  // tip not below base
    } catch (std::invalid_argument const&) {
        threw = true;
    }
    CHECK(threw);
}

// URDF + SRDF -> planning-group model (pik_srdf_group + pik_urdf_tree): two chains off a common torso, a mimic finger
static char const* kTwoArmUrdf =
    "<robot name='two_arm'><link name='base'/><link name='torso_link'/><link name='l1_link'/><link name='l2_link'/>"
    "<link name='l_hand'/><link name='r1_link'/><link name='r2_link'/><link name='r_hand'/><link name='head_link'/>"
    "<joint name='torso' type='revolute'><parent link='base'/><child link='torso_link'/><origin xyz='0 0 0.4'/><axis xyz='0 0 1'/>"
    "<limit lower='-1.5' upper='1.5' velocity='1' effort='1'/></joint>"
    "<joint name='l1' type='revolute'><parent link='torso_link'/><child link='l1_link'/><origin xyz='0 0.2 0.3' rpy='0.3 0 0'/>"
    "<axis xyz='0 0 1'/><limit lower='-2' upper='2' velocity='1.5' effort='1'/></joint>"
    "<joint name='l2' type='revolute'><parent link='l1_link'/><child link='l2_link'/><origin xyz='0.3 0 0' rpy='0 0.5 0'/>"
    "<axis xyz='0 0 1'/><limit lower='-2.2' upper='2.2' velocity='1.5' effort='1'/></joint>"
    "<joint name='l_tool' type='fixed'><parent link='l2_link'/><child link='l_hand'/><origin xyz='0.2 0 0'/></joint>"
    "<joint name='r1' type='revolute'><parent link='torso_link'/><child link='r1_link'/><origin xyz='0 -0.2 0.3' rpy='-0.3 0 0'/>"
    "<axis xyz='0 0 1'/><limit lower='-2' upper='2' velocity='1.5' effort='1'/></joint>"
    "<joint name='r2' type='revolute'><parent link='r1_link'/><child link='r2_link'/><origin xyz='0.3 0 0' rpy='0 0.5 0'/>"
    "<axis xyz='0 0 1'/><limit lower='-2.2' upper='2.2' velocity='1.5' effort='1'/><mimic joint='l2' multiplier='-1' offset='0.1'/></joint>"
    "<joint name='r_tool' type='fixed'><parent link='r2_link'/><child link='r_hand'/><origin xyz='0.2 0 0'/></joint>"
    "<joint name='head' type='revolute'><parent link='torso_link'/><child link='head_link'/><origin xyz='0 0 0.5'/><axis xyz='0 0 1'/>"
    "<limit lower='-1' upper='1' velocity='1' effort='1'/></joint></robot>";
static char const* kTwoArmSrdf =
    "<robot name='two_arm'><group name='left'><chain base_link='base' tip_link='l_hand'/></group>"
    "<group name='right'><chain base_link='base' tip_link='r_hand'/></group>"
    "<group name='both'><group name='left'/><group name='right'/></group></robot>";

static void test_model_from_urdf_srdf(bool have_gpu) {
    std::vector<std::string> tips;
    auto const m = pick_ik_b200::compat::model_from_urdf_srdf(kTwoArmUrdf, kTwoArmSrdf, "both", tips);
    CHECK(tips.size() == 2 && m.model_frame == "base" && m.joints.size() == 7);  // head is not below a tip
    CHECK(m.parent.size() == 7 && m.parent[0] == -1 && m.mimic_of.size() == 7);
    int mimics = 0;
    for (size_t j = 0; j < m.joints.size(); ++j)
        if (m.mimic_of[j] >= 0) {
            ++mimics;
            CHECK(m.joint_names[j] == "r2" && m.joint_names[(size_t)m.mimic_of[j]] == "l2" && m.mimic_factor[j] == -1.0 &&
                  m.mimic_offset[j] == 0.1);
        }
    CHECK(mimics == 1);
    bool threw = false;
    try {
        pick_ik_b200::compat::model_from_urdf_srdf(kTwoArmUrdf, kTwoArmSrdf, "no_such_group", tips);
    } catch (std::invalid_argument const&) {
        threw = true;
    }
    CHECK(threw);
    if (!have_gpu) return;
    PickIKPlugin plugin;
    auto const model = pick_ik_b200::compat::model_from_urdf_srdf(kTwoArmUrdf, kTwoArmSrdf, "both", tips);
    CHECK(plugin.initialize(model, "both", "base", tips, 0.0));
    CHECK(plugin.getJointNames().size() == 4);  // torso, l1, l2, r1; r2 mimics l2
}

// tests/ik_tests.cpp:137-238 through the plugin, local mode, IkTestParams (:78-86)
static void test_rr_ik() {
    PickIKPlugin plugin;
    CHECK(plugin.initialize(make_rr_model_for_ik(), "group", "base", {"ee"}, 0.0));
    CHECK(plugin.getJointNames().size() == 2 && plugin.getLinkNames().size() == 1 && plugin.getLinkNames()[0] == "ee");
    Params p;
    p.mode = "local";
    p.position_threshold = 0.0001;
    p.orientation_threshold = 0.001;
    p.cost_threshold = 0.0001;
    p.rotation_scale = 1.0;
    plugin.setParams(p);
    double const s45 = std::sin(M_PI_4);
    struct Case { Pose goal; double e0, e1, g0, g1; bool solvable; };
    Pose const goal_a = make_pose(3, 0, 0, 1, 0, 0, 0);
    Pose const goal_b = make_pose(s45, 3 * s45, 0, std::cos(0.375 * M_PI), 0, 0, std::sin(0.375 * M_PI));
    std::vector<Case> cases = {
        {goal_a, 0, 0, 0.1, -0.1, true},                      // :140-151
        {goal_a, 0, 0, M_PI_2, -M_PI_2, true},                // :153-164
        {goal_b, M_PI_4, M_PI_2, M_PI_4 + 0.1, M_PI_2 - 0.1, true},  // :166-178
        {goal_b, M_PI_4, M_PI_2, 0, 0, true},                 // :180-192
        {make_pose(0, 0, 0, 1, 0, 0, 0), 0, 0, 0, 0, false},  // :194-203 unreachable
        {make_pose(s45, 3 * s45, 0, 1, 0, 0, 0), 0, 0, 0, 0, false},  // :205-216 position yes, orientation no
    };
    for (auto const& c : cases) {
        std::vector<double> sol;
        MoveItErrorCodes ec;
        bool const ok = plugin.searchPositionIK(c.goal, {c.g0, c.g1}, 0.05, sol, ec);
        CHECK(ok == c.solvable);
        CHECK(ec.val == (c.solvable ? MoveItErrorCodes::SUCCESS : MoveItErrorCodes::NO_IK_SOLUTION));
        if (c.solvable) {
            CHECK(approx(sol[0], c.e0, 0.01) && approx(sol[1], c.e1, 0.01));
        } else {
            CHECK(sol.size() == 2 && sol[0] == c.g0 && sol[1] == c.g1);  // solution = seed on failure, :216
        }
    }
    // :218-233 zero rotation scale makes the last goal solvable
    p.rotation_scale = 0.0;
    plugin.setParams(p);
    std::vector<double> sol;
    MoveItErrorCodes ec;
    CHECK(plugin.searchPositionIK(make_pose(s45, 3 * s45, 0, 1, 0, 0, 0), {M_PI_4 + 0.1, M_PI_2 - 0.1}, 0.05, sol, ec));
    CHECK(approx(sol[0], M_PI_4, 0.01) && approx(sol[1], M_PI_2, 0.01));
    // invalid mode: error + false (src/pick_ik_plugin.cpp:204-207)
    p.mode = "sideways";
    plugin.setParams(p);
    CHECK(!plugin.searchPositionIK(goal_a, {0.1, -0.1}, 0.05, sol, ec));
}

// tests/ik_memetic_tests.cpp:110-206 through the plugin (global mode), plus callback veto and the batch entry
static void test_panda_memetic() {
    PickIKPlugin plugin;
    CHECK(plugin.initialize(make_panda_model(), "panda_arm", "panda_link0", {"panda_hand"}, 0.0));
    CHECK(plugin.getJointNames().size() == 7);
    std::vector<double> const home = {0.0, -M_PI_4, 0.0, -3.0 * M_PI_4, 0.0, M_PI_2, M_PI_4};
    // FK(home) of panda_hand: t = (0.306891, 0, 0.590282), R = Rx(pi) . Rz(...) -> q(w,x,y,z) = (0, 1, 0, 0)
    // (SURVEY.md App. C sanity values)
    Pose const goal = make_pose(0.30689059, 0.0, 0.59028174, 0.0, 1.0, 0.0, 0.0);
    Params p;  // yaml defaults: global, P = 16, E = 4
    p.position_threshold = 0.001;
    plugin.setParams(p);
    std::vector<std::vector<double>> guesses = {home, {0.1, -M_PI_4 - 0.1, 0.0, -3.0 * M_PI_4 + 0.1, -0.1, M_PI_2, M_PI_4 + 0.1},
                                                {0, 0, 0, 0, 0, 0, 0}};
    for (auto const& g : guesses) {
        std::vector<double> sol;
        MoveItErrorCodes ec;
        CHECK(plugin.searchPositionIK(goal, g, 1.0, sol, ec));
        CHECK(ec.val == MoveItErrorCodes::SUCCESS && sol.size() == 7);
    }
    // four species (":174-190 multithreaded"), and with the joint-centering / limit-avoiding goals (:192-213)
    p.memetic_num_threads = 4;
    plugin.setParams(p);
    std::vector<double> sol;
    MoveItErrorCodes ec;
    CHECK(plugin.searchPositionIK(goal, {0, 0, 0, 0, 0, 0, 0}, 1.0, sol, ec));
    p.memetic_num_threads = 1;
    p.center_joints_weight = 0.01;
    p.avoid_joint_limits_weight = 0.01;
    p.cost_threshold = 0.01;
    p.position_threshold = 0.01;
    plugin.setParams(p);
    CHECK(plugin.searchPositionIK(goal, {0, 0, 0, 0, 0, 0, 0}, 1.0, sol, ec));
    // the callback runs only on success and may veto (src/pick_ik_plugin.cpp:270-274)
    int calls = 0;
    auto veto = [&](Pose const&, std::vector<double> const&, MoveItErrorCodes& e) {
        ++calls;
        e.val = MoveItErrorCodes::NO_IK_SOLUTION;
    };
    CHECK(!plugin.searchPositionIK(goal, home, 0.0, sol, veto, ec));
    CHECK(calls == 1 && ec.val == MoveItErrorCodes::NO_IK_SOLUTION);
    // batch entry: 64 copies of the problem from different seeds
    std::vector<Pose> poses(64, goal);
    std::vector<std::vector<double>> seeds(64, home), sols;
    for (int b = 0; b < 64; ++b) seeds[b][0] += 0.01 * b;
    std::vector<MoveItErrorCodes> codes;
    long const solved = plugin.searchPositionIKBatch(poses, seeds, sols, codes);
    CHECK(solved >= 60 && sols.size() == 64 && codes.size() == 64);
    // approximate-solution gating: an unreachable pose fails the (strict) frame tests, seed comes back
    KinematicsQueryOptions opt;
    opt.return_approximate_solution = true;
    p = Params();
    p.memetic_max_generations = 3;
    plugin.setParams(p);
    CHECK(!plugin.searchPositionIK({make_pose(5, 5, 5, 1, 0, 0, 0)}, home, 0.0, {}, sol, compat::IKCallbackFn(), ec, opt));
    CHECK(ec.val == MoveItErrorCodes::NO_IK_SOLUTION && sol == home);
}

// the exact main overload (include/pick_ik/pick_ik_plugin.hpp:31-41): a custom IKCostFn is refused with a defined
// status; an empty one and a context_state are accepted
static void test_cost_function_and_context_state() {
    PickIKPlugin plugin;
    CHECK(plugin.initialize(make_rr_model_for_ik(), "group", "base", {"ee"}, 0.0));
    Params p;
    p.mode = "local";
    plugin.setParams(p);
    std::vector<double> const seed = {0.1, -0.1};
    std::vector<double> sol;
    MoveItErrorCodes ec;
    compat::RobotState context;
    int calls = 0;
    compat::IKCostFn cost = [&](Pose const&, compat::RobotState const&, compat::JointModelGroup const*, std::vector<double> const&) {
        ++calls;
        return 0.0;
    };
    CHECK(!plugin.searchPositionIK({make_pose(3, 0, 0, 1, 0, 0, 0)}, seed, 0.05, {}, sol, compat::IKCallbackFn(), cost, ec,
                                   KinematicsQueryOptions(), &context));
    CHECK(ec.val == MoveItErrorCodes::NO_IK_SOLUTION && sol == seed && calls == 0);
    CHECK(plugin.lastError().find("IKCostFn") != std::string::npos);
    CHECK(plugin.searchPositionIK({make_pose(3, 0, 0, 1, 0, 0, 0)}, seed, 0.05, {}, sol, compat::IKCallbackFn(), compat::IKCostFn(), ec,
                                  KinematicsQueryOptions(), &context));
    CHECK(ec.val == MoveItErrorCodes::SUCCESS);
    // one pose per tip frame
    CHECK(!plugin.searchPositionIK({make_pose(3, 0, 0, 1, 0, 0, 0), make_pose(3, 0, 0, 1, 0, 0, 0)}, seed, 0.05, {}, sol,
                                   compat::IKCallbackFn(), ec));
}

// species inside the engine (src/ik_memetic.cpp:315-370): the cheapest value over all species wins when they all run
// to their end; the batch entry hands the costs back
static void test_species_pick() {
    PickIKPlugin plugin;
    CHECK(plugin.initialize(make_panda_model(), "panda_arm", "panda_link0", {"panda_hand"}, 0.0));
    std::vector<double> const home = {0.0, -M_PI_4, 0.0, -3.0 * M_PI_4, 0.0, M_PI_2, M_PI_4};
    Pose const goal = make_pose(0.30689059, 0.0, 0.59028174, 0.0, 1.0, 0.0, 0.0);
    std::vector<Pose> poses(32, goal);
    std::vector<std::vector<double>> seeds(32, std::vector<double>(7, 0.0)), sols;
    for (int b = 0; b < 32; ++b) seeds[b][1] = -0.1 - 0.02 * b, seeds[b][3] = -1.5;
    std::vector<MoveItErrorCodes> codes;
    std::vector<double> cost1, cost4, cost4all;
    Params p;
    p.memetic_max_generations = 40;
    plugin.setParams(p);
    CHECK(plugin.searchPositionIKBatch(poses, seeds, sols, codes, KinematicsQueryOptions(), &cost1) >= 28);
    p.memetic_num_threads = 4;
    plugin.setParams(p);
    CHECK(plugin.searchPositionIKBatch(poses, seeds, sols, codes, KinematicsQueryOptions(), &cost4) >= 30);
    p.memetic_stop_on_first_solution = false;
    plugin.setParams(p);
    long const solved_all = plugin.searchPositionIKBatch(poses, seeds, sols, codes, KinematicsQueryOptions(), &cost4all);
    CHECK(solved_all >= 30 && cost1.size() == 32 && cost4all.size() == 32);
    int lower = 0;
    for (int b = 0; b < 32; ++b) {
        if (codes[b].val != MoveItErrorCodes::SUCCESS) continue;
        CHECK(cost4all[b] <= cost4[b]);  // every species ran on: the pick is the minimum over at least the same values
        lower += cost4all[b] < cost4[b] ? 1 : 0;
    }
    CHECK(lower > 0);
}

// two tip frames on a tree (src/pick_ik_plugin.cpp:60-68, src/goal.cpp:80-89): one pose per tip
static void test_two_tips() {
    ChainModel m;
    m.group_name = "both_arms";
    m.model_frame = "base";
    add_joint(m, "torso", "torso_link", PIK_JOINT_REVOLUTE, 0, 0, 0.4, 0, 0, 0, -1.5, 1.5, 1.0);
    add_joint(m, "l1", "l1_link", PIK_JOINT_REVOLUTE, 0, 0.2, 0.3, 0.3, 0, 0, -2.0, 2.0, 1.5);
    add_joint(m, "l2", "l2_link", PIK_JOINT_REVOLUTE, 0.3, 0, 0, 0, 0.5, 0, -2.2, 2.2, 1.5);
    add_joint(m, "l_tool", "l_hand", PIK_JOINT_FIXED, 0.2, 0, 0, 0, 0, 0, 0, 0, 0);
    add_joint(m, "r1", "r1_link", PIK_JOINT_REVOLUTE, 0, -0.2, 0.3, -0.3, 0, 0, -2.0, 2.0, 1.5);
    add_joint(m, "r2", "r2_link", PIK_JOINT_REVOLUTE, 0.3, 0, 0, 0, 0.5, 0, -2.2, 2.2, 1.5);
    add_joint(m, "r_tool", "r_hand", PIK_JOINT_FIXED, 0.2, 0, 0, 0, 0, 0, 0, 0, 0);
    add_joint(m, "head", "head_link", PIK_JOINT_REVOLUTE, 0, 0, 0.5, 0, 0, 0, -1.0, 1.0, 1.0);  // not below a tip: unused
    m.parent = {-1, 0, 1, 2, 0, 4, 5, 0};
    PickIKPlugin plugin;
    CHECK(plugin.initialize(m, "both_arms", "base", {"l_hand", "r_hand"}, 0.0));
    CHECK(plugin.getJointNames().size() == 5 && plugin.getLinkNames().size() == 2);
    // targets: FK of a known configuration, through the engine itself
    pik_robot* robot = nullptr;
    std::vector<pik_joint_desc> used(m.joints.begin(), m.joints.begin() + 7);
    int32_t const parent[7] = {-1, 0, 1, 2, 0, 4, 5}, tips[2] = {3, 6};
    CHECK(pik_robot_create_tree(used.data(), 7, parent, tips, 2, nullptr, nullptr, nullptr, &robot) == PIK_OK);
    CHECK(pik_robot_num_variables(robot) == 5 && pik_robot_num_tips(robot) == 2);
    pik_solver* solver = nullptr;
    CHECK(pik_solver_create(robot, 0, nullptr, &solver) == PIK_OK);
    pik_params pp;
    pik_params_default(&pp);
    double const q_goal[5] = {0.4, 0.7, -0.5, -0.6, 0.9};
    double const seed5[5] = {0, 0, 0, 0, 0};
    double const ident[14] = {0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0};
    double tip[14];
    CHECK(pik_eval_cost(solver, &pp, 1, ident, seed5, 0, q_goal, nullptr, nullptr, tip, PIK_MEM_HOST) == PIK_OK);
    std::vector<Pose> goals = {make_pose(tip[0], tip[1], tip[2], tip[3], tip[4], tip[5], tip[6]),
                               make_pose(tip[7], tip[8], tip[9], tip[10], tip[11], tip[12], tip[13])};
    std::vector<double> sol;
    MoveItErrorCodes ec;
    CHECK(plugin.searchPositionIK(goals, {0, 0, 0, 0, 0}, 1.0, {}, sol, compat::IKCallbackFn(), ec));
    CHECK(ec.val == MoveItErrorCodes::SUCCESS && sol.size() == 5);
    // the solution reaches BOTH goals
    int32_t is_solution = 0;
    double goal14[14];
    std::memcpy(goal14, tip, sizeof(goal14));
    CHECK(pik_eval_cost(solver, &pp, 1, goal14, seed5, 0, sol.data(), nullptr, &is_solution, nullptr, PIK_MEM_HOST) == PIK_OK);
    CHECK(is_solution == 1);
    // one pose only: refused
    CHECK(!plugin.searchPositionIK(std::vector<Pose>{goals[0]}, {0, 0, 0, 0, 0}, 0.1, {}, sol, compat::IKCallbackFn(), ec));
    pik_solver_destroy(solver);
    pik_robot_destroy(robot);
}

// searchPositionIK is const and re-entrant (the reference serialises on its FK mutex only): concurrent callers with
// the same inputs get the same answers as a lone caller
static void test_concurrent_callers() {
    PickIKPlugin plugin;
    CHECK(plugin.initialize(make_panda_model(), "panda_arm", "panda_link0", {"panda_hand"}, 0.0));
    Pose const goal = make_pose(0.30689059, 0.0, 0.59028174, 0.0, 1.0, 0.0, 0.0);
    std::vector<std::vector<double>> guesses;
    for (int k = 0; k < 6; ++k) guesses.push_back({0.05 * k, -0.3, 0.0, -1.5 - 0.1 * k, 0.0, 1.2, 0.3});
    std::vector<std::vector<double>> alone(guesses.size());
    for (size_t k = 0; k < guesses.size(); ++k) {
        MoveItErrorCodes ec;
        CHECK(plugin.searchPositionIK(goal, guesses[k], 5.0, alone[k], ec));
    }
    std::vector<std::vector<double>> together(guesses.size());
    std::vector<int> ok(guesses.size(), 0);
    std::vector<std::thread> threads;
    for (size_t k = 0; k < guesses.size(); ++k)
        threads.emplace_back([&, k] {
            for (int rep = 0; rep < 3; ++rep) {
                MoveItErrorCodes ec;
                ok[k] = plugin.searchPositionIK(goal, guesses[k], 5.0, together[k], ec) ? 1 : 0;
            }
        });
    for (auto& t : threads) t.join();
    for (size_t k = 0; k < guesses.size(); ++k) CHECK(ok[k] == 1 && together[k] == alone[k]);
}

int main(int argc, char** argv) {
    bool cpu_only = false;
    std::string yaml = "pick_ik_b200/host/pick_ik_parameters.yaml";
    for (int i = 1; i < argc; ++i) {
        if (std::string(argv[i]) == "--cpu-only") cpu_only = true;
        else yaml = argv[i];
    }
    bool const have_gpu = pik_device_count() > 0;
    test_params(yaml);
    test_initialize_errors(have_gpu);
    test_chain_from_urdf();
    test_model_from_urdf_srdf(have_gpu && !cpu_only);
    if (!cpu_only) {
        if (!have_gpu) {
            std::printf("no CUDA device: the solve sections need one\n");
            return 2;
        }
        test_rr_ik();
        test_panda_memetic();
        test_cost_function_and_context_state();
        test_species_pick();
        test_two_tips();
        test_concurrent_callers();
    }
    std::printf("%d checks, %d failures\n", g_checks, g_failures);
    return g_failures == 0 ? 0 : 1;
}
