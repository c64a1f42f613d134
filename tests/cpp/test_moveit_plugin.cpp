// test_moveit_plugin.cpp -- the real MoveIt translation unit (pick_ik_b200/host/moveit/pick_ik_b200_plugin.cpp)
// compiled against header mocks of the MoveIt 2 API (tests/cpp/mock_moveit), loaded through the pluginlib
// registration macro and driven through kinematics::KinematicsBase the way move_group drives a kinematics plugin.
// `--cpu-only`: everything up to the first solve.
#include <moveit/kinematics_base/kinematics_base.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <pluginlib/class_list_macros.hpp>
#include <string>
#include <vector>

#include "../../include/pik.h"

static int g_failures = 0, g_checks = 0;
#define CHECK(cond)                                                       \
    do {                                                                  \
        ++g_checks;                                                       \
        if (!(cond)) {                                                    \
            ++g_failures;                                                 \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
        }                                                                 \
    } while (0)

using moveit::core::JointModel;

// a model with a virtual fixed joint, a torso, two 2-joint arms with tool frames, a gripper finger that mimics, and a
// head joint outside the planning group
static void add(moveit::core::RobotModel& m, JointModel* joint, std::string const& child, int parent_link, double x, double y,
                double z, double lo = 0, double hi = 0, double vel = 0) {
    joint->index_ = m.joints_.size();
    m.links_.emplace_back(new moveit::core::LinkModel(child));
    auto* link = m.links_.back().get();
    link->origin_.t.v[0] = x; link->origin_.t.v[1] = y; link->origin_.t.v[2] = z;
    link->parent_joint_ = joint;
    joint->child_link_ = link;
    joint->parent_link_ = parent_link >= 0 ? m.links_[(size_t)parent_link].get() : nullptr;
    if (joint->getVariableCount() > 0) {
        auto& b = joint->bounds()[0];
        b.position_bounded_ = lo < hi;
        b.min_position_ = lo; b.max_position_ = hi;
        b.velocity_bounded_ = vel > 0; b.max_velocity_ = vel;
    }
    m.joints_.emplace_back(joint);
    m.joint_ptrs_.push_back(joint);
}

static void build_model(moveit::core::RobotModel& m) {
    using moveit::core::RevoluteJointModel;
    Eigen::Vector3d const Z{{0, 0, 1}}, Y{{0, 1, 0}};
    m.model_frame_ = "world";
    add(m, new JointModel("virtual", JointModel::FIXED, 0), "base", -1, 0, 0, 0);             // link 0
    add(m, new RevoluteJointModel("torso", Z), "torso_link", 0, 0, 0, 0.4, -1.5, 1.5, 1.0);   // link 1
    add(m, new RevoluteJointModel("l1", Y), "l1_link", 1, 0, 0.2, 0.3, -2.0, 2.0, 1.5);       // link 2
    add(m, new RevoluteJointModel("l2", Y), "l2_link", 2, 0.3, 0, 0, -2.2, 2.2, 1.5);         // link 3
    add(m, new JointModel("l_tool", JointModel::FIXED, 0), "l_hand", 3, 0.2, 0, 0);           // link 4
    add(m, new RevoluteJointModel("r1", Y), "r1_link", 1, 0, -0.2, 0.3, -2.0, 2.0, 1.5);      // link 5
    add(m, new RevoluteJointModel("r2", Y), "r2_link", 5, 0.3, 0, 0, -2.2, 2.2, 1.5);         // link 6
    add(m, new JointModel("r_tool", JointModel::FIXED, 0), "r_hand", 6, 0.2, 0, 0);           // link 7
    add(m, new RevoluteJointModel("head", Z), "head_link", 1, 0, 0, 0.5, -1.0, 1.0, 1.0);     // link 8: not in the group
    m.groups_.emplace_back(new moveit::core::JointModelGroup("both_arms"));
    m.groups_.back()->joints_ = {"torso", "l1", "l2", "l_tool", "r1", "r2", "r_tool"};
}

int main(int argc, char** argv) {
    bool const cpu_only = argc > 1 && std::string(argv[1]) == "--cpu-only";
    bool const have_gpu = pik_device_count() > 0;
    // src/pick_ik_plugin.cpp:405: PLUGINLIB_EXPORT_CLASS registered the class under the base type
    auto& registry = pluginlib_mock::registry<kinematics::KinematicsBase>();
    CHECK(registry.count("pick_ik_b200::MoveItPickIKPlugin") == 1);
    std::shared_ptr<kinematics::KinematicsBase> plugin = registry.at("pick_ik_b200::MoveItPickIKPlugin")();
    moveit::core::RobotModel model;
    build_model(model);
    auto node = std::make_shared<rclcpp::Node>();
    CHECK(!plugin->initialize(node, model, "no_such_group", "world", {"l_hand"}, 0.0));
    bool threw = false;
    try {
        plugin->initialize(node, model, "both_arms", "world", {"no_such_link"}, 0.0);
    } catch (std::invalid_argument const&) {
        threw = true;  // src/pick_ik_plugin.cpp:65-67
    }
    CHECK(threw);
    bool const ok = plugin->initialize(node, model, "both_arms", "/world", {"l_hand", "/r_hand"}, 0.0);
    CHECK(ok == have_gpu);  // no device: initialize fails, there is no CPU fallback
    if (!cpu_only && have_gpu) {
        CHECK(plugin->getJointNames().size() == 5 && plugin->getLinkNames().size() == 2 && plugin->getBaseFrame() == "world");
        // parameters live under robot_description_kinematics.<group>, re-read on every call
        node->set_parameter_for_test<std::string>("robot_description_kinematics.both_arms.mode", "global");
        node->set_parameter_for_test<int>("robot_description_kinematics.both_arms.memetic_population_size", 32);
        // goal: both hands where the configuration (0.3, 0.6, -0.4, -0.5, 0.8) puts them -- closed form for this model
        double const q[5] = {0.3, 0.6, -0.4, -0.5, 0.8};
        auto hand = [&](double side_y, double a1, double a2) {
            geometry_msgs::msg::Pose p;
            // torso Rz(q0) at z = 0.4; shoulder at (0, side_y, 0.3); Ry(a1); elbow at x = 0.3; Ry(a2); tool at x = 0.2
            double const lx = 0.3 * std::cos(a1) + 0.2 * std::cos(a1 + a2), lz = -0.3 * std::sin(a1) - 0.2 * std::sin(a1 + a2);
            double const bx = lx, by = side_y;
            p.position.x = std::cos(q[0]) * bx - std::sin(q[0]) * by;
            p.position.y = std::sin(q[0]) * bx + std::cos(q[0]) * by;
            p.position.z = 0.4 + 0.3 + lz;
            // orientation Rz(q0) * Ry(a1 + a2)
            double const hz = q[0] / 2, hy = (a1 + a2) / 2;
            p.orientation.w = std::cos(hz) * std::cos(hy);
            p.orientation.x = -std::sin(hz) * std::sin(hy);
            p.orientation.y = std::cos(hz) * std::sin(hy);
            p.orientation.z = std::sin(hz) * std::cos(hy);
            return p;
        };
        std::vector<geometry_msgs::msg::Pose> goals = {hand(0.2, q[1], q[2]), hand(-0.2, q[3], q[4])};
        std::vector<double> const seed = {0, 0, 0, 0, 0};
        std::vector<double> sol;
        moveit_msgs::msg::MoveItErrorCodes ec;
        int callbacks = 0;
        kinematics::KinematicsBase::IKCallbackFn cb = [&](geometry_msgs::msg::Pose const&, std::vector<double> const&,
                                                          moveit_msgs::msg::MoveItErrorCodes&) { ++callbacks; };
        CHECK(plugin->searchPositionIK(goals, seed, 2.0, {}, sol, cb, ec));
        CHECK(ec.val == moveit_msgs::msg::MoveItErrorCodes::SUCCESS && sol.size() == 5 && callbacks == 1);
        // a custom cost function is refused with a defined status
        kinematics::KinematicsBase::IKCostFn cost = [](geometry_msgs::msg::Pose const&, moveit::core::RobotState const&,
                                                       moveit::core::JointModelGroup const*, std::vector<double> const&) { return 1.0; };
        CHECK(!plugin->searchPositionIK(goals, seed, 0.1, {}, sol, cb, cost, ec));
        CHECK(ec.val == moveit_msgs::msg::MoveItErrorCodes::NO_IK_SOLUTION && sol == seed);
        // the single-pose overloads need a single tip: a second instance on the left arm only
        std::shared_ptr<kinematics::KinematicsBase> left = registry.at("pick_ik_b200::MoveItPickIKPlugin")();
        CHECK(left->initialize(node, model, "both_arms", "world", {"l_hand"}, 0.0));
        CHECK(left->getJointNames().size() == 3);
        CHECK(left->searchPositionIK(goals[0], {0, 0, 0}, 2.0, sol, ec) && sol.size() == 3);
        // invalid mode: error + false (src/pick_ik_plugin.cpp:204-207)
        node->set_parameter_for_test<std::string>("robot_description_kinematics.both_arms.mode", "sideways");
        CHECK(!left->searchPositionIK(goals[0], {0, 0, 0}, 0.1, sol, ec));
        std::vector<geometry_msgs::msg::Pose> fk;
        CHECK(!left->getPositionFK({}, {}, fk));
    }
    std::printf("%d checks, %d failures\n", g_checks, g_failures);
    return g_failures == 0 ? 0 : 1;
}
