// pick_ik_b200_plugin.cpp -- the drop-in artefact for a MoveIt 2 workspace: kinematics::KinematicsBase over the
// GPU engine, registered with pluginlib as pick_ik_b200/PickIkPlugin (pick_ik_b200_kinematics_description.xml).
//
// It replaces pick_ik::PickIKPlugin (include/pick_ik/pick_ik_plugin.hpp:10-102, src/pick_ik_plugin.cpp:22-405):
//   initialize        flattens RobotModel + JointModelGroup + tip frames into the joint table of the C-ABI
//                     (what Robot::from and make_fk_fn read from MoveIt: src/robot.cpp:44-85,122-160,
//                     src/fk_moveit.cpp:11-35) and hands it to pick_ik_b200::PickIKPlugin
//   searchPositionIK  the exact virtual signatures; parameters are re-read from the node on every call under
//                     robot_description_kinematics.<group> with the names, defaults and validators of
//                     src/pick_ik_parameters.yaml (the reference re-reads them through its ParamListener, :86)
//
// MoveIt, rclcpp and pluginlib are not part of this repository's build environment: the translation unit is
// compiled only where <moveit/kinematics_base/kinematics_base.h> exists (a colcon workspace; see INTEGRATION.md
// for the CMake lines), and tests/test_host_plugin.py compiles it against header mocks of the MoveIt API
// (tests/cpp/mock_moveit) so that every `override` is checked against the base-class signatures.
#if __has_include(<moveit/kinematics_base/kinematics_base.h>)

#include <moveit/kinematics_base/kinematics_base.h>
#include <moveit/robot_model/robot_model.h>
#include <moveit/robot_state/robot_state.h>

#include <map>
#include <memory>
#include <pluginlib/class_list_macros.hpp>
#include <rclcpp/rclcpp.hpp>
#include <stdexcept>
#include <string>
#include <vector>

#include "../pick_ik_plugin.hpp"

namespace pick_ik_b200 {
namespace {

auto const LOGGER = rclcpp::get_logger("pick_ik_b200");

// moveit::core::RobotModel -> the joint table: every joint of the model, parents first (RobotModel::getJointModels()
// is in that order), fixed joints included.  A moving joint that is not an active joint of the group keeps the
// value RobotState::setToDefaultValues gives it, which is how the reference's FK sees it (src/fk_moveit.cpp:15-23):
// it enters the table as a fixed joint (default value 0 inside the bounds; anything else is refused).
compat::ChainModel chain_from(moveit::core::RobotModel const& model, moveit::core::JointModelGroup const& jmg) {
    compat::ChainModel m;
    m.group_name = jmg.getName();
    m.model_frame = model.getModelFrame();
    std::map<moveit::core::JointModel const*, int32_t> index;
    for (auto const* joint : model.getJointModels()) {
        pik_joint_desc jd{};
        bool const active = jmg.hasJointModel(joint->getName()) && joint->getMimic() == nullptr &&
                            joint->getType() != moveit::core::JointModel::FIXED;
        bool const mimic_in_group = jmg.hasJointModel(joint->getName()) && joint->getMimic() != nullptr;
        jd.type = PIK_JOINT_FIXED;
        if (active || mimic_in_group) {
            switch (joint->getType()) {
                case moveit::core::JointModel::REVOLUTE: jd.type = PIK_JOINT_REVOLUTE; break;
                case moveit::core::JointModel::PRISMATIC: jd.type = PIK_JOINT_PRISMATIC; break;
                case moveit::core::JointModel::FLOATING: jd.type = PIK_JOINT_FLOATING; break;
                case moveit::core::JointModel::PLANAR: jd.type = PIK_JOINT_PLANAR; break;
                default: throw std::invalid_argument("unsupported joint type: " + joint->getName());
            }
        } else if (joint->getType() != moveit::core::JointModel::FIXED && joint->getVariableCount() > 0) {
            auto const& b = joint->getVariableBounds().front();
            if (b.position_bounded_ && !(b.min_position_ <= 0.0 && 0.0 <= b.max_position_))
                throw std::invalid_argument("joint outside the group with a non-zero default value: " + joint->getName());
        }
        Eigen::Isometry3d const& origin = joint->getChildLinkModel()->getJointOriginTransform();
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) jd.origin_R[3 * r + c] = origin.linear()(r, c);
            jd.origin_t[r] = origin.translation()(r);
        }
        if (auto const* rev = dynamic_cast<moveit::core::RevoluteJointModel const*>(joint)) {
            jd.axis[0] = rev->getAxis().x(); jd.axis[1] = rev->getAxis().y(); jd.axis[2] = rev->getAxis().z();
        } else if (auto const* pri = dynamic_cast<moveit::core::PrismaticJointModel const*>(joint)) {
            jd.axis[0] = pri->getAxis().x(); jd.axis[1] = pri->getAxis().y(); jd.axis[2] = pri->getAxis().z();
        }
        if (jd.type != PIK_JOINT_FIXED) {
            // VariableBounds of the first variable (Robot::from, src/robot.cpp:52-68); for floating / planar joints these
            // are the translation bounds, the library sets the quaternion / angle bounds as MoveIt's joint models do
            auto const& b = joint->getVariableBounds().front();
            jd.bounded = b.position_bounded_ ? 1 : 0;
            jd.min_position = b.min_position_;
            jd.max_position = b.max_position_;
            jd.max_velocity = b.velocity_bounded_ ? b.max_velocity_ : 0.0;
        }
        auto const* parent_link = joint->getParentLinkModel();
        int32_t const parent = parent_link ? index.at(parent_link->getParentJointModel()) : -1;
        index[joint] = static_cast<int32_t>(m.joints.size());
        m.joints.push_back(jd);
        m.joint_names.push_back(joint->getName());
        m.link_names.push_back(joint->getChildLinkModel()->getName());
        m.parent.push_back(parent);
        if (auto const* master = joint->getMimic(); master && jd.type != PIK_JOINT_FIXED) {
            m.mimic_of.resize(m.joints.size(), -1);
            m.mimic_factor.resize(m.joints.size(), 1.0);
            m.mimic_offset.resize(m.joints.size(), 0.0);
            m.mimic_of.back() = index.at(master);
            m.mimic_factor.back() = joint->getMimicFactor();
            m.mimic_offset.back() = joint->getMimicOffset();
        }
    }
    if (!m.mimic_of.empty()) {
        m.mimic_of.resize(m.joints.size(), -1);
        m.mimic_factor.resize(m.joints.size(), 1.0);
        m.mimic_offset.resize(m.joints.size(), 0.0);
    }
    return m;
}

template <class T>
T read_parameter(rclcpp::Node::SharedPtr const& node, std::string const& name, T const& fallback) {
    if (!node->has_parameter(name)) node->declare_parameter<T>(name, fallback);
    T value = fallback;
    node->get_parameter(name, value);
    return value;
}

compat::Pose to_compat(geometry_msgs::msg::Pose const& p) {
    compat::Pose out;
    out.position.x = p.position.x; out.position.y = p.position.y; out.position.z = p.position.z;
    out.orientation.x = p.orientation.x; out.orientation.y = p.orientation.y; out.orientation.z = p.orientation.z;
    out.orientation.w = p.orientation.w;
    return out;
}

}  // namespace

class MoveItPickIKPlugin : public kinematics::KinematicsBase {
    rclcpp::Node::SharedPtr node_;
    std::string parameter_namespace_;
    PickIKPlugin engine_;

    // src/pick_ik_parameters.yaml: same names, defaults and validators (pik_params_validate applies them in the
    // library; an invalid set makes the solve fail with an error, as generate_parameter_library would refuse it)
    Params read_params() const {
        Params p;
        auto const ns = parameter_namespace_ + ".";
#define PIK_READ(field) p.field = read_parameter(node_, ns + #field, p.field)
        PIK_READ(mode); PIK_READ(gd_step_size); PIK_READ(gd_max_iters); PIK_READ(gd_min_cost_delta);
        PIK_READ(position_threshold); PIK_READ(orientation_threshold); PIK_READ(approximate_solution_position_threshold);
        PIK_READ(approximate_solution_orientation_threshold); PIK_READ(approximate_solution_joint_threshold);
        PIK_READ(approximate_solution_cost_threshold); PIK_READ(cost_threshold); PIK_READ(position_scale);
        PIK_READ(rotation_scale); PIK_READ(center_joints_weight); PIK_READ(avoid_joint_limits_weight);
        PIK_READ(minimal_displacement_weight); PIK_READ(stop_optimization_on_valid_solution); PIK_READ(memetic_num_threads);
        PIK_READ(memetic_stop_on_first_solution); PIK_READ(memetic_population_size); PIK_READ(memetic_elite_size);
        PIK_READ(memetic_wipeout_fitness_tol); PIK_READ(memetic_max_generations); PIK_READ(memetic_gd_max_iters);
        PIK_READ(memetic_gd_max_time);
#undef PIK_READ
        return p;
    }

   public:
    // src/pick_ik_plugin.cpp:22-71
    bool initialize(rclcpp::Node::SharedPtr const& node, moveit::core::RobotModel const& robot_model,
                    std::string const& group_name, std::string const& base_frame,
                    std::vector<std::string> const& tip_frames, double search_discretization) override {
        node_ = node;
        parameter_namespace_ = std::string("robot_description_kinematics.").append(group_name);
        storeValues(robot_model, group_name, base_frame, tip_frames, search_discretization);
        auto const* jmg = robot_model_->getJointModelGroup(group_name);
        if (!jmg) {
            RCLCPP_ERROR(LOGGER, "failed to get joint model group %s", group_name.c_str());
            return false;
        }
        int const device = static_cast<int>(read_parameter<int64_t>(node_, parameter_namespace_ + ".cuda_device", 0));
        // throws std::invalid_argument for an unknown tip frame, as the reference does (src/pick_ik_plugin.cpp:65-67)
        if (!engine_.initialize(chain_from(*robot_model_, *jmg), group_name, base_frame_, tip_frames_, search_discretization,
                                device)) {
            RCLCPP_ERROR(LOGGER, "%s", engine_.lastError().c_str());
            return false;
        }
        return true;
    }

    // include/pick_ik/pick_ik_plugin.hpp:31-41, src/pick_ik_plugin.cpp:73-294
    bool searchPositionIK(std::vector<geometry_msgs::msg::Pose> const& ik_poses, std::vector<double> const& ik_seed_state,
                          double timeout, std::vector<double> const& consistency_limits, std::vector<double>& solution,
                          IKCallbackFn const& solution_callback, IKCostFn const& cost_function,
                          moveit_msgs::msg::MoveItErrorCodes& error_code,
                          kinematics::KinematicsQueryOptions const& options = kinematics::KinematicsQueryOptions(),
                          moveit::core::RobotState const* context_state = nullptr) const override {
        (void)context_state;  // not used, as in the reference
        engine_.setParams(read_params());  // re-read on every solve (:86)
        std::vector<compat::Pose> poses;
        for (auto const& p : ik_poses) poses.push_back(to_compat(p));
        compat::KinematicsQueryOptions opt;
        opt.lock_redundant_joints = options.lock_redundant_joints;
        opt.return_approximate_solution = options.return_approximate_solution;
        compat::MoveItErrorCodes ec;
        compat::IKCallbackFn callback;
        if (solution_callback)
            callback = [&](compat::Pose const&, std::vector<double> const& sol, compat::MoveItErrorCodes& code) {
                moveit_msgs::msg::MoveItErrorCodes m;
                m.val = code.val;
                solution_callback(ik_poses.front(), sol, m);  // may veto (src/pick_ik_plugin.cpp:270-274)
                code.val = m.val;
            };
        compat::IKCostFn cost;
        if (cost_function)  // refused by the engine with NO_IK_SOLUTION (a host callback cannot run on the device)
            cost = [](compat::Pose const&, compat::RobotState const&, compat::JointModelGroup const*, std::vector<double> const&) {
                return 0.0;
            };
        bool const found = engine_.searchPositionIK(poses, ik_seed_state, timeout, consistency_limits, solution, callback, cost,
                                                    ec, opt, nullptr);
        error_code.val = ec.val;
        if (!found && !engine_.lastError().empty()) RCLCPP_ERROR(LOGGER, "%s", engine_.lastError().c_str());
        return found;
    }

    // include/pick_ik/pick_ik_plugin.hpp:92-102
    bool searchPositionIK(std::vector<geometry_msgs::msg::Pose> const& ik_poses, std::vector<double> const& ik_seed_state,
                          double timeout, std::vector<double> const& consistency_limits, std::vector<double>& solution,
                          IKCallbackFn const& solution_callback, moveit_msgs::msg::MoveItErrorCodes& error_code,
                          kinematics::KinematicsQueryOptions const& options = kinematics::KinematicsQueryOptions(),
                          moveit::core::RobotState const* context_state = nullptr) const override {
        return searchPositionIK(ik_poses, ik_seed_state, timeout, consistency_limits, solution, solution_callback, IKCostFn(),
                                error_code, options, context_state);
    }

    // include/pick_ik/pick_ik_plugin.hpp:57-90, src/pick_ik_plugin.cpp:314-401
    bool searchPositionIK(geometry_msgs::msg::Pose const& ik_pose, std::vector<double> const& ik_seed_state, double timeout,
                          std::vector<double>& solution, moveit_msgs::msg::MoveItErrorCodes& error_code,
                          kinematics::KinematicsQueryOptions const& options = kinematics::KinematicsQueryOptions()) const override {
        return searchPositionIK(std::vector<geometry_msgs::msg::Pose>{ik_pose}, ik_seed_state, timeout, std::vector<double>(),
                                solution, IKCallbackFn(), error_code, options);
    }
    bool searchPositionIK(geometry_msgs::msg::Pose const& ik_pose, std::vector<double> const& ik_seed_state, double timeout,
                          std::vector<double> const& consistency_limits, std::vector<double>& solution,
                          moveit_msgs::msg::MoveItErrorCodes& error_code,
                          kinematics::KinematicsQueryOptions const& options = kinematics::KinematicsQueryOptions()) const override {
        return searchPositionIK(std::vector<geometry_msgs::msg::Pose>{ik_pose}, ik_seed_state, timeout, consistency_limits,
                                solution, IKCallbackFn(), error_code, options);
    }
    bool searchPositionIK(geometry_msgs::msg::Pose const& ik_pose, std::vector<double> const& ik_seed_state, double timeout,
                          std::vector<double>& solution, IKCallbackFn const& solution_callback,
                          moveit_msgs::msg::MoveItErrorCodes& error_code,
                          kinematics::KinematicsQueryOptions const& options = kinematics::KinematicsQueryOptions()) const override {
        return searchPositionIK(std::vector<geometry_msgs::msg::Pose>{ik_pose}, ik_seed_state, timeout, std::vector<double>(),
                                solution, solution_callback, error_code, options);
    }
    bool searchPositionIK(geometry_msgs::msg::Pose const& ik_pose, std::vector<double> const& ik_seed_state, double timeout,
                          std::vector<double> const& consistency_limits, std::vector<double>& solution,
                          IKCallbackFn const& solution_callback, moveit_msgs::msg::MoveItErrorCodes& error_code,
                          kinematics::KinematicsQueryOptions const& options = kinematics::KinematicsQueryOptions()) const override {
        return searchPositionIK(std::vector<geometry_msgs::msg::Pose>{ik_pose}, ik_seed_state, timeout, consistency_limits,
                                solution, solution_callback, error_code, options);
    }

    std::vector<std::string> const& getJointNames() const override { return engine_.getJointNames(); }  // :296
    std::vector<std::string> const& getLinkNames() const override { return engine_.getLinkNames(); }    // :298
    // src/pick_ik_plugin.cpp:300-312: not implemented by the reference either
    bool getPositionFK(std::vector<std::string> const&, std::vector<double> const&,
                       std::vector<geometry_msgs::msg::Pose>&) const override {
        return false;
    }
    bool getPositionIK(geometry_msgs::msg::Pose const&, std::vector<double> const&, std::vector<double>&,
                       moveit_msgs::msg::MoveItErrorCodes&, kinematics::KinematicsQueryOptions const&) const override {
        return false;
    }
};

}  // namespace pick_ik_b200

// src/pick_ik_plugin.cpp:405
PLUGINLIB_EXPORT_CLASS(pick_ik_b200::MoveItPickIKPlugin, kinematics::KinematicsBase)

#endif  // __has_include(<moveit/kinematics_base/kinematics_base.h>)
