// pick_ik_plugin.hpp -- host-side mirror of the reference's plugin surface over the C-ABI (include/pik.h).
//
// The reference's drop-in entry is pick_ik::PickIKPlugin : kinematics::KinematicsBase
// (include/pick_ik/pick_ik_plugin.hpp:10-102, src/pick_ik_plugin.cpp:22-401).  MoveIt, rclcpp and the ROS
// message packages are not available in this environment, so the MoveIt types the interface names are
// declared here as minimal stand-ins with the same member names (namespace pick_ik_b200::compat); a MoveIt
// build replaces this header's `compat` types by the real ones.  Method names, argument
// order and meaning, return values and error codes follow the reference; every method cites what it mirrors.
//
// What is behind it differs: the solve runs on the GPU through pik_solve_batch, and the class also
// offers the batched entry point the reference does not have (searchPositionIKBatch).
#pragma once

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/pik.h"

namespace pick_ik_b200 {

namespace compat {

// geometry_msgs::msg::Pose
struct Point {
    double x = 0.0, y = 0.0, z = 0.0;
};
struct Quaternion {
    double x = 0.0, y = 0.0, z = 0.0, w = 1.0;
};
struct Pose {
    Point position;
    Quaternion orientation;
};

// moveit_msgs::msg::MoveItErrorCodes (the two values the plugin writes, src/pick_ik_plugin.cpp:212,215)
struct MoveItErrorCodes {
    static constexpr int SUCCESS = 1;
    static constexpr int NO_IK_SOLUTION = -31;
    int val = 0;
};

// kinematics::KinematicsQueryOptions (only return_approximate_solution is read, src/pick_ik_plugin.cpp:187,199,222)
struct KinematicsQueryOptions {
    bool lock_redundant_joints = false;
    bool return_approximate_solution = false;
};

// kinematics::KinematicsBase::IKCallbackFn
using IKCallbackFn = std::function<void(Pose const&, std::vector<double> const&, MoveItErrorCodes&)>;

// moveit::core::RobotState / JointModelGroup: only named by the interface (context_state is ignored by the reference,
// src/pick_ik_plugin.cpp:83; an IKCostFn receives them)
struct RobotState {};
struct JointModelGroup {};
// kinematics::KinematicsBase::IKCostFn
using IKCostFn = std::function<double(Pose const&, RobotState const&, JointModelGroup const*, std::vector<double> const&)>;

// What the plugin reads from moveit::core::RobotModel + JointModelGroup (src/pick_ik_plugin.cpp:42-68,
// src/robot.cpp:44-85,107-160): the joints of the model from the root down to the tip links of the group -- a serial
// chain (parent empty) or a tree (parent[j] = index of the joint above joint j, -1 = the model root, parents first).
struct ChainModel {
    std::string group_name;
    std::string model_frame;               // RobotModel::getModelFrame()
    std::vector<pik_joint_desc> joints;    // parents before children (a chain: chain order), fixed joints included
    std::vector<std::string> joint_names;  // one per entry of `joints`
    std::vector<std::string> link_names;   // child link of each joint (a chain: the last one is the tip link)
    std::vector<int32_t> parent;           // empty: a serial chain in joint order
    std::vector<int32_t> mimic_of;         // empty: no mimic joints; else per joint the joint it follows or -1
    std::vector<double> mimic_factor, mimic_offset;
};

// Stand-alone replacement for the RobotModel: the chain base_link -> tip_link of a URDF document
// (pik_urdf_chain; urdfdom / MoveIt semantics).  Throws std::invalid_argument when the document is malformed,
// the tip is not below the base, or the chain has a joint the engine does not support (floating, planar, mimic).
ChainModel chain_from_urdf(std::string const& urdf_xml, std::string const& group_name, std::string const& base_link,
                           std::string const& tip_link);

// URDF + SRDF -> the planning group's model: the group's <chain> entries give the base link and the tip links
// (pik_srdf_group), the URDF the joints between them with mimic joints (pik_urdf_tree).  tip_frames receives the tip
// links of the group (what a MoveIt configuration passes to initialize).  Throws std::invalid_argument like
// chain_from_urdf.
ChainModel model_from_urdf_srdf(std::string const& urdf_xml, std::string const& srdf_xml, std::string const& group_name,
                                std::vector<std::string>& tip_frames);

}  // namespace compat

// pick_ik::Params (generated from src/pick_ik_parameters.yaml by generate_parameter_library): same member
// names, same defaults.  The reference re-reads them on every solve (src/pick_ik_plugin.cpp:86); here the
// caller sets them on the plugin (setParams / loadParams) and they are read on every solve.
struct Params {
    std::string mode = "global";
    double gd_step_size = 0.0001;
    int gd_max_iters = 100;
    double gd_min_cost_delta = 1.0e-12;
    double position_threshold = 0.001;
    double orientation_threshold = 0.001;
    double approximate_solution_position_threshold = 0.05;
    double approximate_solution_orientation_threshold = 0.05;
    double approximate_solution_joint_threshold = 0.0;
    double approximate_solution_cost_threshold = 0.0;
    double cost_threshold = 0.001;
    double position_scale = 1.0;
    double rotation_scale = 0.5;
    double center_joints_weight = 0.0;
    double avoid_joint_limits_weight = 0.0;
    double minimal_displacement_weight = 0.0;
    bool stop_optimization_on_valid_solution = true;
    int memetic_num_threads = 1;
    bool memetic_stop_on_first_solution = true;
    int memetic_population_size = 16;
    int memetic_elite_size = 4;
    double memetic_wipeout_fitness_tol = 0.00001;
    int memetic_max_generations = 100;
    int memetic_gd_max_iters = 25;
    double memetic_gd_max_time = 0.005;
    unsigned long long rng_seed = 0x5EED;  // not a reference parameter: the reference's RNG is unseeded
};

// Sets one parameter by its YAML name from text ("true"/"false" for bools).  Returns false for an unknown
// name or an unparsable value.
bool set_param(Params& p, std::string const& name, std::string const& value);
// Reads "name: value" lines (a flat override file, or the `default_value` entries of
// pick_ik_parameters.yaml).  Returns the number of parameters set, or -1 on a malformed line.
int load_params(Params& p, std::string const& text);
// The YAML -> C-ABI mapping (mode string -> enum etc.).  Returns false when `mode` is invalid.
bool to_pik_params(Params const& p, bool return_approximate_solution, pik_params& out);

class PickIKPlugin {
   public:
    PickIKPlugin();
    ~PickIKPlugin();
    PickIKPlugin(PickIKPlugin const&) = delete;
    PickIKPlugin& operator=(PickIKPlugin const&) = delete;

    // kinematics::KinematicsBase::initialize (include/pick_ik/pick_ik_plugin.hpp:24-29,
    // src/pick_ik_plugin.cpp:22-71).  Throws std::invalid_argument when a tip frame is not a link of the
    // chain (src/pick_ik_plugin.cpp:65-67).  `device` selects the GPU (no reference counterpart).
    bool initialize(compat::ChainModel const& robot_model, std::string const& group_name,
                    std::string const& base_frame, std::vector<std::string> const& tip_frames,
                    double search_discretization, int device = 0);

    // The main overload, signature as include/pick_ik/pick_ik_plugin.hpp:31-41 (src/pick_ik_plugin.cpp:73-294): one
    // pose per tip frame.  consistency_limits and context_state are ignored, as in the reference.  A non-empty
    // cost_function is a host callback over a RobotState (src/goal.cpp:146-161) and cannot be evaluated by the device
    // kernels: the call logs the reason, sets NO_IK_SOLUTION, hands the seed back and returns false.  The method is
    // const and re-entrant: concurrent callers get solvers of their own from a pool.
    bool searchPositionIK(std::vector<compat::Pose> const& ik_poses, std::vector<double> const& ik_seed_state,
                          double timeout, std::vector<double> const& consistency_limits,
                          std::vector<double>& solution, compat::IKCallbackFn const& solution_callback,
                          compat::IKCostFn const& cost_function, compat::MoveItErrorCodes& error_code,
                          compat::KinematicsQueryOptions const& options = compat::KinematicsQueryOptions(),
                          compat::RobotState const* context_state = nullptr) const;
    // include/pick_ik/pick_ik_plugin.hpp:92-102: the same without a cost function
    bool searchPositionIK(std::vector<compat::Pose> const& ik_poses, std::vector<double> const& ik_seed_state,
                          double timeout, std::vector<double> const& consistency_limits,
                          std::vector<double>& solution, compat::IKCallbackFn const& solution_callback,
                          compat::MoveItErrorCodes& error_code,
                          compat::KinematicsQueryOptions const& options = compat::KinematicsQueryOptions(),
                          compat::RobotState const* context_state = nullptr) const;

    // Forwarding overloads (include/pick_ik/pick_ik_plugin.hpp:57-102, src/pick_ik_plugin.cpp:314-401)
    bool searchPositionIK(compat::Pose const& ik_pose, std::vector<double> const& ik_seed_state, double timeout,
                          std::vector<double>& solution, compat::MoveItErrorCodes& error_code,
                          compat::KinematicsQueryOptions const& options = compat::KinematicsQueryOptions()) const;
    bool searchPositionIK(compat::Pose const& ik_pose, std::vector<double> const& ik_seed_state, double timeout,
                          std::vector<double> const& consistency_limits, std::vector<double>& solution,
                          compat::MoveItErrorCodes& error_code,
                          compat::KinematicsQueryOptions const& options = compat::KinematicsQueryOptions()) const;
    bool searchPositionIK(compat::Pose const& ik_pose, std::vector<double> const& ik_seed_state, double timeout,
                          std::vector<double>& solution, compat::IKCallbackFn const& solution_callback,
                          compat::MoveItErrorCodes& error_code,
                          compat::KinematicsQueryOptions const& options = compat::KinematicsQueryOptions()) const;
    bool searchPositionIK(compat::Pose const& ik_pose, std::vector<double> const& ik_seed_state, double timeout,
                          std::vector<double> const& consistency_limits, std::vector<double>& solution,
                          compat::IKCallbackFn const& solution_callback, compat::MoveItErrorCodes& error_code,
                          compat::KinematicsQueryOptions const& options = compat::KinematicsQueryOptions()) const;

    // Batched form of the main overload (new): B independent problems solved in one pik_solve_batch call.  ik_poses
    // holds the poses of problem b at [b * n_tips, (b + 1) * n_tips), one per tip frame; one seed per problem or one for
    // all.  solutions [B][n]; error_codes [B]; costs (optional) [B] the cost of the best individual found.  rng_seed:
    // the random stream of this call (the reference's RNG is unseeded; Params::rng_seed when 0).  Returns the number
    // of problems solved, or -1 on an invalid mode / engine error.  Applies the same approximate-solution gating per
    // problem.  memetic_num_threads species per problem run inside the engine (src/ik_memetic.cpp:315-370).
    long searchPositionIKBatch(std::vector<compat::Pose> const& ik_poses,
                               std::vector<std::vector<double>> const& ik_seed_states,
                               std::vector<std::vector<double>>& solutions,
                               std::vector<compat::MoveItErrorCodes>& error_codes,
                               compat::KinematicsQueryOptions const& options = compat::KinematicsQueryOptions(),
                               std::vector<double>* costs = nullptr, unsigned long long rng_seed = 0) const;

    std::vector<std::string> const& getJointNames() const;  // src/pick_ik_plugin.cpp:296
    std::vector<std::string> const& getLinkNames() const;   // src/pick_ik_plugin.cpp:298
    // src/pick_ik_plugin.cpp:300-312: not implemented by the reference either
    bool getPositionFK(std::vector<std::string> const&, std::vector<double> const&, std::vector<compat::Pose>&) const;
    bool getPositionIK(compat::Pose const&, std::vector<double> const&, std::vector<double>&,
                       compat::MoveItErrorCodes&, compat::KinematicsQueryOptions const&) const;

    void setParams(Params const& p) const;  // (const: the parameters are re-read before every solve, src/pick_ik_plugin.cpp:86)
    Params const& getParams() const;
    std::string const& lastError() const;

   private:
    struct Impl;
    std::unique_ptr<Impl> impl_;
};

}  // namespace pick_ik_b200
