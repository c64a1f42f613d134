// pick_ik_plugin.cpp -- host-side mirror of pick_ik::PickIKPlugin over the C-ABI.  See pick_ik_plugin.hpp.
#include "pick_ik_plugin.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <sstream>
#include <stdexcept>

namespace pick_ik_b200 {

using compat::KinematicsQueryOptions;
using compat::MoveItErrorCodes;
using compat::Pose;

// ------------------------------------------------------------------------------------------------
// parameters
// ------------------------------------------------------------------------------------------------
namespace {

bool parse_bool(std::string const& v, bool& out) {
    if (v == "true" || v == "True" || v == "1") { out = true; return true; }
    if (v == "false" || v == "False" || v == "0") { out = false; return true; }
    return false;
}
bool parse_double(std::string const& v, double& out) {
    char* end = nullptr;
    out = std::strtod(v.c_str(), &end);
    return end != v.c_str() && *end == '\0';
}
bool parse_int(std::string const& v, int& out) {
    char* end = nullptr;
    long x = std::strtol(v.c_str(), &end, 10);
    out = static_cast<int>(x);
    return end != v.c_str() && *end == '\0';
}
std::string trim(std::string s) {
    auto const ws = " \t\r\n\"'";
    auto a = s.find_first_not_of(ws);
    if (a == std::string::npos) return "";
    auto b = s.find_last_not_of(ws);
    return s.substr(a, b - a + 1);
}

}  // namespace

bool set_param(Params& p, std::string const& name, std::string const& value) {
#define PIK_D(field) if (name == #field) return parse_double(value, p.field)
#define PIK_I(field) if (name == #field) return parse_int(value, p.field)
#define PIK_B(field) if (name == #field) return parse_bool(value, p.field)
    if (name == "mode") { p.mode = value; return true; }
    PIK_D(gd_step_size); PIK_I(gd_max_iters); PIK_D(gd_min_cost_delta); PIK_D(position_threshold);
    PIK_D(orientation_threshold); PIK_D(approximate_solution_position_threshold);
    PIK_D(approximate_solution_orientation_threshold); PIK_D(approximate_solution_joint_threshold);
    PIK_D(approximate_solution_cost_threshold); PIK_D(cost_threshold); PIK_D(position_scale); PIK_D(rotation_scale);
    PIK_D(center_joints_weight); PIK_D(avoid_joint_limits_weight); PIK_D(minimal_displacement_weight);
    PIK_B(stop_optimization_on_valid_solution); PIK_I(memetic_num_threads); PIK_B(memetic_stop_on_first_solution);
    PIK_I(memetic_population_size); PIK_I(memetic_elite_size); PIK_D(memetic_wipeout_fitness_tol);
    PIK_I(memetic_max_generations); PIK_I(memetic_gd_max_iters); PIK_D(memetic_gd_max_time);
#undef PIK_D
#undef PIK_I
#undef PIK_B
    if (name == "rng_seed") {
        char* end = nullptr;
        p.rng_seed = std::strtoull(value.c_str(), &end, 0);
        return end != value.c_str() && *end == '\0';
    }
    return false;
}

int load_params(Params& p, std::string const& text) {
    std::istringstream in(text);
    std::string line;
    int count = 0;
    while (std::getline(in, line)) {
        auto hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        if (trim(line).empty()) continue;
        auto colon = line.find(':');
        if (colon == std::string::npos) return -1;
        std::string name = trim(line.substr(0, colon));
        std::string rest = trim(line.substr(colon + 1));
        if (rest.empty()) continue;  // a section header such as "pick_ik:"
        if (rest.front() == '{') {   // generate_parameter_library entry: take default_value
            auto dv = rest.find("default_value:");
            if (dv == std::string::npos) return -1;
            auto tail = rest.substr(dv + std::strlen("default_value:"));
            auto stop = tail.find_first_of(",}");
            rest = trim(tail.substr(0, stop));
        }
        if (!set_param(p, name, rest)) return -1;
        ++count;
    }
    return count;
}

bool to_pik_params(Params const& p, bool return_approximate_solution, pik_params& out) {
    pik_params_default(&out);
    if (p.mode == "global") {
        out.mode = PIK_MODE_GLOBAL;
    } else if (p.mode == "local") {
        out.mode = PIK_MODE_LOCAL;
    } else {
        return false;  // src/pick_ik_plugin.cpp:204-207
    }
    out.gd_step_size = p.gd_step_size;
    out.gd_max_iters = p.gd_max_iters;
    out.gd_min_cost_delta = p.gd_min_cost_delta;
    out.position_threshold = p.position_threshold;
    out.orientation_threshold = p.orientation_threshold;
    out.approximate_solution_position_threshold = p.approximate_solution_position_threshold;
    out.approximate_solution_orientation_threshold = p.approximate_solution_orientation_threshold;
    out.approximate_solution_joint_threshold = p.approximate_solution_joint_threshold;
    out.approximate_solution_cost_threshold = p.approximate_solution_cost_threshold;
    out.cost_threshold = p.cost_threshold;
    out.position_scale = p.position_scale;
    out.rotation_scale = p.rotation_scale;
    out.center_joints_weight = p.center_joints_weight;
    out.avoid_joint_limits_weight = p.avoid_joint_limits_weight;
    out.minimal_displacement_weight = p.minimal_displacement_weight;
    out.stop_optimization_on_valid_solution = p.stop_optimization_on_valid_solution ? 1 : 0;
    out.memetic_num_threads = p.memetic_num_threads;
    out.memetic_stop_on_first_solution = p.memetic_stop_on_first_solution ? 1 : 0;
    out.memetic_population_size = p.memetic_population_size;
    out.memetic_elite_size = p.memetic_elite_size;
    out.memetic_wipeout_fitness_tol = p.memetic_wipeout_fitness_tol;
    out.memetic_max_generations = p.memetic_max_generations;
    out.memetic_gd_max_iters = p.memetic_gd_max_iters;
    out.memetic_gd_max_time = p.memetic_gd_max_time;
    out.return_approximate_solution = return_approximate_solution ? 1 : 0;
    out.rng_seed = p.rng_seed;
    return true;
}

// ------------------------------------------------------------------------------------------------
// plugin
// ------------------------------------------------------------------------------------------------
struct PickIKPlugin::Impl {
    std::string group_name, base_frame;
    std::vector<std::string> tip_frames;
    std::vector<std::string> joint_names, link_names;
    pik_robot* robot = nullptr;
    pik_solver* solver = nullptr;
    int n = 0;
    // constant transform model frame <- base frame (identity unless base_frame is a link behind fixed joints)
    double base_R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double base_t[3] = {0, 0, 0};
    bool base_identity = true;
    Params params;
    mutable std::mutex mutex;  // one in-flight call per solver handle (the reference serialises FK on fk_mutex_)
    mutable std::string last_error;

    ~Impl() {
        if (solver) pik_solver_destroy(solver);
        if (robot) pik_robot_destroy(robot);
    }
};

namespace {

void quat_to_matrix(double w, double x, double y, double z, double* R) {  // Eigen toRotationMatrix, no normalisation
    double const tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    double const twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
    double const tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}

void matrix_to_quat(double const* R, double* q) {  // Eigen Quaterniond(Matrix3d): w x y z
    double t = (R[0] + R[4]) + R[8];
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        q[0] = 0.5 * t;
        t = 0.5 / t;
        q[1] = (R[7] - R[5]) * t; q[2] = (R[2] - R[6]) * t; q[3] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        int const j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(((R[4 * i] - R[4 * j]) - R[4 * k]) + 1.0);
        q[1 + i] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R[3 * k + j] - R[3 * j + k]) * t;
        q[1 + j] = (R[3 * j + i] + R[3 * i + j]) * t;
        q[1 + k] = (R[3 * k + i] + R[3 * i + k]) * t;
    }
}

}  // namespace

PickIKPlugin::PickIKPlugin() : impl_(new Impl) {}
PickIKPlugin::~PickIKPlugin() = default;

compat::ChainModel compat::chain_from_urdf(std::string const& urdf_xml, std::string const& group_name,
                                           std::string const& base_link, std::string const& tip_link) {
    int32_t n = 0;
    int rc = pik_urdf_chain(urdf_xml.c_str(), base_link.c_str(), tip_link.c_str(), nullptr, 0, &n, nullptr, nullptr);
    if (rc != PIK_OK) throw std::invalid_argument(std::string("chain_from_urdf: ") + pik_status_string(rc));
    ChainModel m;
    m.group_name = group_name;
    m.model_frame = base_link;
    m.joints.resize((size_t)n);
    std::vector<char> jn((size_t)n * PIK_URDF_NAME_BYTES + 1), ln((size_t)n * PIK_URDF_NAME_BYTES + 1);
    rc = pik_urdf_chain(urdf_xml.c_str(), base_link.c_str(), tip_link.c_str(), m.joints.data(), n, &n, jn.data(), ln.data());
    if (rc != PIK_OK) throw std::invalid_argument(std::string("chain_from_urdf: ") + pik_status_string(rc));
    for (int32_t k = 0; k < n; ++k) {
        m.joint_names.emplace_back(jn.data() + (size_t)k * PIK_URDF_NAME_BYTES);
        m.link_names.emplace_back(ln.data() + (size_t)k * PIK_URDF_NAME_BYTES);
    }
    return m;
}

bool PickIKPlugin::initialize(compat::ChainModel const& model, std::string const& group_name,
                              std::string const& base_frame, std::vector<std::string> const& tip_frames,
                              double /*search_discretization*/, int device) {
    Impl& d = *impl_;
    d.last_error.clear();
    d.group_name = group_name;
    d.base_frame = base_frame;
    d.tip_frames = tip_frames;
    if (model.group_name != group_name) {  // src/pick_ik_plugin.cpp:45-49: unknown joint model group
        d.last_error = "failed to get joint model group " + group_name;
        return false;
    }
    if (model.joints.empty() || model.joints.size() != model.joint_names.size() ||
        model.joints.size() != model.link_names.size()) {
        d.last_error = "malformed chain model";
        return false;
    }
    // link_names_ = tip_frames_ (src/pick_ik_plugin.cpp:62); every tip must be a link of the model, else
    // std::invalid_argument (src/pick_ik_plugin.cpp:65-67 via get_link_indices, src/robot.cpp:107-120)
    d.link_names = tip_frames;
    std::vector<size_t> tip_idx;
    for (auto const& tip : tip_frames) {
        auto it = std::find(model.link_names.begin(), model.link_names.end(), tip);
        if (it == model.link_names.end()) throw std::invalid_argument("link not found: " + tip);
        tip_idx.push_back(static_cast<size_t>(it - model.link_names.begin()));
    }
    if (tip_idx.size() != 1) {
        d.last_error = "exactly one tip frame is supported by the batched engine";
        return false;
    }
    size_t const n_joints = tip_idx[0] + 1;  // the chain up to and including the tip link's parent joint
    // joint names of the group's active joints (src/pick_ik_plugin.cpp:52-58)
    d.joint_names.clear();
    for (size_t j = 0; j < n_joints; ++j)
        if (model.joints[j].type != PIK_JOINT_FIXED) d.joint_names.push_back(model.joint_names[j]);
    // base frame: the model frame itself, or a chain link that only fixed joints separate from it
    d.base_identity = true;
    if (!base_frame.empty() && base_frame != model.model_frame) {
        auto it = std::find(model.link_names.begin(), model.link_names.end(), base_frame);
        if (it == model.link_names.end()) {
            d.last_error = "base frame " + base_frame + " is not a link of the chain";
            return false;
        }
        size_t const upto = static_cast<size_t>(it - model.link_names.begin());
        double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
        for (size_t j = 0; j <= upto; ++j) {
            auto const& jd = model.joints[j];
            if (jd.type != PIK_JOINT_FIXED) {
                d.last_error = "base frame " + base_frame + " moves with the group: unsupported";
                return false;
            }
            double nR[9], nt[3];
            for (int r = 0; r < 3; ++r) {
                nt[r] = R[3 * r] * jd.origin_t[0] + R[3 * r + 1] * jd.origin_t[1] + R[3 * r + 2] * jd.origin_t[2] + t[r];
                for (int c = 0; c < 3; ++c)
                    nR[3 * r + c] = R[3 * r] * jd.origin_R[c] + R[3 * r + 1] * jd.origin_R[3 + c] + R[3 * r + 2] * jd.origin_R[6 + c];
            }
            std::memcpy(R, nR, sizeof(R));
            std::memcpy(t, nt, sizeof(t));
        }
        std::memcpy(d.base_R, R, sizeof(R));
        std::memcpy(d.base_t, t, sizeof(t));
        d.base_identity = false;
    }
    if (d.solver) { pik_solver_destroy(d.solver); d.solver = nullptr; }
    if (d.robot) { pik_robot_destroy(d.robot); d.robot = nullptr; }
    int rc = pik_robot_create(model.joints.data(), static_cast<int32_t>(n_joints), &d.robot);  // Robot::from, :68
    if (rc != PIK_OK) {
        d.last_error = std::string("pik_robot_create: ") + pik_status_string(rc);
        return false;
    }
    d.n = pik_robot_num_variables(d.robot);
    rc = pik_solver_create(d.robot, device, nullptr, &d.solver);
    if (rc != PIK_OK) {
        d.last_error = std::string("pik_solver_create: ") + pik_status_string(rc);
        return false;
    }
    return true;
}

namespace {

// pose given in the base frame -> px py pz qw qx qy qz in the model frame (transform_poses_to_frames,
// src/robot.cpp:169-181: base_frame_T * (Translation * Quaterniond(w,x,y,z)), quaternion not normalised)
void pose_to_goal(bool identity, double const* bR, double const* bt, Pose const& pose, double* g7) {
    if (identity) {
        g7[0] = pose.position.x; g7[1] = pose.position.y; g7[2] = pose.position.z;
        g7[3] = pose.orientation.w; g7[4] = pose.orientation.x; g7[5] = pose.orientation.y; g7[6] = pose.orientation.z;
        return;
    }
    double R[9], M[9];
    quat_to_matrix(pose.orientation.w, pose.orientation.x, pose.orientation.y, pose.orientation.z, R);
    double const p[3] = {pose.position.x, pose.position.y, pose.position.z};
    for (int r = 0; r < 3; ++r) {
        g7[r] = bR[3 * r] * p[0] + bR[3 * r + 1] * p[1] + bR[3 * r + 2] * p[2] + bt[r];
        for (int c = 0; c < 3; ++c) M[3 * r + c] = bR[3 * r] * R[c] + bR[3 * r + 1] * R[3 + c] + bR[3 * r + 2] * R[6 + c];
    }
    matrix_to_quat(M, g7 + 3);
}

}  // namespace

long PickIKPlugin::searchPositionIKBatch(std::vector<Pose> const& ik_poses,
                                         std::vector<std::vector<double>> const& seeds,
                                         std::vector<std::vector<double>>& solutions,
                                         std::vector<MoveItErrorCodes>& error_codes,
                                         KinematicsQueryOptions const& options) const {
    Impl const& d = *impl_;
    std::lock_guard<std::mutex> lock(d.mutex);
    d.last_error.clear();
    if (!d.solver) { d.last_error = "not initialized"; return -1; }
    int64_t const B = static_cast<int64_t>(ik_poses.size());
    if (static_cast<int64_t>(seeds.size()) != B && seeds.size() != 1) { d.last_error = "one seed per pose (or one for all)"; return -1; }
    pik_params pp;
    if (!to_pik_params(d.params, options.return_approximate_solution, pp)) {
        d.last_error = "Invalid solver mode: " + d.params.mode;
        std::fprintf(stderr, "[pick_ik] %s\n", d.last_error.c_str());
        return -1;
    }
    int const n = d.n;
    std::vector<double> goal(static_cast<size_t>(B) * 7), seed(seeds.size() * n), sol(static_cast<size_t>(B) * n), cost(B);
    std::vector<int32_t> err(B), its(B);
    for (int64_t b = 0; b < B; ++b) pose_to_goal(d.base_identity, d.base_R, d.base_t, ik_poses[b], &goal[7 * b]);
    for (size_t b = 0; b < seeds.size(); ++b) {
        if (static_cast<int>(seeds[b].size()) != n) { d.last_error = "seed size != number of variables"; return -1; }
        std::copy(seeds[b].begin(), seeds[b].end(), seed.begin() + b * n);
    }
    int64_t const stride = seeds.size() == 1 ? 0 : n;
    int rc = pik_solve_batch(d.solver, &pp, B, 0, goal.data(), seed.data(), stride, sol.data(), err.data(), cost.data(),
                             its.data(), PIK_MEM_HOST);
    if (rc != PIK_OK) {
        d.last_error = std::string("pik_solve_batch: ") + pik_status_string(rc) + " " + pik_solver_last_error(d.solver);
        return -1;
    }
    // approximate-solution gating (src/pick_ik_plugin.cpp:222-267), per problem
    std::vector<int32_t> approx_ok;
    if (options.return_approximate_solution && B > 0) {
        pik_params ap = pp;
        ap.cost_threshold = d.params.approximate_solution_cost_threshold;
        if (d.params.approximate_solution_cost_threshold <= 0.0)  // goals.clear(), :240-242
            ap.center_joints_weight = ap.avoid_joint_limits_weight = ap.minimal_displacement_weight = 0.0;
        // the reference tests `frame_tests` (the strict thresholds), not approx_frame_tests (:244-248): replicated
        approx_ok.resize(B);
        rc = pik_eval_cost(d.solver, &ap, B, goal.data(), seed.data(), stride, sol.data(), nullptr, approx_ok.data(),
                           nullptr, PIK_MEM_HOST);
        if (rc != PIK_OK) { d.last_error = std::string("pik_eval_cost: ") + pik_status_string(rc); return -1; }
    }
    solutions.assign(B, std::vector<double>());
    error_codes.assign(B, MoveItErrorCodes());
    long solved = 0;
    for (int64_t b = 0; b < B; ++b) {
        double const* sd = &seed[stride ? b * n : 0];
        solutions[b].assign(sol.begin() + b * n, sol.begin() + (b + 1) * n);
        error_codes[b].val = err[b] == PIK_SUCCESS ? MoveItErrorCodes::SUCCESS : MoveItErrorCodes::NO_IK_SOLUTION;
        if (options.return_approximate_solution) {
            bool valid = approx_ok[b] != 0;
            if (valid && d.params.approximate_solution_joint_threshold > 0.0)
                for (int i = 0; i < n; ++i)
                    if (std::fabs(solutions[b][i] - sd[i]) > d.params.approximate_solution_joint_threshold) { valid = false; break; }
            if (!valid) {
                error_codes[b].val = MoveItErrorCodes::NO_IK_SOLUTION;
                solutions[b].assign(sd, sd + n);
            }
        }
        if (error_codes[b].val == MoveItErrorCodes::SUCCESS) ++solved;
    }
    return solved;
}

bool PickIKPlugin::searchPositionIK(std::vector<Pose> const& ik_poses, std::vector<double> const& ik_seed_state,
                                    double timeout, std::vector<double> const& /*consistency_limits*/,
                                    std::vector<double>& solution, compat::IKCallbackFn const& solution_callback,
                                    MoveItErrorCodes& error_code, KinematicsQueryOptions const& options) const {
    Impl& d = *impl_;
    if (ik_poses.size() != d.tip_frames.size()) {  // one pose per tip frame (assert in src/goal.cpp:169)
        d.last_error = "one pose per tip frame expected";
        return false;
    }
    Params const base = d.params;
    bool found_valid_solution = false;
    auto const t0 = std::chrono::steady_clock::now();
    // Optimize until a valid solution or the timeout (src/pick_ik_plugin.cpp:147-150,162-291).  The
    // reference retries from the same seed; only its unseeded RNG differs between attempts, which the
    // attempt number stands in for here.  Local mode is deterministic: one attempt.
    for (unsigned attempt = 0;; ++attempt) {
        // memetic_num_threads species (src/ik_memetic.cpp:312-371) = replicas with distinct random streams
        int const species = (base.mode == "global" && base.memetic_num_threads > 1) ? base.memetic_num_threads : 1;
        std::vector<Pose> poses(species, ik_poses.front());
        std::vector<std::vector<double>> seeds(1, ik_seed_state), sols;
        std::vector<MoveItErrorCodes> codes;
        {
            std::lock_guard<std::mutex> lock(d.mutex);
            d.params.rng_seed = base.rng_seed + 0x9E3779B97F4A7C15ull * attempt;
        }
        long const solved = searchPositionIKBatch(poses, seeds, sols, codes, options);
        {
            std::lock_guard<std::mutex> lock(d.mutex);
            d.params.rng_seed = base.rng_seed;
        }
        if (solved < 0) return false;  // invalid mode / engine error (src/pick_ik_plugin.cpp:204-207)
        int pick = 0;
        for (int s = 0; s < species; ++s)
            if (codes[s].val == MoveItErrorCodes::SUCCESS) { pick = s; break; }
        error_code.val = codes[pick].val;
        solution = codes[pick].val == MoveItErrorCodes::SUCCESS ? sols[pick] : ik_seed_state;  // :209-217
        if (solution_callback && error_code.val == MoveItErrorCodes::SUCCESS)  // :270-274; the callback may veto
            solution_callback(ik_poses.front(), solution, error_code);
        found_valid_solution = error_code.val == MoveItErrorCodes::SUCCESS;
        double const elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (found_valid_solution || elapsed >= timeout || base.mode != "global") break;
    }
    return found_valid_solution;
}

bool PickIKPlugin::searchPositionIK(Pose const& ik_pose, std::vector<double> const& ik_seed_state, double timeout,
                                    std::vector<double>& solution, MoveItErrorCodes& error_code,
                                    KinematicsQueryOptions const& options) const {
    return searchPositionIK(std::vector<Pose>{ik_pose}, ik_seed_state, timeout, {}, solution, compat::IKCallbackFn(), error_code, options);
}

bool PickIKPlugin::searchPositionIK(Pose const& ik_pose, std::vector<double> const& ik_seed_state, double timeout,
                                    std::vector<double> const& consistency_limits, std::vector<double>& solution,
                                    MoveItErrorCodes& error_code, KinematicsQueryOptions const& options) const {
    return searchPositionIK(std::vector<Pose>{ik_pose}, ik_seed_state, timeout, consistency_limits, solution, compat::IKCallbackFn(),
                            error_code, options);
}

bool PickIKPlugin::searchPositionIK(Pose const& ik_pose, std::vector<double> const& ik_seed_state, double timeout,
                                    std::vector<double>& solution, compat::IKCallbackFn const& solution_callback,
                                    MoveItErrorCodes& error_code, KinematicsQueryOptions const& options) const {
    return searchPositionIK(std::vector<Pose>{ik_pose}, ik_seed_state, timeout, {}, solution, solution_callback, error_code, options);
}

bool PickIKPlugin::searchPositionIK(Pose const& ik_pose, std::vector<double> const& ik_seed_state, double timeout,
                                    std::vector<double> const& consistency_limits, std::vector<double>& solution,
                                    compat::IKCallbackFn const& solution_callback, MoveItErrorCodes& error_code,
                                    KinematicsQueryOptions const& options) const {
    return searchPositionIK(std::vector<Pose>{ik_pose}, ik_seed_state, timeout, consistency_limits, solution, solution_callback,
                            error_code, options);
}

std::vector<std::string> const& PickIKPlugin::getJointNames() const { return impl_->joint_names; }
std::vector<std::string> const& PickIKPlugin::getLinkNames() const { return impl_->link_names; }

bool PickIKPlugin::getPositionFK(std::vector<std::string> const&, std::vector<double> const&, std::vector<Pose>&) const {
    return false;
}
bool PickIKPlugin::getPositionIK(Pose const&, std::vector<double> const&, std::vector<double>&, MoveItErrorCodes&,
                                 KinematicsQueryOptions const&) const {
    return false;
}

void PickIKPlugin::setParams(Params const& p) {
    std::lock_guard<std::mutex> lock(impl_->mutex);
    impl_->params = p;
}
Params const& PickIKPlugin::getParams() const { return impl_->params; }
std::string const& PickIKPlugin::lastError() const { return impl_->last_error; }

}  // namespace pick_ik_b200
