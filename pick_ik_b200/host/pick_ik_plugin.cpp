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
    int device = 0;
    int n = 0;
    int n_tips = 1;
    // searchPositionIK is const and re-entrant in the reference (it serialises on the FK mutex only): every call in
    // flight takes a solver (own stream, own device buffers) from this pool and puts it back
    mutable std::vector<pik_solver*> idle_solvers;
    // constant transform model frame <- base frame (identity unless base_frame is a link behind fixed joints)
    double base_R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double base_t[3] = {0, 0, 0};
    bool base_identity = true;
    Params params;
    mutable std::mutex mutex;  // guards params, the pool and last_error
    mutable std::string last_error;

    pik_solver* take_solver() const {
        {
            std::lock_guard<std::mutex> lock(mutex);
            if (!idle_solvers.empty()) {
                pik_solver* s = idle_solvers.back();
                idle_solvers.pop_back();
                return s;
            }
        }
        pik_solver* s = nullptr;
        int const rc = pik_solver_create(robot, device, nullptr, &s);
        if (rc != PIK_OK) {
            set_error(std::string("pik_solver_create: ") + pik_status_string(rc));
            return nullptr;
        }
        return s;
    }
    void give_solver(pik_solver* s) const {
        std::lock_guard<std::mutex> lock(mutex);
        idle_solvers.push_back(s);
    }
    void set_error(std::string const& e) const {
        std::lock_guard<std::mutex> lock(mutex);
        last_error = e;
    }
    void drop_solvers() {
        for (pik_solver* s : idle_solvers) pik_solver_destroy(s);
        idle_solvers.clear();
    }

    ~Impl() {
        drop_solvers();
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

compat::ChainModel compat::model_from_urdf_srdf(std::string const& urdf_xml, std::string const& srdf_xml,
                                                std::string const& group_name, std::vector<std::string>& tip_frames) {
    int32_t n_tips = 0;
    char base[PIK_URDF_NAME_BYTES];
    int rc = pik_srdf_group(srdf_xml.c_str(), group_name.c_str(), base, nullptr, 0, &n_tips);
    if (rc != PIK_OK) throw std::invalid_argument(std::string("model_from_urdf_srdf: group ") + group_name + ": " + pik_status_string(rc));
    std::vector<char> tips((size_t)n_tips * PIK_URDF_NAME_BYTES + 1);
    rc = pik_srdf_group(srdf_xml.c_str(), group_name.c_str(), base, tips.data(), n_tips, &n_tips);
    if (rc != PIK_OK) throw std::invalid_argument(std::string("model_from_urdf_srdf: ") + pik_status_string(rc));
    tip_frames.clear();
    std::vector<char const*> tip_ptrs;
    for (int32_t t = 0; t < n_tips; ++t) {
        tip_frames.emplace_back(tips.data() + (size_t)t * PIK_URDF_NAME_BYTES);
        tip_ptrs.push_back(tips.data() + (size_t)t * PIK_URDF_NAME_BYTES);
    }
    int32_t n = 0;
    rc = pik_urdf_tree(urdf_xml.c_str(), base, tip_ptrs.data(), n_tips, nullptr, 0, &n, nullptr, nullptr, nullptr, nullptr, nullptr,
                       nullptr, nullptr);
    if (rc != PIK_OK) throw std::invalid_argument(std::string("model_from_urdf_srdf: ") + pik_status_string(rc));
    ChainModel m;
    m.group_name = group_name;
    m.model_frame = base;
    m.joints.resize((size_t)n);
    m.parent.resize((size_t)n);
    m.mimic_of.resize((size_t)n);
    m.mimic_factor.resize((size_t)n);
    m.mimic_offset.resize((size_t)n);
    std::vector<int32_t> tip_joint((size_t)n_tips);
    std::vector<char> jn((size_t)n * PIK_URDF_NAME_BYTES + 1), ln((size_t)n * PIK_URDF_NAME_BYTES + 1);
    rc = pik_urdf_tree(urdf_xml.c_str(), base, tip_ptrs.data(), n_tips, m.joints.data(), n, &n, m.parent.data(), tip_joint.data(),
                       m.mimic_of.data(), m.mimic_factor.data(), m.mimic_offset.data(), jn.data(), ln.data());
    if (rc != PIK_OK) throw std::invalid_argument(std::string("model_from_urdf_srdf: ") + pik_status_string(rc));
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
    size_t const nj = model.joints.size();
    if (nj == 0 || nj != model.joint_names.size() || nj != model.link_names.size() ||
        (!model.parent.empty() && model.parent.size() != nj) ||
        (!model.mimic_of.empty() && (model.mimic_of.size() != nj || model.mimic_factor.size() != nj || model.mimic_offset.size() != nj))) {
        d.last_error = "malformed chain model";
        return false;
    }
    if (tip_frames.empty() || tip_frames.size() > PIK_MAX_TIPS) {
        d.last_error = "between 1 and " + std::to_string(PIK_MAX_TIPS) + " tip frames are supported";
        return false;
    }
    // link_names_ = tip_frames_ (src/pick_ik_plugin.cpp:62); every tip must be a link of the model, else
    // std::invalid_argument (src/pick_ik_plugin.cpp:65-67 via get_link_indices, src/robot.cpp:107-120)
    d.link_names = tip_frames;
    std::vector<int32_t> tip_joint;
    for (auto const& tip : tip_frames) {
        auto it = std::find(model.link_names.begin(), model.link_names.end(), tip);
        if (it == model.link_names.end()) throw std::invalid_argument("link not found: " + tip);
        tip_joint.push_back(static_cast<int32_t>(it - model.link_names.begin()));
    }
    auto parent_of = [&](size_t j) { return model.parent.empty() ? static_cast<int32_t>(j) - 1 : model.parent[j]; };
    // the joints between the tips and the root are the ones in use (get_active_variable_indices, src/robot.cpp:122-160)
    std::vector<char> used(nj, 0);
    for (int32_t t : tip_joint)
        for (int32_t j = t; j >= 0; j = parent_of(static_cast<size_t>(j))) used[static_cast<size_t>(j)] = 1;
    // compact to the used joints (order kept: parents stay in front of their children)
    std::vector<pik_joint_desc> joints;
    std::vector<int32_t> parent, mimic_of, new_index(nj, -1);
    std::vector<double> mimic_factor, mimic_offset;
    bool any_mimic = false;
    d.joint_names.clear();
    for (size_t j = 0; j < nj; ++j) {
        if (!used[j]) continue;
        new_index[j] = static_cast<int32_t>(joints.size());
        joints.push_back(model.joints[j]);
        int32_t const p = parent_of(j);
        parent.push_back(p >= 0 ? new_index[static_cast<size_t>(p)] : -1);
        int32_t const m = model.mimic_of.empty() ? -1 : model.mimic_of[j];
        if (m >= 0 && (static_cast<size_t>(m) >= nj || !used[static_cast<size_t>(m)] || new_index[static_cast<size_t>(m)] < 0)) {
            d.last_error = "joint " + model.joint_names[j] + " mimics a joint outside the chains of the tip frames";
            return false;
        }
        mimic_of.push_back(m >= 0 ? new_index[static_cast<size_t>(m)] : -1);
        mimic_factor.push_back(m >= 0 ? model.mimic_factor[j] : 1.0);
        mimic_offset.push_back(m >= 0 ? model.mimic_offset[j] : 0.0);
        any_mimic = any_mimic || m >= 0;
        // joint names of the group's active joints (src/pick_ik_plugin.cpp:52-58)
        if (model.joints[j].type != PIK_JOINT_FIXED && m < 0) d.joint_names.push_back(model.joint_names[j]);
    }
    for (int32_t& t : tip_joint) t = new_index[static_cast<size_t>(t)];
    // base frame: the model frame itself, or a link that only fixed joints separate from it
    d.base_identity = true;
    if (!base_frame.empty() && base_frame != model.model_frame) {
        auto it = std::find(model.link_names.begin(), model.link_names.end(), base_frame);
        if (it == model.link_names.end()) {
            d.last_error = "base frame " + base_frame + " is not a link of the model";
            return false;
        }
        std::vector<size_t> path;
        for (int32_t j = static_cast<int32_t>(it - model.link_names.begin()); j >= 0; j = parent_of(static_cast<size_t>(j)))
            path.push_back(static_cast<size_t>(j));
        double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
        for (size_t a = path.size(); a-- > 0;) {
            auto const& jd = model.joints[path[a]];
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
    d.drop_solvers();
    if (d.robot) { pik_robot_destroy(d.robot); d.robot = nullptr; }
    // Robot::from, src/pick_ik_plugin.cpp:68
    int rc = pik_robot_create_tree(joints.data(), static_cast<int32_t>(joints.size()), parent.data(), tip_joint.data(),
                                   static_cast<int32_t>(tip_joint.size()), any_mimic ? mimic_of.data() : nullptr,
                                   any_mimic ? mimic_factor.data() : nullptr, any_mimic ? mimic_offset.data() : nullptr,
                                   &d.robot);
    if (rc != PIK_OK) {
        d.last_error = std::string("pik_robot_create_tree: ") + pik_status_string(rc);
        return false;
    }
    d.n = pik_robot_num_variables(d.robot);
    d.n_tips = pik_robot_num_tips(d.robot);
    d.device = device;
    pik_solver* first = d.take_solver();  // fails here, not in the first solve, when there is no device
    if (!first) return false;
    d.give_solver(first);
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
                                         KinematicsQueryOptions const& options, std::vector<double>* costs,
                                         unsigned long long rng_seed) const {
    Impl const& d = *impl_;
    if (!d.robot) { d.set_error("not initialized"); return -1; }
    int const n = d.n, T = d.n_tips;
    if (ik_poses.size() % static_cast<size_t>(T) != 0) { d.set_error("one pose per tip frame and problem expected"); return -1; }
    int64_t const B = static_cast<int64_t>(ik_poses.size()) / T;
    if (static_cast<int64_t>(seeds.size()) != B && seeds.size() != 1) { d.set_error("one seed per problem (or one for all)"); return -1; }
    Params params;
    {
        std::lock_guard<std::mutex> lock(d.mutex);  // the reference re-reads its parameters on every solve (:86)
        params = d.params;
    }
    pik_params pp;
    if (!to_pik_params(params, options.return_approximate_solution, pp)) {
        d.set_error("Invalid solver mode: " + params.mode);
        std::fprintf(stderr, "[pick_ik] Invalid solver mode: %s\n", params.mode.c_str());
        return -1;
    }
    if (rng_seed != 0) pp.rng_seed = rng_seed;
    std::vector<double> goal(static_cast<size_t>(B) * 7 * T), seed(seeds.size() * n), sol(static_cast<size_t>(B) * n), cost(B);
    std::vector<int32_t> err(B), its(B);
    for (size_t k = 0; k < ik_poses.size(); ++k) pose_to_goal(d.base_identity, d.base_R, d.base_t, ik_poses[k], &goal[7 * k]);
    for (size_t b = 0; b < seeds.size(); ++b) {
        if (static_cast<int>(seeds[b].size()) != n) { d.set_error("seed size != number of variables"); return -1; }
        std::copy(seeds[b].begin(), seeds[b].end(), seed.begin() + b * n);
    }
    int64_t const stride = seeds.size() == 1 ? 0 : n;
    pik_solver* solver = d.take_solver();
    if (!solver) return -1;
    int rc = pik_solve_batch(solver, &pp, B, 0, goal.data(), seed.data(), stride, sol.data(), err.data(), cost.data(),
                             its.data(), PIK_MEM_HOST);
    if (rc != PIK_OK) {
        d.set_error(std::string("pik_solve_batch: ") + pik_status_string(rc) + " " + pik_solver_last_error(solver));
        d.give_solver(solver);
        return -1;
    }
    // approximate-solution gating (src/pick_ik_plugin.cpp:222-267), per problem
    std::vector<int32_t> approx_ok;
    if (options.return_approximate_solution && B > 0) {
        pik_params ap = pp;
        ap.cost_threshold = params.approximate_solution_cost_threshold;
        if (params.approximate_solution_cost_threshold <= 0.0)  // goals.clear(), :240-242
            ap.center_joints_weight = ap.avoid_joint_limits_weight = ap.minimal_displacement_weight = 0.0;
        // the reference tests `frame_tests` (the strict thresholds), not approx_frame_tests (:244-248): replicated
        approx_ok.resize(B);
        rc = pik_eval_cost(solver, &ap, B, goal.data(), seed.data(), stride, sol.data(), nullptr, approx_ok.data(),
                           nullptr, PIK_MEM_HOST);
        if (rc != PIK_OK) {
            d.set_error(std::string("pik_eval_cost: ") + pik_status_string(rc));
            d.give_solver(solver);
            return -1;
        }
    }
    d.give_solver(solver);
    solutions.assign(B, std::vector<double>());
    error_codes.assign(B, MoveItErrorCodes());
    if (costs) costs->assign(cost.begin(), cost.end());
    long solved = 0;
    for (int64_t b = 0; b < B; ++b) {
        double const* sd = &seed[stride ? b * n : 0];
        solutions[b].assign(sol.begin() + b * n, sol.begin() + (b + 1) * n);
        error_codes[b].val = err[b] == PIK_SUCCESS ? MoveItErrorCodes::SUCCESS : MoveItErrorCodes::NO_IK_SOLUTION;
        if (options.return_approximate_solution) {
            bool valid = approx_ok[b] != 0;
            if (valid && params.approximate_solution_joint_threshold > 0.0)
                for (int i = 0; i < n; ++i)
                    if (std::fabs(solutions[b][i] - sd[i]) > params.approximate_solution_joint_threshold) { valid = false; break; }
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
                                    compat::IKCostFn const& cost_function, MoveItErrorCodes& error_code,
                                    KinematicsQueryOptions const& options, compat::RobotState const* /*context_state*/) const {
    Impl const& d = *impl_;
    if (ik_poses.size() != d.tip_frames.size()) {  // one pose per tip frame (assert in src/goal.cpp:169)
        d.set_error("one pose per tip frame expected");
        error_code.val = MoveItErrorCodes::NO_IK_SOLUTION;
        solution = ik_seed_state;
        return false;
    }
    if (cost_function) {
        // src/pick_ik_plugin.cpp:130-135 adds one goal per pose that calls back into host code with a RobotState
        // (src/goal.cpp:146-161); the device kernels cannot evaluate it and silently ignoring it would change what is
        // optimised, so the call is refused with a defined status
        d.set_error("a custom IKCostFn is a host callback and is not supported by the GPU solver");
        std::fprintf(stderr, "[pick_ik] %s\n", "a custom IKCostFn is a host callback and is not supported by the GPU solver");
        error_code.val = MoveItErrorCodes::NO_IK_SOLUTION;
        solution = ik_seed_state;
        return false;
    }
    Params base;
    {
        std::lock_guard<std::mutex> lock(d.mutex);
        base = d.params;
    }
    bool found_valid_solution = false;
    auto const t0 = std::chrono::steady_clock::now();
    // Optimize until a valid solution or the timeout (src/pick_ik_plugin.cpp:147-150,162-291).  The reference retries
    // from the same seed; only its unseeded RNG differs between attempts, which the attempt number stands in for here
    // (passed down as the random stream of the call: nothing shared is modified).  Local mode is deterministic: one
    // attempt.
    for (unsigned attempt = 0;; ++attempt) {
        std::vector<std::vector<double>> seeds(1, ik_seed_state), sols;
        std::vector<MoveItErrorCodes> codes;
        unsigned long long const stream = base.rng_seed + 0x9E3779B97F4A7C15ull * attempt;
        long const solved = searchPositionIKBatch(ik_poses, seeds, sols, codes, options, nullptr, stream ? stream : 1);
        if (solved < 0) {  // invalid mode / engine error (src/pick_ik_plugin.cpp:204-207)
            error_code.val = MoveItErrorCodes::NO_IK_SOLUTION;
            solution = ik_seed_state;
            return false;
        }
        error_code.val = codes[0].val;
        solution = codes[0].val == MoveItErrorCodes::SUCCESS ? sols[0] : ik_seed_state;  // :209-217
        if (solution_callback && error_code.val == MoveItErrorCodes::SUCCESS)  // :270-274; the callback may veto
            solution_callback(ik_poses.front(), solution, error_code);
        found_valid_solution = error_code.val == MoveItErrorCodes::SUCCESS;
        double const elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (found_valid_solution || elapsed >= timeout || base.mode != "global") break;
    }
    return found_valid_solution;
}

bool PickIKPlugin::searchPositionIK(std::vector<Pose> const& ik_poses, std::vector<double> const& ik_seed_state,
                                    double timeout, std::vector<double> const& consistency_limits,
                                    std::vector<double>& solution, compat::IKCallbackFn const& solution_callback,
                                    MoveItErrorCodes& error_code, KinematicsQueryOptions const& options,
                                    compat::RobotState const* context_state) const {
    return searchPositionIK(ik_poses, ik_seed_state, timeout, consistency_limits, solution, solution_callback,
                            compat::IKCostFn(), error_code, options, context_state);
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

void PickIKPlugin::setParams(Params const& p) const {
    std::lock_guard<std::mutex> lock(impl_->mutex);
    impl_->params = p;
}
Params const& PickIKPlugin::getParams() const { return impl_->params; }
std::string const& PickIKPlugin::lastError() const { return impl_->last_error; }

}  // namespace pick_ik_b200
