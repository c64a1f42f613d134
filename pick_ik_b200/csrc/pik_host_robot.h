// pik_host_robot.h -- host-side flattening of a joint chain into the device robot table, and the
// YAML -> solver parameter mapping.
//
// Reference: Robot::from (src/robot.cpp:44-85) for the variable table; make_fk_fn
// (src/fk_moveit.cpp:11-35) + MoveIt's RobotState chain walk for what the table must reproduce;
// pick_ik_plugin.cpp:97-129,166-196 for the parameter mapping.
#pragma once

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "../../include/pik.h"
#include "pik_types.h"

namespace pik {

inline void host_mat_mul(const double* A, const double* B, double* C) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            C[3 * r + c] = std::fma(A[3 * r + 2], B[6 + c], std::fma(A[3 * r + 1], B[3 + c], A[3 * r] * B[c]));
}

inline void host_mat_vec_add(const double* A, const double* v, const double* t, double* out) {
    for (int r = 0; r < 3; ++r)
        out[r] = std::fma(A[3 * r + 2], v[2], std::fma(A[3 * r + 1], v[1], std::fma(A[3 * r], v[0], t[r])));
}

// does the constant rotation R have the sparsity pattern `cls` (exact tests: the device skips only entries that
// ARE 0 or 1)?  An identity has every pattern.
inline bool origin_has_pattern(const double* R, int cls) {
    auto z = [&](int i) { return R[i] == 0.0; };
    switch (cls) {
        case kOrgGeneral: return true;
        case kOrgRotX: return R[0] == 1.0 && z(1) && z(2) && z(3) && z(6);
        case kOrgRotY: return R[4] == 1.0 && z(1) && z(3) && z(5) && z(7);
        case kOrgRotZ: return R[8] == 1.0 && z(2) && z(5) && z(6) && z(7);
        case kOrgIdentity: return origin_has_pattern(R, kOrgRotX) && origin_has_pattern(R, kOrgRotY) && origin_has_pattern(R, kOrgRotZ);
        default: return false;
    }
}

inline int classify_origin(const double* R) {
    for (int cls : {kOrgIdentity, kOrgRotX, kOrgRotY, kOrgRotZ})
        if (origin_has_pattern(R, cls)) return cls;
    return kOrgGeneral;
}

// acc <- product, root to leaf, of the origins of the fixed joints between joint j and its nearest moving ancestor
// (exclusive), then the origin of j itself.  Returns that ancestor (-1: the model root).
inline int host_fold_origins(const pik_joint_desc* joints, const int32_t* parent, int j, double* accR, double* acct) {
    int path[64], np = 0;
    path[np++] = j;
    int k = parent ? parent[j] : j - 1;
    while (k >= 0 && joints[k].type == PIK_JOINT_FIXED && np < 64) {
        path[np++] = k;
        k = parent ? parent[k] : k - 1;
    }
    for (int a = np - 1; a >= 0; --a) {
        const pik_joint_desc& jd = joints[path[a]];
        if (a == np - 1) {
            std::memcpy(accR, jd.origin_R, 9 * sizeof(double));
            std::memcpy(acct, jd.origin_t, 3 * sizeof(double));
        } else {
            double nR[9], nt[3];
            host_mat_vec_add(accR, jd.origin_t, acct, nt);
            host_mat_mul(accR, jd.origin_R, nR);
            std::memcpy(accR, nR, sizeof(nR));
            std::memcpy(acct, nt, sizeof(nt));
        }
    }
    return k;
}

// Flattens a kinematic tree (Robot::from, src/robot.cpp:44-85; get_active_variable_indices, :122-160; the FK of
// src/fk_moveit.cpp:20-34) into the device table.  Fixed joints are folded into the constant origin of the moving
// joint below them (or into a tip transform); each moving joint becomes one step.  parent == nullptr: a serial
// chain in joint order; mimic_of == nullptr: no mimic joints.  rcp_out (optional) [kMaxVars]: 1 / max velocity per
// variable.  A serial single-tip chain of one-variable joints gets the table the chain kernels read (is_tree = 0).
inline int build_dev_robot_tree(const pik_joint_desc* joints, int n_joints, const int32_t* parent, const int32_t* tip_joint,
                                int n_tips, const int32_t* mimic_of, const double* mimic_factor, const double* mimic_offset,
                                DevRobot* out, double* rcp_out) {
    std::memset(out, 0, sizeof(*out));
    if (!joints || n_joints <= 0 || n_joints > 64 || !tip_joint) return PIK_E_INVALID_ROBOT;
    if (n_tips < 1 || n_tips > kMaxTips) return PIK_E_UNSUPPORTED;
    int step_of[64];
    int n = 0, ns = 0;
    bool simple = n_tips == 1;
    double max_velocity_rcp[kMaxVars] = {0};
    for (int j = 0; j < n_joints; ++j) {
        const pik_joint_desc& jd = joints[j];
        step_of[j] = -1;
        if (parent && !(parent[j] >= -1 && parent[j] < j)) return PIK_E_INVALID_ROBOT;  // parents precede children
        if (jd.type == PIK_JOINT_FIXED) continue;
        if (jd.type < PIK_JOINT_FIXED || jd.type > PIK_JOINT_PLANAR) return PIK_E_INVALID_ROBOT;
        if (ns >= kMaxVars) return PIK_E_UNSUPPORTED;
        const int up = host_fold_origins(joints, parent, j, out->R[ns], out->t[ns]);
        out->parent[ns] = up >= 0 ? step_of[up] : -1;
        if (up >= 0 && step_of[up] < 0) return PIK_E_INVALID_ROBOT;
        if (out->parent[ns] != ns - 1) simple = false;
        // unit axis (RevoluteJointModel / PrismaticJointModel::setAxis normalise theirs); pik_robot_create admits
        // |axis|^2 within 1e-6 of 1
        double x = jd.axis[0], y = jd.axis[1], z = jd.axis[2];
        const double a2 = x * x + y * y + z * z;
        if (std::fabs(a2 - 1.0) > 1.0e-14 && a2 > 0.0) {  /* an axis that is unit to rounding is kept as given */
            const double nrm = std::sqrt(a2);
            x /= nrm; y /= nrm; z /= nrm;
        }
        out->axis[ns][0] = x; out->axis[ns][1] = y; out->axis[ns][2] = z;
        out->axis_sq[ns][0] = x * x; out->axis_sq[ns][1] = y * y; out->axis_sq[ns][2] = z * z;
        out->axis_sq[ns][3] = x * y; out->axis_sq[ns][4] = x * z; out->axis_sq[ns][5] = y * z;
        out->sign[ns] = 1.0;
        out->mimic_factor[ns] = 1.0;
        out->mimic_offset[ns] = 0.0;
        int n_vars = 1;
        if (jd.type == PIK_JOINT_PRISMATIC) {
            out->kind[ns] = kPrismatic;
            if (std::fabs(x) == 1.0 && y == 0.0 && z == 0.0) { out->kind[ns] = kPrisX; out->sign[ns] = x; }
            if (x == 0.0 && std::fabs(y) == 1.0 && z == 0.0) { out->kind[ns] = kPrisY; out->sign[ns] = y; }
            if (x == 0.0 && y == 0.0 && std::fabs(z) == 1.0) { out->kind[ns] = kPrisZ; out->sign[ns] = z; }
        } else if (jd.type == PIK_JOINT_REVOLUTE) {
            out->kind[ns] = kRevGeneral;
            if (std::fabs(x) == 1.0 && y == 0.0 && z == 0.0) { out->kind[ns] = kRevX; out->sign[ns] = x; }
            if (x == 0.0 && std::fabs(y) == 1.0 && z == 0.0) { out->kind[ns] = kRevY; out->sign[ns] = y; }
            if (x == 0.0 && y == 0.0 && std::fabs(z) == 1.0) { out->kind[ns] = kRevZ; out->sign[ns] = z; }
        } else if (jd.type == PIK_JOINT_FLOATING) {
            out->kind[ns] = kFloating;
            n_vars = 7;
            simple = false;
        } else {
            out->kind[ns] = kPlanar;
            n_vars = 3;
            simple = false;
        }
        const int m = mimic_of ? mimic_of[j] : -1;
        if (m >= 0) {
            // a mimic joint follows the variable of its master and owns none (src/robot.cpp:145-147)
            if (n_vars != 1 || m >= j || step_of[m] < 0 || joints[m].type > PIK_JOINT_PRISMATIC) return PIK_E_INVALID_ROBOT;
            const int ms = step_of[m];
            out->var0[ns] = out->var0[ms];
            out->mimic_factor[ns] = mimic_factor[j] * out->mimic_factor[ms];
            out->mimic_offset[ns] = mimic_factor[j] * out->mimic_offset[ms] + mimic_offset[j];
            simple = false;
        } else {
            if (n + n_vars > kMaxVars) return PIK_E_UNSUPPORTED;
            out->var0[ns] = n;
            for (int k = 0; k < n_vars; ++k) {
                // VariableBounds as MoveIt's joint models set them: the translation variables of floating / planar joints
                // take the description's bounds (unbounded by default), quaternion components [-1, 1], the planar
                // angle is unbounded with the nominal range -pi .. pi
                int bounded = jd.bounded ? 1 : 0;
                double lo = jd.min_position, hi = jd.max_position;
                if (jd.type == PIK_JOINT_FLOATING && k >= 3) { bounded = 1; lo = -1.0; hi = 1.0; }
                if (jd.type == PIK_JOINT_PLANAR && k == 2) { bounded = 0; lo = -M_PI; hi = M_PI; }
                // Robot::from, robot.cpp:52-72
                out->bounded[n + k] = bounded;
                out->vmin[n + k] = lo;
                out->vmax[n + k] = hi;
                out->vmid[n + k] = 0.5 * (lo + hi);
                out->vhalf[n + k] = bounded ? (hi - lo) / 2.0 : M_PI;
                max_velocity_rcp[n + k] = jd.max_velocity > 0.0 ? 1.0 / jd.max_velocity : 0.0;
            }
            n += n_vars;
        }
        step_of[j] = ns++;
    }
    if (n == 0) return PIK_E_INVALID_ROBOT;
    out->n = n;
    out->n_steps = ns;
    out->n_tips = n_tips;
    for (int i = 0; i < n; ++i)
        if (!out->bounded[i]) out->any_unbounded = 1;
    for (int t = 0; t < n_tips; ++t) {
        const int j = tip_joint[t];
        if (j < 0 || j >= n_joints) return PIK_E_INVALID_ROBOT;
        if (joints[j].type != PIK_JOINT_FIXED) {
            out->tip_step[t] = step_of[j];
            out->tip_has[t] = 0;
            const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            std::memcpy(out->tips_R[t], I, sizeof(I));
        } else {
            const int up = host_fold_origins(joints, parent, j, out->tips_R[t], out->tips_t[t]);
            out->tip_step[t] = up >= 0 ? step_of[up] : -1;
            out->tip_has[t] = 1;
        }
    }
    if (out->tip_step[0] != ns - 1) simple = false;
    // where each step finds its parent frame while the tree is walked in step order
    int n_saved = 0;
    for (int k = 0; k < ns; ++k) out->save_slot[k] = -1;
    for (int k = 0; k < ns; ++k) {
        const int p = out->parent[k];
        if (p == k - 1 && k > 0) {
            out->load_slot[k] = -1;
        } else if (p < 0) {
            out->load_slot[k] = -2;
        } else {
            if (out->save_slot[p] < 0) {
                if (n_saved >= kMaxSavedFrames) return PIK_E_UNSUPPORTED;
                out->save_slot[p] = n_saved++;
            }
            out->load_slot[k] = out->save_slot[p];
        }
    }
    out->is_tree = simple ? 0 : 1;
    // the table of the chain kernels: the tip transform as entry n of the origin table (identity when the chain ends in a
    // moving joint)
    out->has_tip = out->tip_has[0];
    if (simple) {
        std::memcpy(out->tip_R, out->tips_R[0], sizeof(out->tip_R));
        std::memcpy(out->tip_t, out->tips_t[0], sizeof(out->tip_t));
        std::memcpy(out->R[n], out->tips_R[0], 9 * sizeof(double));
        std::memcpy(out->t[n], out->tips_t[0], 3 * sizeof(double));
        for (int i = 0; i <= n; ++i) out->ocls[i] = classify_origin(out->R[i]);
    }
    // robot.cpp:69-82
    double divisor = 0.0;
    for (int i = 0; i < n; ++i) {
        out->vfac[i] = 1.0 / static_cast<double>(n);
        divisor += max_velocity_rcp[i];
    }
    if (divisor > 0.0)
        for (int i = 0; i < n; ++i) out->vfac[i] = max_velocity_rcp[i] / divisor;
    if (rcp_out)
        for (int i = 0; i < kMaxVars; ++i) rcp_out[i] = max_velocity_rcp[i];
    return PIK_OK;
}

// serial chain model root -> tip link, joints in chain order
inline int build_dev_robot(const pik_joint_desc* joints, int n_joints, DevRobot* out, double* rcp_out = nullptr) {
    if (n_joints <= 0) return PIK_E_INVALID_ROBOT;
    const int32_t tip = n_joints - 1;
    return build_dev_robot_tree(joints, n_joints, nullptr, &tip, 1, nullptr, nullptr, nullptr, out, rcp_out);
}

// pick_ik_plugin.cpp:97-129 (which tests / goals exist) and :166-196 (solver structs)
inline DevParams make_dev_params(const pik_params& p) {
    DevParams d;
    std::memset(&d, 0, sizeof(d));
    d.step_size = p.gd_step_size;
    d.min_cost_delta = p.gd_min_cost_delta;
    d.position_threshold = p.position_threshold;
    d.orientation_threshold = p.orientation_threshold;
    d.cost_threshold_sq = p.cost_threshold * p.cost_threshold;
    d.position_scale = p.position_scale;
    d.rotation_scale = p.rotation_scale;
    d.w2_center = p.center_joints_weight > 0.0 ? p.center_joints_weight * p.center_joints_weight : 0.0;
    d.w2_avoid = p.avoid_joint_limits_weight > 0.0 ? p.avoid_joint_limits_weight * p.avoid_joint_limits_weight : 0.0;
    d.w2_mindisp =
        p.minimal_displacement_weight > 0.0 ? p.minimal_displacement_weight * p.minimal_displacement_weight : 0.0;
    d.wipeout_tol = p.memetic_wipeout_fitness_tol;
    d.gd_max_iters = (p.mode == PIK_MODE_GLOBAL) ? p.memetic_gd_max_iters : p.gd_max_iters;
    d.stop_on_valid = p.stop_optimization_on_valid_solution ? 1 : 0;
    d.approx = p.return_approximate_solution ? 1 : 0;
    d.P = p.memetic_population_size;
    d.E = p.memetic_elite_size;
    d.max_generations = p.memetic_max_generations;
    d.debug = std::getenv("PIK_DEBUG_PHASES") ? 1 : 0;
    d.lockstep = std::getenv("PIK_NO_LOCKSTEP") ? 0 : (std::getenv("PIK_LOCKSTEP_MASK") ? std::atoi(std::getenv("PIK_LOCKSTEP_MASK")) : 1);  // (bit 1, the reproduce rounds: +1 % with one 16-warp CTA per SM)
    return d;
}

// Philox4x32-10 key schedule of rng_seed (travels in the kernel parameter block, SolveBuffers::round_key)
inline void make_round_keys(uint64_t rng_seed, uint32_t* round_key) {
    const uint32_t lo = static_cast<uint32_t>(rng_seed), hi = static_cast<uint32_t>(rng_seed >> 32);
    for (int r = 0; r < 10; ++r) {
        round_key[2 * r] = lo + 0x9E3779B9u * static_cast<uint32_t>(r);
        round_key[2 * r + 1] = hi + 0xBB67AE85u * static_cast<uint32_t>(r);
    }
}

}  // namespace pik
