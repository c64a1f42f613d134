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

// Fixed joints are folded into the constant origin of the next moving joint (or into the tip
// transform); each moving joint becomes one chain step.
inline int build_dev_robot(const pik_joint_desc* joints, int n_joints, DevRobot* out) {
    std::memset(out, 0, sizeof(*out));
    if (!joints || n_joints <= 0) return PIK_E_INVALID_ROBOT;
    double accR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, acct[3] = {0, 0, 0};
    bool have_acc = false;
    int n = 0;
    double max_velocity_rcp[kMaxVars];
    for (int j = 0; j < n_joints; ++j) {
        const pik_joint_desc& jd = joints[j];
        if (!have_acc) {
            std::memcpy(accR, jd.origin_R, sizeof(accR));
            std::memcpy(acct, jd.origin_t, sizeof(acct));
            have_acc = true;
        } else {
            double nR[9], nt[3];
            host_mat_vec_add(accR, jd.origin_t, acct, nt);
            host_mat_mul(accR, jd.origin_R, nR);
            std::memcpy(accR, nR, sizeof(nR));
            std::memcpy(acct, nt, sizeof(nt));
        }
        if (jd.type == PIK_JOINT_FIXED) continue;
        if (jd.type != PIK_JOINT_REVOLUTE && jd.type != PIK_JOINT_PRISMATIC) return PIK_E_INVALID_ROBOT;
        if (n >= kMaxVars) return PIK_E_INVALID_ROBOT;
        std::memcpy(out->R[n], accR, sizeof(accR));
        std::memcpy(out->t[n], acct, sizeof(acct));
        const double x = jd.axis[0], y = jd.axis[1], z = jd.axis[2];
        out->axis[n][0] = x; out->axis[n][1] = y; out->axis[n][2] = z;
        out->axis_sq[n][0] = x * x; out->axis_sq[n][1] = y * y; out->axis_sq[n][2] = z * z;
        out->axis_sq[n][3] = x * y; out->axis_sq[n][4] = x * z; out->axis_sq[n][5] = y * z;
        out->sign[n] = 1.0;
        if (jd.type == PIK_JOINT_PRISMATIC) {
            out->kind[n] = kPrismatic;
            if (std::fabs(x) == 1.0 && y == 0.0 && z == 0.0) { out->kind[n] = kPrisX; out->sign[n] = x; }
            if (x == 0.0 && std::fabs(y) == 1.0 && z == 0.0) { out->kind[n] = kPrisY; out->sign[n] = y; }
            if (x == 0.0 && y == 0.0 && std::fabs(z) == 1.0) { out->kind[n] = kPrisZ; out->sign[n] = z; }
        } else {
            out->kind[n] = kRevGeneral;
            if (std::fabs(x) == 1.0 && y == 0.0 && z == 0.0) { out->kind[n] = kRevX; out->sign[n] = x; }
            if (x == 0.0 && std::fabs(y) == 1.0 && z == 0.0) { out->kind[n] = kRevY; out->sign[n] = y; }
            if (x == 0.0 && y == 0.0 && std::fabs(z) == 1.0) { out->kind[n] = kRevZ; out->sign[n] = z; }
        }
        // Robot::from, robot.cpp:52-72
        out->bounded[n] = jd.bounded ? 1 : 0;
        out->vmin[n] = jd.min_position;
        out->vmax[n] = jd.max_position;
        out->vmid[n] = 0.5 * (jd.min_position + jd.max_position);
        out->vhalf[n] = jd.bounded ? (jd.max_position - jd.min_position) / 2.0 : M_PI;
        max_velocity_rcp[n] = jd.max_velocity > 0.0 ? 1.0 / jd.max_velocity : 0.0;
        ++n;
        have_acc = false;
    }
    if (n == 0) return PIK_E_INVALID_ROBOT;
    out->n = n;
    out->has_tip = have_acc ? 1 : 0;
    for (int i = 0; i < n; ++i)
        if (!out->bounded[i]) out->any_unbounded = 1;
    if (have_acc) {
        std::memcpy(out->tip_R, accR, sizeof(accR));
        std::memcpy(out->tip_t, acct, sizeof(acct));
    } else {
        const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        std::memcpy(accR, I, sizeof(I));
        acct[0] = acct[1] = acct[2] = 0.0;
    }
    // the tip transform as entry n of the origin table (identity when the chain ends in a moving joint)
    std::memcpy(out->R[n], accR, sizeof(accR));
    std::memcpy(out->t[n], acct, sizeof(acct));
    for (int i = 0; i <= n; ++i) out->ocls[i] = classify_origin(out->R[i]);
    // robot.cpp:69-82
    double divisor = 0.0;
    for (int i = 0; i < n; ++i) {
        out->vfac[i] = 1.0 / static_cast<double>(n);
        divisor += max_velocity_rcp[i];
    }
    if (divisor > 0.0)
        for (int i = 0; i < n; ++i) out->vfac[i] = max_velocity_rcp[i] / divisor;
    return PIK_OK;
}

inline void host_max_velocity_rcp(const pik_joint_desc* joints, int n_joints, double* out) {
    int n = 0;
    for (int j = 0; j < n_joints && n < kMaxVars; ++j) {
        if (joints[j].type == PIK_JOINT_FIXED) continue;
        out[n++] = joints[j].max_velocity > 0.0 ? 1.0 / joints[j].max_velocity : 0.0;
    }
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
    d.lockstep = std::getenv("PIK_NO_LOCKSTEP") ? 0 : (std::getenv("PIK_LOCKSTEP_MASK") ? std::atoi(std::getenv("PIK_LOCKSTEP_MASK")) : 3);
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
