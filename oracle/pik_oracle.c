/*
 * pik_oracle.c -- CPU ORACLE for the pick_ik hot path.  TEST INFRASTRUCTURE ONLY.
 * See pik_oracle.h for scope, pinning status and the arithmetic contract.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -mfma -fPIC -shared pik_oracle.c -o _build/liboracle.so -lm -lpthread
 * (-ffp-contract=off: no fused operations other than the fma() calls spelled out below.)
 *
 * Every function cites the reference lines it restates (pick_ik @ 8c99999).
 */
#include "pik_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------
 * Deterministic elementary functions.  The reference calls libm sin/cos (inside moveit_core's
 * RevoluteJointModel::computeTransform) and atan2 (Eigen angularDistance).  libm and CUDA's libm
 * differ by ulps, which would make flags and GD trajectories diverge, so both the oracle and the
 * kernels use these: Cody-Waite reduction + fdlibm-style minimax kernels, written only with
 * + - * / fma.  |error| vs the exact function is ~1 ulp for |x| < 1e5.
 * ------------------------------------------------------------------------------------------ */
static const double INV_PIO2 = 0x1.45f306dc9c883p-1;
static const double PIO2_1 = 0x1.921fb54442d18p+0;
static const double PIO2_2 = 0x1.1a62633145c07p-54;
static const double PIO2_3 = -0x1.f1976b7ed8fbcp-110;
static const double S1 = -0x1.5555555555549p-3, S2 = 0x1.111111110f8a6p-7, S3 = -0x1.a01a019c161d5p-13,
                    S4 = 0x1.71de357b1fe7dp-19, S5 = -0x1.ae5e68a2b9cebp-26, S6 = 0x1.5d93a5acfd57cp-33;
static const double C1 = 0x1.555555555554cp-5, C2 = -0x1.6c16c16c15177p-10, C3 = 0x1.a01a019cb1590p-16,
                    C4 = -0x1.27e4f809c52adp-22, C5 = 0x1.1ee9ebdb4b1c4p-29, C6 = -0x1.8fae9be8838d4p-37;

static double orc_nan(void) {
    union { uint64_t u; double d; } v;
    v.u = 0x7ff8000000000000ull;
    return v.d;
}

void orc_sincos(double x, double* s, double* c) {
    if (!(fabs(x) < 1.0e15)) { /* also catches NaN / inf */
        *s = orc_nan();
        *c = orc_nan();
        return;
    }
    double k = rint(x * INV_PIO2);
    double r = fma(-k, PIO2_1, x);
    r = fma(-k, PIO2_2, r);
    r = fma(-k, PIO2_3, r);
    long long q = (long long)k;
    double z = r * r;
    double ps = fma(z, S6, S5);
    ps = fma(z, ps, S4);
    ps = fma(z, ps, S3);
    ps = fma(z, ps, S2);
    ps = fma(z, ps, S1);
    double sr = fma(r * z, ps, r);
    double pc = fma(z, C6, C5);
    pc = fma(z, pc, C4);
    pc = fma(z, pc, C3);
    pc = fma(z, pc, C2);
    pc = fma(z, pc, C1);
    double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
    switch ((int)(q & 3)) {
        case 0: *s = sr; *c = cr; break;
        case 1: *s = cr; *c = -sr; break;
        case 2: *s = -sr; *c = -cr; break;
        default: *s = -cr; *c = sr; break;
    }
}

static const double AT0 = 0x1.555555555550dp-2, AT1 = -0x1.999999998ebc4p-3, AT2 = 0x1.24924920083ffp-3,
                    AT3 = -0x1.c71c6fe231671p-4, AT4 = 0x1.745cdc54c206ep-4, AT5 = -0x1.3b0f2af749a6dp-4,
                    AT6 = 0x1.10d66a0d03d51p-4, AT7 = -0x1.dde2d52defd9ap-5, AT8 = 0x1.97b4b24760debp-5,
                    AT9 = -0x1.2b4442c6a6c2fp-5, AT10 = 0x1.0ad3ae322da11p-6;
static const double TAN_PIO8 = 0x1.a827999fcef34p-2;
static const double PIO4_HI = 0x1.921fb54442d18p-1, PIO4_LO = 0x1.1a62633145c07p-55;
static const double PIO2_HI = 0x1.921fb54442d18p+0, PIO2_LO = 0x1.1a62633145c07p-54;
static const double PI_HI = 0x1.921fb54442d18p+1, PI_LO = 0x1.1a62633145c07p-53;

/* atan of a in [0, 1] */
static double atan_unit(double a) {
    double t = a, hi = 0.0, lo = 0.0;
    if (a > TAN_PIO8) {
        t = (a - 1.0) / (a + 1.0);
        hi = PIO4_HI;
        lo = PIO4_LO;
    }
    double z = t * t;
    double w = z * z;
    /* fdlibm split: even and odd coefficient chains */
    double s1 = fma(w, AT10, AT8);
    s1 = fma(w, s1, AT6);
    s1 = fma(w, s1, AT4);
    s1 = fma(w, s1, AT2);
    s1 = fma(w, s1, AT0);
    s1 = z * s1;
    double s2 = fma(w, AT9, AT7);
    s2 = fma(w, s2, AT5);
    s2 = fma(w, s2, AT3);
    s2 = fma(w, s2, AT1);
    s2 = w * s2;
    double r = fma(-t, s1 + s2, t); /* t - t*(s1+s2) */
    return hi + (r + lo);
}

double orc_atan2(double y, double x) {
    if (x != x || y != y) return orc_nan();
    double ax = fabs(x), ay = fabs(y);
    double mx = ax > ay ? ax : ay;
    double mn = ax > ay ? ay : ax;
    double r;
    if (mx == 0.0) {
        r = 0.0;
    } else {
        double a = mn / mx; /* inf/inf -> NaN, documented */
        r = atan_unit(a);
        if (ay > ax) r = PIO2_HI - (r - PIO2_LO);
    }
    if (x < 0.0) r = PI_HI - (r - PI_LO);
    return (y < 0.0) ? -r : r;
}

/* ------------------------------------------------------------------------------------------
 * RNG.  The reference uses rsl::uniform_real / uniform_int over a thread-local std::mt19937 seeded
 * from random_device (src/robot.cpp:25-28, src/ik_memetic.cpp:131-159): unseeded, PARITY-UNPINNED.
 * We define Philox4x32-10 word streams: key = rng_seed, counter = (block, individual,
 * purpose << 28 | epoch, problem).  uniform_real = 53-bit (as generate_canonical<double,53> over two
 * 32-bit draws), uniform_int = Lemire multiply-shift with rejection (as libstdc++ >= 11 for 32-bit
 * URBGs).
 * ------------------------------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { STREAM_INIT = 1, STREAM_REPRODUCE = 2, STREAM_TARGET = 3, STREAM_RANDOM_CHILD = 4 };

/* A stream is identified by (seed, problem, purpose, epoch, individual); its words are addressed
 * by block: counter = (block, individual, purpose << 28 | epoch, problem), key = seed.  Every
 * consumer below reads FIXED (block, word) positions, so a draw never depends on how many words an
 * earlier draw consumed -- which is what lets the CUDA kernels produce children in any order. */
typedef struct {
    uint32_t key[2];
    uint32_t c1, c2, c3;
} orc_stream;

static void stream_init(orc_stream* s, uint64_t seed, uint32_t problem, uint32_t purpose, uint32_t epoch,
                        uint32_t individual) {
    s->key[0] = (uint32_t)seed;
    s->key[1] = (uint32_t)(seed >> 32);
    s->c1 = individual;
    s->c2 = (purpose << 28) | (epoch & 0x0fffffffu);
    s->c3 = problem;
}

static void stream_block(const orc_stream* s, uint32_t block, uint32_t out[4]) {
    uint32_t ctr[4] = {block, s->c1, s->c2, s->c3};
    orc_philox4x32_10(ctr, s->key, out);
}

/* u in [0,1), 53 bits, as generate_canonical<double,53> over two 32-bit draws: first word = low half */
static double unit_from_words(uint32_t lo, uint32_t hi) {
    return (double)(((((uint64_t)hi) << 32) | (uint64_t)lo) >> 11) * 0x1.0p-53;
}

/* rsl::uniform_real(a, b): a + (b - a) * u */
static double uniform_real_words(double a, double b, uint32_t lo, uint32_t hi) {
    return a + (b - a) * unit_from_words(lo, hi);
}

/* Word list feeding the parent-index draws of one child (rsl::uniform_int, ik_memetic.cpp:131-135):
 * block0.w0..w3, block1.w2, block1.w3, then the words of blocks first_overflow, first_overflow+1, ... */
typedef struct {
    const orc_stream* s;
    uint32_t head[6];
    uint32_t ovf[4];
    uint32_t ovf_block;
    int pos;
} orc_index_words;

static uint32_t index_words_next(orc_index_words* w) {
    uint32_t v;
    if (w->pos < 6) {
        v = w->head[w->pos];
    } else {
        int k = (w->pos - 6) & 3;
        if (k == 0) {
            stream_block(w->s, w->ovf_block, w->ovf);
            w->ovf_block += 1;
        }
        v = w->ovf[k];
    }
    w->pos += 1;
    return v;
}

/* rsl::uniform_int<size_t>(0, m - 1): Lemire multiply-shift with rejection (libstdc++ >= 11) */
static uint32_t uniform_int_words(orc_index_words* w, uint32_t m) {
    uint64_t prod = (uint64_t)index_words_next(w) * m;
    uint32_t low = (uint32_t)prod;
    if (low < m) {
        uint32_t thr = (0u - m) % m;
        while (low < thr) {
            prod = (uint64_t)index_words_next(w) * m;
            low = (uint32_t)prod;
        }
    }
    return (uint32_t)(prod >> 32);
}

/* ------------------------------------------------------------------------------------------
 * Frames (Eigen arithmetic restated)
 * ------------------------------------------------------------------------------------------ */
void orc_quat_to_matrix(const double q[4], double R[9]) {
    /* Eigen QuaternionBase::toRotationMatrix; q = (w,x,y,z); no normalisation (tf2::fromMsg) */
    double w = q[0], x = q[1], y = q[2], z = q[3];
    double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    double twx = tx * w, twy = ty * w, twz = tz * w;
    double txx = tx * x, txy = ty * x, txz = tz * x;
    double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0 - (tyy + tzz);
    R[1] = txy - twz;
    R[2] = txz + twy;
    R[3] = txy + twz;
    R[4] = 1.0 - (txx + tzz);
    R[5] = tyz - twx;
    R[6] = txz - twy;
    R[7] = tyz + twx;
    R[8] = 1.0 - (txx + tyy);
}

void orc_matrix_to_quat(const double R[9], double q[4]) {
    /* Eigen quaternionbase_assign_impl<Matrix3d> */
    double t = (R[0] + R[4]) + R[8];
    if (t > 0.0) {
        t = sqrt(t + 1.0);
        q[0] = 0.5 * t;
        t = 0.5 / t;
        q[1] = (R[7] - R[5]) * t;
        q[2] = (R[2] - R[6]) * t;
        q[3] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        int j = (i + 1) % 3;
        int k = (j + 1) % 3;
        t = sqrt(((R[4 * i] - R[4 * j]) - R[4 * k]) + 1.0);
        q[1 + i] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R[3 * k + j] - R[3 * j + k]) * t;
        q[1 + j] = (R[3 * j + i] + R[3 * i + j]) * t;
        q[1 + k] = (R[3 * k + i] + R[3 * i + k]) * t;
    }
}

/* goal.cpp:17-19 */
double orc_linear_distance(const double t1[3], const double t2[3]) {
    double dx = t1[0] - t2[0], dy = t1[1] - t2[1], dz = t1[2] - t2[2];
    return sqrt((dx * dx + dy * dy) + dz * dz);
}

/* goal.cpp:21-25: q_2.angularDistance(q_1), q_1 = goal, q_2 = tip.
 * Eigen >= 3.3: d = q_2 * conj(q_1); 2 * atan2(|d.vec|, |d.w|).  Products written out in
 * Eigen's quaternion-product order with conj(q_1) = (w, -x, -y, -z) substituted. */
double orc_angular_distance_q(const double g[4], const double R_tip[9]) {
    double a[4];
    orc_matrix_to_quat(R_tip, a);
    double aw = a[0], ax = a[1], ay = a[2], az = a[3];
    double bw = g[0], bx = -g[1], by = -g[2], bz = -g[3];
    double dw = ((aw * bw - ax * bx) - ay * by) - az * bz;
    double dx = ((aw * bx + ax * bw) + ay * bz) - az * by;
    double dy = ((aw * by + ay * bw) + az * bx) - ax * bz;
    double dz = ((aw * bz + az * bw) + ax * by) - ay * bx;
    double vn = sqrt((dx * dx + dy * dy) + dz * dz);
    return 2.0 * orc_atan2(vn, fabs(dw));
}

double orc_angular_distance(const double R_goal[9], const double R_tip[9]) {
    double g[4];
    orc_matrix_to_quat(R_goal, g);
    return orc_angular_distance_q(g, R_tip);
}

/* goal.cpp:27-36.  threshold < 0 == std::nullopt */
int orc_frame_test(const double goal_t[3], const double goal_R[9], const double tip_t[3],
                   const double tip_R[9], double position_threshold, double orientation_threshold) {
    if (position_threshold >= 0.0 && !(orc_linear_distance(goal_t, tip_t) <= position_threshold))
        return 0;
    if (orientation_threshold >= 0.0 &&
        !(fabs(orc_angular_distance(goal_R, tip_R)) <= orientation_threshold))
        return 0;
    return 1;
}

static double pose_cost_q(const double goal_t[3], const double goal_q[4], const double tip_t[3],
                          const double tip_R[9], double ps, double rs) {
    /* goal.cpp:51-78; std::pow(x, 2) == x * x */
    double cost = 0.0;
    if (ps > 0.0) {
        double d = orc_linear_distance(goal_t, tip_t) * ps;
        if (rs > 0.0) {
            double a = orc_angular_distance_q(goal_q, tip_R) * rs;
            cost = d * d + a * a;
        } else {
            cost = d * d;
        }
    } else if (rs > 0.0) {
        double a = orc_angular_distance_q(goal_q, tip_R) * rs;
        cost = a * a;
    }
    return cost;
}

double orc_pose_cost(const double goal_t[3], const double goal_R[9], const double tip_t[3],
                     const double tip_R[9], double position_scale, double rotation_scale) {
    double g[4];
    orc_matrix_to_quat(goal_R, g);
    return pose_cost_q(goal_t, g, tip_t, tip_R, position_scale, rotation_scale);
}

/* ------------------------------------------------------------------------------------------
 * Robot table (src/robot.cpp:44-85) and chain description
 * ------------------------------------------------------------------------------------------ */
static void mat_mul(const double A[9], const double B[9], double C[9]) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            C[3 * r + c] = fma(A[3 * r + 2], B[6 + c], fma(A[3 * r + 1], B[3 + c], A[3 * r] * B[c]));
}
static void mat_vec_add(const double A[9], const double v[3], const double t[3], double out[3]) {
    for (int r = 0; r < 3; ++r)
        out[r] = fma(A[3 * r + 2], v[2], fma(A[3 * r + 1], v[1], fma(A[3 * r], v[0], t[r])));
}

static void fill_variable(orc_variable* v, int bounded, double lo, double hi, double max_velocity) {
    /* Robot::from, robot.cpp:52-72 */
    v->bounded = bounded;
    v->min = lo;
    v->max = hi;
    v->mid = 0.5 * (v->min + v->max);
    v->half_span = v->bounded ? (v->max - v->min) / 2.0 : M_PI;
    v->max_velocity_rcp = max_velocity > 0.0 ? 1.0 / max_velocity : 0.0;
}

/* acc <- product, root to leaf, of the origins of the fixed joints between joint j and its nearest moving
 * ancestor (exclusive), then the origin of j itself.  Returns that ancestor (-1: the model root). */
static int fold_origins(const orc_joint_desc* joints, const int32_t* parent, int j, double accR[9], double acct[3]) {
    int path[64], np = 0;
    path[np++] = j;
    int k = parent ? parent[j] : j - 1;
    while (k >= 0 && joints[k].type == ORC_JOINT_FIXED && np < 64) {
        path[np++] = k;
        k = parent ? parent[k] : k - 1;
    }
    for (int a = np - 1; a >= 0; --a) {
        const orc_joint_desc* jd = &joints[path[a]];
        if (a == np - 1) {
            memcpy(accR, jd->origin_R, 9 * sizeof(double));
            memcpy(acct, jd->origin_t, 3 * sizeof(double));
        } else {
            double nR[9], nt[3];
            mat_vec_add(accR, jd->origin_t, acct, nt);
            mat_mul(accR, jd->origin_R, nR);
            memcpy(accR, nR, sizeof(nR));
            memcpy(acct, nt, sizeof(nt));
        }
    }
    return k;
}

int orc_robot_build_tree(const orc_joint_desc* joints, int n_joints, const int32_t* parent,
                         const int32_t* tip_joint, int n_tips, const int32_t* mimic_of,
                         const double* mimic_factor, const double* mimic_offset, orc_robot* out) {
    memset(out, 0, sizeof(*out));
    if (n_joints <= 0 || n_joints > 64 || n_tips < 1 || n_tips > ORC_MAX_TIPS) return -1;
    int step_of[64], n = 0, ns = 0;
    double max_velocity[ORC_MAX_VARS];
    for (int j = 0; j < n_joints; ++j) {
        const orc_joint_desc* jd = &joints[j];
        step_of[j] = -1;
        if (parent && !(parent[j] >= -1 && parent[j] < j)) return -4; /* parents precede children */
        if (jd->type == ORC_JOINT_FIXED) continue;
        if (ns >= ORC_MAX_STEPS) return -1;
        orc_step* st = &out->steps[ns];
        const int up = fold_origins(joints, parent, j, st->R, st->t);
        st->parent = up >= 0 ? step_of[up] : -1;
        /* unit axis: moveit_core RevoluteJointModel::setAxis / PrismaticJointModel::setAxis normalise */
        double x = jd->axis[0], y = jd->axis[1], z = jd->axis[2];
        const double a2 = x * x + y * y + z * z;
        if (fabs(a2 - 1.0) > 1.0e-14 && a2 > 0.0) {  /* an axis that is unit to rounding is kept as given */
            const double nrm = sqrt(a2);
            x /= nrm; y /= nrm; z /= nrm;
        }
        st->axis[0] = x; st->axis[1] = y; st->axis[2] = z;
        st->axis_sq[0] = x * x; st->axis_sq[1] = y * y; st->axis_sq[2] = z * z;
        st->axis_sq[3] = x * y; st->axis_sq[4] = x * z; st->axis_sq[5] = y * z;
        st->sign = 1.0;
        st->mimic_factor = 1.0;
        st->mimic_offset = 0.0;
        int n_vars = 1;
        if (jd->type == ORC_JOINT_PRISMATIC) {
            st->kind = ORC_STEP_PRISMATIC;
        } else if (jd->type == ORC_JOINT_REVOLUTE) {
            st->kind = ORC_STEP_REV_GENERAL;
            if (fabs(x) == 1.0 && y == 0.0 && z == 0.0) { st->kind = ORC_STEP_REV_X; st->sign = x; }
            if (x == 0.0 && fabs(y) == 1.0 && z == 0.0) { st->kind = ORC_STEP_REV_Y; st->sign = y; }
            if (x == 0.0 && y == 0.0 && fabs(z) == 1.0) { st->kind = ORC_STEP_REV_Z; st->sign = z; }
        } else if (jd->type == ORC_JOINT_FLOATING) {
            st->kind = ORC_STEP_FLOATING;
            n_vars = 7;
        } else if (jd->type == ORC_JOINT_PLANAR) {
            st->kind = ORC_STEP_PLANAR;
            n_vars = 3;
        } else {
            return -2;
        }
        const int m = mimic_of ? mimic_of[j] : -1;
        if (m >= 0) {
            /* a mimic joint follows the variable of its master and owns none (robot.cpp:145-147) */
            if (n_vars != 1 || m >= j || step_of[m] < 0 || joints[m].type > ORC_JOINT_PRISMATIC) return -5;
            st->var0 = out->steps[step_of[m]].var0;
            st->mimic_factor = mimic_factor[j] * out->steps[step_of[m]].mimic_factor;
            st->mimic_offset = mimic_factor[j] * out->steps[step_of[m]].mimic_offset + mimic_offset[j];
        } else {
            if (n + n_vars > ORC_MAX_VARS) return -1;
            st->var0 = n;
            /* variable bounds as MoveIt's joint models set them: the translation variables of floating / planar
             * joints take the description's bounds (unbounded by default), quaternion components [-1, 1], the
             * planar angle is unbounded with the nominal range -pi .. pi */
            for (int k = 0; k < n_vars; ++k) {
                int bounded = jd->bounded;
                double lo = jd->min_position, hi = jd->max_position;
                if (jd->type == ORC_JOINT_FLOATING && k >= 3) { bounded = 1; lo = -1.0; hi = 1.0; }
                if (jd->type == ORC_JOINT_PLANAR && k == 2) { bounded = 0; lo = -M_PI; hi = M_PI; }
                fill_variable(&out->vars[n + k], bounded, lo, hi, jd->max_velocity);
                max_velocity[n + k] = jd->max_velocity;
            }
            n += n_vars;
        }
        step_of[j] = ns++;
    }
    if (n == 0) return -3;
    out->n = n;
    out->n_steps = ns;
    out->n_tips = n_tips;
    for (int t = 0; t < n_tips; ++t) {
        const int j = tip_joint[t];
        if (j < 0 || j >= n_joints) return -6;
        if (joints[j].type != ORC_JOINT_FIXED) {
            out->tip_step[t] = step_of[j];
            out->tip_has[t] = 0;
        } else {
            const int up = fold_origins(joints, parent, j, out->tip_R[t], out->tip_t[t]);
            out->tip_step[t] = up >= 0 ? step_of[up] : -1;
            out->tip_has[t] = 1;
        }
    }
    out->has_tip = out->tip_has[0];
    /* robot.cpp:69-82 */
    double divisor = 0.0;
    for (int i = 0; i < n; ++i) {
        out->vars[i].minimal_displacement_factor = 1.0 / (double)n;
        divisor += out->vars[i].max_velocity_rcp;
    }
    if (divisor > 0.0)
        for (int i = 0; i < n; ++i)
            out->vars[i].minimal_displacement_factor = out->vars[i].max_velocity_rcp / divisor;
    (void)max_velocity;
    return 0;
}

int orc_robot_build(const orc_joint_desc* joints, int n_joints, orc_robot* out) {
    const int32_t tip = n_joints - 1;
    return orc_robot_build_tree(joints, n_joints, NULL, &tip, 1, NULL, NULL, NULL, out);
}

/* Rotate columns (a, b) of R: the product R * Rot_axis(angle) for an axis-aligned joint.
 * col_a' = col_a * c + col_b * s ; col_b' = col_b * c - col_a * s */
static void rotate_cols(double R[9], int a, int b, double s, double c) {
    for (int r = 0; r < 3; ++r) {
        double va = R[3 * r + a], vb = R[3 * r + b];
        R[3 * r + a] = fma(vb, s, va * c);
        R[3 * r + b] = fma(vb, c, -(va * s));
    }
}

/* One moving joint applied to frame (R, t) that already includes the folded origin; qv = the variables of the
 * joint.  Revolute: RevoluteJointModel::computeTransform (same rotation as
 * forward_kinematics.cpp:48-57); prismatic: forward_kinematics.cpp:58-63; floating:
 * Translation3d(v0 v1 v2) * Quaterniond(w = v6, x = v3, y = v4, z = v5), quaternion used as given
 * (forward_kinematics.cpp:64-70); planar: Translation3d(x, y, 0) * rotation about z by theta (what
 * PlanarJointModel::computeTransform returns, forward_kinematics.cpp:71-79). */
static void apply_joint(const orc_step* st, const double* qv, double R[9], double t[3]) {
    if (st->kind == ORC_STEP_FLOATING || st->kind == ORC_STEP_PLANAR) {
        double J[9], d[3], nR[9], nt[3];
        if (st->kind == ORC_STEP_FLOATING) {
            const double quat[4] = {qv[6], qv[3], qv[4], qv[5]};
            orc_quat_to_matrix(quat, J);
            d[0] = qv[0]; d[1] = qv[1]; d[2] = qv[2];
        } else {
            double s, c;
            orc_sincos(qv[2], &s, &c);
            J[0] = c; J[1] = -s; J[2] = 0.0;
            J[3] = s; J[4] = c; J[5] = 0.0;
            J[6] = 0.0; J[7] = 0.0; J[8] = 1.0;
            d[0] = qv[0]; d[1] = qv[1]; d[2] = 0.0;
        }
        mat_vec_add(R, d, t, nt);
        mat_mul(R, J, nR);
        memcpy(R, nR, sizeof(nR));
        memcpy(t, nt, sizeof(nt));
        return;
    }
    /* (a joint that is not a mimic has factor 1 and offset 0: the value is the variable itself) */
    const double q = (st->mimic_factor == 1.0 && st->mimic_offset == 0.0) ? qv[0] : qv[0] * st->mimic_factor + st->mimic_offset;
    if (st->kind == ORC_STEP_PRISMATIC) {
        double d[3] = {st->axis[0] * q, st->axis[1] * q, st->axis[2] * q};
        double nt[3];
        mat_vec_add(R, d, t, nt);
        t[0] = nt[0]; t[1] = nt[1]; t[2] = nt[2];
        return;
    }
    double s, c;
    orc_sincos(q, &s, &c);
    switch (st->kind) {
        case ORC_STEP_REV_X: rotate_cols(R, 1, 2, st->sign * s, c); break;
        case ORC_STEP_REV_Y: rotate_cols(R, 2, 0, st->sign * s, c); break;
        case ORC_STEP_REV_Z: rotate_cols(R, 0, 1, st->sign * s, c); break;
        default: {
            double x = st->axis[0], y = st->axis[1], z = st->axis[2];
            const double* a2 = st->axis_sq;
            double t1 = 1.0 - c;
            double J[9], N[9];
            J[0] = fma(t1, a2[0], c);
            J[1] = fma(t1, a2[3], -(z * s));
            J[2] = fma(t1, a2[4], y * s);
            J[3] = fma(t1, a2[3], z * s);
            J[4] = fma(t1, a2[1], c);
            J[5] = fma(t1, a2[5], -(x * s));
            J[6] = fma(t1, a2[4], -(y * s));
            J[7] = fma(t1, a2[5], x * s);
            J[8] = fma(t1, a2[2], c);
            mat_mul(R, J, N);
            memcpy(R, N, sizeof(N));
        }
    }
}

/* fk_moveit.cpp:20-34 -> the tip frames.  Every link frame = parent link frame * folded origin * joint motion,
 * root to leaf (RobotState::updateLinkTransforms). */
void orc_fk_tips(const orc_robot* robot, const double* q, double* R_out, double* t_out) {
    double FR[ORC_MAX_STEPS][9], FT[ORC_MAX_STEPS][3];
    for (int k = 0; k < robot->n_steps; ++k) {
        const orc_step* st = &robot->steps[k];
        if (st->parent < 0) {
            memcpy(FR[k], st->R, 9 * sizeof(double));
            memcpy(FT[k], st->t, 3 * sizeof(double));
        } else {
            mat_vec_add(FR[st->parent], st->t, FT[st->parent], FT[k]);
            mat_mul(FR[st->parent], st->R, FR[k]);
        }
        apply_joint(st, q + st->var0, FR[k], FT[k]);
    }
    for (int i = 0; i < robot->n_tips; ++i) {
        double* R = R_out + 9 * i;
        double* t = t_out + 3 * i;
        const int k = robot->tip_step[i];
        if (k < 0) { /* a tip the group does not move: the constant transform from the model root */
            static const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            memcpy(R, robot->tip_has[i] ? robot->tip_R[i] : I, 9 * sizeof(double));
            t[0] = t[1] = t[2] = 0.0;
            if (robot->tip_has[i]) memcpy(t, robot->tip_t[i], 3 * sizeof(double));
        } else if (robot->tip_has[i]) {
            mat_vec_add(FR[k], robot->tip_t[i], FT[k], t);
            mat_mul(FR[k], robot->tip_R[i], R);
        } else {
            memcpy(R, FR[k], 9 * sizeof(double));
            memcpy(t, FT[k], 3 * sizeof(double));
        }
    }
}

void orc_fk(const orc_robot* robot, const double* q, double R[9], double t[3]) {
    double Rs[ORC_MAX_TIPS][9], ts[ORC_MAX_TIPS][3];
    orc_fk_tips(robot, q, &Rs[0][0], &ts[0][0]);
    memcpy(R, Rs[0], 9 * sizeof(double));
    memcpy(t, ts[0], 3 * sizeof(double));
}

/* robot.cpp:36-42 (std::clamp(v, lo, hi) = v < lo ? lo : hi < v ? hi : v) */
double orc_clamp_to_limits(const orc_variable* v, double val) {
    double lo = v->bounded ? v->min : val - v->half_span;
    double hi = v->bounded ? v->max : val + v->half_span;
    return (val < lo) ? lo : ((hi < val) ? hi : val);
}

/* robot.cpp:32-34, 97-105 */
int orc_is_valid_configuration(const orc_robot* robot, const double* q) {
    for (int i = 0; i < robot->n; ++i) {
        const orc_variable* v = &robot->vars[i];
        if (!(!v->bounded || (q[i] <= v->max && q[i] >= v->min))) return 0;
    }
    return 1;
}

/* robot.cpp:23-30, 87-95.  Variable i draws its uniform from block i >> 1, word pair i & 1. */
static void set_random_valid_configuration(const orc_robot* robot, const orc_stream* st, double* config) {
    uint32_t w[4] = {0, 0, 0, 0};
    for (int i = 0; i < robot->n; ++i) {
        const orc_variable* v = &robot->vars[i];
        if ((i & 1) == 0) stream_block(st, (uint32_t)(i >> 1), w);
        uint32_t lo = w[2 * (i & 1)], hi = w[2 * (i & 1) + 1];
        if (v->bounded)
            config[i] = uniform_real_words(v->min, v->max, lo, hi);
        else
            config[i] = uniform_real_words(config[i] - M_PI, config[i] + M_PI, lo, hi);
    }
}

/* ------------------------------------------------------------------------------------------
 * Goals, cost, solution test
 * ------------------------------------------------------------------------------------------ */
void orc_params_default(orc_params* p) {
    /* src/pick_ik_parameters.yaml defaults */
    memset(p, 0, sizeof(*p));
    p->mode = 0;
    p->gd_step_size = 0.0001;
    p->gd_max_iters = 100;
    p->gd_min_cost_delta = 1.0e-12;
    p->position_threshold = 0.001;
    p->orientation_threshold = 0.001;
    p->cost_threshold = 0.001;
    p->position_scale = 1.0;
    p->rotation_scale = 0.5;
    p->center_joints_weight = 0.0;
    p->avoid_joint_limits_weight = 0.0;
    p->minimal_displacement_weight = 0.0;
    p->stop_optimization_on_valid_solution = 1;
    p->memetic_population_size = 16;
    p->memetic_elite_size = 4;
    p->memetic_wipeout_fitness_tol = 0.00001;
    p->memetic_max_generations = 100;
    p->memetic_gd_max_iters = 25;
    p->return_approximate_solution = 0;
    p->memetic_num_threads = 1;
    p->memetic_stop_on_first_solution = 1;
    p->rng_seed = 0x5EED;
}

void orc_problem_init(orc_problem* pb, const orc_robot* robot, const orc_params* params,
                      const double* goal_pose, const double* seed) {
    memset(pb, 0, sizeof(*pb));
    pb->robot = robot;
    pb->params = params;
    for (int i = 0; i < robot->n_tips; ++i) {
        const double* gp = goal_pose + 7 * i;
        pb->goal_t[i][0] = gp[0];
        pb->goal_t[i][1] = gp[1];
        pb->goal_t[i][2] = gp[2];
        orc_quat_to_matrix(gp + 3, pb->goal_R[i]);      /* tf2::fromMsg, robot.cpp:175-176 */
        orc_matrix_to_quat(pb->goal_R[i], pb->goal_q[i]); /* goal.cpp:22 */
    }
    for (int i = 0; i < ORC_MAX_VARS; ++i) pb->seed[i] = (i < robot->n) ? seed[i] : 0.0;
}

/* goal.cpp:91-108 */
double orc_center_joints_cost(const orc_robot* robot, const double* q) {
    double sum = 0.0;
    for (int i = 0; i < robot->n; ++i) {
        const orc_variable* v = &robot->vars[i];
        if (!v->bounded) continue;
        double mid = (v->min + v->max) * 0.5;
        double e = (q[i] - mid) * v->minimal_displacement_factor;
        sum += e * e;
    }
    return sum;
}

/* goal.cpp:110-129 */
double orc_avoid_joint_limits_cost(const orc_robot* robot, const double* q) {
    double sum = 0.0;
    for (int i = 0; i < robot->n; ++i) {
        const orc_variable* v = &robot->vars[i];
        if (!v->bounded) continue;
        double x = fabs(q[i] - v->mid) * 2.0 - v->half_span;
        double m = (x > 0.0) ? x : 0.0; /* std::fmax(0.0, x): NaN -> 0 */
        double e = m * v->minimal_displacement_factor;
        sum += e * e;
    }
    return sum;
}

/* goal.cpp:131-144 */
double orc_minimal_displacement_cost(const orc_robot* robot, const double* q, const double* seed) {
    double sum = 0.0;
    for (int i = 0; i < robot->n; ++i) {
        double e = (q[i] - seed[i]) * robot->vars[i].minimal_displacement_factor;
        sum += e * e;
    }
    return sum;
}

/* goals in plugin order (pick_ik_plugin.cpp:118-129); returns count, fills weighted costs */
static int goal_costs(const orc_problem* pb, const double* q, double out[3]) {
    const orc_params* p = pb->params;
    int n = 0;
    if (p->center_joints_weight > 0.0)
        out[n++] = orc_center_joints_cost(pb->robot, q) * (p->center_joints_weight * p->center_joints_weight);
    if (p->avoid_joint_limits_weight > 0.0)
        out[n++] = orc_avoid_joint_limits_cost(pb->robot, q) *
                   (p->avoid_joint_limits_weight * p->avoid_joint_limits_weight);
    if (p->minimal_displacement_weight > 0.0)
        out[n++] = orc_minimal_displacement_cost(pb->robot, q, pb->seed) *
                   (p->minimal_displacement_weight * p->minimal_displacement_weight);
    return n;
}

/* goal.cpp:188-203.  fk_mutex (optional): the lock the reference's FK closure takes around the shared RobotState
 * (fk_moveit.cpp:21); only the reference-structure baseline passes one. */
static double cost_locked(const orc_problem* pb, const double* q, pthread_mutex_t* fk_mutex) {
    double R[ORC_MAX_TIPS][9], t[ORC_MAX_TIPS][3];
    if (fk_mutex) pthread_mutex_lock(fk_mutex);
    orc_fk_tips(pb->robot, q, &R[0][0], &t[0][0]);
    if (fk_mutex) pthread_mutex_unlock(fk_mutex);
    double pose_cost = 0.0; /* std::accumulate over the tips, goal.cpp:192-196 */
    for (int i = 0; i < pb->robot->n_tips; ++i)
        pose_cost = pose_cost + pose_cost_q(pb->goal_t[i], pb->goal_q[i], t[i], R[i], pb->params->position_scale,
                                            pb->params->rotation_scale);
    double g[3];
    int ng = goal_costs(pb, q, g);
    double goal_cost = 0.0;
    for (int i = 0; i < ng; ++i) goal_cost = goal_cost + g[i];
    return pose_cost + goal_cost;
}

double orc_cost(const orc_problem* pb, const double* q) { return cost_locked(pb, q, NULL); }

/* goal.cpp:163-186 with thresholds enabled as pick_ik_plugin.cpp:97-106 */
int orc_is_solution(const orc_problem* pb, const double* q) {
    const orc_params* p = pb->params;
    double R[ORC_MAX_TIPS][9], t[ORC_MAX_TIPS][3];
    orc_fk_tips(pb->robot, q, &R[0][0], &t[0][0]);
    for (int i = 0; i < pb->robot->n_tips; ++i) { /* every frame test, goal.cpp:169-175 */
        if (p->position_scale > 0.0 && !(orc_linear_distance(pb->goal_t[i], t[i]) <= p->position_threshold))
            return 0;
        if (p->rotation_scale > 0.0 &&
            !(fabs(orc_angular_distance_q(pb->goal_q[i], R[i])) <= p->orientation_threshold))
            return 0;
    }
    double thr_sq = p->cost_threshold * p->cost_threshold;
    double g[3];
    int ng = goal_costs(pb, q, g);
    for (int i = 0; i < ng; ++i)
        if (g[i] >= thr_sq) return 0;
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * Gradient descent (src/ik_gradient.cpp)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const orc_problem* pb;
    uint64_t evals;
    uint32_t gd_steps;
    pthread_mutex_t* fk_mutex; /* reference-structure baseline only */
} orc_ctx;

static double cost_counted(orc_ctx* cx, const double* q) {
    cx->evals++;
    return cost_locked(cx->pb, q, cx->fk_mutex);
}

/* ik_gradient.cpp:24-94 */
static int gd_step(orc_ctx* cx, double* gradient, double* working, double* local, double* best,
                   double* local_cost, double* best_cost) {
    const orc_robot* robot = cx->pb->robot;
    const double step_size = cx->pb->params->gd_step_size;
    const int n = robot->n;
    cx->gd_steps++;
    for (int i = 0; i < n; ++i) {
        working[i] = local[i] - step_size;
        double p1 = cost_counted(cx, working);
        working[i] = local[i] + step_size;
        double p3 = cost_counted(cx, working);
        working[i] = local[i];
        gradient[i] = p3 - p1;
    }
    double sum = step_size;
    for (int i = 0; i < n; ++i) sum = sum + fabs(gradient[i]);
    double f = 1.0 / sum * step_size;
    for (int i = 0; i < n; ++i) gradient[i] = gradient[i] * f;

    for (int i = 0; i < n; ++i) working[i] = local[i] - gradient[i];
    double p1 = cost_counted(cx, working);
    for (int i = 0; i < n; ++i) working[i] = local[i] + gradient[i];
    double p3 = cost_counted(cx, working);
    double p2 = (p1 + p3) * 0.5;
    double cost_diff = (p3 - p1) * 0.5;
    double joint_diff = p2 / cost_diff;
    if (!isfinite(joint_diff)) joint_diff = 0.0;

    for (int i = 0; i < n; ++i) {
        double updated = local[i] - gradient[i] * joint_diff;
        working[i] = orc_clamp_to_limits(&robot->vars[i], updated);
    }
    for (int i = 0; i < n; ++i) local[i] = working[i];
    *local_cost = cost_counted(cx, local);
    if (*local_cost < *best_cost) {
        for (int i = 0; i < n; ++i) best[i] = local[i];
        *best_cost = *local_cost;
        return 1;
    }
    return 0;
}

int orc_gd_step(const orc_problem* pb, double* gradient, double* working, double* local,
                double* best, double* local_cost, double* best_cost) {
    orc_ctx cx = {pb, 0, 0, NULL};
    return gd_step(&cx, gradient, working, local, best, local_cost, best_cost);
}

/* ik_gradient.cpp:96-139, max_time = inf */
void orc_ik_gradient(const orc_problem* pb, const double* initial_guess, orc_result* out) {
    const orc_params* p = pb->params;
    const int n = pb->robot->n;
    memset(out, 0, sizeof(*out));
    orc_ctx cx = {pb, 0, 0, NULL};
    if (p->stop_optimization_on_valid_solution && orc_is_solution(pb, initial_guess)) {
        out->found = 1;
        memcpy(out->solution, initial_guess, n * sizeof(double));
        out->cost = orc_cost(pb, initial_guess);
        return;
    }
    double gradient[ORC_MAX_VARS] = {0}, working[ORC_MAX_VARS], local[ORC_MAX_VARS], best[ORC_MAX_VARS];
    double local_cost = cost_counted(&cx, initial_guess), best_cost = local_cost;
    for (int i = 0; i < n; ++i) working[i] = local[i] = best[i] = initial_guess[i];

    int num_iterations = 0, found = 0;
    double previous_cost = 0.0;
    while (num_iterations < p->gd_max_iters) {
        if (gd_step(&cx, gradient, working, local, best, &local_cost, &best_cost)) {
            if (p->stop_optimization_on_valid_solution && orc_is_solution(pb, best)) {
                found = 1;
                break;
            }
        }
        if (fabs(local_cost - previous_cost) <= p->gd_min_cost_delta) break;
        previous_cost = local_cost;
        num_iterations++;
    }
    if (!found && !p->stop_optimization_on_valid_solution && orc_is_solution(pb, best)) found = 1;
    if (!found && p->return_approximate_solution) found = 1;
    out->found = found;
    out->iterations = num_iterations;
    out->cost = best_cost;
    out->evals = cx.evals;
    out->gd_steps = cx.gd_steps;
    memcpy(out->solution, best, n * sizeof(double));
}

/* ------------------------------------------------------------------------------------------
 * Memetic solver (src/ik_memetic.cpp), single species, max_time = inf
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    double genes[ORC_MAX_VARS];
    double gradient[ORC_MAX_VARS];
    double fitness;
    double extinction;
} orc_individual;

typedef struct {
    orc_ctx cx;
    int n, P, E;
    uint32_t problem_index;
    uint32_t species; /* replica index (MemeticIkParams::num_threads); high half of the individual word */
    orc_individual* pop;
    orc_individual* scratch;
    int* order;
    orc_individual best, best_curr;
    int has_previous;
    double previous_fitness;
    uint32_t init_epoch;
    uint32_t wipeouts;
} orc_memetic;

/* ik_memetic.cpp:57-64 */
static void compute_extinctions(orc_memetic* m) {
    double min_fitness = m->pop[0].fitness;
    double max_fitness = m->pop[m->P - 1].fitness;
    for (int i = 0; i < m->P; ++i) {
        double grading = (double)i / (double)(m->P - 1); /* ik_memetic.cpp:36-39 */
        m->pop[i].extinction = (m->pop[i].fitness + min_fitness * (grading - 1.0)) / max_fitness;
    }
}

/* ik_memetic.cpp:93-117.  RNG stream: (STREAM_INIT, epoch, elite index). */
static void init_population(orc_memetic* m, const double* initial_guess) {
    const orc_robot* robot = m->cx.pb->robot;
    for (int i = 0; i < m->E; ++i) {
        orc_individual* ind = &m->pop[i];
        memset(ind, 0, sizeof(*ind));
        memcpy(ind->genes, initial_guess, m->n * sizeof(double));
        if (i > 0) {
            orc_stream st;
            stream_init(&st, m->cx.pb->params->rng_seed, m->problem_index, STREAM_INIT, m->init_epoch,
                        (uint32_t)i | (m->species << 16));
            set_random_valid_configuration(robot, &st, ind->genes);
        }
        ind->fitness = cost_counted(&m->cx, ind->genes);
        ind->extinction = 1.0;
    }
    for (int i = m->E; i < m->P; ++i) {
        orc_individual* ind = &m->pop[i];
        memset(ind, 0, sizeof(*ind));
        memcpy(ind->genes, initial_guess, m->n * sizeof(double));
        ind->extinction = 1.0;
    }
    for (int i = 0; i < m->P; ++i) m->pop[i].fitness = cost_counted(&m->cx, m->pop[i].genes);
    compute_extinctions(m);
    m->has_previous = 0;
    m->init_epoch++;
}

/* ik_memetic.cpp:66-91 */
static void gradient_descent_ctx(orc_memetic* m, int i, orc_ctx* cx);
static void gradient_descent(orc_memetic* m, int i) { gradient_descent_ctx(m, i, &m->cx); }

static void gradient_descent_ctx(orc_memetic* m, int i, orc_ctx* cx) {
    const orc_params* p = cx->pb->params;
    orc_individual* ind = &m->pop[i];
    const int n = m->n;
    double gradient[ORC_MAX_VARS] = {0}, working[ORC_MAX_VARS], local[ORC_MAX_VARS], best[ORC_MAX_VARS];
    double local_cost = cost_counted(cx, ind->genes), best_cost = local_cost;
    for (int k = 0; k < n; ++k) working[k] = local[k] = best[k] = ind->genes[k];
    int num_iterations = 0;
    double previous_cost = 0.0;
    while (num_iterations < p->memetic_gd_max_iters) {
        gd_step(cx, gradient, working, local, best, &local_cost, &best_cost);
        if (fabs(local_cost - previous_cost) <= p->gd_min_cost_delta) break;
        previous_cost = local_cost;
        num_iterations++;
    }
    for (int k = 0; k < n; ++k) ind->genes[k] = best[k];
    ind->fitness = cost_counted(cx, ind->genes);
    for (int k = 0; k < n; ++k) ind->gradient[k] = gradient[k];
}

/* ik_memetic.cpp:119-190.  Stream of child i in generation g: (STREAM_REPRODUCE, g, i).  Fixed
 * word layout: parent indices from the index word list (block0.w0-3, block1.w2-3, overflow blocks
 * 2n+2, ...), idxA first, then idxB's rejection loop; mix = block1.(w0,w1); gene j: r_A =
 * block(2+2j).(w0,w1), r_B = block(2+2j).(w2,w3) (the reference leaves the order of these two
 * draws to the compiler: fixed here), r_mut = block(3+2j).(w0,w1), r_amp = block(3+2j).(w2,w3). */
static void reproduce(orc_memetic* m, uint32_t generation) {
    const orc_robot* robot = m->cx.pb->robot;
    const int n = m->n;
    const double inverse_gene_size = 1.0 / (double)n;
    int pool[ORC_MAX_VARS * 64];
    int pool_size = m->E;
    for (int i = 0; i < m->E; ++i) pool[i] = i;

    for (int i = m->E; i < m->P; ++i) {
        orc_individual* child = &m->pop[i];
        if (pool_size > 0) {
            orc_stream st;
            stream_init(&st, m->cx.pb->params->rng_seed, m->problem_index, STREAM_REPRODUCE, generation,
                        (uint32_t)i | (m->species << 16));
            uint32_t b0[4], b1[4];
            stream_block(&st, 0, b0);
            stream_block(&st, 1, b1);
            orc_index_words iw;
            iw.s = &st;
            iw.head[0] = b0[0]; iw.head[1] = b0[1]; iw.head[2] = b0[2]; iw.head[3] = b0[3];
            iw.head[4] = b1[2]; iw.head[5] = b1[3];
            iw.ovf_block = (uint32_t)(2 * n + 2);
            iw.pos = 0;
            uint32_t idxA = uniform_int_words(&iw, (uint32_t)pool_size);
            uint32_t idxB = idxA;
            while (idxB == idxA && pool_size > 1) idxB = uniform_int_words(&iw, (uint32_t)pool_size);
            const int ia = pool[idxA], ib = pool[idxB];
            const orc_individual* parentA = &m->pop[ia];
            const orc_individual* parentB = &m->pop[ib];

            double extinction = 0.5 * (parentA->extinction + parentB->extinction);
            double mutation_prob = extinction * (1.0 - inverse_gene_size) + inverse_gene_size;
            double mix_ratio = uniform_real_words(0.0, 1.0, b1[0], b1[1]);
            for (int j = 0; j < n; ++j) {
                const orc_variable* joint = &robot->vars[j];
                uint32_t wa[4], wm[4];
                stream_block(&st, (uint32_t)(2 + 2 * j), wa);
                stream_block(&st, (uint32_t)(3 + 2 * j), wm);
                double gene = mix_ratio * parentA->genes[j] + (1.0 - mix_ratio) * parentB->genes[j];
                double rA = uniform_real_words(0.0, 1.0, wa[0], wa[1]);
                double rB = uniform_real_words(0.0, 1.0, wa[2], wa[3]);
                gene += rA * parentA->gradient[j] + rB * parentB->gradient[j];
                double original_gene = gene;
                if (uniform_real_words(0.0, 1.0, wm[0], wm[1]) < mutation_prob)
                    gene += extinction * joint->half_span * uniform_real_words(-1.0, 1.0, wm[2], wm[3]);
                gene = orc_clamp_to_limits(joint, gene);
                child->genes[j] = gene;
                child->gradient[j] = gene - original_gene;
            }
            child->fitness = cost_counted(&m->cx, child->genes);
            /* parents referenced by identity; A first, then B */
            if (child->fitness < parentA->fitness) {
                for (int k = 0; k < pool_size; ++k)
                    if (pool[k] == ia) {
                        for (int l = k; l + 1 < pool_size; ++l) pool[l] = pool[l + 1];
                        pool_size--;
                        break;
                    }
            }
            if (child->fitness < parentB->fitness) {
                for (int k = 0; k < pool_size; ++k)
                    if (pool[k] == ib) {
                        for (int l = k; l + 1 < pool_size; ++l) pool[l] = pool[l + 1];
                        pool_size--;
                        break;
                    }
            }
        } else {
            orc_stream st;
            stream_init(&st, m->cx.pb->params->rng_seed, m->problem_index, STREAM_RANDOM_CHILD, generation,
                        (uint32_t)i | (m->species << 16));
            set_random_valid_configuration(robot, &st, child->genes);
            child->fitness = cost_counted(&m->cx, child->genes);
            for (int j = 0; j < n; ++j) child->gradient[j] = 0.0;
        }
    }
}

/* NaN sorts last; ties keep pre-sort order (the reference's std::sort is unstable: defined here) */
static int fit_less(double a, double b) {
    if (a != a) return 0;
    if (b != b) return 1;
    return a < b;
}

/* ik_memetic.cpp:200-209 */
static void sort_population(orc_memetic* m) {
    for (int i = 0; i < m->P; ++i) m->order[i] = i;
    for (int i = 1; i < m->P; ++i) { /* stable insertion sort on indices */
        int oi = m->order[i];
        double f = m->pop[oi].fitness;
        int k = i;
        while (k > 0 && fit_less(f, m->pop[m->order[k - 1]].fitness)) {
            m->order[k] = m->order[k - 1];
            --k;
        }
        m->order[k] = oi;
    }
    for (int i = 0; i < m->P; ++i) m->scratch[i] = m->pop[m->order[i]];
    memcpy(m->pop, m->scratch, (size_t)m->P * sizeof(orc_individual));
    compute_extinctions(m);
    m->best_curr = m->pop[0];
    if (m->best_curr.fitness < m->best.fitness) m->best = m->best_curr;
}

/* ik_memetic.cpp:43-55 */
static int check_wipeout(orc_memetic* m) {
    if (m->has_previous) {
        int improved = m->best_curr.fitness <
                       m->previous_fitness - m->cx.pb->params->memetic_wipeout_fitness_tol;
        if (!improved) return 1;
    }
    m->has_previous = 1;
    m->previous_fitness = m->best_curr.fitness;
    return 0;
}

/* MemeticIk::from + the first initPopulation (ik_memetic.cpp:18-41, 221-223) */
static void memetic_begin(orc_memetic* m, const orc_problem* pb, const double* initial_guess,
                          uint32_t problem_index, uint32_t species) {
    const orc_params* p = pb->params;
    const int n = pb->robot->n;
    memset(m, 0, sizeof(*m));
    m->cx.pb = pb;
    m->n = n;
    m->P = p->memetic_population_size;
    m->E = p->memetic_elite_size;
    m->problem_index = problem_index;
    m->species = species;
    m->pop = (orc_individual*)calloc((size_t)m->P, sizeof(orc_individual));
    m->scratch = (orc_individual*)calloc((size_t)m->P, sizeof(orc_individual));
    m->order = (int*)calloc((size_t)m->P, sizeof(int));
    memset(&m->best, 0, sizeof(m->best));
    memcpy(m->best.genes, initial_guess, n * sizeof(double));
    m->best.fitness = cost_counted(&m->cx, initial_guess);
    m->best_curr = m->best;
    init_population(m, initial_guess);
}

static void memetic_end(orc_memetic* m) {
    free(m->pop);
    free(m->scratch);
    free(m->order);
}

/* One pass of the loop body of ik_memetic_impl up to the `terminate` test (ik_memetic.cpp:229-262).
 * Returns 1 when the species returns its best individual at the solution test. */
static int memetic_generation(orc_memetic* m, int iter) {
    const orc_params* p = m->cx.pb->params;
    for (int i = 0; i < m->E; ++i) gradient_descent(m, i);
    reproduce(m, (uint32_t)iter);
    sort_population(m);
    if (p->stop_optimization_on_valid_solution && orc_is_solution(m->cx.pb, m->best.genes)) return 1;
    if (check_wipeout(m)) {
        m->wipeouts++;
        init_population(m, m->best.genes);
    }
    return 0;
}

/* the tail of ik_memetic_impl (ik_memetic.cpp:272-282) */
static int memetic_tail_found(const orc_memetic* m) {
    const orc_params* p = m->cx.pb->params;
    if (!p->stop_optimization_on_valid_solution && orc_is_solution(m->cx.pb, m->best.genes)) return 1;
    return p->return_approximate_solution ? 1 : 0;
}

/* ik_memetic.cpp:285-296 + 211-283, one species */
void orc_ik_memetic(const orc_problem* pb, const double* initial_guess, uint32_t problem_index,
                    orc_result* out) {
    const orc_params* p = pb->params;
    const int n = pb->robot->n;
    memset(out, 0, sizeof(*out));
    if (p->stop_optimization_on_valid_solution && orc_is_solution(pb, initial_guess)) {
        out->found = 1;
        memcpy(out->solution, initial_guess, n * sizeof(double));
        out->cost = orc_cost(pb, initial_guess);
        return;
    }
    orc_memetic m;
    memetic_begin(&m, pb, initial_guess, problem_index, 0);
    int iter = 0, found = 0;
    while (iter < p->memetic_max_generations) {
        if (memetic_generation(&m, iter)) {
            found = 1;
            break;
        }
        iter++;
    }
    if (!found) found = memetic_tail_found(&m);

    out->found = found;
    out->iterations = iter;
    out->cost = m.best.fitness;
    out->evals = m.cx.evals;
    out->wipeouts = m.wipeouts;
    out->gd_steps = m.cx.gd_steps;
    memcpy(out->solution, m.best.genes, n * sizeof(double));
    memetic_end(&m);
}

/* ik_memetic.cpp:285-373 with num_threads species.  The reference races its species in OS threads and takes
 * their results in the order they arrive, which no test pins; defined here (and in the CUDA product) as the
 * lockstep schedule: every live species executes generation g before any executes g + 1, and results arrive in
 * the order (generations executed, returned at the solution test before returned on `terminate`, species
 * index).  `terminate` (set when the first arrival holds a value and stop_on_first_soln, :334-346) is seen by
 * the other species at the test that ends the generation in which it was set (:264-268): they leave the loop
 * before iter++ and return through the tail (:272-282).  A flag set in the last generation is not observed
 * (the loop ends anyway; only the reported generation count would differ).  The pick (:352-370): the first
 * arrival, if it holds a value and stop_on_first_soln, unconditionally; then every value whose fitness is
 * strictly below the current minimum.  Species s draws from the streams of the problem with the individual
 * word | s << 16.  Without any value: found = 0, the lowest best fitness and the largest generation count. */
void orc_ik_memetic_species(const orc_problem* pb, const double* initial_guess, uint32_t problem_index,
                            int n_species, int stop_on_first, orc_result* out) {
    const orc_params* p = pb->params;
    const int n = pb->robot->n;
    const int G = p->memetic_max_generations;
    memset(out, 0, sizeof(*out));
    if (p->stop_optimization_on_valid_solution && orc_is_solution(pb, initial_guess)) {
        out->found = 1;
        memcpy(out->solution, initial_guess, n * sizeof(double));
        out->cost = orc_cost(pb, initial_guess);
        return;
    }
    const int S = n_species < 1 ? 1 : n_species;
    orc_memetic* ms = (orc_memetic*)calloc((size_t)S, sizeof(orc_memetic));
    int* done = (int*)calloc((size_t)S, sizeof(int));
    int* found = (int*)calloc((size_t)S, sizeof(int));
    int* iters = (int*)calloc((size_t)S, sizeof(int));
    int* phase = (int*)calloc((size_t)S, sizeof(int));
    for (int s = 0; s < S; ++s) memetic_begin(&ms[s], pb, initial_guess, problem_index, (uint32_t)s);
    for (int iter = 0; iter < G; ++iter) {
        int value_this_generation = 0, live = 0;
        for (int s = 0; s < S; ++s) {
            if (done[s]) continue;
            if (memetic_generation(&ms[s], iter)) {
                done[s] = 1; found[s] = 1; iters[s] = iter; phase[s] = 0;
                value_this_generation = 1;
            } else {
                live++;
            }
        }
        if (stop_on_first && value_this_generation && iter + 1 < G) {
            for (int s = 0; s < S; ++s) {
                if (done[s]) continue;
                done[s] = 1; found[s] = memetic_tail_found(&ms[s]); iters[s] = iter; phase[s] = 1;
            }
            live = 0;
        }
        if (!live) break;
    }
    for (int s = 0; s < S; ++s) {
        if (done[s]) continue;
        done[s] = 1; found[s] = memetic_tail_found(&ms[s]); iters[s] = G; phase[s] = 0;
    }
    /* arrival order */
    int pick = -1;
    double min_cost = 1.7976931348623157e308;
    int prev_it = -1, prev_ph = -1, prev_s = -1;
    double fail_cost = 0.0;
    int fail_it = 0;
    for (int k = 0; k < S; ++k) {
        int cur = -1;
        for (int s = 0; s < S; ++s) {
            int after_prev = iters[s] > prev_it ||
                             (iters[s] == prev_it && (phase[s] > prev_ph || (phase[s] == prev_ph && s > prev_s)));
            int before_cur = cur < 0 || iters[s] < iters[cur] ||
                             (iters[s] == iters[cur] && (phase[s] < phase[cur] || (phase[s] == phase[cur] && s < cur)));
            if (after_prev && before_cur) cur = s;
        }
        prev_it = iters[cur]; prev_ph = phase[cur]; prev_s = cur;
        double c = ms[cur].best.fitness;
        if (k == 0 || fit_less(c, fail_cost)) fail_cost = c;
        if (iters[cur] > fail_it) fail_it = iters[cur];
        if (found[cur] && ((k == 0 && stop_on_first) || c < min_cost)) {
            pick = cur;
            min_cost = c;
        }
    }
    for (int s = 0; s < S; ++s) {
        out->evals += ms[s].cx.evals;
        out->wipeouts += ms[s].wipeouts;
        out->gd_steps += ms[s].cx.gd_steps;
    }
    if (pick >= 0) {
        out->found = 1;
        out->iterations = iters[pick];
        out->cost = ms[pick].best.fitness;
        memcpy(out->solution, ms[pick].best.genes, n * sizeof(double));
    } else {
        out->found = 0;
        out->iterations = fail_it;
        out->cost = fail_cost;
        memcpy(out->solution, initial_guess, n * sizeof(double));
    }
    for (int s = 0; s < S; ++s) memetic_end(&ms[s]);
    free(ms); free(done); free(found); free(iters); free(phase);
}

/* ------------------------------------------------------------------------------------------
 * Baseline A: the reference's own execution structure, for timing only.  One solve at a time; in
 * every generation the elites' gradient descents run in E freshly created threads (ik_memetic.cpp:230-239) that
 * share one FK mutex (fk_moveit.cpp:21, pick_ik_plugin.hpp:21); reproduce, sort and the tests run on the calling
 * thread, whose cost evaluations take the same lock.  The results are those of orc_ik_memetic (the elites are
 * independent of one another).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    orc_memetic* m;
    int elite;
    orc_ctx cx;
} elite_job;

static void* elite_thread(void* arg) {
    elite_job* job = (elite_job*)arg;
    gradient_descent_ctx(job->m, job->elite, &job->cx);
    return NULL;
}

static void ik_memetic_reference_structure(const orc_problem* pb, const double* initial_guess,
                                           uint32_t problem_index, orc_result* out) {
    const orc_params* p = pb->params;
    const int n = pb->robot->n;
    memset(out, 0, sizeof(*out));
    if (p->stop_optimization_on_valid_solution && orc_is_solution(pb, initial_guess)) {
        out->found = 1;
        memcpy(out->solution, initial_guess, n * sizeof(double));
        out->cost = orc_cost(pb, initial_guess);
        return;
    }
    pthread_mutex_t fk_mutex;
    pthread_mutex_init(&fk_mutex, NULL);
    orc_memetic m;
    memetic_begin(&m, pb, initial_guess, problem_index, 0);
    m.cx.fk_mutex = &fk_mutex;
    elite_job jobs[32];
    pthread_t threads[32];
    int iter = 0, found = 0;
    while (iter < p->memetic_max_generations) {
        for (int i = 0; i < m.E; ++i) {
            jobs[i].m = &m;
            jobs[i].elite = i;
            memset(&jobs[i].cx, 0, sizeof(orc_ctx));
            jobs[i].cx.pb = pb;
            jobs[i].cx.fk_mutex = &fk_mutex;
            pthread_create(&threads[i], NULL, elite_thread, &jobs[i]);
        }
        for (int i = 0; i < m.E; ++i) {
            pthread_join(threads[i], NULL);
            m.cx.evals += jobs[i].cx.evals;
            m.cx.gd_steps += jobs[i].cx.gd_steps;
        }
        reproduce(&m, (uint32_t)iter);
        sort_population(&m);
        if (p->stop_optimization_on_valid_solution && orc_is_solution(pb, m.best.genes)) {
            found = 1;
            break;
        }
        if (check_wipeout(&m)) {
            m.wipeouts++;
            init_population(&m, m.best.genes);
        }
        iter++;
    }
    if (!found) found = memetic_tail_found(&m);
    out->found = found;
    out->iterations = iter;
    out->cost = m.best.fitness;
    out->evals = m.cx.evals;
    out->wipeouts = m.wipeouts;
    out->gd_steps = m.cx.gd_steps;
    memcpy(out->solution, m.best.genes, n * sizeof(double));
    memetic_end(&m);
    pthread_mutex_destroy(&fk_mutex);
}

void orc_solve_batch_reference_structure(const orc_robot* robot, const orc_params* params, int64_t B,
                                         int64_t first_problem_index, const double* goal_pose, const double* seed,
                                         int64_t seed_stride, double* solution, int32_t* error_code, double* cost,
                                         int32_t* iterations) {
    const int n = robot->n;
    for (int64_t b = 0; b < B; ++b) {
        const double* sd = seed + b * seed_stride;
        orc_problem pb;
        orc_problem_init(&pb, robot, params, goal_pose + 7 * robot->n_tips * b, sd);
        orc_result res;
        ik_memetic_reference_structure(&pb, sd, (uint32_t)(first_problem_index + b), &res);
        error_code[b] = res.found ? 1 : -31;
        memcpy(solution + b * n, res.found ? res.solution : sd, n * sizeof(double));
        if (cost) cost[b] = res.cost;
        if (iterations) iterations[b] = res.iterations;
    }
}

/* ------------------------------------------------------------------------------------------
 * Batch drivers
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const orc_robot* robot;
    const orc_params* params;
    int64_t B, first;
    const double* goal_pose;
    const double* seed;
    int64_t seed_stride;
    double* solution;
    int32_t* error_code;
    double* cost;
    int32_t* iterations;
    uint64_t evals;
    int64_t next; /* atomic work counter */
    pthread_mutex_t mu;
} batch_job;

static void solve_one(batch_job* job, int64_t b, uint64_t* evals) {
    const int n = job->robot->n;
    const double* seed = job->seed + b * job->seed_stride;
    orc_problem pb;
    orc_problem_init(&pb, job->robot, job->params, job->goal_pose + 7 * job->robot->n_tips * b, seed);
    orc_result res;
    if (job->params->mode == 0 && job->params->memetic_num_threads > 1)
        orc_ik_memetic_species(&pb, seed, (uint32_t)(job->first + b), job->params->memetic_num_threads,
                               job->params->memetic_stop_on_first_solution, &res);
    else if (job->params->mode == 0)
        orc_ik_memetic(&pb, seed, (uint32_t)(job->first + b), &res);
    else
        orc_ik_gradient(&pb, seed, &res);
    /* pick_ik_plugin.cpp:209-217 */
    if (res.found) {
        job->error_code[b] = 1;
        memcpy(job->solution + b * n, res.solution, n * sizeof(double));
    } else {
        job->error_code[b] = -31;
        memcpy(job->solution + b * n, seed, n * sizeof(double));
    }
    if (job->cost) job->cost[b] = res.cost;
    if (job->iterations) job->iterations[b] = res.iterations;
    *evals += res.evals;
}

static void* batch_worker(void* arg) {
    batch_job* job = (batch_job*)arg;
    uint64_t evals = 0;
    for (;;) {
        int64_t b = __atomic_fetch_add(&job->next, 1, __ATOMIC_RELAXED);
        if (b >= job->B) break;
        solve_one(job, b, &evals);
    }
    pthread_mutex_lock(&job->mu);
    job->evals += evals;
    pthread_mutex_unlock(&job->mu);
    return NULL;
}

void orc_solve_batch(const orc_robot* robot, const orc_params* params, int64_t B,
                     int64_t first_problem_index, const double* goal_pose, const double* seed,
                     int64_t seed_stride, double* solution, int32_t* error_code, double* cost,
                     int32_t* iterations, uint64_t* evals_total, int n_threads) {
    batch_job job;
    memset(&job, 0, sizeof(job));
    job.robot = robot; job.params = params; job.B = B; job.first = first_problem_index;
    job.goal_pose = goal_pose; job.seed = seed; job.seed_stride = seed_stride;
    job.solution = solution; job.error_code = error_code; job.cost = cost; job.iterations = iterations;
    pthread_mutex_init(&job.mu, NULL);
    if (n_threads <= 0) n_threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (n_threads > B) n_threads = (int)(B > 0 ? B : 1);
    if (n_threads <= 1) {
        batch_worker(&job);
    } else {
        pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
        for (int i = 0; i < n_threads; ++i) pthread_create(&th[i], NULL, batch_worker, &job);
        for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
        free(th);
    }
    pthread_mutex_destroy(&job.mu);
    if (evals_total) *evals_total = job.evals;
}

void orc_eval_cost_batch(const orc_robot* robot, const orc_params* params, int64_t B,
                         const double* goal_pose, const double* seed, int64_t seed_stride,
                         const double* q, double* cost, int32_t* is_solution, double* tip_pose) {
    const int n = robot->n, T = robot->n_tips;
    for (int64_t b = 0; b < B; ++b) {
        orc_problem pb;
        orc_problem_init(&pb, robot, params, goal_pose + 7 * T * b, seed + b * seed_stride);
        if (cost) cost[b] = orc_cost(&pb, q + b * n);
        if (is_solution) is_solution[b] = orc_is_solution(&pb, q + b * n);
        if (tip_pose) orc_poses_from_fk(robot, q + b * n, tip_pose + 7 * T * b);
    }
}

/* SURVEY 8(d) synthetic inputs: q* ~ U(min, max) per variable (unbounded: U(-pi, pi)) */
void orc_random_configuration(const orc_robot* robot, uint64_t gen_seed, uint32_t problem_index,
                              double* q) {
    orc_stream st;
    stream_init(&st, gen_seed, problem_index, STREAM_TARGET, 0, 0);
    uint32_t w[4] = {0, 0, 0, 0};
    for (int i = 0; i < robot->n; ++i) {
        const orc_variable* v = &robot->vars[i];
        if ((i & 1) == 0) stream_block(&st, (uint32_t)(i >> 1), w);
        uint32_t lo = w[2 * (i & 1)], hi = w[2 * (i & 1) + 1];
        q[i] = v->bounded ? uniform_real_words(v->min, v->max, lo, hi) : uniform_real_words(-M_PI, M_PI, lo, hi);
    }
}

void orc_poses_from_fk(const orc_robot* robot, const double* q, double* pose) {
    double R[ORC_MAX_TIPS][9], t[ORC_MAX_TIPS][3];
    orc_fk_tips(robot, q, &R[0][0], &t[0][0]);
    for (int i = 0; i < robot->n_tips; ++i) {
        pose[7 * i + 0] = t[i][0]; pose[7 * i + 1] = t[i][1]; pose[7 * i + 2] = t[i][2];
        orc_matrix_to_quat(R[i], pose + 7 * i + 3);
    }
}

void orc_pose_from_fk(const orc_robot* robot, const double* q, double pose[7]) {
    double R[9], t[3];
    orc_fk(robot, q, R, t);
    pose[0] = t[0]; pose[1] = t[1]; pose[2] = t[2];
    orc_matrix_to_quat(R, pose + 3);
}
