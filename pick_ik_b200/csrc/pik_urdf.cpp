// pik_urdf.cpp -- URDF -> chain table: the stand-alone replacement for what the reference
// takes from a MoveIt RobotModel (Robot::from, src/robot.cpp:44-85; get_active_variable_indices,
// src/robot.cpp:122-160; the joint origin transforms walked by src/fk_moveit.cpp:20-34).  Reads the <joint>
// elements of a URDF document and emits the serial chain base_link -> tip_link as pik_joint_desc[] in
// root-to-tip order, ready for pik_robot_create.
//
// Semantics follow urdfdom + MoveIt:
//   * origin rpy -> quaternion (urdfdom Rotation::setFromRPY, normalised) -> Eigen toRotationMatrix
//   * axis defaults to (1, 0, 0) and is normalised (RevoluteJointModel::setAxis)
//   * bounds (jointBoundsFromURDF): <limit lower upper>, intersected with the <safety_controller> soft limits
//     when present; max_velocity = |velocity|; a continuous joint is unbounded with the nominal range -pi..pi
//   * joint types: revolute, continuous, prismatic, fixed, floating, planar; <mimic joint multiplier offset>.
//     pik_urdf_chain (one tip, no mimic outputs) reports floating / planar / mimic joints as PIK_E_UNSUPPORTED;
//     pik_urdf_tree (several tips, mimic arrays) takes them all.
//   * SRDF: pik_srdf_group resolves a planning group to its base link and tip links (<chain base_link tip_link>
//     entries, sub-groups) -- the part of the JointModelGroup get_active_variable_indices needs.
// No exception leaves the extern "C" functions (std::bad_alloc -> PIK_E_OUT_OF_MEMORY), and a joint or link name
// that does not fit PIK_URDF_NAME_BYTES is an error (PIK_E_INVALID_ARGUMENT), never a silent truncation.
// The XML reader is a minimal tokenizer (elements, attributes, comments, declarations, CDATA skipped): URDF
// carries all of its data in attributes.  Host-only code: no CUDA in this file.
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "../../include/pik.h"

namespace {

struct Element {
    std::string name;
    std::map<std::string, std::string> attr;
    int depth = 0;       // 0 = document root element
    bool open = false;   // true: start tag (children follow), false: self-closing
};

// Calls on_open(element) for every start / self-closing tag and on_close(name, depth) for every end tag.
template <class Open, class Close>
bool scan_xml(const char* xml, Open&& on_open, Close&& on_close) {
    const char* p = xml;
    int depth = 0;
    std::vector<std::string> stack;
    while (*p) {
        const char* lt = std::strchr(p, '<');
        if (!lt) break;
        p = lt + 1;
        if (!std::strncmp(p, "!--", 3)) {
            const char* e = std::strstr(p + 3, "-->");
            if (!e) return false;
            p = e + 3;
            continue;
        }
        if (!std::strncmp(p, "![CDATA[", 8)) {
            const char* e = std::strstr(p + 8, "]]>");
            if (!e) return false;
            p = e + 3;
            continue;
        }
        if (*p == '?' || *p == '!') {  // declaration / doctype
            const char* e = std::strchr(p, '>');
            if (!e) return false;
            p = e + 1;
            continue;
        }
        if (*p == '/') {
            ++p;
            const char* e = std::strchr(p, '>');
            if (!e) return false;
            std::string name(p, e);
            while (!name.empty() && std::isspace((unsigned char)name.back())) name.pop_back();
            if (stack.empty() || stack.back() != name) return false;
            stack.pop_back();
            --depth;
            on_close(name, depth);
            p = e + 1;
            continue;
        }
        Element el;
        const char* q = p;
        while (*q && !std::isspace((unsigned char)*q) && *q != '>' && *q != '/') ++q;
        el.name.assign(p, q);
        if (el.name.empty()) return false;
        p = q;
        bool self_closing = false;
        for (;;) {
            while (*p && std::isspace((unsigned char)*p)) ++p;
            if (!*p) return false;
            if (*p == '/') {
                self_closing = true;
                ++p;
                continue;
            }
            if (*p == '>') {
                ++p;
                break;
            }
            const char* k = p;
            while (*p && *p != '=' && !std::isspace((unsigned char)*p) && *p != '>' && *p != '/') ++p;
            std::string key(k, p);
            while (*p && std::isspace((unsigned char)*p)) ++p;
            if (*p != '=') return false;
            ++p;
            while (*p && std::isspace((unsigned char)*p)) ++p;
            const char quote = *p;
            if (quote != '"' && quote != '\'') return false;
            ++p;
            const char* v = p;
            while (*p && *p != quote) ++p;
            if (!*p) return false;
            el.attr[key] = std::string(v, p);
            ++p;
        }
        el.depth = depth;
        el.open = !self_closing;
        on_open(el);
        if (!self_closing) {
            stack.push_back(el.name);
            ++depth;
        }
    }
    return stack.empty();
}

bool parse_doubles(const std::string& s, double* out, int n) {
    const char* p = s.c_str();
    for (int i = 0; i < n; ++i) {
        char* end = nullptr;
        out[i] = std::strtod(p, &end);
        if (end == p) return false;
        p = end;
    }
    while (*p && std::isspace((unsigned char)*p)) ++p;
    return *p == 0;
}

bool parse_double(const std::string& s, double* out) { return parse_doubles(s, out, 1); }

struct UrdfJoint {
    std::string name, type, parent, child, mimic;
    double xyz[3] = {0, 0, 0}, rpy[3] = {0, 0, 0}, axis[3] = {1, 0, 0};
    bool has_limit = false, has_safety = false;
    double lower = 0, upper = 0, velocity = 0, soft_lower = 0, soft_upper = 0;
    double mimic_multiplier = 1.0, mimic_offset = 0.0;
};

// urdfdom Rotation::setFromRPY -> normalise -> Eigen Quaterniond::toRotationMatrix (row-major)
void rpy_to_matrix(const double* rpy, double* R) {
    const double phi = rpy[0] / 2.0, the = rpy[1] / 2.0, psi = rpy[2] / 2.0;
    double x = std::sin(phi) * std::cos(the) * std::cos(psi) - std::cos(phi) * std::sin(the) * std::sin(psi);
    double y = std::cos(phi) * std::sin(the) * std::cos(psi) + std::sin(phi) * std::cos(the) * std::sin(psi);
    double z = std::cos(phi) * std::cos(the) * std::sin(psi) - std::sin(phi) * std::sin(the) * std::cos(psi);
    double w = std::cos(phi) * std::cos(the) * std::cos(psi) + std::sin(phi) * std::sin(the) * std::sin(psi);
    const double nrm = std::sqrt(x * x + y * y + z * z + w * w);
    x /= nrm; y /= nrm; z /= nrm; w /= nrm;
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}

// every <joint> of the document
int parse_urdf(const char* urdf_xml, std::vector<UrdfJoint>& joints) {
    bool in_robot = false, bad = false;
    int joint_depth = -1;
    const bool ok = scan_xml(
        urdf_xml,
        [&](const Element& el) {
            if (el.depth == 0 && el.name == "robot") in_robot = true;
            if (!in_robot) return;
            if (el.depth == 1 && el.name == "joint") {
                UrdfJoint j;
                auto it = el.attr.find("name");
                if (it != el.attr.end()) j.name = it->second;
                it = el.attr.find("type");
                if (it != el.attr.end()) j.type = it->second;
                joints.push_back(j);
                joint_depth = el.open ? 1 : -1;
                return;
            }
            if (joint_depth == 1 && el.depth == 2 && !joints.empty()) {
                UrdfJoint& j = joints.back();
                auto get = [&](const char* k) -> const std::string* {
                    auto f = el.attr.find(k);
                    return f == el.attr.end() ? nullptr : &f->second;
                };
                if (el.name == "parent") {
                    if (auto* v = get("link")) j.parent = *v;
                } else if (el.name == "child") {
                    if (auto* v = get("link")) j.child = *v;
                } else if (el.name == "origin") {
                    if (auto* v = get("xyz")) bad |= !parse_doubles(*v, j.xyz, 3);
                    if (auto* v = get("rpy")) bad |= !parse_doubles(*v, j.rpy, 3);
                } else if (el.name == "axis") {
                    if (auto* v = get("xyz")) bad |= !parse_doubles(*v, j.axis, 3);
                } else if (el.name == "limit") {
                    j.has_limit = true;
                    if (auto* v = get("lower")) bad |= !parse_double(*v, &j.lower);
                    if (auto* v = get("upper")) bad |= !parse_double(*v, &j.upper);
                    if (auto* v = get("velocity")) bad |= !parse_double(*v, &j.velocity);
                } else if (el.name == "safety_controller") {
                    auto* lo = get("soft_lower_limit");
                    auto* hi = get("soft_upper_limit");
                    if (lo && hi) {
                        j.has_safety = true;
                        bad |= !parse_double(*lo, &j.soft_lower);
                        bad |= !parse_double(*hi, &j.soft_upper);
                    }
                } else if (el.name == "mimic") {
                    if (auto* v = get("joint")) j.mimic = *v;
                    if (auto* v = get("multiplier")) bad |= !parse_double(*v, &j.mimic_multiplier);
                    if (auto* v = get("offset")) bad |= !parse_double(*v, &j.mimic_offset);
                }
            }
        },
        [&](const std::string& name, int depth) {
            if (depth == 1 && name == "joint") joint_depth = -1;
            if (depth == 0 && name == "robot") in_robot = false;
        });
    return (ok && !bad) ? PIK_OK : PIK_E_INVALID_ROBOT;
}

// one <joint> as the C-ABI describes it
int joint_to_desc(const UrdfJoint& j, bool allow_multi_variable, pik_joint_desc& d) {
    const double pi = 3.14159265358979323846;
    std::memset(&d, 0, sizeof(d));
    rpy_to_matrix(j.rpy, d.origin_R);
    for (int i = 0; i < 3; ++i) d.origin_t[i] = j.xyz[i];
    for (int i = 0; i < 3; ++i) d.axis[i] = j.axis[i];
    if (j.type == "fixed") {
        d.type = PIK_JOINT_FIXED;
    } else if (j.type == "revolute" || j.type == "continuous" || j.type == "prismatic") {
        d.type = j.type == "prismatic" ? PIK_JOINT_PRISMATIC : PIK_JOINT_REVOLUTE;
        const double a2 = j.axis[0] * j.axis[0] + j.axis[1] * j.axis[1] + j.axis[2] * j.axis[2];
        if (!(a2 > 0.0)) return PIK_E_INVALID_ROBOT;
        const double nrm = std::sqrt(a2);
        for (int i = 0; i < 3; ++i) d.axis[i] = j.axis[i] / nrm;
        if (j.type == "continuous") {
            d.bounded = 0;
            d.min_position = -pi;
            d.max_position = pi;
        } else {
            if (!j.has_limit) return PIK_E_INVALID_ROBOT;  // URDF requires <limit> here
            double lo = j.lower, hi = j.upper;
            if (j.has_safety) {  // jointBoundsFromURDF: soft limits intersected with the hard ones
                lo = j.soft_lower > lo ? j.soft_lower : lo;
                hi = j.soft_upper < hi ? j.soft_upper : hi;
            }
            if (!(lo <= hi)) return PIK_E_INVALID_ROBOT;
            d.bounded = 1;
            d.min_position = lo;
            d.max_position = hi;
        }
        d.max_velocity = std::fabs(j.velocity);
    } else if ((j.type == "floating" || j.type == "planar") && allow_multi_variable) {
        // MoveIt's FloatingJointModel / PlanarJointModel: translation unbounded, no velocity limit from the URDF
        d.type = j.type == "floating" ? PIK_JOINT_FLOATING : PIK_JOINT_PLANAR;
        d.bounded = 0;
        d.min_position = -pi;
        d.max_position = pi;
        d.max_velocity = std::fabs(j.velocity);
    } else if (j.type == "floating" || j.type == "planar") {
        return PIK_E_UNSUPPORTED;
    } else {
        return PIK_E_INVALID_ROBOT;
    }
    return PIK_OK;
}

int put_name(char* names, size_t k, const std::string& name) {
    if (!names) return PIK_OK;
    if (name.size() > PIK_URDF_NAME_BYTES - 1) return PIK_E_INVALID_ARGUMENT;  // never truncate silently
    char* dst = names + k * PIK_URDF_NAME_BYTES;
    std::memset(dst, 0, PIK_URDF_NAME_BYTES);
    std::memcpy(dst, name.c_str(), name.size());
    return PIK_OK;
}

int parent_joint_map(const std::vector<UrdfJoint>& joints, std::map<std::string, int>& parent_joint) {
    for (int i = 0; i < (int)joints.size(); ++i) {
        if (joints[i].child.empty() || joints[i].parent.empty()) return PIK_E_INVALID_ROBOT;
        if (parent_joint.count(joints[i].child)) return PIK_E_INVALID_ROBOT;  // a link with two parents
        parent_joint[joints[i].child] = i;
    }
    return PIK_OK;
}

int urdf_chain_impl(const char* urdf_xml, const char* base_link, const char* tip_link, pik_joint_desc* out, int32_t capacity,
                    int32_t* n_joints, char* joint_names, char* link_names) {
    std::vector<UrdfJoint> joints;
    int rc = parse_urdf(urdf_xml, joints);
    if (rc != PIK_OK) return rc;
    // walk tip -> base through the (unique) parent joint of every link
    std::map<std::string, int> parent_joint;
    if ((rc = parent_joint_map(joints, parent_joint)) != PIK_OK) return rc;
    std::vector<int> chain;
    std::string link = tip_link;
    while (link != base_link) {
        auto it = parent_joint.find(link);
        if (it == parent_joint.end()) return PIK_E_INVALID_ROBOT;  // tip is not below base
        chain.push_back(it->second);
        link = joints[it->second].parent;
        if (chain.size() > joints.size()) return PIK_E_INVALID_ROBOT;  // cycle
    }
    *n_joints = (int32_t)chain.size();
    if ((int)chain.size() > capacity) return out ? PIK_E_INVALID_ARGUMENT : PIK_OK;  // capacity 0: size query
    for (size_t k = 0; k < chain.size(); ++k) {
        const UrdfJoint& j = joints[chain[chain.size() - 1 - k]];
        if (!j.mimic.empty() && j.type != "fixed") return PIK_E_UNSUPPORTED;  // pik_urdf_tree carries mimic joints
        if ((rc = joint_to_desc(j, false, out[k])) != PIK_OK) return rc;
        if ((rc = put_name(joint_names, k, j.name)) != PIK_OK || (rc = put_name(link_names, k, j.child)) != PIK_OK) return rc;
    }
    return PIK_OK;
}

int urdf_tree_impl(const char* urdf_xml, const char* base_link, const char* const* tip_links, int32_t n_tips, pik_joint_desc* out,
                   int32_t capacity, int32_t* n_joints, int32_t* parent, int32_t* tip_joint, int32_t* mimic_of,
                   double* mimic_factor, double* mimic_offset, char* joint_names, char* link_names) {
    std::vector<UrdfJoint> joints;
    int rc = parse_urdf(urdf_xml, joints);
    if (rc != PIK_OK) return rc;
    std::map<std::string, int> parent_joint, by_name;
    if ((rc = parent_joint_map(joints, parent_joint)) != PIK_OK) return rc;
    for (int i = 0; i < (int)joints.size(); ++i) by_name[joints[i].name] = i;
    // the joints between every tip and the base (get_active_variable_indices, src/robot.cpp:122-143), plus the masters
    // of mimic joints among them when they lie below the base
    std::vector<char> used(joints.size(), 0);
    std::vector<int> tip_of(n_tips, -1);
    for (int t = 0; t < n_tips; ++t) {
        std::string link = tip_links[t];
        size_t steps = 0;
        while (link != base_link) {
            auto it = parent_joint.find(link);
            if (it == parent_joint.end()) return PIK_E_INVALID_ROBOT;  // tip is not below base
            if (tip_of[t] < 0) tip_of[t] = it->second;
            used[it->second] = 1;
            link = joints[it->second].parent;
            if (++steps > joints.size()) return PIK_E_INVALID_ROBOT;  // cycle
        }
        if (tip_of[t] < 0) return PIK_E_INVALID_ROBOT;  // the tip is the base link itself
    }
    // parents first, every branch contiguous: depth-first from the base, children in document order (a walk of the tree
    // in this order keeps one running frame per branch and saves a frame only where the tree forks)
    std::vector<int> order, stack;
    for (int i = (int)joints.size() - 1; i >= 0; --i)
        if (used[i] && joints[i].parent == base_link) stack.push_back(i);
    while (!stack.empty()) {
        const int i = stack.back();
        stack.pop_back();
        order.push_back(i);
        if (order.size() > joints.size()) return PIK_E_INVALID_ROBOT;
        for (int c = (int)joints.size() - 1; c >= 0; --c)
            if (used[c] && joints[c].parent == joints[i].child) stack.push_back(c);
    }
    *n_joints = (int32_t)order.size();
    if ((int)order.size() > capacity) return out ? PIK_E_INVALID_ARGUMENT : PIK_OK;  // capacity 0: size query
    std::map<int, int> new_index;
    for (size_t k = 0; k < order.size(); ++k) new_index[order[k]] = (int)k;
    for (size_t k = 0; k < order.size(); ++k) {
        const UrdfJoint& j = joints[order[k]];
        if ((rc = joint_to_desc(j, true, out[k])) != PIK_OK) return rc;
        if (parent) parent[k] = j.parent == base_link ? -1 : new_index[parent_joint[j.parent]];
        int master = -1;
        if (!j.mimic.empty() && j.type != "fixed") {
            auto it = by_name.find(j.mimic);
            if (it == by_name.end() || !used[it->second]) return PIK_E_UNSUPPORTED;  // master outside the tips' chains
            master = new_index[it->second];
            if (master >= (int)k) return PIK_E_UNSUPPORTED;  // a master must precede its mimic
        }
        if (master >= 0 && !mimic_of) return PIK_E_INVALID_ARGUMENT;
        if (mimic_of) mimic_of[k] = master;
        if (mimic_factor) mimic_factor[k] = master >= 0 ? j.mimic_multiplier : 1.0;
        if (mimic_offset) mimic_offset[k] = master >= 0 ? j.mimic_offset : 0.0;
        if ((rc = put_name(joint_names, k, j.name)) != PIK_OK || (rc = put_name(link_names, k, j.child)) != PIK_OK) return rc;
    }
    if (tip_joint)
        for (int t = 0; t < n_tips; ++t) tip_joint[t] = new_index[tip_of[t]];
    return PIK_OK;
}

// SRDF <group name=...>: <chain base_link tip_link/> entries and <group name=.../> sub-groups
int srdf_group_impl(const char* srdf_xml, const char* group, char* base_link, char* tip_links, int32_t capacity, int32_t* n_tips) {
    struct Group {
        std::vector<std::pair<std::string, std::string>> chains;
        std::vector<std::string> subgroups;
    };
    std::map<std::string, Group> groups;
    std::string current;
    const bool ok = scan_xml(
        srdf_xml,
        [&](const Element& el) {
            if (el.depth == 1 && el.name == "group") {
                auto it = el.attr.find("name");
                current = it == el.attr.end() ? std::string() : it->second;
                if (!current.empty()) groups[current];
                if (!el.open) current.clear();
            } else if (el.depth == 2 && !current.empty() && el.name == "chain") {
                auto b = el.attr.find("base_link"), t = el.attr.find("tip_link");
                if (b != el.attr.end() && t != el.attr.end()) groups[current].chains.emplace_back(b->second, t->second);
            } else if (el.depth == 2 && !current.empty() && el.name == "group") {
                auto it = el.attr.find("name");
                if (it != el.attr.end()) groups[current].subgroups.push_back(it->second);
            }
        },
        [&](const std::string& name, int depth) {
            if (depth == 1 && name == "group") current.clear();
        });
    if (!ok) return PIK_E_INVALID_ROBOT;
    std::vector<std::pair<std::string, std::string>> chains;
    std::vector<std::string> todo{group}, seen;
    while (!todo.empty()) {
        const std::string g = todo.back();
        todo.pop_back();
        bool dup = false;
        for (auto& s : seen) dup = dup || s == g;
        if (dup) continue;
        seen.push_back(g);
        auto it = groups.find(g);
        if (it == groups.end()) return PIK_E_INVALID_ROBOT;  // unknown group
        for (auto& c : it->second.chains) chains.push_back(c);
        for (auto& sub : it->second.subgroups) todo.push_back(sub);
    }
    if (chains.empty()) return PIK_E_UNSUPPORTED;  // a group given by <joint> / <link> lists only: no chain to resolve
    for (auto& c : chains)
        if (c.first != chains[0].first) return PIK_E_UNSUPPORTED;  // chains of one group must share their base here
    *n_tips = (int32_t)chains.size();
    int rc = put_name(base_link, 0, chains[0].first);
    if (rc != PIK_OK) return rc;
    if ((int)chains.size() > capacity) return tip_links ? PIK_E_INVALID_ARGUMENT : PIK_OK;
    for (size_t k = 0; k < chains.size(); ++k)
        if ((rc = put_name(tip_links, k, chains[k].second)) != PIK_OK) return rc;
    return PIK_OK;
}

}  // namespace

extern "C" int pik_urdf_chain(const char* urdf_xml, const char* base_link, const char* tip_link, pik_joint_desc* out,
                              int32_t capacity, int32_t* n_joints, char* joint_names, char* link_names) {
    if (n_joints) *n_joints = 0;
    if (!urdf_xml || !base_link || !tip_link || !n_joints || capacity < 0 || (capacity > 0 && !out))
        return PIK_E_INVALID_ARGUMENT;
    try {
        return urdf_chain_impl(urdf_xml, base_link, tip_link, out, capacity, n_joints, joint_names, link_names);
    } catch (const std::bad_alloc&) {
        return PIK_E_OUT_OF_MEMORY;
    } catch (...) {
        return PIK_E_INVALID_ROBOT;
    }
}

extern "C" int pik_urdf_tree(const char* urdf_xml, const char* base_link, const char* const* tip_links, int32_t n_tips,
                             pik_joint_desc* out, int32_t capacity, int32_t* n_joints, int32_t* parent, int32_t* tip_joint,
                             int32_t* mimic_of, double* mimic_factor, double* mimic_offset, char* joint_names,
                             char* link_names) {
    if (n_joints) *n_joints = 0;
    if (!urdf_xml || !base_link || !tip_links || n_tips < 1 || n_tips > PIK_MAX_TIPS || !n_joints || capacity < 0 ||
        (capacity > 0 && !out))
        return PIK_E_INVALID_ARGUMENT;
    for (int t = 0; t < n_tips; ++t)
        if (!tip_links[t]) return PIK_E_INVALID_ARGUMENT;
    try {
        return urdf_tree_impl(urdf_xml, base_link, tip_links, n_tips, out, capacity, n_joints, parent, tip_joint, mimic_of,
                              mimic_factor, mimic_offset, joint_names, link_names);
    } catch (const std::bad_alloc&) {
        return PIK_E_OUT_OF_MEMORY;
    } catch (...) {
        return PIK_E_INVALID_ROBOT;
    }
}

extern "C" int pik_srdf_group(const char* srdf_xml, const char* group, char* base_link, char* tip_links, int32_t capacity,
                              int32_t* n_tips) {
    if (n_tips) *n_tips = 0;
    if (!srdf_xml || !group || !base_link || !n_tips || capacity < 0 || (capacity > 0 && !tip_links)) return PIK_E_INVALID_ARGUMENT;
    try {
        return srdf_group_impl(srdf_xml, group, base_link, tip_links, capacity, n_tips);
    } catch (const std::bad_alloc&) {
        return PIK_E_OUT_OF_MEMORY;
    } catch (...) {
        return PIK_E_INVALID_ROBOT;
    }
}
