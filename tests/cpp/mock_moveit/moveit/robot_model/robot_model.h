// Header mock (tests only) of the moveit::core model classes the plugin reads: the member functions, names and
// semantics of moveit_core's RobotModel / JointModel / LinkModel / JointModelGroup that
// src/robot.cpp, src/fk_moveit.cpp and src/pick_ik_plugin.cpp of the reference call.
#pragma once
#include <Eigen/Geometry>
#include <algorithm>
#include <memory>
#include <string>
#include <vector>
namespace moveit::core {
struct VariableBounds {
    double min_position_ = 0, max_position_ = 0;
    bool position_bounded_ = false;
    double max_velocity_ = 0;
    bool velocity_bounded_ = false;
};
class LinkModel;
class JointModel {
   public:
    enum JointType { UNKNOWN, REVOLUTE, PRISMATIC, PLANAR, FLOATING, FIXED };
    JointModel(std::string name, JointType type, size_t variables) : name_(std::move(name)), type_(type), bounds_(variables) {}
    virtual ~JointModel() = default;
    std::string const& getName() const { return name_; }
    JointType getType() const { return type_; }
    size_t getVariableCount() const { return bounds_.size(); }
    std::vector<VariableBounds> const& getVariableBounds() const { return bounds_; }
    std::vector<VariableBounds>& bounds() { return bounds_; }
    LinkModel const* getParentLinkModel() const { return parent_link_; }
    LinkModel const* getChildLinkModel() const { return child_link_; }
    JointModel const* getMimic() const { return mimic_; }
    double getMimicFactor() const { return mimic_factor_; }
    double getMimicOffset() const { return mimic_offset_; }
    void setMimic(JointModel const* m, double factor, double offset) { mimic_ = m; mimic_factor_ = factor; mimic_offset_ = offset; }
    size_t getJointIndex() const { return index_; }
    LinkModel const* parent_link_ = nullptr;
    LinkModel const* child_link_ = nullptr;
    size_t index_ = 0;
   private:
    std::string name_;
    JointType type_;
    std::vector<VariableBounds> bounds_;
    JointModel const* mimic_ = nullptr;
    double mimic_factor_ = 1.0, mimic_offset_ = 0.0;
};
class RevoluteJointModel : public JointModel {
   public:
    RevoluteJointModel(std::string name, Eigen::Vector3d axis) : JointModel(std::move(name), REVOLUTE, 1), axis_(axis) {}
    Eigen::Vector3d const& getAxis() const { return axis_; }
   private:
    Eigen::Vector3d axis_;
};
class PrismaticJointModel : public JointModel {
   public:
    PrismaticJointModel(std::string name, Eigen::Vector3d axis) : JointModel(std::move(name), PRISMATIC, 1), axis_(axis) {}
    Eigen::Vector3d const& getAxis() const { return axis_; }
   private:
    Eigen::Vector3d axis_;
};
class LinkModel {
   public:
    explicit LinkModel(std::string name) : name_(std::move(name)) {}
    std::string const& getName() const { return name_; }
    Eigen::Isometry3d const& getJointOriginTransform() const { return origin_; }
    JointModel const* getParentJointModel() const { return parent_joint_; }
    Eigen::Isometry3d origin_;
    JointModel const* parent_joint_ = nullptr;
   private:
    std::string name_;
};
class JointModelGroup {
   public:
    explicit JointModelGroup(std::string name) : name_(std::move(name)) {}
    std::string const& getName() const { return name_; }
    bool hasJointModel(std::string const& joint) const { return std::find(joints_.begin(), joints_.end(), joint) != joints_.end(); }
    std::vector<std::string> joints_;
   private:
    std::string name_;
};
class RobotModel {
   public:
    std::string const& getModelFrame() const { return model_frame_; }
    std::vector<JointModel const*> const& getJointModels() const { return joint_ptrs_; }
    JointModelGroup const* getJointModelGroup(std::string const& name) const {
        for (auto const& g : groups_) if (g->getName() == name) return g.get();
        return nullptr;
    }
    // test construction helpers
    std::string model_frame_;
    std::vector<std::unique_ptr<JointModel>> joints_;
    std::vector<std::unique_ptr<LinkModel>> links_;
    std::vector<std::unique_ptr<JointModelGroup>> groups_;
    std::vector<JointModel const*> joint_ptrs_;
};
}  // namespace moveit::core
