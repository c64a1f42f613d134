// Header mock (tests only): moveit::core::RobotState is only named by the plugin interface (context_state).
#pragma once
#include <moveit/robot_model/robot_model.h>
namespace moveit::core {
class RobotState {};
}  // namespace moveit::core
