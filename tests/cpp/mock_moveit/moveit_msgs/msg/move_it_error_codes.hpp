// Header mock (tests only) of moveit_msgs/msg/move_it_error_codes.hpp
#pragma once
namespace moveit_msgs::msg {
struct MoveItErrorCodes {
    static constexpr int SUCCESS = 1;
    static constexpr int NO_IK_SOLUTION = -31;
    int val = 0;
};
}  // namespace moveit_msgs::msg
