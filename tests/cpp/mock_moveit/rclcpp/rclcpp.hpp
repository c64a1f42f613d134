// Header mock (tests only) of the rclcpp calls the plugin makes: parameters and logging.
#pragma once
#include <cstdio>
#include <map>
#include <memory>
#include <sstream>
#include <string>
namespace rclcpp {
struct Logger { std::string name; };
inline Logger get_logger(std::string const& name) { return Logger{name}; }
class Node {
    std::map<std::string, std::string> values_;
    template <class T> static std::string to_text(T const& v) { std::ostringstream s; s.precision(17); s << v; return s.str(); }
    template <class T> static void from_text(std::string const& t, T& v) { std::istringstream s(t); s >> std::boolalpha >> v; if (s.fail()) { std::istringstream s2(t); s2 >> v; } }
   public:
    using SharedPtr = std::shared_ptr<Node>;
    bool has_parameter(std::string const& name) const { return values_.count(name) != 0; }
    template <class T> T declare_parameter(std::string const& name, T const& fallback) { values_[name] = to_text(fallback); return fallback; }
    template <class T> bool get_parameter(std::string const& name, T& out) const {
        auto it = values_.find(name);
        if (it == values_.end()) return false;
        from_text(it->second, out);
        return true;
    }
    template <class T> void set_parameter_for_test(std::string const& name, T const& v) { values_[name] = to_text(v); }
};
template <> inline void Node::from_text<std::string>(std::string const& t, std::string& v) { v = t; }
}  // namespace rclcpp
#define RCLCPP_ERROR(logger, ...) do { std::fprintf(stderr, "[%s] ", (logger).name.c_str()); std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define RCLCPP_WARN(logger, ...) RCLCPP_ERROR(logger, __VA_ARGS__)
