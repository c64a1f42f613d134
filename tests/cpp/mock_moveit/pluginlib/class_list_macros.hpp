// Header mock (tests only): PLUGINLIB_EXPORT_CLASS registers a factory the test can call.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
namespace pluginlib_mock {
template <class Base> std::map<std::string, std::function<std::shared_ptr<Base>()>>& registry() {
    static std::map<std::string, std::function<std::shared_ptr<Base>()>> r;
    return r;
}
template <class Derived, class Base> struct Registrar {
    explicit Registrar(char const* name) { registry<Base>()[name] = [] { return std::static_pointer_cast<Base>(std::make_shared<Derived>()); }; }
};
}  // namespace pluginlib_mock
#define PLUGINLIB_MOCK_CAT2(a, b) a##b
#define PLUGINLIB_MOCK_CAT(a, b) PLUGINLIB_MOCK_CAT2(a, b)
#define PLUGINLIB_EXPORT_CLASS(Derived, Base) \
    static pluginlib_mock::Registrar<Derived, Base> PLUGINLIB_MOCK_CAT(pluginlib_mock_registrar_, __LINE__)(#Derived);
