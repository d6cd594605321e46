// Stand-in for the one helper of noether_tpp/serialization.h:10-16 the shim uses; see plugin/compat/README.md.
#pragma once
#include <yaml-cpp/yaml.h>

#include <stdexcept>
#include <string>

namespace YAML
{
template <typename T>
T getMember(const Node& node, const std::string& key)
{
  if (!node[key])
    throw std::runtime_error("Required member '" + key + "' not found in YAML node");
  return node[key].as<T>();
}
}  // namespace YAML
