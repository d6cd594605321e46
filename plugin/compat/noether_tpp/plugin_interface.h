// Stand-in declaring noether::Plugin<T>, noether::Factory and the EXPORT_*_PLUGIN macros
// (noether_tpp/plugin_interface.h:17-168; section names from noether_tpp/CMakeLists.txt:126-131);
// see plugin/compat/README.md.
#pragma once
#include <noether_tpp/core/mesh_modifier.h>
#include <noether_tpp/core/tool_path_modifier.h>
#include <noether_tpp/core/tool_path_planner.h>
#include <noether_tpp/tool_path_planners/raster/raster_planner.h>

#include <boost_plugin_loader/plugin_loader.h>
#include <yaml-cpp/yaml.h>

#include <memory>
#include <string>

#ifndef NOETHER_TPP_SECTION
#define NOETHER_TPP_SECTION tpp
#define NOETHER_DIRECTION_GENERATOR_SECTION dg
#define NOETHER_ORIGIN_GENERATOR_SECTION og
#define NOETHER_TOOL_PATH_MODIFIER_SECTION mod
#define NOETHER_MESH_MODIFIER_SECTION mesh
#endif

namespace noether
{
class Factory;

template <typename T>
class Plugin
{
public:
  typedef T ComponentT;
  using Ptr = std::shared_ptr<Plugin>;
  virtual ~Plugin() = default;
  virtual std::unique_ptr<ComponentT> create(const YAML::Node& config, std::shared_ptr<const Factory> factory) const = 0;
};

using ToolPathPlannerPlugin = Plugin<ToolPathPlanner>;
using DirectionGeneratorPlugin = Plugin<DirectionGenerator>;
using OriginGeneratorPlugin = Plugin<OriginGenerator>;
using ToolPathModifierPlugin = Plugin<ToolPathModifier>;
using MeshModifierPlugin = Plugin<MeshModifier>;

class Factory
{
public:
  explicit Factory(std::shared_ptr<const boost_plugin_loader::PluginLoader> loader) : loader_(std::move(loader)) {}

  MeshModifier::Ptr createMeshModifier(const YAML::Node& config) const { return create<MeshModifierPlugin>(config); }
  ToolPathPlanner::Ptr createToolPathPlanner(const YAML::Node& config) const { return create<ToolPathPlannerPlugin>(config); }
  DirectionGenerator::Ptr createDirectionGenerator(const YAML::Node& config) const
  {
    return create<DirectionGeneratorPlugin>(config);
  }
  OriginGenerator::Ptr createOriginGenerator(const YAML::Node& config) const { return create<OriginGeneratorPlugin>(config); }
  ToolPathModifier::Ptr createToolPathModifier(const YAML::Node& config) const { return create<ToolPathModifierPlugin>(config); }

protected:
  std::shared_ptr<const boost_plugin_loader::PluginLoader> loader_;

  /** plugin_interface.cpp:63-70: look the alias up, hand the plugin a non-owning pointer to this factory */
  template <typename PluginT>
  std::unique_ptr<typename PluginT::ComponentT> create(const YAML::Node& config) const
  {
    if (!config["name"])
      throw std::runtime_error("Required member 'name' not found in YAML node");
    const auto name = config["name"].template as<std::string>();
    const PluginT* plugin = loader_->createInstance<PluginT>(name);
    return plugin->create(config, std::shared_ptr<const Factory>(this, [](const Factory*) {}));
  }
};
}  // namespace noether

#include <boost_plugin_loader/macros.h>

#define EXPORT_MESH_MODIFIER_PLUGIN(DERIVED_CLASS, ALIAS)                                                              \
  EXPORT_CLASS_SECTIONED(DERIVED_CLASS, ALIAS, NOETHER_MESH_MODIFIER_SECTION)
#define EXPORT_TOOL_PATH_PLANNER_PLUGIN(DERIVED_CLASS, ALIAS) EXPORT_CLASS_SECTIONED(DERIVED_CLASS, ALIAS, NOETHER_TPP_SECTION)
#define EXPORT_DIRECTION_GENERATOR_PLUGIN(DERIVED_CLASS, ALIAS)                                                        \
  EXPORT_CLASS_SECTIONED(DERIVED_CLASS, ALIAS, NOETHER_DIRECTION_GENERATOR_SECTION)
#define EXPORT_ORIGIN_GENERATOR_PLUGIN(DERIVED_CLASS, ALIAS)                                                           \
  EXPORT_CLASS_SECTIONED(DERIVED_CLASS, ALIAS, NOETHER_ORIGIN_GENERATOR_SECTION)
#define EXPORT_TOOL_PATH_MODIFIER_PLUGIN(DERIVED_CLASS, ALIAS)                                                         \
  EXPORT_CLASS_SECTIONED(DERIVED_CLASS, ALIAS, NOETHER_TOOL_PATH_MODIFIER_SECTION)
