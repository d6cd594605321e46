// Minimal stand-in (see plugin/compat/README.md): resolves a plugin by dlsym(alias) over a list of library handles.
#pragma once
#include <dlfcn.h>

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost_plugin_loader
{
template <typename T>
struct has_getSection
{
};

class PluginLoader
{
public:
  /** dlopen handles searched in order (the real loader searches library names / paths) */
  std::vector<void*> libraries;

  template <typename PluginT>
  const PluginT* createInstance(const std::string& name) const
  {
    for (void* lib : libraries)
      if (void* sym = dlsym(lib, name.c_str()))
        return static_cast<const PluginT*>(sym);
    throw std::runtime_error("Failed to create plugin instance '" + name + "'");
  }
};
}  // namespace boost_plugin_loader
