// Stand-in declaring noether::ToolPathModifier; see plugin/compat/README.md.
#pragma once
#include <noether_tpp/core/types.h>

#include <memory>

namespace noether
{
struct ToolPathModifier
{
  using Ptr = std::unique_ptr<ToolPathModifier>;
  using ConstPtr = std::unique_ptr<const ToolPathModifier>;
  virtual ~ToolPathModifier() = default;
  virtual ToolPaths modify(ToolPaths tool_paths) const = 0;
};
}  // namespace noether
