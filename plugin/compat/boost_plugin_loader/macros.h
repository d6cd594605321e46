// Minimal stand-in (see plugin/compat/README.md).  boost_plugin_loader exports a plugin as a default-constructed global
// object of the plugin class, `extern "C"` under the alias as symbol name, placed in a named ELF section.
#pragma once
#define NB2_COMPAT_STR_(x) #x
#define NB2_COMPAT_STR(x) NB2_COMPAT_STR_(x)
#define EXPORT_CLASS_SECTIONED(DERIVED_CLASS, ALIAS, SECTION)                                                          \
  extern "C" __attribute__((visibility("default"))) DERIVED_CLASS ALIAS;                                              \
  __attribute__((section(NB2_COMPAT_STR(SECTION)), used)) DERIVED_CLASS ALIAS;
