// Minimal stand-in for yaml-cpp's YAML::Node (see plugin/compat/README.md): scalars, maps and sequences built in code.
#pragma once
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace YAML
{
class Node
{
public:
  Node() = default;
  Node(double v) { setScalar(v); }
  Node(int v) { setScalar(v); }
  Node(bool v) : data_(std::make_shared<Data>()) { data_->scalar = v ? "true" : "false"; }
  Node(const char* v) : data_(std::make_shared<Data>()) { data_->scalar = v; }
  Node(const std::string& v) : data_(std::make_shared<Data>()) { data_->scalar = v; }

  explicit operator bool() const { return static_cast<bool>(data_); }
  bool operator!() const { return !data_; }

  /** map access; a missing key yields an undefined node (as yaml-cpp does on a const node) */
  Node operator[](const std::string& key) const
  {
    if (!data_)
      return Node();
    auto it = data_->map.find(key);
    return it == data_->map.end() ? Node() : it->second;
  }
  Node& operator[](const std::string& key)
  {
    if (!data_)
      data_ = std::make_shared<Data>();
    return data_->map[key];
  }
  void push_back(const Node& n)
  {
    if (!data_)
      data_ = std::make_shared<Data>();
    data_->seq.push_back(n);
  }
  std::size_t size() const { return data_ ? data_->seq.size() : 0; }
  Node operator[](std::size_t i) const { return data_->seq.at(i); }
  /** sequence iteration */
  std::vector<Node>::const_iterator begin() const { return data_ ? data_->seq.begin() : empty().begin(); }
  std::vector<Node>::const_iterator end() const { return data_ ? data_->seq.end() : empty().end(); }

  template <typename T>
  T as() const;

private:
  struct Data
  {
    std::string scalar;
    std::map<std::string, Node> map;
    std::vector<Node> seq;
  };
  template <typename T>
  void setScalar(T v)
  {
    data_ = std::make_shared<Data>();
    std::ostringstream ss;
    ss.precision(17);
    ss << v;
    data_->scalar = ss.str();
  }
  static const std::vector<Node>& empty()
  {
    static const std::vector<Node> none;
    return none;
  }
  std::shared_ptr<Data> data_;
};

template <>
inline Node Node::as<Node>() const
{
  return *this;
}
template <>
inline std::string Node::as<std::string>() const
{
  if (!data_)
    throw std::runtime_error("bad conversion");
  return data_->scalar;
}
template <>
inline double Node::as<double>() const
{
  if (!data_)
    throw std::runtime_error("bad conversion");
  return std::stod(data_->scalar);
}
template <>
inline float Node::as<float>() const
{
  if (!data_)
    throw std::runtime_error("bad conversion");
  return std::stof(data_->scalar);
}
template <>
inline int Node::as<int>() const
{
  if (!data_)
    throw std::runtime_error("bad conversion");
  return std::stoi(data_->scalar);
}
template <>
inline bool Node::as<bool>() const
{
  if (!data_)
    throw std::runtime_error("bad conversion");
  return data_->scalar == "true" || data_->scalar == "True" || data_->scalar == "1";
}
}  // namespace YAML
