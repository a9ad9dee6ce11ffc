// TEST INFRASTRUCTURE — not part of the product.
//
// Minimal stand-in for the toml++ header (marzer/tomlplusplus; the reference's
// CMakeLists.txt:6-18 locates it through $TOMLCPP_DIR, no version pinned; it is
// NOT vendored under /root/reference and not installed in this image).  It
// exposes only the API subset the reference sources touch, so that those
// sources compile UNMODIFIED, where they lie, into oracle/_ref/:
//   toml::table, toml::array, toml::node, toml::node_view<T>,
//   toml::parse_file, toml::parse_error,
//   operator[] lookup, .value<T>(), .as_array(), array::size()/at().
// Key lookup only — no arithmetic of the hot path lives in toml++.
// Parsing is delegated to the repo's own toml_lite reader.
#ifndef ORACLE_SHIM_TOMLPP_HPP
#define ORACLE_SHIM_TOMLPP_HPP

#include <cstdint>
#include <optional>
#include <ostream>
#include <stdexcept>
#include <string>
#include <string_view>
#include <type_traits>

#include "../../../lattice-boltzmann-method_b200/host/toml_lite.hpp"

namespace toml
{

class parse_error : public std::runtime_error
{
public:
  explicit parse_error(const std::string& w) : std::runtime_error(w) {}
};

inline std::ostream& operator<<(std::ostream& os, const parse_error& e) { return os << e.what(); }

class array;
class table;

// A node is a thin handle on a toml_lite::value.
class node
{
protected:
  toml_lite::value_ptr v_;

public:
  node() = default;
  explicit node(toml_lite::value_ptr v) : v_(std::move(v)) {}
  const toml_lite::value* raw() const { return v_.get(); }
  const toml_lite::value_ptr& ptr() const { return v_; }

  template <typename T>
  std::optional<T> value() const
  {
    if (!v_) return std::nullopt;
    if constexpr (std::is_same_v<T, bool>)
      return v_->as_bool();
    else if constexpr (std::is_floating_point_v<T>)
    {
      auto d = v_->as_double();
      if (d) return static_cast<T>(*d);
      return std::nullopt;
    }
    else if constexpr (std::is_integral_v<T>)
    {
      auto i = v_->as_int();
      if (i) return static_cast<T>(*i);
      return std::nullopt;
    }
    else if constexpr (std::is_same_v<T, std::string> || std::is_same_v<T, std::string_view>)
    {
      auto s = v_->as_string();
      if (s) return T(*s);
      return std::nullopt;
    }
    else
      return std::nullopt;
  }

};

// toml++ hands out `const toml::array*` pointers that live as long as the
// parsed document (src/ibm.cpp:78-79 keeps them across statements).  Here an
// array pointer is an overlay on the toml_lite::value that the document owns:
// the class has no data members and is never instantiated.
class array
{
  array() = delete;
  const toml_lite::value* self() const { return reinterpret_cast<const toml_lite::value*>(this); }

public:
  std::size_t size() const { return self()->arr.size(); }
  node at(std::size_t i) const { return node(self()->arr.at(i)); }
  node operator[](std::size_t i) const { return node(self()->arr[i]); }
};

template <typename ViewedType>
class node_view
{
  toml_lite::value_ptr v_;

public:
  node_view() = default;
  explicit node_view(toml_lite::value_ptr v) : v_(std::move(v)) {}
  explicit operator bool() const { return static_cast<bool>(v_); }

  node_view operator[](std::string_view key) const
  {
    if (!v_ || v_->kind != toml_lite::value::TABLE) return node_view();
    auto it = v_->tbl.find(std::string(key));
    if (it == v_->tbl.end()) return node_view();
    return node_view(it->second);
  }

  template <typename T>
  std::optional<T> value() const
  {
    return node(v_).template value<T>();
  }

  const array* as_array() const
  {
    if (!v_ || v_->kind != toml_lite::value::ARRAY) return nullptr;
    return reinterpret_cast<const array*>(v_.get());
  }
};

class table : public node
{
public:
  table()
  {
    v_ = std::make_shared<toml_lite::value>();
    v_->kind = toml_lite::value::TABLE;
  }
  explicit table(toml_lite::value_ptr v) : node(std::move(v)) {}

  node_view<node> operator[](std::string_view key)
  {
    auto it = v_->tbl.find(std::string(key));
    if (it == v_->tbl.end()) return node_view<node>();
    return node_view<node>(it->second);
  }
  node_view<const node> operator[](std::string_view key) const
  {
    auto it = v_->tbl.find(std::string(key));
    if (it == v_->tbl.end()) return node_view<const node>();
    return node_view<const node>(it->second);
  }
};

inline table parse_file(std::string_view path)
{
  try
  {
    return table(toml_lite::parse_file(std::string(path)));
  }
  catch (const toml_lite::parse_error& e)
  {
    throw parse_error(e.what());
  }
}

inline table parse(std::string_view text)
{
  try
  {
    return table(toml_lite::parse_string(std::string(text)));
  }
  catch (const toml_lite::parse_error& e)
  {
    throw parse_error(e.what());
  }
}

}  // namespace toml

#endif
