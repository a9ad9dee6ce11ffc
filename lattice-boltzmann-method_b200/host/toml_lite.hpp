// toml_lite.hpp — a small, dependency-free TOML reader for the host side.
//
// The reference reads its run configuration through toml++ (a third-party,
// header-only library that is NOT vendored under /root/reference and is not
// installed in this image).  Only key lookup lives there; no arithmetic.  This
// header provides what the host drivers need: tables, dotted/bare/quoted keys,
// integers, floats (incl. exponent forms such as 1.0533E-6), strings,
// booleans, (multi-line, nested) arrays and inline tables.
//
// Reference call sites this serves: src/params.cpp:7-120, src/colour.cpp:11-47,
// src/ibm.cpp:78-102, test/mrtcg_rayleigh_taylor.cpp:103-117,360-369.
#ifndef LBM_TOML_LITE_HPP
#define LBM_TOML_LITE_HPP

#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace toml_lite
{

struct parse_error : std::runtime_error
{
  int line;
  parse_error(const std::string& what, int line_)
      : std::runtime_error(what + " (line " + std::to_string(line_) + ")"), line(line_) {}
};

struct value;
using value_ptr = std::shared_ptr<value>;

struct value
{
  enum kind_t { NONE, INTEGER, FLOAT, STRING, BOOLEAN, ARRAY, TABLE } kind = NONE;
  std::int64_t i = 0;
  double d = 0.0;
  bool b = false;
  std::string s;
  std::vector<value_ptr> arr;
  std::map<std::string, value_ptr> tbl;
  bool defined_inline = false;

  bool is_table() const { return kind == TABLE; }
  bool is_array() const { return kind == ARRAY; }

  const value* find(const std::string& key) const
  {
    if (kind != TABLE) return nullptr;
    auto it = tbl.find(key);
    return it == tbl.end() ? nullptr : it->second.get();
  }

  // Same conversion rules the reference relies on in toml++'s value<T>():
  // an integer node can be read as double (x_multiplier = 9), a float node can
  // be read as an integer only when it holds an integral value.
  std::optional<double> as_double() const
  {
    if (kind == FLOAT) return d;
    if (kind == INTEGER) return static_cast<double>(i);
    return std::nullopt;
  }
  std::optional<std::int64_t> as_int() const
  {
    if (kind == INTEGER) return i;
    if (kind == FLOAT && std::floor(d) == d && std::isfinite(d)) return static_cast<std::int64_t>(d);
    return std::nullopt;
  }
  std::optional<std::string> as_string() const
  {
    if (kind == STRING) return s;
    return std::nullopt;
  }
  std::optional<bool> as_bool() const
  {
    if (kind == BOOLEAN) return b;
    return std::nullopt;
  }
};

class parser
{
public:
  explicit parser(const std::string& text) : t_(text) {}

  value_ptr parse()
  {
    auto root = std::make_shared<value>();
    root->kind = value::TABLE;
    value* current = root.get();
    while (true)
    {
      skip_ws_comments_newlines();
      if (eof()) break;
      if (peek() == '[')
      {
        bool is_array_table = false;
        ++p_;
        if (!eof() && peek() == '[') { is_array_table = true; ++p_; }
        skip_inline_ws();
        std::vector<std::string> path = parse_key_path();
        skip_inline_ws();
        expect(']');
        if (is_array_table) expect(']');
        current = descend(root.get(), path, is_array_table);
        end_of_line();
      }
      else
      {
        std::vector<std::string> path = parse_key_path();
        skip_inline_ws();
        expect('=');
        skip_inline_ws();
        value_ptr v = parse_value();
        assign(current, path, v);
        end_of_line();
      }
    }
    return root;
  }

private:
  const std::string& t_;
  std::size_t p_ = 0;
  int line_ = 1;

  bool eof() const { return p_ >= t_.size(); }
  char peek() const { return t_[p_]; }
  [[noreturn]] void fail(const std::string& m) const { throw parse_error(m, line_); }
  void expect(char c)
  {
    if (eof() || peek() != c) fail(std::string("expected '") + c + "'");
    ++p_;
  }
  void skip_inline_ws()
  {
    while (!eof() && (peek() == ' ' || peek() == '\t')) ++p_;
  }
  void skip_comment()
  {
    if (!eof() && peek() == '#')
      while (!eof() && peek() != '\n') ++p_;
  }
  void skip_ws_comments_newlines()
  {
    while (!eof())
    {
      char c = peek();
      if (c == ' ' || c == '\t' || c == '\r') ++p_;
      else if (c == '\n') { ++p_; ++line_; }
      else if (c == '#') skip_comment();
      else break;
    }
  }
  void end_of_line()
  {
    skip_inline_ws();
    skip_comment();
    if (eof()) return;
    if (peek() == '\r') ++p_;
    if (eof()) return;
    if (peek() != '\n') fail("unexpected characters after value");
    ++p_;
    ++line_;
  }

  static bool bare_key_char(char c)
  {
    return std::isalnum(static_cast<unsigned char>(c)) || c == '_' || c == '-';
  }

  std::string parse_simple_key()
  {
    if (eof()) fail("key expected");
    if (peek() == '"') return parse_basic_string();
    if (peek() == '\'') return parse_literal_string();
    std::size_t b = p_;
    while (!eof() && bare_key_char(peek())) ++p_;
    if (b == p_) fail("key expected");
    return t_.substr(b, p_ - b);
  }

  std::vector<std::string> parse_key_path()
  {
    std::vector<std::string> path;
    path.push_back(parse_simple_key());
    while (true)
    {
      skip_inline_ws();
      if (!eof() && peek() == '.')
      {
        ++p_;
        skip_inline_ws();
        path.push_back(parse_simple_key());
      }
      else break;
    }
    return path;
  }

  value* descend(value* root, const std::vector<std::string>& path, bool array_table)
  {
    value* cur = root;
    for (std::size_t k = 0; k < path.size(); ++k)
    {
      const bool last = (k + 1 == path.size());
      auto it = cur->tbl.find(path[k]);
      if (it == cur->tbl.end())
      {
        auto nv = std::make_shared<value>();
        if (last && array_table)
        {
          nv->kind = value::ARRAY;
          auto elem = std::make_shared<value>();
          elem->kind = value::TABLE;
          nv->arr.push_back(elem);
          cur->tbl[path[k]] = nv;
          cur = elem.get();
        }
        else
        {
          nv->kind = value::TABLE;
          cur->tbl[path[k]] = nv;
          cur = nv.get();
        }
      }
      else
      {
        value* v = it->second.get();
        if (v->kind == value::ARRAY)
        {
          if (last && array_table)
          {
            auto elem = std::make_shared<value>();
            elem->kind = value::TABLE;
            v->arr.push_back(elem);
            cur = elem.get();
          }
          else
          {
            if (v->arr.empty() || v->arr.back()->kind != value::TABLE) fail("key is not a table: " + path[k]);
            cur = v->arr.back().get();
          }
        }
        else if (v->kind == value::TABLE)
        {
          if (v->defined_inline) fail("cannot extend inline table: " + path[k]);
          cur = v;
        }
        else fail("key is not a table: " + path[k]);
      }
    }
    return cur;
  }

  void assign(value* tbl, const std::vector<std::string>& path, const value_ptr& v)
  {
    value* cur = tbl;
    for (std::size_t k = 0; k + 1 < path.size(); ++k)
    {
      auto it = cur->tbl.find(path[k]);
      if (it == cur->tbl.end())
      {
        auto nv = std::make_shared<value>();
        nv->kind = value::TABLE;
        cur->tbl[path[k]] = nv;
        cur = nv.get();
      }
      else
      {
        if (it->second->kind != value::TABLE) fail("key is not a table: " + path[k]);
        cur = it->second.get();
      }
    }
    if (cur->tbl.count(path.back())) fail("duplicate key: " + path.back());
    cur->tbl[path.back()] = v;
  }

  std::string parse_basic_string()
  {
    expect('"');
    // multi-line basic string
    if (t_.compare(p_, 2, "\"\"") == 0)
    {
      p_ += 2;
      if (!eof() && peek() == '\r') ++p_;
      if (!eof() && peek() == '\n') { ++p_; ++line_; }
      std::string out;
      while (true)
      {
        if (eof()) fail("unterminated multi-line string");
        if (t_.compare(p_, 3, "\"\"\"") == 0) { p_ += 3; break; }
        char c = t_[p_++];
        if (c == '\n') ++line_;
        if (c == '\\') out += parse_escape(true);
        else out += c;
      }
      return out;
    }
    std::string out;
    while (true)
    {
      if (eof() || peek() == '\n') fail("unterminated string");
      char c = t_[p_++];
      if (c == '"') break;
      if (c == '\\') out += parse_escape(false);
      else out += c;
    }
    return out;
  }

  std::string parse_escape(bool multiline)
  {
    if (eof()) fail("bad escape");
    char c = t_[p_++];
    switch (c)
    {
      case 'b': return "\b";
      case 't': return "\t";
      case 'n': return "\n";
      case 'f': return "\f";
      case 'r': return "\r";
      case '"': return "\"";
      case '\\': return "\\";
      case 'u': case 'U':
      {
        int n = (c == 'u') ? 4 : 8;
        if (p_ + n > t_.size()) fail("bad unicode escape");
        unsigned long cp = std::strtoul(t_.substr(p_, n).c_str(), nullptr, 16);
        p_ += n;
        return encode_utf8(cp);
      }
      case '\n': case ' ': case '\t': case '\r':
        if (multiline)
        {
          if (c == '\n') ++line_;
          while (!eof() && (peek() == ' ' || peek() == '\t' || peek() == '\n' || peek() == '\r'))
          {
            if (peek() == '\n') ++line_;
            ++p_;
          }
          return "";
        }
        [[fallthrough]];
      default: fail("bad escape");
    }
  }

  static std::string encode_utf8(unsigned long cp)
  {
    std::string o;
    if (cp < 0x80) o += static_cast<char>(cp);
    else if (cp < 0x800) { o += static_cast<char>(0xC0 | (cp >> 6)); o += static_cast<char>(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000)
    {
      o += static_cast<char>(0xE0 | (cp >> 12));
      o += static_cast<char>(0x80 | ((cp >> 6) & 0x3F));
      o += static_cast<char>(0x80 | (cp & 0x3F));
    }
    else
    {
      o += static_cast<char>(0xF0 | (cp >> 18));
      o += static_cast<char>(0x80 | ((cp >> 12) & 0x3F));
      o += static_cast<char>(0x80 | ((cp >> 6) & 0x3F));
      o += static_cast<char>(0x80 | (cp & 0x3F));
    }
    return o;
  }

  std::string parse_literal_string()
  {
    expect('\'');
    if (t_.compare(p_, 2, "''") == 0)
    {
      p_ += 2;
      if (!eof() && peek() == '\r') ++p_;
      if (!eof() && peek() == '\n') { ++p_; ++line_; }
      std::size_t e = t_.find("'''", p_);
      if (e == std::string::npos) fail("unterminated multi-line literal string");
      std::string out = t_.substr(p_, e - p_);
      for (char c : out) if (c == '\n') ++line_;
      p_ = e + 3;
      return out;
    }
    std::size_t b = p_;
    while (!eof() && peek() != '\'' && peek() != '\n') ++p_;
    if (eof() || peek() != '\'') fail("unterminated literal string");
    std::string out = t_.substr(b, p_ - b);
    ++p_;
    return out;
  }

  value_ptr parse_value()
  {
    if (eof()) fail("value expected");
    auto v = std::make_shared<value>();
    char c = peek();
    if (c == '"') { v->kind = value::STRING; v->s = parse_basic_string(); return v; }
    if (c == '\'') { v->kind = value::STRING; v->s = parse_literal_string(); return v; }
    if (c == '[') return parse_array();
    if (c == '{') return parse_inline_table();
    if (t_.compare(p_, 4, "true") == 0 && !is_token_char(p_ + 4))
    { p_ += 4; v->kind = value::BOOLEAN; v->b = true; return v; }
    if (t_.compare(p_, 5, "false") == 0 && !is_token_char(p_ + 5))
    { p_ += 5; v->kind = value::BOOLEAN; v->b = false; return v; }
    return parse_number();
  }

  bool is_token_char(std::size_t q) const
  {
    if (q >= t_.size()) return false;
    char c = t_[q];
    return std::isalnum(static_cast<unsigned char>(c)) || c == '_' || c == '.' || c == '+' || c == '-' || c == ':';
  }

  value_ptr parse_number()
  {
    std::size_t b = p_;
    while (is_token_char(p_)) ++p_;
    if (b == p_) fail("value expected");
    std::string raw = t_.substr(b, p_ - b);
    std::string tok;
    for (char ch : raw) if (ch != '_') tok += ch;
    auto v = std::make_shared<value>();

    std::string body = tok;
    if (!body.empty() && (body[0] == '+' || body[0] == '-')) body = body.substr(1);
    if (body == "inf" || body == "nan")
    {
      v->kind = value::FLOAT;
      v->d = (body == "inf") ? HUGE_VAL : std::nan("");
      if (tok[0] == '-') v->d = -v->d;
      return v;
    }
    if (body.size() > 2 && body[0] == '0' && (body[1] == 'x' || body[1] == 'o' || body[1] == 'b'))
    {
      int base = body[1] == 'x' ? 16 : (body[1] == 'o' ? 8 : 2);
      char* end = nullptr;
      errno = 0;
      unsigned long long u = std::strtoull(body.c_str() + 2, &end, base);
      if (*end != '\0' || errno) fail("bad integer: " + raw);
      v->kind = value::INTEGER;
      v->i = static_cast<std::int64_t>(u);
      return v;
    }
    bool is_float = false;
    for (char ch : body)
    {
      if (ch == '.' || ch == 'e' || ch == 'E') is_float = true;
      else if (ch == ':' ) fail("date/time values are not supported: " + raw);
    }
    if (body.empty() || !std::isdigit(static_cast<unsigned char>(body[0]))) fail("bad value: " + raw);
    char* end = nullptr;
    errno = 0;
    if (is_float)
    {
      double d = std::strtod(tok.c_str(), &end);
      if (*end != '\0') fail("bad float: " + raw);
      v->kind = value::FLOAT;
      v->d = d;
    }
    else
    {
      long long ll = std::strtoll(tok.c_str(), &end, 10);
      if (*end != '\0' || errno) fail("bad integer: " + raw);
      v->kind = value::INTEGER;
      v->i = ll;
    }
    return v;
  }

  value_ptr parse_array()
  {
    expect('[');
    auto v = std::make_shared<value>();
    v->kind = value::ARRAY;
    while (true)
    {
      skip_ws_comments_newlines();
      if (eof()) fail("unterminated array");
      if (peek() == ']') { ++p_; break; }
      v->arr.push_back(parse_value());
      skip_ws_comments_newlines();
      if (eof()) fail("unterminated array");
      if (peek() == ',') { ++p_; continue; }
      if (peek() == ']') { ++p_; break; }
      fail("expected ',' or ']' in array");
    }
    return v;
  }

  value_ptr parse_inline_table()
  {
    expect('{');
    auto v = std::make_shared<value>();
    v->kind = value::TABLE;
    v->defined_inline = true;
    skip_inline_ws();
    if (!eof() && peek() == '}') { ++p_; return v; }
    while (true)
    {
      skip_inline_ws();
      std::vector<std::string> path = parse_key_path();
      skip_inline_ws();
      expect('=');
      skip_inline_ws();
      value_ptr item = parse_value();
      assign(v.get(), path, item);
      skip_inline_ws();
      if (eof()) fail("unterminated inline table");
      if (peek() == ',') { ++p_; continue; }
      if (peek() == '}') { ++p_; break; }
      fail("expected ',' or '}' in inline table");
    }
    return v;
  }
};

inline value_ptr parse_string(const std::string& text)
{
  parser p(text);
  return p.parse();
}

inline value_ptr parse_file(const std::string& path)
{
  std::ifstream in(path, std::ios::binary);
  if (!in) throw parse_error("cannot open file '" + path + "'", 0);
  std::stringstream ss;
  ss << in.rdbuf();
  std::string text = ss.str();
  return parse_string(text);
}

}  // namespace toml_lite

#endif
