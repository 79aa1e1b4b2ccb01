// Minimal stand-in for the header-only serialisation library cereal (USCiLab/cereal v1.2.2),
// which the reference fetches at configure time and which is absent from this image.
// TEST INFRASTRUCTURE ONLY: used solely to compile the unmodified reference sources into
// oracle/_ref/ (see oracle/build_ref.sh). It implements only the subset the reference's
// index (de)serialisation touches:
//   * portable-binary layout: arithmetic = raw little-endian bytes, containers = u64 count + items
//   * JSON: name/value pairs of string / bool / integer / enum / vector<string>, top-level "value0"
// Written from cereal's documented on-disk format, not from its sources.
#pragma once
#include <cstdint>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

namespace cereal {

struct Exception : public std::runtime_error {
  explicit Exception(const std::string& w) : std::runtime_error(w) {}
};

template <class T> struct NameValuePair {
  const char* name;
  T& value;
};
template <class T> inline NameValuePair<T> make_nvp(const char* n, T& v) { return {n, v}; }
template <class T> inline NameValuePair<const T> make_nvp(const char* n, const T& v) { return {n, v}; }
template <class T> inline NameValuePair<T> make_nvp(const std::string& n, T& v) { return {n.c_str(), v}; }

struct SizeTag { uint64_t size; uint64_t* target; };
inline SizeTag make_size_tag(const size_t& s) { return {static_cast<uint64_t>(s), nullptr}; }
inline SizeTag make_size_tag(size_t& s) { return {static_cast<uint64_t>(s), reinterpret_cast<uint64_t*>(&s)}; }

struct BinaryData { void* data; uint64_t size; };
template <class T> inline BinaryData binary_data(T* p, size_t n) {
  return {const_cast<void*>(static_cast<const void*>(p)), static_cast<uint64_t>(n)};
}

namespace detail {
template <class...> using void_t = void;
template <class T, class A, class = void> struct has_member_save : std::false_type {};
template <class T, class A>
struct has_member_save<T, A, void_t<decltype(std::declval<const T&>().save(std::declval<A&>()))>> : std::true_type {};
template <class T, class A, class = void> struct has_member_load : std::false_type {};
template <class T, class A>
struct has_member_load<T, A, void_t<decltype(std::declval<T&>().load(std::declval<A&>()))>> : std::true_type {};
template <class T, class A, class = void> struct has_member_serialize : std::false_type {};
template <class T, class A>
struct has_member_serialize<T, A, void_t<decltype(std::declval<T&>().serialize(std::declval<A&>()))>> : std::true_type {};
} // namespace detail

// ----------------------------------------------------------------------------------------------
// Binary archives
// ----------------------------------------------------------------------------------------------
class BinaryOutputArchive {
public:
  explicit BinaryOutputArchive(std::ostream& os) : os_(os) {}
  template <class... Ts> BinaryOutputArchive& operator()(Ts&&... vs) {
    int dummy[] = {0, (put(vs), 0)...};
    (void)dummy;
    return *this;
  }
  void saveBinary(const void* p, size_t n) {
    os_.write(static_cast<const char*>(p), static_cast<std::streamsize>(n));
    if (!os_) throw Exception("Failed to write " + std::to_string(n) + " bytes to output stream!");
  }

private:
  template <class T> typename std::enable_if<std::is_arithmetic<T>::value>::type put(const T& v) { saveBinary(&v, sizeof(T)); }
  template <class T> typename std::enable_if<std::is_enum<T>::value>::type put(const T& v) {
    auto u = static_cast<typename std::underlying_type<T>::type>(v);
    saveBinary(&u, sizeof(u));
  }
  void put(const SizeTag& s) { saveBinary(&s.size, 8); }
  void put(const BinaryData& b) { saveBinary(b.data, b.size); }
  template <class T> void put(const NameValuePair<T>& nv) { put(nv.value); }
  void put(const std::string& s) {
    uint64_t n = s.size();
    saveBinary(&n, 8);
    saveBinary(s.data(), n);
  }
  template <class T, class A> void put(const std::vector<T, A>& v) { putVec(v, std::is_arithmetic<T>()); }
  template <class T, class A> void putVec(const std::vector<T, A>& v, std::true_type) {
    uint64_t n = v.size();
    saveBinary(&n, 8);
    saveBinary(v.data(), n * sizeof(T));
  }
  template <class T, class A> void putVec(const std::vector<T, A>& v, std::false_type) {
    uint64_t n = v.size();
    saveBinary(&n, 8);
    for (const auto& e : v) put(e);
  }
  template <class A, class B> void put(const std::pair<A, B>& p) { put(p.first); put(p.second); }
  template <class K, class V, class H, class E, class Al> void put(const std::unordered_map<K, V, H, E, Al>& m) {
    uint64_t n = m.size();
    saveBinary(&n, 8);
    for (const auto& kv : m) { put(kv.first); put(kv.second); }
  }
  template <class T>
  typename std::enable_if<std::is_class<T>::value && detail::has_member_save<T, BinaryOutputArchive>::value>::type
  put(const T& v) { v.save(*this); }
  template <class T>
  typename std::enable_if<std::is_class<T>::value && !detail::has_member_save<T, BinaryOutputArchive>::value &&
                          detail::has_member_serialize<T, BinaryOutputArchive>::value>::type
  put(const T& v) { const_cast<T&>(v).serialize(*this); }
  std::ostream& os_;
};

class BinaryInputArchive {
public:
  explicit BinaryInputArchive(std::istream& is) : is_(is) {}
  template <class... Ts> BinaryInputArchive& operator()(Ts&&... vs) {
    int dummy[] = {0, (get(vs), 0)...};
    (void)dummy;
    return *this;
  }
  void loadBinary(void* p, size_t n) {
    is_.read(static_cast<char*>(p), static_cast<std::streamsize>(n));
    if (static_cast<size_t>(is_.gcount()) != n)
      throw Exception("Failed to read " + std::to_string(n) + " bytes from input stream! Read " +
                      std::to_string(is_.gcount()));
  }

private:
  template <class T> typename std::enable_if<std::is_arithmetic<T>::value>::type get(T& v) { loadBinary(&v, sizeof(T)); }
  template <class T> typename std::enable_if<std::is_enum<T>::value>::type get(T& v) {
    typename std::underlying_type<T>::type u;
    loadBinary(&u, sizeof(u));
    v = static_cast<T>(u);
  }
  void get(SizeTag& s) { loadBinary(&s.size, 8); if (s.target) *s.target = s.size; }
  void get(SizeTag&& s) { get(s); }
  void get(BinaryData& b) { loadBinary(b.data, b.size); }
  void get(BinaryData&& b) { loadBinary(b.data, b.size); }
  template <class T> void get(NameValuePair<T>& nv) { get(nv.value); }
  template <class T> void get(NameValuePair<T>&& nv) { get(nv.value); }
  void get(std::string& s) {
    uint64_t n;
    loadBinary(&n, 8);
    s.resize(n);
    if (n) loadBinary(&s[0], n);
  }
  template <class T, class A> void get(std::vector<T, A>& v) { getVec(v, std::is_arithmetic<T>()); }
  template <class T, class A> void getVec(std::vector<T, A>& v, std::true_type) {
    uint64_t n;
    loadBinary(&n, 8);
    v.resize(n);
    if (n) loadBinary(v.data(), n * sizeof(T));
  }
  template <class T, class A> void getVec(std::vector<T, A>& v, std::false_type) {
    uint64_t n;
    loadBinary(&n, 8);
    v.resize(n);
    for (auto& e : v) get(e);
  }
  template <class A, class B> void get(std::pair<A, B>& p) { get(p.first); get(p.second); }
  template <class K, class V, class H, class E, class Al> void get(std::unordered_map<K, V, H, E, Al>& m) {
    uint64_t n;
    loadBinary(&n, 8);
    m.clear();
    for (uint64_t i = 0; i < n; ++i) { K k; V v; get(k); get(v); m.emplace(std::move(k), std::move(v)); }
  }
  template <class T>
  typename std::enable_if<std::is_class<T>::value && detail::has_member_load<T, BinaryInputArchive>::value>::type
  get(T& v) { v.load(*this); }
  template <class T>
  typename std::enable_if<std::is_class<T>::value && !detail::has_member_load<T, BinaryInputArchive>::value &&
                          detail::has_member_serialize<T, BinaryInputArchive>::value>::type
  get(T& v) { v.serialize(*this); }
  std::istream& is_;
};

// ----------------------------------------------------------------------------------------------
// JSON archives (small subset)
// ----------------------------------------------------------------------------------------------
namespace json_detail {
struct Value {
  enum Kind { NUL, BOOL, NUM, STR, ARR, OBJ } kind{NUL};
  bool b{false};
  std::string s; // string payload or number text
  std::vector<Value> arr;
  std::vector<std::pair<std::string, Value>> obj;
  const Value* find(const std::string& k) const {
    for (auto& kv : obj) if (kv.first == k) return &kv.second;
    return nullptr;
  }
};
class Parser {
public:
  explicit Parser(const std::string& t) : t_(t) {}
  Value parse() { ws(); Value v = value(); return v; }
private:
  void ws() { while (p_ < t_.size() && (t_[p_] == ' ' || t_[p_] == '\n' || t_[p_] == '\t' || t_[p_] == '\r')) ++p_; }
  char peek() { if (p_ >= t_.size()) throw Exception("JSON: unexpected end of input"); return t_[p_]; }
  void expect(char c) { if (peek() != c) throw Exception(std::string("JSON: expected '") + c + "'"); ++p_; }
  std::string str() {
    expect('"');
    std::string o;
    while (peek() != '"') {
      char c = t_[p_++];
      if (c == '\\') {
        char e = peek(); ++p_;
        switch (e) {
          case 'n': o += '\n'; break; case 't': o += '\t'; break; case 'r': o += '\r'; break;
          case 'b': o += '\b'; break; case 'f': o += '\f'; break;
          case 'u': { unsigned cp = std::stoul(t_.substr(p_, 4), nullptr, 16); p_ += 4; o += static_cast<char>(cp); break; }
          default: o += e;
        }
      } else o += c;
    }
    ++p_;
    return o;
  }
  Value value() {
    ws();
    Value v;
    char c = peek();
    if (c == '{') {
      v.kind = Value::OBJ; ++p_; ws();
      if (peek() == '}') { ++p_; return v; }
      while (true) {
        ws(); std::string k = str(); ws(); expect(':');
        Value e = value(); v.obj.emplace_back(std::move(k), std::move(e)); ws();
        if (peek() == ',') { ++p_; continue; }
        expect('}'); break;
      }
    } else if (c == '[') {
      v.kind = Value::ARR; ++p_; ws();
      if (peek() == ']') { ++p_; return v; }
      while (true) {
        v.arr.push_back(value()); ws();
        if (peek() == ',') { ++p_; continue; }
        expect(']'); break;
      }
    } else if (c == '"') { v.kind = Value::STR; v.s = str(); }
    else if (t_.compare(p_, 4, "true") == 0) { v.kind = Value::BOOL; v.b = true; p_ += 4; }
    else if (t_.compare(p_, 5, "false") == 0) { v.kind = Value::BOOL; v.b = false; p_ += 5; }
    else if (t_.compare(p_, 4, "null") == 0) { p_ += 4; }
    else {
      v.kind = Value::NUM;
      size_t b = p_;
      while (p_ < t_.size() && (std::isdigit(static_cast<unsigned char>(t_[p_])) || t_[p_] == '-' || t_[p_] == '+' ||
                                t_[p_] == '.' || t_[p_] == 'e' || t_[p_] == 'E')) ++p_;
      if (b == p_) throw Exception("JSON: bad value");
      v.s = t_.substr(b, p_ - b);
    }
    return v;
  }
  const std::string& t_;
  size_t p_{0};
};
inline std::string escape(const std::string& s) {
  std::string o;
  for (char c : s) {
    switch (c) {
      case '"': o += "\\\""; break; case '\\': o += "\\\\"; break; case '\n': o += "\\n"; break;
      case '\t': o += "\\t"; break; case '\r': o += "\\r"; break; default: o += c;
    }
  }
  return o;
}
} // namespace json_detail

class JSONOutputArchive {
public:
  explicit JSONOutputArchive(std::ostream& os) : os_(os) { os_ << "{"; first_.push_back(true); counter_.push_back(0); }
  ~JSONOutputArchive() { os_ << "\n}"; os_.flush(); }
  template <class... Ts> JSONOutputArchive& operator()(Ts&&... vs) {
    int dummy[] = {0, (put(vs), 0)...};
    (void)dummy;
    return *this;
  }
private:
  void indent() { os_ << "\n"; for (size_t i = 0; i < first_.size(); ++i) os_ << "    "; }
  void key(const char* name) {
    if (!first_.back()) os_ << ",";
    first_.back() = false;
    indent();
    if (name) os_ << "\"" << json_detail::escape(name) << "\": ";
    else os_ << "\"value" << counter_.back()++ << "\": ";
  }
  template <class T> void put(const NameValuePair<T>& nv) { key(nv.name); emit(nv.value); }
  template <class T> void put(const T& v) { key(nullptr); emit(v); }
  template <class T> typename std::enable_if<std::is_integral<T>::value && !std::is_same<T, bool>::value>::type
  emit(const T& v) { os_ << +v; }
  template <class T> typename std::enable_if<std::is_floating_point<T>::value>::type emit(const T& v) { os_ << v; }
  template <class T> typename std::enable_if<std::is_enum<T>::value>::type emit(const T& v) {
    os_ << +static_cast<typename std::underlying_type<T>::type>(v);
  }
  void emit(const bool& v) { os_ << (v ? "true" : "false"); }
  void emit(const std::string& v) { os_ << "\"" << json_detail::escape(v) << "\""; }
  template <class T, class A> void emit(const std::vector<T, A>& v) {
    os_ << "[";
    first_.push_back(true); counter_.push_back(0);
    for (auto& e : v) { if (!first_.back()) os_ << ","; first_.back() = false; indent(); emit(e); }
    first_.pop_back(); counter_.pop_back();
    if (!v.empty()) indent();
    os_ << "]";
  }
  template <class T> typename std::enable_if<std::is_class<T>::value>::type emit(const T& v) {
    os_ << "{";
    first_.push_back(true); counter_.push_back(0);
    v.save(*this);
    first_.pop_back(); counter_.pop_back();
    indent();
    os_ << "}";
  }
  std::ostream& os_;
  std::vector<bool> first_;
  std::vector<int> counter_;
};

class JSONInputArchive {
public:
  explicit JSONInputArchive(std::istream& is) {
    std::stringstream ss; ss << is.rdbuf();
    text_ = ss.str();
    root_ = json_detail::Parser(text_).parse();
    if (root_.kind != json_detail::Value::OBJ) throw Exception("JSON: document root is not an object");
    stack_.push_back(&root_); counter_.push_back(0);
  }
  template <class... Ts> JSONInputArchive& operator()(Ts&&... vs) {
    int dummy[] = {0, (get(vs), 0)...};
    (void)dummy;
    return *this;
  }
private:
  const json_detail::Value& lookup(const char* name) {
    std::string k = name ? std::string(name) : ("value" + std::to_string(counter_.back()++));
    const json_detail::Value* v = stack_.back()->find(k);
    if (!v) throw Exception("JSON Parsing failed - provided NVP (" + k + ") not found");
    return *v;
  }
  template <class T> void get(NameValuePair<T>& nv) { read(lookup(nv.name), nv.value); }
  template <class T> void get(NameValuePair<T>&& nv) { read(lookup(nv.name), nv.value); }
  template <class T> void get(T& v) { read(lookup(nullptr), v); }
  template <class T> typename std::enable_if<std::is_integral<T>::value && !std::is_same<T, bool>::value>::type
  read(const json_detail::Value& j, T& v) {
    if (j.kind != json_detail::Value::NUM) throw Exception("JSON: expected number");
    v = static_cast<T>(std::stoll(j.s));
  }
  template <class T> typename std::enable_if<std::is_floating_point<T>::value>::type
  read(const json_detail::Value& j, T& v) { v = static_cast<T>(std::stod(j.s)); }
  template <class T> typename std::enable_if<std::is_enum<T>::value>::type read(const json_detail::Value& j, T& v) {
    v = static_cast<T>(std::stoll(j.s));
  }
  void read(const json_detail::Value& j, bool& v) {
    if (j.kind != json_detail::Value::BOOL) throw Exception("JSON: expected bool");
    v = j.b;
  }
  void read(const json_detail::Value& j, std::string& v) {
    if (j.kind != json_detail::Value::STR) throw Exception("JSON: expected string");
    v = j.s;
  }
  template <class T, class A> void read(const json_detail::Value& j, std::vector<T, A>& v) {
    if (j.kind != json_detail::Value::ARR) throw Exception("JSON: expected array");
    v.resize(j.arr.size());
    for (size_t i = 0; i < j.arr.size(); ++i) read(j.arr[i], v[i]);
  }
  template <class T> typename std::enable_if<std::is_class<T>::value>::type read(const json_detail::Value& j, T& v) {
    if (j.kind != json_detail::Value::OBJ) throw Exception("JSON: expected object");
    stack_.push_back(&j); counter_.push_back(0);
    v.load(*this);
    stack_.pop_back(); counter_.pop_back();
  }
  std::string text_;
  json_detail::Value root_;
  std::vector<const json_detail::Value*> stack_;
  std::vector<int> counter_;
};

} // namespace cereal
