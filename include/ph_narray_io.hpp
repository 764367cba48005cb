// ph_narray_io.hpp -- the data formats either side of the hot path for a compiled host:
// ph-core's own `{"shape": [...], "elements": [...]}` JSON / YAML (src/n_array.cr:807-912; goldens
// spec/n_array_spec.cr:520-558), elements in flat lexicographic order, and the binary dump the Python mirror
// writes (`ph-core_b200/io.py`: magic, one JSON header line with shape and numpy dtype string, raw
// row-major little-endian bytes) so checkpoints are interchangeable between the two hosts.
//
// The text <-> (shape, elements) functions are pure host code (namespace Phase::IO::host, usable without a
// GPU); the DeviceNArray wrappers add the explicit D2H / H2D transfer.
#ifndef PH_NARRAY_IO_HPP
#define PH_NARRAY_IO_HPP

#include <cctype>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "ph_narray.hpp"

namespace Phase {
namespace IO {

struct ParseError : std::runtime_error { using std::runtime_error::runtime_error; };   // JSON::ParseException / YAML::ParseException

namespace host {

template <class T> struct NumpyName;
#define PH_NPY_(T, s) template <> struct NumpyName<T> { static const char* value() { return s; } }
PH_NPY_(float, "<f4");   PH_NPY_(double, "<f8");  PH_NPY_(int8_t, "|i1");  PH_NPY_(int16_t, "<i2"); PH_NPY_(int32_t, "<i4");
PH_NPY_(int64_t, "<i8"); PH_NPY_(uint8_t, "|u1"); PH_NPY_(uint16_t, "<u2"); PH_NPY_(uint32_t, "<u4"); PH_NPY_(uint64_t, "<u8");
#undef PH_NPY_

// One element as Crystal's to_json / to_yaml prints it: integers plain, floats as the shortest text that
// round-trips, always with a fraction or exponent ("2.0", not "2").  JSON has no NaN / Infinity: Crystal's
// Float#to_json raises, and so does this.
template <class T>
inline std::string format_element(T v) {
  char buf[64];
  if constexpr (std::is_floating_point<T>::value) {
    // Crystal's Float#to_s (Float::Printer, Crystal 1.0.0 stdlib; restated, the Python mirror's io.format_float
    // applies the same rule): shortest digits that round-trip in T's own width; positional while the decimal
    // point sits in [-3, 15], otherwise d.ddde+X with at least one fraction digit and an unpadded exponent.
    if (!std::isfinite(v)) throw std::invalid_argument("NaN and Infinity cannot be written as JSON / YAML numbers");
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);   // shortest, "d.ddde[+-]XX"
    std::string sci(buf, r.ptr);
    const size_t epos = sci.find('e');
    std::string mant = sci.substr(0, epos);
    const int exp10 = std::stoi(sci.substr(epos + 1));
    std::string sign;
    if (!mant.empty() && mant[0] == '-') { sign = "-"; mant.erase(0, 1); }
    std::string digits;
    for (char c : mant) if (c != '.') digits += c;
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    if (digits == "0") return sign + "0.0";
    const int point = exp10 + 1, n = (int)digits.size();                 // value = 0.DIGITS x 10^point
    if (point > 15 || point < -3) {
      const int e = point - 1;
      return sign + digits.substr(0, 1) + "." + (n > 1 ? digits.substr(1) : std::string("0")) + "e" + (e > 0 ? "+" : "") + std::to_string(e);
    }
    if (point <= 0) return sign + "0." + std::string((size_t)(-point), '0') + digits;
    if (point >= n) return sign + digits + std::string((size_t)(point - n), '0') + ".0";
    return sign + digits.substr(0, (size_t)point) + "." + digits.substr((size_t)point);
  } else {
    auto r = std::to_chars(buf, buf + sizeof(buf), (typename std::conditional<std::is_signed<T>::value, long long, unsigned long long>::type)v);
    return std::string(buf, r.ptr);
  }
}

template <class T>
inline std::string join(const std::vector<T>& v, const char* sep) {
  std::string out;
  for (size_t i = 0; i < v.size(); i++) { if (i) out += sep; out += format_element(v[i]); }
  return out;
}

// NArray#to_json (n_array.cr:807-818): compact separators
template <class T>
inline std::string to_json(const Shape& shape, const std::vector<T>& elements) {
  return std::string("{\"shape\":[") + join(shape, ",") + "],\"elements\":[" + join(elements, ",") + "]}";
}
// NArray#to_yaml (n_array.cr:853-869): document start marker, flow sequences
template <class T>
inline std::string to_yaml(const Shape& shape, const std::vector<T>& elements) {
  return std::string("---\nshape: [") + join(shape, ", ") + "]\nelements: [" + join(elements, ", ") + "]\n";
}

// ---- a scanner for exactly this schema: two keys, each a flat list of numbers (or true / false)
struct Scanner {
  const std::string& s;
  size_t i = 0;
  explicit Scanner(const std::string& text) : s(text) {}
  void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) i++; }
  bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { i++; return true; } return false; }
  void need(char c) { if (!eat(c)) throw ParseError(std::string("expected '") + c + "' at offset " + std::to_string(i)); }
  std::string word() {   // a bare or double-quoted key
    ws();
    std::string out;
    if (i < s.size() && s[i] == '"') {
      for (i++; i < s.size() && s[i] != '"'; i++) out += s[i];
      if (i >= s.size()) throw ParseError("unterminated string");
      i++;
    } else {
      while (i < s.size() && (std::isalnum((unsigned char)s[i]) || s[i] == '_')) out += s[i++];
    }
    return out;
  }
  template <class T> T number() {
    ws();
    if (s.compare(i, 4, "true") == 0) { i += 4; return (T)1; }
    if (s.compare(i, 5, "false") == 0) { i += 5; return (T)0; }
    size_t start = i;
    while (i < s.size() && (std::isdigit((unsigned char)s[i]) || s[i] == '-' || s[i] == '+' || s[i] == '.' || s[i] == 'e' || s[i] == 'E')) i++;
    if (start == i) throw ParseError("expected a number at offset " + std::to_string(i));
    T v{};
    if constexpr (std::is_floating_point<T>::value) {
      auto r = std::from_chars(s.data() + start, s.data() + i, v);
      if (r.ec != std::errc() || r.ptr != s.data() + i) throw ParseError("malformed number '" + s.substr(start, i - start) + "'");
    } else {
      // an integer array may hold "3.0"-style text only if it is integral: parse the integer part strictly
      auto r = std::from_chars(s.data() + start, s.data() + i, v);
      if (r.ec != std::errc() || r.ptr != s.data() + i) throw ParseError("'" + s.substr(start, i - start) + "' is not a valid element of this integer type");
    }
    return v;
  }
  template <class T> std::vector<T> list() {
    std::vector<T> out;
    need('[');
    if (eat(']')) return out;
    do { out.push_back(number<T>()); } while (eat(','));
    need(']');
    return out;
  }
};

template <class T>
inline void finish(bool has_shape, bool has_elements, const Shape& shape, const std::vector<T>& elements, const char* what) {
  if (!has_shape || !has_elements)   // n_array.cr:836-838 / :897-899
    throw ParseError(std::string("Could not read NArray from ") + what + ": 'shape' and/or 'elements' were missing.");
  for (int64_t d : shape) if (d < 0) throw DimensionError("Cannot create NArray: One or more of the provided dimensions was negative.");
  if (shape_to_size(shape) != (int64_t)elements.size())
    throw ShapeError(std::string("Could not read NArray from ") + what + ": " + std::to_string(elements.size()) + " elements for shape " + shape_str(shape));
}

// NArray(T).from_json (n_array.cr:820-851): keys in any order, unknown keys rejected
template <class T>
inline void from_json(const std::string& text, Shape& shape, std::vector<T>& elements) {
  Scanner sc(text);
  bool has_shape = false, has_elements = false;
  sc.need('{');
  if (!sc.eat('}')) {
    do {
      std::string key = sc.word();
      sc.need(':');
      if (key == "shape") { shape = sc.list<int64_t>(); has_shape = true; }
      else if (key == "elements") { elements = sc.list<T>(); has_elements = true; }
      else throw ParseError("unexpected key '" + key + "' in NArray JSON");
    } while (sc.eat(','));
    sc.need('}');
  }
  finish(has_shape, has_elements, shape, elements, "JSON");
}

// NArray(T).from_yaml (n_array.cr:871-912) for the flow form to_yaml writes
template <class T>
inline void from_yaml(const std::string& text, Shape& shape, std::vector<T>& elements) {
  Scanner sc(text);
  bool has_shape = false, has_elements = false;
  sc.ws();
  if (text.compare(sc.i, 3, "---") == 0) sc.i += 3;
  for (;;) {
    sc.ws();
    if (sc.i >= text.size()) break;
    std::string key = sc.word();
    if (key.empty()) throw ParseError("expected a key at offset " + std::to_string(sc.i));
    sc.need(':');
    if (key == "shape") { shape = sc.list<int64_t>(); has_shape = true; }
    else if (key == "elements") { elements = sc.list<T>(); has_elements = true; }
    else throw ParseError("unexpected key '" + key + "' in NArray YAML");
  }
  finish(has_shape, has_elements, shape, elements, "YAML");
}

// ---- binary dump: "PHNARR1\n", {"shape": [..], "dtype": "<f4"}\n, raw bytes
static const char kMagic[] = "PHNARR1\n";

template <class T>
inline void dump(const std::string& path, const Shape& shape, const std::vector<T>& elements) {
  std::ofstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path + " for writing");
  std::string header = std::string("{\"shape\": [") + join(shape, ", ") + "], \"dtype\": \"" + NumpyName<T>::value() + "\"}\n";
  f.write(kMagic, 8);
  f.write(header.data(), (std::streamsize)header.size());
  if (!elements.empty()) f.write(reinterpret_cast<const char*>(elements.data()), (std::streamsize)(elements.size() * sizeof(T)));
  if (!f) throw std::runtime_error("short write to " + path);
}

template <class T>
inline void load(const std::string& path, Shape& shape, std::vector<T>& elements) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  char magic[8];
  f.read(magic, 8);
  if (!f || std::memcmp(magic, kMagic, 8) != 0) throw ParseError(path + " is not a ph-core binary dump");
  std::string header;
  std::getline(f, header);
  // the header is JSON with a string-valued "dtype": scan it by hand
  size_t sp = header.find("\"shape\""), dp = header.find("\"dtype\"");
  if (sp == std::string::npos || dp == std::string::npos) throw ParseError("binary dump header lacks shape / dtype");
  Scanner sc(header);
  sc.i = header.find('[', sp);
  if (sc.i == std::string::npos) throw ParseError("binary dump header lacks a shape list");
  shape = sc.list<int64_t>();
  size_t q0 = header.find('"', header.find(':', dp)), q1 = header.find('"', q0 + 1);
  std::string dtype = header.substr(q0 + 1, q1 - q0 - 1);
  std::string want = NumpyName<T>::value();
  // numpy writes "|b1" for Bool and may use '=' for native order; one-byte types carry '|'
  if (dtype != want && !(sizeof(T) == 1 && dtype == "|b1" && want == "|u1") && !(dtype.size() == 3 && dtype[0] == '=' && dtype.substr(1) == want.substr(1)))
    throw ParseError("binary dump holds dtype " + dtype + ", expected " + want);
  const int64_t n = shape_to_size(shape);
  elements.resize((size_t)n);
  if (n) f.read(reinterpret_cast<char*>(elements.data()), (std::streamsize)((size_t)n * sizeof(T)));
  if (f.gcount() != (std::streamsize)((size_t)n * sizeof(T)) && n) throw ShapeError("binary dump is shorter than its header's shape " + shape_str(shape));
  f.peek();
  if (!f.eof()) throw ShapeError("binary dump is longer than its header's shape " + shape_str(shape));
}

}  // namespace host

// ---- DeviceNArray wrappers: the transfer is explicit in the name of the game (to_host / from_host inside)
template <class T> inline std::string to_json(const MultiIndexable<T>& a) { return host::to_json(a.shape(), a.to_host()); }
template <class T> inline std::string to_yaml(const MultiIndexable<T>& a) { return host::to_yaml(a.shape(), a.to_host()); }
template <class T> inline DeviceNArray<T> from_json(const std::string& text) {
  Shape shape; std::vector<T> el;
  host::from_json(text, shape, el);
  return DeviceNArray<T>::from_host(shape, el);
}
template <class T> inline DeviceNArray<T> from_yaml(const std::string& text) {
  Shape shape; std::vector<T> el;
  host::from_yaml(text, shape, el);
  return DeviceNArray<T>::from_host(shape, el);
}
template <class T> inline void dump(const MultiIndexable<T>& a, const std::string& path) { host::dump(path, a.shape(), a.to_host()); }
template <class T> inline DeviceNArray<T> load(const std::string& path) {
  Shape shape; std::vector<T> el;
  host::load(path, shape, el);
  return DeviceNArray<T>::from_host(shape, el);
}

}  // namespace IO
}  // namespace Phase
#endif  // PH_NARRAY_IO_HPP
