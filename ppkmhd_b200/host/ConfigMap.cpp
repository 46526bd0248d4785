#include "ConfigMap.h"

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <vector>

namespace {
const size_t kMaxLine = 200, kMaxSection = 50, kMaxName = 50;  // inih/ini.cpp:15-17

std::string rstrip(std::string s) {
  while (!s.empty() && isspace((unsigned char)s.back())) s.pop_back();
  return s;
}
std::string lskip(const std::string &s) {
  size_t p = 0;
  while (p < s.size() && isspace((unsigned char)s[p])) ++p;
  return s.substr(p);
}
// position of `c`, or of a ';' that follows whitespace, or npos (inih find_char_or_comment)
size_t find_char_or_comment(const std::string &s, char c, size_t from = 0) {
  bool was_space = false;
  for (size_t p = from; p < s.size(); ++p) {
    if (s[p] == c || (was_space && s[p] == ';')) return p;
    was_space = isspace((unsigned char)s[p]) != 0;
  }
  return std::string::npos;
}
}  // namespace

ConfigMap::ConfigMap(const std::string &filename) {
  std::ifstream in(filename.c_str());
  if (!in) {
    error_ = -1;
    return;
  }
  parse_stream(in, true);
}

ConfigMap::ConfigMap(const char *buffer, int buffer_size) {
  std::istringstream in(std::string(buffer, buffer + (buffer_size > 0 ? buffer_size : 0)));
  parse_stream(in, false);
}

void ConfigMap::parse_stream(std::istream &in, bool clip_lines) {
  std::string section, prev_name, raw;
  int lineno = 0;
  while (std::getline(in, raw)) {
    // the file reader of inih uses fgets on a 200-byte buffer: longer lines arrive in pieces
    std::vector<std::string> pieces;
    if (clip_lines && raw.size() >= kMaxLine) {
      for (size_t p = 0; p < raw.size(); p += kMaxLine - 1) pieces.push_back(raw.substr(p, kMaxLine - 1));
    } else {
      pieces.push_back(raw);
    }
    for (const std::string &piece : pieces) {
      ++lineno;
      const std::string line = rstrip(piece);
      const std::string start = lskip(line);
      const bool indented = !start.empty() && start.size() < line.size();
      if (!prev_name.empty() && indented) {
        values_[make_key(section, prev_name)] = start;  // continuation line overwrites the value
      } else if (start.empty() || start[0] == ';' || start[0] == '#') {
        // blank or comment
      } else if (start[0] == '[') {
        const size_t end = find_char_or_comment(start, ']', 1);
        if (end != std::string::npos && start[end] == ']') {
          section = start.substr(1, end - 1).substr(0, kMaxSection - 1);
          prev_name.clear();
        } else if (!error_) {
          error_ = lineno;
        }
      } else {
        const size_t eq = find_char_or_comment(start, '=');
        if (eq != std::string::npos && start[eq] == '=') {
          const std::string name = rstrip(start.substr(0, eq));
          std::string value = lskip(start.substr(eq + 1));
          const size_t cm = find_char_or_comment(value, '\0');
          if (cm != std::string::npos && value[cm] == ';') value = value.substr(0, cm);
          value = rstrip(value);
          prev_name = name.substr(0, kMaxName - 1);
          values_[make_key(section, name)] = value;
        } else if (!error_) {
          error_ = lineno;
        }
      }
    }
  }
}

std::string ConfigMap::make_key(const std::string &section, const std::string &name) {
  std::string key = section + "." + name;
  for (char &ch : key) ch = (char)tolower((unsigned char)ch);
  return key;
}

std::string ConfigMap::getString(const std::string &section, const std::string &name, const std::string &dflt) const {
  auto it = values_.find(make_key(section, name));
  return it == values_.end() ? dflt : it->second;
}
void ConfigMap::setString(const std::string &section, const std::string &name, const std::string &value) {
  values_[make_key(section, name)] = value;
}
long ConfigMap::getInteger(const std::string &section, const std::string &name, long dflt) const {
  const std::string v = getString(section, name, "");
  char *end = nullptr;
  const long n = strtol(v.c_str(), &end, 0);  // decimal, hex, octal
  return end > v.c_str() ? n : dflt;
}
void ConfigMap::setInteger(const std::string &section, const std::string &name, long value) {
  setString(section, name, std::to_string(value));
}
float ConfigMap::getFloat(const std::string &section, const std::string &name, float dflt) const {
  const std::string v = getString(section, name, "");
  char *end = nullptr;
  const float f = strtof(v.c_str(), &end);  // float precision on purpose (ConfigMap.cpp:37-46)
  return end > v.c_str() ? f : dflt;
}
void ConfigMap::setFloat(const std::string &section, const std::string &name, float value) {
  std::ostringstream ss;
  ss << value;
  setString(section, name, ss.str());
}
bool ConfigMap::getBool(const std::string &section, const std::string &name, bool dflt) const {
  const std::string v = getString(section, name, "");
  if (v == "1" || v == "yes" || v == "true" || v == "on") return true;
  if (v == "0" || v == "no" || v == "false" || v == "off") return false;
  return dflt;
}
void ConfigMap::setBool(const std::string &section, const std::string &name, bool value) {
  setString(section, name, value ? "true" : "false");
}

ConfigMap broadcast_parameters(const std::string &filename) { return ConfigMap(filename); }
