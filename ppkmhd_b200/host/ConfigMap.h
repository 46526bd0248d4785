// Settings contract of the reference: an .ini file read into a case-insensitive "section.name" ->
// string map with typed getters.  Mirrors ConfigMap / INIReader of the reference
// (src/utils/config/ConfigMap.h:21-52, src/utils/config/inih/INIReader.cpp:43-106, inih/ini.cpp)
// including its quirks: getFloat returns *float* (strtof), inline comments start with " ;",
// indented lines continue the previous key, names/sections are truncated at 49 characters.
#pragma once
#include <map>
#include <string>

class ConfigMap {
public:
  ConfigMap() = default;
  explicit ConfigMap(const std::string &filename);  // ConfigMap.cpp:21-23
  ConfigMap(const char *buffer, int buffer_size);   // ConfigMap.cpp:27-29 (the broadcast path)

  int ParseError() const { return error_; }  // 0 ok, -1 cannot open, else first bad line

  std::string getString(const std::string &section, const std::string &name, const std::string &default_value) const;
  void setString(const std::string &section, const std::string &name, const std::string &value);
  long getInteger(const std::string &section, const std::string &name, long default_value) const;
  void setInteger(const std::string &section, const std::string &name, long value);
  float getFloat(const std::string &section, const std::string &name, float default_value) const;
  void setFloat(const std::string &section, const std::string &name, float value);
  bool getBool(const std::string &section, const std::string &name, bool default_value) const;
  void setBool(const std::string &section, const std::string &name, bool value);

  const std::map<std::string, std::string> &values() const { return values_; }

private:
  void parse_stream(std::istream &in, bool clip_lines);
  static std::string make_key(const std::string &section, const std::string &name);
  std::map<std::string, std::string> values_;
  int error_ = 0;
};

// rank 0 reads, everybody gets the same map (src/utils/config/ConfigMap.cpp:99-154). In this
// single-node build every process reads the file itself.
ConfigMap broadcast_parameters(const std::string &filename);
