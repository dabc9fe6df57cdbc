// oracle/eigen_standin/yaml-cpp/yaml.h -- TEST INFRASTRUCTURE, not yaml-cpp.
// yaml-cpp is absent from this image.  The reference touches it in one place, TargetManager::loadYamlFile
// (src/target_manager.cpp:18-104): YAML::LoadFile, node[key].as<std::vector<double>>(), node["type"].as<std::string>() and
// catch (YAML::ParserException&).  This stand-in reads the flat `key: value` / `key: [a, b, ...]` documents of models/*.yaml.
#pragma once
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace YAML {

class Exception : public std::runtime_error {
 public:
  explicit Exception(const std::string& m) : std::runtime_error(m) {}
};
class ParserException : public Exception { public: using Exception::Exception; };
class BadFile : public Exception { public: using Exception::Exception; };
class BadConversion : public Exception { public: using Exception::Exception; };
class InvalidNode : public Exception { public: using Exception::Exception; };

class Node {
 public:
  Node() : defined_(true) {}
  const Node operator[](const std::string& key) const {
    auto it = map_.find(key);
    if (it == map_.end()) { Node n; n.defined_ = false; return n; }
    Node n;
    n.scalar_ = it->second;
    return n;
  }
  template <class T> T as() const;
  bool IsDefined() const { return defined_; }
  std::map<std::string, std::string> map_;
  std::string scalar_;
  bool defined_;
};
template <> inline std::string Node::as<std::string>() const {
  if (!defined_) throw InvalidNode("invalid node; this may result from using a map iterator as a sequence iterator, or vice-versa");
  std::string s = scalar_;
  while (!s.empty() && (s.back() == ' ' || s.back() == '\r' || s.back() == '"' || s.back() == '\'')) s.pop_back();
  size_t b = 0;
  while (b < s.size() && (s[b] == ' ' || s[b] == '"' || s[b] == '\'')) ++b;
  return s.substr(b);
}
template <> inline std::vector<double> Node::as<std::vector<double>>() const {
  if (!defined_) throw InvalidNode("invalid node; this may result from using a map iterator as a sequence iterator, or vice-versa");
  std::vector<double> v;
  std::string s = scalar_;
  for (char& c : s)
    if (c == '[' || c == ']' || c == ',') c = ' ';
  std::istringstream is(s);
  std::string tok;
  while (is >> tok) {
    char* end = nullptr;
    const double d = std::strtod(tok.c_str(), &end);
    if (end == tok.c_str() || *end != '\0') throw BadConversion("bad conversion");
    v.push_back(d);
  }
  return v;
}
inline Node LoadFile(const std::string& path) {
  std::ifstream f(path.c_str());
  if (!f.is_open()) throw BadFile("bad file: " + path);
  Node n;
  std::string line, key;
  while (std::getline(f, line)) {
    const size_t hash = line.find('#');
    if (hash != std::string::npos) line = line.substr(0, hash);
    const size_t colon = line.find(':');
    if (colon != std::string::npos && line.find_first_not_of(" \t") != std::string::npos && line[line.find_first_not_of(" \t")] != '-' &&
        line.substr(0, colon).find('[') == std::string::npos) {
      key = line.substr(line.find_first_not_of(" \t"), colon - line.find_first_not_of(" \t"));
      n.map_[key] = line.substr(colon + 1);
    } else if (!key.empty()) {
      n.map_[key] += " " + line;   // a flow sequence continued on the next line
    }
  }
  return n;
}

}  // namespace YAML
