#include "service_env.h"

#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace tepdist {

ServiceEnv::ServiceEnv() {
#define X(name, def, help)   \
  values_[#name] = def;      \
  help_[#name] = help;       \
  order_.push_back(#name);
  TEPDIST_SERVICE_OPTIONS(X)
#undef X
}

ServiceEnv* ServiceEnv::Instance() {
  static ServiceEnv env;
  return &env;
}

namespace {
// minimal flat-JSON reader: {"KEY": value, ...} with string / number / bool values
std::map<std::string, std::string> ParseFlatJson(const std::string& s) {
  std::map<std::string, std::string> out;
  size_t i = 0;
  auto skip = [&] { while (i < s.size() && (isspace((unsigned char)s[i]) || s[i] == ',' || s[i] == '{' || s[i] == '}')) ++i; };
  while (true) {
    skip();
    if (i >= s.size() || s[i] != '"') break;
    size_t e = s.find('"', i + 1);
    if (e == std::string::npos) break;
    std::string key = s.substr(i + 1, e - i - 1);
    i = s.find(':', e);
    if (i == std::string::npos) break;
    ++i;
    while (i < s.size() && isspace((unsigned char)s[i])) ++i;
    std::string val;
    if (i < s.size() && s[i] == '"') {
      size_t e2 = s.find('"', i + 1);
      val = s.substr(i + 1, e2 - i - 1);
      i = e2 + 1;
    } else {
      size_t e2 = i;
      while (e2 < s.size() && s[e2] != ',' && s[e2] != '}' && !isspace((unsigned char)s[e2])) ++e2;
      val = s.substr(i, e2 - i);
      i = e2;
    }
    out[key] = val;
  }
  return out;
}
}  // namespace

std::vector<std::string> ServiceEnv::Load(const std::string& config_file) {
  std::vector<std::string> warnings;
  // every load starts from the defaults: a key that was set by an earlier load (or Set) and is no longer in the file or
  // the environment must not survive a reload
#define X(name, def, help) values_[#name] = def;
  TEPDIST_SERVICE_OPTIONS(X)
#undef X
  std::string path = config_file;
  if (path.empty()) {
    const char* e = std::getenv("CONFIG_FILE");
    path = e ? e : "config.json";
  }
  std::ifstream f(path);
  if (f) {
    std::stringstream ss;
    ss << f.rdbuf();
    for (auto& kv : ParseFlatJson(ss.str())) {
      if (values_.count(kv.first)) values_[kv.first] = kv.second;
      else warnings.push_back("unknown option in " + path + ": " + kv.first);
    }
  }
  for (auto& k : order_) {
    const char* e = std::getenv(k.c_str());
    if (e) {
      if (values_[k] != e) warnings.push_back("env overrides " + k + ": " + values_[k] + " -> " + e);
      values_[k] = e;
    }
  }
  return warnings;
}

std::string ServiceEnv::Get(const std::string& key) const {
  auto it = values_.find(key);
  if (it == values_.end()) throw std::out_of_range("unknown ServiceEnv option " + key);
  return it->second;
}
long long ServiceEnv::GetInt(const std::string& key) const { return std::stoll(Get(key)); }
double ServiceEnv::GetDouble(const std::string& key) const { return std::stod(Get(key)); }
bool ServiceEnv::GetBool(const std::string& key) const {
  std::string v = Get(key);
  return v == "1" || v == "true" || v == "True" || v == "TRUE" || v == "on";
}
void ServiceEnv::Set(const std::string& key, const std::string& value) {
  if (!values_.count(key)) throw std::out_of_range("unknown ServiceEnv option " + key);
  values_[key] = value;
}
std::vector<std::string> ServiceEnv::Keys() const { return order_; }
std::string ServiceEnv::Dump() const {
  std::ostringstream o;
  for (auto& k : order_) o << k << "=" << values_.at(k) << "    # " << help_.at(k) << "\n";
  return o.str();
}

}  // namespace tepdist
