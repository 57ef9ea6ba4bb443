// flags.hpp -- minimal stand-in for gflags (the reference's tools use
// DEFINE_int32 / DEFINE_uint64 / DEFINE_string, tool_query.cpp:26-36): accepts
// --name value, --name=value and -name value; --help lists the flags.
#ifndef PQT_B200_HOST_FLAGS_HPP
#define PQT_B200_HOST_FLAGS_HPP

#include <cstdlib>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

class Flags {
 public:
  void add(const std::string& name, const std::string& def, const std::string& help) {
    order_.push_back(name);
    vals_[name] = def;
    help_[name] = help;
  }
  // returns false if --help was requested
  bool parse(int argc, char** argv, const std::string& usage) {
    for (int i = 1; i < argc; i++) {
      std::string a = argv[i];
      if (a == "--help" || a == "-help" || a == "-h") {
        std::cout << usage << "\nflags:\n";
        for (const auto& n : order_)
          std::cout << "  --" << n << " (" << help_[n] << ") default: " << vals_[n] << "\n";
        return false;
      }
      if (a.size() < 2 || a[0] != '-') throw std::runtime_error("unexpected argument " + a);
      a = a.substr(a[1] == '-' ? 2 : 1);
      std::string v;
      size_t eq = a.find('=');
      if (eq != std::string::npos) {
        v = a.substr(eq + 1);
        a = a.substr(0, eq);
      } else {
        if (i + 1 >= argc) throw std::runtime_error("flag --" + a + " needs a value");
        v = argv[++i];
      }
      if (!vals_.count(a)) throw std::runtime_error("unknown flag --" + a);
      vals_[a] = v;
    }
    return true;
  }
  std::string str(const std::string& n) const { return vals_.at(n); }
  long long num(const std::string& n) const { return std::atoll(vals_.at(n).c_str()); }

 private:
  std::vector<std::string> order_;
  std::map<std::string, std::string> vals_, help_;
};

#endif
