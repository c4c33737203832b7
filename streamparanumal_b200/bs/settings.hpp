// Command line of the BS executables — the reference's settings_t as the drivers use it
// (include/settings.hpp; newSetting / getSetting / compareSetting / report; options declared in
// libs/core/platformSettings.cpp:33-47, libs/mesh/meshSettings.cpp:27-50, BS/BS1/src/bs1Settings.cpp:35-59,
// BS/BS0/src/bs0Settings.cpp:35-59, BS/BS6/src/bs6Settings.cpp:33-42).  Same flags, same names,
// same defaults; an unknown flag or a value outside the allowed set aborts like the reference.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "platform.hpp"

namespace spn {

class settings_t {
 public:
  explicit settings_t(comm_t comm_) : comm(comm_) {}
  comm_t comm;

  void newSetting(const std::string& shortkey, const std::string& longkey, const std::string& name,
                  const std::string& val, const std::string& description = "",
                  const std::vector<std::string>& options = {}) {
    SPN_ABORT("settings: duplicate setting " + name, entries_.count(name) != 0);
    entries_[name] = Entry{shortkey, longkey, val, description, options, false};
    order_.push_back(name);
  }

  void parseSettings(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
      const std::string arg = argv[i];
      if (arg == "-h" || arg == "--help") {
        if (comm.rank() == 0) printUsage(argv[0]);
        std::exit(0);
      }
      Entry* hit = nullptr;
      std::string hitName;
      for (auto& kv : entries_)
        if (arg == kv.second.shortkey || arg == kv.second.longkey) { hit = &kv.second; hitName = kv.first; }
      SPN_ABORT("Unrecognized setting: " + arg, hit == nullptr);
      if (hit->options.size() == 2 && hit->options[0] == "TRUE" && hit->options[1] == "FALSE") {
        hit->value = "TRUE";  // toggle (e.g. -v)
        hit->given = true;
        continue;
      }
      SPN_ABORT("Missing value for setting: " + arg, i + 1 >= argc);
      changeSetting(hitName, argv[++i]);
      hit->given = true;
    }
  }

  void changeSetting(const std::string& name, const std::string& val) {
    auto it = entries_.find(name);
    SPN_ABORT("Unable to find setting: " + name, it == entries_.end());
    if (!it->second.options.empty()) {
      bool ok = false;
      for (auto& o : it->second.options) ok = ok || (o == val);
      SPN_ABORT("Value: \"" + val + "\" not valid for setting " + name, !ok);
    }
    it->second.value = val;
  }

  template <typename T>
  void getSetting(const std::string& name, T& value) const {
    auto it = entries_.find(name);
    SPN_ABORT("Unable to find setting: " + name, it == entries_.end());
    std::stringstream ss(it->second.value);
    ss >> value;
  }
  std::string getSetting(const std::string& name) const {
    auto it = entries_.find(name);
    SPN_ABORT("Unable to find setting: " + name, it == entries_.end());
    return it->second.value;
  }
  bool compareSetting(const std::string& name, const std::string& token) const { return getSetting(name) == token; }
  bool given(const std::string& name) const {
    auto it = entries_.find(name);
    return it != entries_.end() && it->second.given;
  }

  void reportSetting(const std::string& name) const {
    auto it = entries_.find(name);
    if (it == entries_.end()) return;
    std::cout << "Name:     [" << name << "]\nCL keys:  [" << it->second.shortkey << ", " << it->second.longkey
              << "]\nValue:    " << it->second.value << "\n\n";
  }
  void report() const {
    std::cout << "Settings:\n\n";
    for (auto& n : order_) reportSetting(n);
  }

 private:
  struct Entry {
    std::string shortkey, longkey, value, description;
    std::vector<std::string> options;
    bool given;
  };
  void printUsage(const char* prog) const {
    std::cout << "Usage: " << prog << " [options]\n";
    for (auto& n : order_) {
      const Entry& e = entries_.at(n);
      std::cout << "  " << e.shortkey << ", " << e.longkey << "  " << e.description << " [" << e.value << "]\n";
    }
  }
  std::map<std::string, Entry> entries_;
  std::vector<std::string> order_;
};

// libs/core/platformSettings.cpp:33-47
inline void platformAddSettings(settings_t& s) {
  s.newSetting("-m", "--mode", "THREAD MODEL", "CUDA", "Parallel execution platform", {"Serial", "OpenMP", "CUDA", "HIP", "OpenCL"});
  s.newSetting("-pl", "--platform", "PLATFORM NUMBER", "0", "Parallel platform number (used in OpenCL mode)");
  s.newSetting("-d", "--device", "DEVICE NUMBER", "0", "Parallel device number");
}
// libs/mesh/meshSettings.cpp:27-50
inline void meshAddSettings(settings_t& s) {
  s.newSetting("-e", "--elements", "ELEMENT TYPE", "Hex", "Type of mesh elements", {"Tri", "Quad", "Tet", "Hex"});
  s.newSetting("-nx", "--dimx", "BOX NX", "50", "Number of elements in X-dimension per rank");
  s.newSetting("-ny", "--dimy", "BOX NY", "50", "Number of elements in Y-dimension per rank");
  s.newSetting("-nz", "--dimz", "BOX NZ", "50", "Number of elements in Z-dimension per rank");
  s.newSetting("-p", "--degree", "POLYNOMIAL DEGREE", "3", "Degree of polynomial finite element space",
               {"1", "2", "3", "4", "5", "6", "7", "8", "9", "10", "11", "12", "13", "14", "15"});
}
inline void verboseAddSetting(settings_t& s) {
  s.newSetting("-v", "--verbose", "VERBOSE", "FALSE", "Enable verbose output", {"TRUE", "FALSE"});
}

// platform_t(settings) of the reference (libs/core/platform.cpp, platformDeviceConfig.cpp:84-159): only the CUDA
// thread model exists in this backend — Serial / OpenMP live in the oracle build, HIP / OpenCL nowhere.
inline platform_t makePlatform(const settings_t& settings) {
  const std::string mode = settings.getSetting("THREAD MODEL");
  SPN_ABORT("mode '" + mode + "' is not available: this backend is sm_100a CUDA only (no multi-backend dispatch, no CPU fallback)",
            mode != "CUDA");
  int device = -1;  // local rank, as the reference does for multi-rank runs
  if (settings.given("DEVICE NUMBER") && settings.comm.size() == 1) settings.getSetting("DEVICE NUMBER", device);
  return platform_t(settings.comm, device);
}

}  // namespace spn
