// Handle semantics of the host layer that need no device ("Null handles" and "Memory
// ownership"): at P=1 the reference routinely passes uninitialised deviceMemory handles through conversions,
// operator+ and zero-count copies (include/memory.hpp:296-303,319-324,338,372,400,434) and expects no-ops.
// Also: properties_t path access, settings_t parsing, kernel_t on an unbuilt kernel, Factor3 / RankDecomp3 and the
// box labels.  Exit code 0 = all checks passed.  Built and run by tests/test_cpp_host.py.
#include <cstdio>
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include "kernel.hpp"
#include "ogs.hpp"
#include "settings.hpp"

using namespace spn;

static int failures = 0;
#define CHECK(cond)                                                         \
  do {                                                                      \
    if (!(cond)) { ++failures; fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
  } while (0)
#define CHECK_THROWS(expr)                                                  \
  do {                                                                      \
    bool threw = false;                                                     \
    try { expr; } catch (const std::exception&) { threw = true; }           \
    if (!threw) { ++failures; fprintf(stderr, "FAILED %s:%d: no exception from %s\n", __FILE__, __LINE__, #expr); } \
  } while (0)

int main() {
  // ---- null deviceMemory handles ----
  deviceMemory<dfloat> o_null;
  CHECK(!o_null.isInitialized());
  CHECK(o_null.ptr() == nullptr);
  CHECK(o_null.length() == 0 && o_null.size() == 0);
  deviceMemory<dfloat> o_off = o_null + 5;           // operator+ on a null handle returns a null handle
  CHECK(!o_off.isInitialized() && o_off.ptr() == nullptr);
  o_off += 3;
  CHECK(!o_off.isInitialized());
  deviceMemory<char> o_bytes;                         // deviceMemory<char> -> deviceMemory<T> reinterpretation of null
  deviceMemory<hlong> o_cast = o_bytes;
  CHECK(!o_cast.isInitialized() && o_cast.length() == 0);
  dfloat host[4] = {1, 2, 3, 4};
  o_null.copyFrom(host, 0);                           // zero-count copies are no-ops, even on null handles
  o_null.copyTo(host, 0);
  o_null.copyFrom(o_off, 0);
  o_null.copyTo(o_off, 0);
  CHECK(host[0] == 1 && host[3] == 4);
  o_null.free();
  pinnedMemory<dfloat> h_null;
  CHECK(!h_null.isInitialized() && h_null.ptr() == nullptr && h_null.length() == 0);

  // ---- kernel_t / properties_t ----
  kernel_t k;
  CHECK(!k.isInitialized());
  CHECK_THROWS(k(1, o_null, o_null));                 // calling an unbuilt kernel raises, it does not compute on the host
  properties_t props;
  props["defines/p_blockSize"] = 256;
  props["defines/T"] = "double";
  props["defines/OGS_OP(a,b)"] = "a+=b";
  props["compiler_flags"] += " -O3";
  CHECK(props.get("defines/p_blockSize") == "256");
  CHECK(props.get("defines/T") == "double" && props.has("defines/OGS_OP(a,b)"));
  CHECK(props.get("missing", "dflt") == "dflt");
  properties_t async("async", true);
  CHECK(async.get("async") == "true");

  // ---- settings_t: the reference flags ----
  {
    settings_t s{comm_t()};
    platformAddSettings(s);
    meshAddSettings(s);
    verboseAddSetting(s);
    const char* argv[] = {"BS8", "-nx", "64", "--dimy", "32", "-p", "7", "-v", "-m", "CUDA"};
    s.parseSettings(10, const_cast<char**>(argv));
    int nx = 0, ny = 0, nz = 0, p = 0;
    s.getSetting("BOX NX", nx); s.getSetting("BOX NY", ny); s.getSetting("BOX NZ", nz); s.getSetting("POLYNOMIAL DEGREE", p);
    CHECK(nx == 64 && ny == 32 && nz == 50 && p == 7);   // -nz keeps its default of 50 (meshSettings.cpp:40-42)
    CHECK(s.compareSetting("VERBOSE", "TRUE") && s.compareSetting("ELEMENT TYPE", "Hex"));
    CHECK(s.given("BOX NX") && !s.given("BOX NZ"));
    const char* bad[] = {"BS8", "-p", "16"};
    CHECK_THROWS(s.parseSettings(3, const_cast<char**>(bad)));
    const char* unknown[] = {"BS8", "--frobnicate"};
    CHECK_THROWS(s.parseSettings(2, const_cast<char**>(unknown)));
    const char* serial[] = {"BS8", "-m", "Serial"};
    s.parseSettings(3, const_cast<char**>(serial));
    CHECK_THROWS(makePlatform(s));                    // only the CUDA thread model exists: no CPU fallback
  }

  // ---- process grid and box labels (libs/core/rankDecomp.cpp:55-91, meshSetupBoxHex3D.cpp) ----
  {
    int sx, sy, sz;
    Factor3(8, sx, sy, sz); CHECK(sx == 2 && sy == 2 && sz == 2);
    Factor3(4, sx, sy, sz); CHECK(sx * sy * sz == 4 && sx == 2 && sy == 2 && sz == 1);
    Factor3(2, sx, sy, sz); CHECK(sx == 2 && sy == 1 && sz == 1);
    Factor3(6, sx, sy, sz); CHECK(sx * sy * sz == 6);
    std::set<hlong> all;
    hlong expected = 0;
    const int nx = 2, ny = 3, nz = 2, p = 3, Np = (p + 1) * (p + 1) * (p + 1);
    for (int r = 0; r < 4; ++r) {
      std::vector<hlong> ids((size_t)nx * ny * nz * Np);
      BoxHexGlobalIds(r, 4, nx, ny, nz, p, ids.data(), &expected);
      for (hlong id : ids) { CHECK(id > 0); all.insert(id); }
    }
    Factor3(4, sx, sy, sz);
    CHECK(expected == (hlong)(sx * nx * p + 1) * (sy * ny * p + 1) * (sz * nz * p + 1));
    CHECK((hlong)all.size() == expected);              // every lattice point is covered exactly once as a label
  }

  // ---- ogs enums equal the reference's and the ABI's numeric values ----
  CHECK((int)ogs::Float == SPN_FLOAT && (int)ogs::Double == SPN_DOUBLE && (int)ogs::Int32 == SPN_INT32 && (int)ogs::Int64 == SPN_INT64);
  CHECK((int)ogs::Add == SPN_ADD && (int)ogs::Mul == SPN_MUL && (int)ogs::Max == SPN_MAX && (int)ogs::Min == SPN_MIN);
  CHECK(ogs::Sizeof(ogs::Dfloat) == 8 && ogs::Sizeof(ogs::Dlong) == 4 && ogs::Sizeof(ogs::Hlong) == 8);

  if (failures) { fprintf(stderr, "%d check(s) failed\n", failures); return 1; }
  printf("host handle semantics ok\n");
  return 0;
}
