// Self-test of include/shim/occa.hpp against a RECORDING back end defined in this file (no device, no kernels): the
// semantics of the OCCA handle API that the reference's first-party code relies on:
//   occa::json      path-addressed operator[], = / += of int / double / string / bool, asObject, json(key, value), const lookup
//   occa::device    setup string parsing, mode(), malloc<T> (+ initial copy), zero-sized malloc stays uninitialised,
//                   host (pinned) allocations via props["host"], current stream get / set / create, finish on the current stream
//   occa::memory    ptr / size / length in elements of the dtype, operator+ / slice in elements, bounds, swap, free,
//                   copyFrom / copyTo for {pointer, memory} x {bytes, offsets, async}: kind of copy, addresses, async flag,
//                   zero-byte copies are no-ops
//   occa::kernel    variadic call packing int / long long / float / double / memory, argument count check, uninitialised kernel
// Prints "occa shim ok"; a failed check aborts with its line number.
#include <occa.hpp>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CHECK(cond)                                                            \
  do {                                                                         \
    if (!(cond)) {                                                             \
      fprintf(stderr, "check failed at line %d: %s\n", __LINE__, #cond);       \
      abort();                                                                 \
    }                                                                          \
  } while (0)

namespace {
struct CopyRecord { void* dst; const void* src; size_t bytes; occa::backend::copyKind kind; void* stream; bool async; };
struct LaunchRecord { std::string name; std::vector<occa::kernelArg> args; void* stream; };
std::vector<CopyRecord> g_copies;
std::vector<LaunchRecord> g_launches;
std::vector<void*> g_synced;
int g_device_allocs = 0, g_host_allocs = 0, g_streams = 0;
std::string g_setup_mode;
int g_setup_id = -1;
const char* kNames[] = {"probe", "probe3"};
}  // namespace

namespace occa {
namespace backend {
std::string deviceSetup(const std::string& mode, int device_id) {
  g_setup_mode = mode;
  g_setup_id = device_id;
  if (mode != "CUDA") throw occa::exception("mode not implemented: " + mode);
  return mode;
}
int deviceCount(const std::string&) { return 3; }
void* deviceMalloc(size_t bytes) { ++g_device_allocs; return malloc(bytes); }
void deviceFree(void* p) { --g_device_allocs; free(p); }
void* hostMalloc(size_t bytes) { ++g_host_allocs; return malloc(bytes); }
void hostFree(void* p) { --g_host_allocs; free(p); }
void copy(void* dst, const void* src, size_t bytes, copyKind kind, void* stream, bool async) {
  g_copies.push_back({dst, src, bytes, kind, stream, async});
  memmove(dst, src, bytes);
}
void* streamCreate() { return reinterpret_cast<void*>((size_t)(0x1000 + 16 * ++g_streams)); }
void streamSync(void* stream) { g_synced.push_back(stream); }
static void record(const void* ctx, const kernelArg* args, int nargs, void* stream) {
  LaunchRecord r;
  r.name = static_cast<const char*>(ctx);
  r.args.assign(args, args + nargs);
  r.stream = stream;
  g_launches.push_back(r);
}
kernelEntry lookupKernel(const std::string&, const std::string& name, const json& props) {
  kernelEntry e;
  if (name == "probe") { e.launch = record; e.ctx = kNames[0]; e.nargs = 0; }                     // any argument count
  if (name == "probe3" && props.get("defines/T") == "double") { e.launch = record; e.ctx = kNames[1]; e.nargs = 3; }
  return e;  // unknown kernels stay uninitialised
}
}  // namespace backend
}  // namespace occa

static void test_json() {
  occa::json p;
  CHECK(!p.isInitialized() && !p.has("defines/T"));
  p["defines/T"] = "double";
  p["defines/p_blockSize"] = 256;
  p["defines"]["p_NnodesPerBlock"] = 512;
  p["compiler_flags"] += " -O3";
  p["compiler_flags"] += std::string(" --use_fast_math");
  p["okl/enabled"] = false;
  p["alpha"] = 0.1;
  CHECK(p.kind() == occa::json::object_ && p.has("defines") && p["defines"].children().size() == 3);
  CHECK(p.get("defines/T") == "double" && p.get("defines/p_blockSize") == "256" && p.get("defines/p_NnodesPerBlock") == "512");
  CHECK(p.get("compiler_flags") == " -O3 --use_fast_math");
  CHECK(p.get("okl/enabled", true) == false && p.get("missing/flag", true) == true && p.get("missing") == "");
  CHECK(std::stod(p.get("alpha")) == 0.1);
  p["defines/T"] = "float";  // overwrite keeps siblings
  CHECK(p.get("defines/T") == "float" && p.get("defines/p_blockSize") == "256");
  const occa::json& c = p;
  CHECK(!c["nothing/here"].isInitialized() && !p.has("nothing"));  // a const lookup does not create nodes
  occa::json async("async", true);
  CHECK(async.get("async", false) == true);
  occa::json copy = p;  // value semantics
  copy["defines/T"] = "int";
  CHECK(p.get("defines/T") == "float" && copy.get("defines/T") == "int");
  occa::json arr;
  arr["list"].asArray();
  arr["obj"].asObject();
  CHECK(arr["list"].kind() == occa::json::array_ && arr["obj"].kind() == occa::json::object_);
}

static void test_device_and_memory() {
  occa::device dev;
  CHECK(!dev.isInitialized());
  bool threw = false;
  try { dev.setup("{mode: 'Serial'}"); } catch (const occa::exception&) { threw = true; }
  CHECK(threw && g_setup_mode == "Serial");
  dev.setup("{mode: 'CUDA', device_id: 2}");
  CHECK(dev.isInitialized() && dev.mode() == "CUDA" && g_setup_id == 2);
  CHECK(occa::getDeviceCount("{mode: 'CUDA'}") == 3);
  occa::env::setOccaCacheDir("/nowhere");

  // streams: the default stream is an initialised handle with a null native stream; finish syncs the CURRENT one
  occa::stream def = dev.getStream();
  CHECK(def.isInitialized() && def.handle() == nullptr);
  occa::stream side = dev.createStream();
  CHECK(side.isInitialized() && side.handle() != nullptr && !(side == def));
  CHECK(!occa::stream().isInitialized());
  dev.finish();
  dev.setStream(side);
  CHECK(dev.getStream() == side);
  dev.finish();
  CHECK(g_synced.size() == 2 && g_synced[0] == nullptr && g_synced[1] == side.handle());

  // allocation: zero-sized stays uninitialised; typed malloc with initial data; pinned host memory
  occa::memory none = dev.malloc<double>(0);
  CHECK(!none.isInitialized() && none.ptr() == nullptr && none.size() == 0 && none.length() == 0 && g_device_allocs == 0);
  none.copyFrom((const void*)nullptr, 0);  // zero bytes: no-ops even on a null handle
  none.copyTo((void*)nullptr, 0);
  CHECK(g_copies.empty());
  std::vector<double> init(10);
  for (int i = 0; i < 10; ++i) init[i] = i + 0.5;
  occa::memory a = dev.malloc<double>(10, init.data());
  CHECK(a.isInitialized() && a.size() == 80 && a.length() == 10 && !a.isHost() && g_device_allocs == 1);
  CHECK(g_copies.size() == 1 && g_copies[0].kind == occa::backend::h2d && g_copies[0].bytes == 80 && !g_copies[0].async &&
        g_copies[0].stream == side.handle());
  occa::json hostProps;
  hostProps["host"] = true;
  occa::memory h = dev.malloc<char>(80, hostProps);
  CHECK(h.isHost() && g_host_allocs == 1 && h.size() == 80);
  h.setDtype(occa::dtype::get<double>());
  CHECK(h.length() == 10);

  // slicing in elements of the dtype; a slice shares the buffer
  occa::memory tail = a + 6;
  CHECK(tail.ptr() == static_cast<char*>(a.ptr()) + 48 && tail.length() == 4);
  occa::memory mid = a.slice(2, 3);
  CHECK(mid.ptr() == static_cast<char*>(a.ptr()) + 16 && mid.size() == 24);
  occa::memory inner = mid + 1;
  CHECK(inner.ptr() == static_cast<char*>(a.ptr()) + 24 && inner.length() == 2);
  CHECK((a + 10).length() == 0);
  threw = false;
  try { a.slice(8, 3); } catch (const occa::exception&) { threw = true; }
  CHECK(threw);
  threw = false;
  try { occa::memory().slice(0); } catch (const occa::exception&) { threw = true; }
  CHECK(threw);

  // copies: bytes and offsets are BYTES at this level (deviceMemory<T> multiplies by sizeof(T): include/memory.hpp:331-453)
  g_copies.clear();
  double out[4] = {0, 0, 0, 0};
  tail.copyTo(out);
  CHECK(g_copies.back().kind == occa::backend::d2h && g_copies.back().bytes == 32 && out[0] == 6.5 && out[3] == 9.5);
  double two[2] = {-1.0, -2.0};
  a.copyFrom(two, 16, 8, occa::json("async", true));
  CHECK(g_copies.back().kind == occa::backend::h2d && g_copies.back().async && g_copies.back().dst == static_cast<char*>(a.ptr()) + 8);
  h.copyFrom(a, 24, 8, 16);  // device -> pinned host, 3 doubles from element 2 to element 1
  CHECK(g_copies.back().kind == occa::backend::d2h && g_copies.back().bytes == 24 &&
        g_copies.back().src == static_cast<char*>(a.ptr()) + 16 && g_copies.back().dst == static_cast<char*>(h.ptr()) + 8);
  CHECK(static_cast<double*>(h.ptr())[1] == -2.0 && static_cast<double*>(h.ptr())[2] == 3.5);
  h.copyTo(a, 8, 72, 8, occa::json("async", true));  // pinned host -> device
  CHECK(g_copies.back().kind == occa::backend::h2d && g_copies.back().async && static_cast<double*>(a.ptr())[9] == -2.0);
  occa::memory b = dev.malloc<double>(10);
  b.copyFrom(a);
  CHECK(g_copies.back().kind == occa::backend::d2d && g_copies.back().bytes == 80 && static_cast<double*>(b.ptr())[9] == -2.0);
  occa::memory h2 = dev.malloc<double>(10, hostProps);
  h2.copyFrom(h);
  CHECK(g_copies.back().kind == occa::backend::h2h);
  const size_t before = g_copies.size();
  a.copyFrom(two, 0);
  b.copyTo(a, 0);
  CHECK(g_copies.size() == before);
  threw = false;
  try { a.copyFrom(two, 16, 72); } catch (const occa::exception&) { threw = true; }
  CHECK(threw);

  // swap and free; the buffer lives until its last handle goes
  occa::memory a2 = a;
  a.swap(b);
  CHECK(static_cast<double*>(b.ptr())[0] == 0.5 && a2.ptr() == b.ptr());
  b.free();
  CHECK(!b.isInitialized() && g_device_allocs == 2 && a2.isInitialized());
  a2.free(); tail.free(); mid.free(); inner.free();
  CHECK(g_device_allocs == 1);
  a.free();
  CHECK(g_device_allocs == 0);
  h.free(); h2.free();
  CHECK(g_host_allocs == 0);
  dev.setStream(def);
}

static void test_kernel() {
  occa::device dev;
  occa::json props;
  occa::kernel nothing = dev.buildKernel("x.okl", "absent", props);
  CHECK(!nothing.isInitialized() && !occa::kernel().isInitialized());
  bool threw = false;
  try { nothing(1, 2); } catch (const occa::exception&) { threw = true; }
  CHECK(threw);
  props["defines/T"] = "float";
  CHECK(!dev.buildKernel("x.okl", "probe3", props).isInitialized());  // resolved by (file, name, props)
  props["defines/T"] = "double";
  occa::kernel k3 = dev.buildKernel("x.okl", "probe3", props);
  occa::kernel k = dev.buildKernel("x.okl", "probe", props);
  CHECK(k3.isInitialized() && k.isInitialized() && k3.name() == "probe3");

  occa::memory m = dev.malloc<double>(4);
  occa::memory off = m + 1;
  const int N = 7;
  const long long big = 1LL << 40;
  const double alpha = 0.37;
  const float beta = 0.5f;
  occa::stream side = dev.createStream();
  dev.setStream(side);
  k(N, big, alpha, beta, m, off);
  CHECK(g_launches.size() == 1 && g_launches[0].name == "probe" && g_launches[0].stream == side.handle());
  const std::vector<occa::kernelArg>& a = g_launches[0].args;
  CHECK(a.size() == 6);
  CHECK(a[0].kind == occa::kernelArg::i32 && a[0].asInt() == 7);
  CHECK(a[1].kind == occa::kernelArg::i64 && a[1].asInt() == big);
  CHECK(a[2].kind == occa::kernelArg::f64 && a[2].asDouble() == 0.37);
  CHECK(a[3].kind == occa::kernelArg::f32 && a[3].asDouble() == 0.5);
  CHECK(a[4].kind == occa::kernelArg::pointer && a[4].asPointer() == m.ptr());
  CHECK(a[5].asPointer() == static_cast<char*>(m.ptr()) + 8);
  occa::memory null;
  k(N, null);  // an uninitialised handle travels as a null pointer (the P = 1 exchange buffers: libs/ogs/ogsPairwise.cpp:389-402)
  CHECK(g_launches[1].args[1].kind == occa::kernelArg::pointer && g_launches[1].args[1].asPointer() == nullptr);
  k();
  CHECK(g_launches[2].args.empty());
  k3(N, m, off);
  threw = false;
  try { k3(N, m); } catch (const occa::exception&) { threw = true; }
  CHECK(threw && g_launches.size() == 4);
}

int main() {
  test_json();
  test_device_and_memory();
  test_kernel();
  printf("occa shim ok\n");
  return 0;
}
