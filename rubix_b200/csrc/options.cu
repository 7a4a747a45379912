// Tuning / test switches of the library (include/rubix_b200.h: rbx_set_option / rbx_get_option).
// Values live in a table of atomics that the launch path reads; the environment (RBX_<NAME>) is consulted ONCE,
// when the library is loaded, to seed the table -- nothing on the launch path calls getenv().
#include <cctype>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace rbx {

static const char *const kOptNames[OPT_COUNT] = {
    "psub", "sort_bits", "fused_force_lut", "fused_force_cas", "fused_impl", "fused_chs", "fused_no_skew",
    "fused_warps", "prep_blocks", "small_shift", "tail_shift", "host_chunks", "march_no_bulk", "sort_impl",
    "fused_variant", "host_ratio", "fused_tr"};

static std::atomic<int64_t> g_opts[OPT_COUNT];

static int64_t parse_value(const char *s) {
  if (!strcmp(s, "group")) return 1;   // RBX_FUSED_IMPL=group
  if (!strcmp(s, "warp") || !strcmp(s, "auto")) return 0;
  if (!strcmp(s, "cub")) return 1;     // RBX_SORT_IMPL=cub
  char *end = nullptr;
  const long long v = strtoll(s, &end, 10);
  return end == s ? 1 : (int64_t)v;    // a bare flag ("RBX_FUSED_NO_SKEW=yes") counts as 1
}

struct OptInit {
  OptInit() {
    for (int o = 0; o < OPT_COUNT; ++o) {
      g_opts[o].store(-1);
      std::string env = "RBX_";
      for (const char *c = kOptNames[o]; *c; ++c) env.push_back((char)toupper(*c));
      if (const char *e = getenv(env.c_str())) g_opts[o].store(parse_value(e));
    }
  }
};
static OptInit g_opt_init;

int64_t opt(Opt o) { return g_opts[o].load(std::memory_order_relaxed); }

}  // namespace rbx

using namespace rbx;

extern "C" int rbx_set_option(const char *name, int64_t value) {
  RBX_REQUIRE(name, "rbx_set_option: null name");
  for (int o = 0; o < OPT_COUNT; ++o)
    if (!strcmp(name, kOptNames[o])) {
      g_opts[o].store(value < 0 ? -1 : value);
      return RBX_OK;
    }
  set_error(std::string("rbx_set_option: unknown option ") + name);
  return RBX_ERR_INVALID_ARGUMENT;
}

extern "C" int rbx_get_option(const char *name, int64_t *value) {
  RBX_REQUIRE(name && value, "rbx_get_option: null argument");
  for (int o = 0; o < OPT_COUNT; ++o)
    if (!strcmp(name, kOptNames[o])) {
      *value = g_opts[o].load();
      return RBX_OK;
    }
  set_error(std::string("rbx_get_option: unknown option ") + name);
  return RBX_ERR_INVALID_ARGUMENT;
}
