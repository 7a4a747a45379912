// Shared declarations for the rubix_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "rubix_b200.h"

namespace rbx {

void set_error(const std::string &msg);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define RBX_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      rbx::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                   \
      return RBX_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define RBX_REQUIRE(cond, msg)                                                              \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      rbx::set_error(msg);                                                                  \
      return RBX_ERR_INVALID_ARGUMENT;                                                      \
    }                                                                                       \
  } while (0)

#define RBX_LAUNCH_OK()                                                                     \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      rbx::set_error(std::string("kernel launch: ") + cudaGetErrorString(_e));              \
      return RBX_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

// Tuning / test switches (options.cu).  -1 = unset (the library's own choice).  Seeded once from the environment
// (RBX_<NAME>) when the library is loaded, changed with rbx_set_option(); the launch path only reads atomics.
enum Opt {
  OPT_PSUB = 0, OPT_SORT_BITS, OPT_FUSED_FORCE_LUT, OPT_FUSED_FORCE_CAS, OPT_FUSED_IMPL, OPT_FUSED_CHS,
  OPT_FUSED_NO_SKEW, OPT_FUSED_WARPS, OPT_PREP_BLOCKS, OPT_SMALL_SHIFT, OPT_TAIL_SHIFT, OPT_HOST_CHUNKS,
  OPT_MARCH_NO_BULK, OPT_SORT_IMPL, OPT_FUSED_VARIANT, OPT_HOST_RATIO, OPT_FUSED_TR, OPT_COUNT
};
int64_t opt(Opt o);
inline bool opt_on(Opt o) { return opt(o) > 0; }

constexpr int kMaxLutBuckets = 8192;  // channel-lookup buckets the fused kernel keeps in shared memory
constexpr float kSpeedOfLight = 299792.458f;  // rubix/config/rubix_config.yml:8

// Device view of a plan (passed by value to kernels).
struct PlanView {
  int nz, na, L, Lp, W, method, vel_comp;
  int nb;                      // number of channel-lookup buckets
  float tmin, tmax, trange;    // telescope band; trange = tmax - tmin
  float lut_scale;             // (2*nb - 1) / trange, see bucket_offset()
  float tref;                  // reference wavelength of the q table (band centre)
  int affine;                  // t[w] == fadd(fmul(w, tdelta), t0) for every channel (numpy/jnp arange grids)
  float t0, tdelta, tinv;      // affine grid parameters (tinv = 1 / tdelta)
  const float *zgrid, *agrid;  // SSP metallicity / age axes
  const uint16_t *alut;        // (alut_n) start index for the age-axis search of a query in bucket b (plan.cu)
  int alut_n;
  float alut_scale;            // bucket = (age - agrid[0]) * alut_scale
  const float *tab[4];         // f, fx, fy, fxy: (nz*na, Lp) float32, rows 16-byte aligned
  const float *wt[4];          // window tables of the warp cube kernel (plan.cu: window_table_kernel) or nullptr:
                               // (nz*na, 256) float32, the knots wt_jbase .. wt_jbase + 255 of every row (clamped to
                               // the SSP grid like jnp.interp's end values), laid out so that the two 16-byte loads
                               // of a lane (its knots 8l .. 8l+3 and 8l+4 .. 8l+7) are each contiguous across the warp
  int wt_jbase;                // first knot of the window (a multiple of 4, may be negative)
  const float *lamz;           // (L)  (1+z)*wavelength                       rubix/spectra/ifu.py:80
  const float *rdl;            // (L)  1/(lamz[j+1]-lamz[j]); 0 for the last knot or zero width
  const float *ka, *kb;        // (L)  knot position in channel units u' = ka * (d - 1) + kb (affine grids; fused.cu: knot_ab)
  const float *t;              // (W)  telescope wave_seq
  const float *dt;             // (W)  diff0(t): [0, t1-t0, ...]                rubix/spectra/ifu.py:84-102
  const uint16_t *lut;         // (nb)  number of channels whose bucket is smaller than b
  const float2 *tt;            // (W)   (t[max(w-1,0)], t[w])
  const float2 *q;             // (W+1) double-float prefix sums of dt[w]*(t[w]-tref) over w < k
};

// Monotone map wavelength -> 2 * bucket (a byte offset into the uint16 lut), without F2I: the
// rounded product lands in the mantissa of 1.5 * 2^23.  The lut is built on the device with this
// very function (plan.cu), so host/device rounding differences cannot arise.
__device__ __forceinline__ int bucket_offset(float x, float tmin, float trange, float scale) {
  float xs = fminf(fmaxf(__fsub_rn(x, tmin), 0.f), trange);
  float y = __fmaf_rn(xs, scale, 12582912.f);
  return __float_as_int(y) & 0x3FFFFE;
}

}  // namespace rbx

struct rbx_plan {
  rbx::PlanView v;
  std::vector<void *> allocs;
  std::vector<float> h_lamz, h_t;
  int device;
  int lut_ok;        // every telescope channel has a bucket of its own
  float min_dt, max_dt;
};

namespace rbx {

// Radix sort of (key, particle index) pairs (sort.cu).
constexpr int kSortMaxBits = 8;     // digit width: at most 256 bins per pass
constexpr int kSortMaxPasses = 4;   // keys have < 32 significant bits
struct SortPlan {
  int npass, ntiles;
  int bits[kSortMaxPasses], shift[kSortMaxPasses];
};
SortPlan make_sort_plan(int64_t n, int end_bit);
size_t sort_state_words(const SortPlan &sp);   // uint32 words of state: per pass [256 digit counts][256: ticket][ntiles][256]
int radix_sort_pairs(const SortPlan &sp, uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n,
                     uint32_t *d_state, uint32_t **keys_sorted, uint32_t **vals_sorted, cudaStream_t stream);

// Everything the rbx_*build_cube* entry points hand to build_cube_impl (fused.cu).
struct CubeBuild {
  const float *vel = nullptr;      // Doppler component of particle 0 ...
  int vstride = 3;                 // ... and floats between particles ((n, 3) arrays: 3; a line-of-sight array: 1)
  const float *mass = nullptr, *met = nullptr, *age = nullptr;
  int32_t *pixel = nullptr;        // spaxel ids: input when cx == nullptr, else an optional output
  const float *cx = nullptr, *cy = nullptr;   // x / y of particle 0 (spaxel assignment inside prep_kernel) or nullptr
  int cstride = 3;
  const float *edges = nullptr;    // spatial bin edges (with cx)
  int n_edges = 0;
  int mark_outside = 0;            // aperture filter: particles outside the edges get pixel -1
  int accumulate = 0;              // cube += instead of cube =
  int nslab = 1, halo = 0;         // slab-major cube (rbx_assign_build_cube_slabs)
};
int build_cube_impl(const rbx_plan *plan, const CubeBuild &b, int64_t n, int num_spaxels, float *d_cube, void *d_ws,
                    size_t ws_bytes, void *stream);

// ---- packed float32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): one issue slot for two IEEE operations, each lane
// an ordinary round-to-nearest op, so results are bit-identical to the scalar form -------------------------
__device__ __forceinline__ void ffma2s(float &d0, float &d1, float a0, float a1, float b0, float b1) {  // d = a * b + d
  asm("{\n.reg .b64 ra, rb, rc;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmov.b64 rc, {%0, %1};\n"
      "fma.rn.f32x2 rc, ra, rb, rc;\nmov.b64 {%0, %1}, rc;\n}"
      : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void ffma2o(float &d0, float &d1, float a0, float a1, float b0, float b1, float c0, float c1) {  // d = a * b + c
  asm("{\n.reg .b64 ra, rb, rc;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmov.b64 rc, {%6, %7};\n"
      "fma.rn.f32x2 rc, ra, rb, rc;\nmov.b64 {%0, %1}, rc;\n}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fmul2s(float &d0, float &d1, float a0, float a1, float b0, float b1) {  // d = a * b
  asm("{\n.reg .b64 ra, rb, rc;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\n"
      "mul.rn.f32x2 rc, ra, rb;\nmov.b64 {%0, %1}, rc;\n}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fadd2s(float &d0, float &d1, float a0, float a1, float b0, float b1) {  // d = a + b
  asm("{\n.reg .b64 ra, rb, rc;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\n"
      "add.rn.f32x2 rc, ra, rb;\nmov.b64 {%0, %1}, rc;\n}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// ---- small device helpers ---------------------------------------------------------------------
// searchsorted(a, v, side='right') on a sorted array: number of elements <= v (NaN v -> n, like numpy/XLA
// where NaN sorts last).
__device__ __forceinline__ int ss_right(const float *__restrict__ a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (!(a[mid] > v)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// The same count, found by walking from a guess k (any value; a good one costs one or two compares): exact.
__device__ __forceinline__ int ss_right_from(const float *__restrict__ a, int n, float v, int k) {
  if (v != v) return n;
  k = min(max(k, 0), n);
  while (k < n && !(a[k] > v)) ++k;
  while (k > 0 && a[k - 1] > v) --k;
  return k;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// nan_to_num(x, nan=0): NaN -> 0, +-inf -> +-FLT_MAX   (rubix/spectra/ifu.py:253-255)
__device__ __forceinline__ float nan_to_num0(float x) {
  if (x != x) return 0.f;
  if (isinf(x)) return x > 0 ? 3.4028234664e38f : -3.4028234664e38f;
  return x;
}

// Interpolation weights of one (Z, age) query: up to 16 (row offset, weight) pairs such that
// spectrum[l] = sum_q w[q] * tab[tq[q]][row[q] * Lp + l].   Restates interpax.interp2d
// (method linear / cubic, extrap=0) as bound at rubix/spectra/ssp/grid.py:113-120.
// Returns the number of terms (0 when the query is outside the grid -> zero spectrum).
struct SspTerms {
  int n;
  int row[16];   // row index iz*na+ia
  int tabid[16]; // which table
  float w[16];
};

__device__ __forceinline__ void ssp_cell(const PlanView &p, float zq, float aq, int &i, int &j, bool &inside) {
  inside = (zq >= p.zgrid[0]) && (zq <= p.zgrid[p.nz - 1]) && (aq >= p.agrid[0]) && (aq <= p.agrid[p.na - 1]);
  i = min(max(ss_right(p.zgrid, p.nz, zq), 1), p.nz - 1);
  j = min(max(ss_right(p.agrid, p.na, aq), 1), p.na - 1);
}

__device__ inline void ssp_terms_at(const PlanView &p, float zq, float aq, float mass, int i, int j, bool inside,
                                    SspTerms &o);

__device__ inline void ssp_terms(const PlanView &p, float zq, float aq, float mass, SspTerms &o) {
  int i, j;
  bool inside;
  ssp_cell(p, zq, aq, i, j, inside);
  ssp_terms_at(p, zq, aq, mass, i, j, inside, o);
}

// the same with the cell (i, j) of ssp_cell() already known
__device__ inline void ssp_terms_at(const PlanView &p, float zq, float aq, float mass, int i, int j, bool inside,
                                    SspTerms &o) {
  if (!inside) { o.n = 0; return; }
  float x0 = p.zgrid[i - 1], x1 = p.zgrid[i], y0 = p.agrid[j - 1], y1 = p.agrid[j];
  float dx = x1 - x0, dy = y1 - y0;
  float dxi = dx == 0.f ? 0.f : 1.f / dx, dyi = dy == 0.f ? 0.f : 1.f / dy;
  if (p.method == RBX_METHOD_LINEAR) {
    float tx0 = x1 - zq, tx1 = zq - x0, ty0 = y1 - aq, ty1 = aq - y0, sc = dxi * dyi * mass;
    o.n = 4;
    o.row[0] = (i - 1) * p.na + (j - 1); o.w[0] = sc * (tx0 * ty0);
    o.row[1] = (i - 1) * p.na + j;       o.w[1] = sc * (tx0 * ty1);
    o.row[2] = i * p.na + (j - 1);       o.w[2] = sc * (tx1 * ty0);
    o.row[3] = i * p.na + j;             o.w[3] = sc * (tx1 * ty1);
#pragma unroll
    for (int q = 0; q < 4; ++q) o.tabid[q] = 0;
    return;
  }
  float tx = (zq - x0) * dxi, ty = (aq - y0) * dyi;
  float tx2 = tx * tx, tx3 = tx2 * tx, ty2 = ty * ty, ty3 = ty2 * ty;
  float hx[4] = {2.f * tx3 - 3.f * tx2 + 1.f, -2.f * tx3 + 3.f * tx2, tx3 - 2.f * tx2 + tx, tx3 - tx2};
  float hy[4] = {2.f * ty3 - 3.f * ty2 + 1.f, -2.f * ty3 + 3.f * ty2, ty3 - 2.f * ty2 + ty, ty3 - ty2};
  o.n = 16;
  int k = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    int sx = (t & 1) ? 2 : 0, sy = (t & 2) ? 2 : 0;
    float scale = ((t & 1) ? dx : 1.f) * ((t & 2) ? dy : 1.f) * mass;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
      for (int ii = 0; ii < 2; ++ii) {
        o.w[k] = hx[sx + ii] * hy[sy + jj] * scale;
        o.row[k] = (i - 1 + ii) * p.na + (j - 1 + jj);
        o.tabid[k] = t;
        ++k;
      }
  }
}

}  // namespace rbx
