// Fused a1..a5: particles -> (S*S, W) cube without materialising (n, L) or (n, W).
//
// Pipeline (all on the caller's stream, no host synchronisation):
//   prep_kernel      (optional) spaxel assignment + aperture filter, validity, template cell, sort key =
//                    spaxel * ncell + cell, spaxel histogram (MUSE-size cubes), min/max Doppler factor,
//                    per-particle record {d, 1/d, template row, weights * mass}
//   cub radix sort   (key, particle index) pairs -- stable, so the summation order is deterministic
//   count_runs_kernel  (large cubes) spaxel counts from the run lengths of the sorted keys
//   segment_kernel   one block: segment starts, work items (bulk items of psub particles, quarter-size tail
//                    items behind them in the queue), SSP knot window for the observed Doppler range
//   fused_cube_warp_kernel  persistent CTAs (one per SM), one warp per work item and particle, the spaxel's
//                    spectrum accumulated in the warp's own shared-memory cells (no atomics), written once;
//                    fused_cube_kernel (warp groups) for non-arange grids / SSP grids finer than the telescope's
//   reduce_partials_kernel  spaxels that were split over several items: fixed-order sum of partial rows
//                    (and NaN poisoning of the cube when a configuration did not fit)
//
// The algorithm of fused_cube_kernel is described above the kernel.  In short: for one particle the
// reference evaluates p_w = jnp.interp(t_w, lam' = lam_z * d, s) for every telescope channel and rescales
// by total/new (rubix/spectra/ifu.py:241-260).  p is piecewise linear in t with break points at the
// Doppler-shifted SSP knots, so a particle is fully described by its ~220 knots: work per particle is
// O(#knots in band), not O(W log L), and nothing per-channel is touched until the spaxel is finished.
#include <cub/device/device_radix_sort.cuh>

#include <cstring>

#include "common.cuh"

namespace rbx {

#ifndef RBX_NB
#define RBX_NB 4
#endif
constexpr int NB = RBX_NB;            // particles per batch (between group barriers)
constexpr int KPL = 2;                // knot slots per lane
constexpr int OWN = 30;               // owner lanes per warp
constexpr int kCtaThreads = NB == 4 ? 512 : 640;   // NB = 2 needs fewer registers: one more group per SM
constexpr int kCtaWarps = kCtaThreads / 32;
constexpr int kMaxGroups = NB == 4 ? 4 : 5;
constexpr int kMaxGroupWarps = 8;
constexpr int kRegionSlack = 64;      // spare cells per warp region
constexpr int kRedStride = 12;        // floats per reduction slot (8 warps + padding: conflict-free LDS.128)
constexpr int kMaxKnots = KPL * OWN * kMaxGroupWarps - 3;  // widest knot window [ja, jb) a group can hold


struct Item { int start, count, spaxel, slot; };

enum Ctrl { C_NITEMS = 0, C_JA, C_JB, C_ERROR, C_WORK, C_NVALID, C_DMIN, C_DMAX, C_NSPLIT, C_IMPL, C_CHS, C_GCHS, C_NSPLITSEG, C_COUNT };
// ctrl[C_IMPL]: which cube kernel runs, decided on the device by segment_kernel from the knot window and the
// Doppler range actually present (both kernels are launched; the one not selected returns at once)
enum Impl { IMPL_WARP = 0, IMPL_GROUP = 1, IMPL_WARP_TR = 2 };   // TR: the warp kernel with the transposed cell layout
// ctrl[C_CHS] / ctrl[C_GCHS]: log2(channels per chunk) of the warp / group kernel, the smallest value (not below the
// host's choice from the grids) whose chunk geometry holds for the Doppler range present
// ctrl[C_ERROR]: 0 ok, 1 work-item tables too small, 2 knot window wider than the group kernel holds,
// 3 no chunk size fits the Doppler range present (rbx_build_cube_status reports it; the cube is NaN)

// Where the cube rows go.  Standard: (nseg, W) row-major (nslab = 1, ws = W).  Slab-major (multi-GPU):
// nslab wavelength slabs of wslab channels, each stored as its own (nseg, ws) block with `halo` extra channels on
// both sides (ws = wslab + 2 halo), so that one reduce-scatter hands every rank its summed slab INCLUDING the LSF
// halo.  A channel within `halo` of a slab border is stored twice (its own slab and the neighbour's halo);
// channels outside [0, W) stay zero from the initial memset (the zero padding of the 'same' convolution).
struct CubeLayout {
  int nslab, wslab, halo, ws;
  long long slab_stride;   // nseg * ws
};

__device__ __forceinline__ void cube_put(float *__restrict__ cube, const CubeLayout &cl, int spaxel, int ch, float v,
                                         bool add) {
  if (cl.nslab == 1) {
    float *q = cube + (size_t)spaxel * cl.ws + ch;
    *q = add ? *q + v : v;
    return;
  }
  const int r = ch / cl.wslab, o = ch - r * cl.wslab;
  float *row = cube + (size_t)r * cl.slab_stride + (size_t)spaxel * cl.ws;
  float *q = row + cl.halo + o;
  const float nv = add ? *q + v : v;
  *q = nv;
  if (o < cl.halo && r > 0) row[cl.halo + o - cl.slab_stride + cl.wslab] = nv;                   // right halo of slab r - 1
  if (o >= cl.wslab - cl.halo && r + 1 < cl.nslab) row[cl.slab_stride + o - (cl.wslab - cl.halo)] = nv;   // left halo of slab r + 1
}
__device__ __forceinline__ float cube_get(const float *__restrict__ cube, const CubeLayout &cl, int spaxel, int ch) {
  if (cl.nslab == 1) return cube[(size_t)spaxel * cl.ws + ch];
  const int r = ch / cl.wslab, o = ch - r * cl.wslab;
  return cube[(size_t)r * cl.slab_stride + (size_t)spaxel * cl.ws + cl.halo + o];
}

struct FusedWs {
  uint32_t *keys_in, *keys_out, *idx_in, *idx_out;
  float *rec;       // (n, rec_stride) sorted particle records
  int *counts;      // (nseg + 1)
  int *seg_start;   // (nseg + 1)
  int *item_start;  // (nseg + 1)
  int *split_list;  // (nseg) spaxels cut into several items, in spaxel order (reduce_partials_kernel walks this list)
  int *ctrl;        // C_COUNT
  Item *items;      // (max_items)
  float *partials;  // (max_split, Wp)
  void *cub_temp;
  size_t cub_bytes;
  SortPlan sp;           // own radix sort (sort.cu)
  uint32_t *sort_state;  // its digit histograms / tickets / tile states, zeroed together with counts and ctrl
  size_t zero_bytes;     // counts + ctrl + sort state: one memset
  int rec_stride, max_items, max_split, Wp, psub, end_bit, ncell;
  int cell_bits, cell_shift;  // sort key = spaxel << cell_bits | (template cell >> cell_shift)
};

__device__ __forceinline__ int rec_stride_for(int method) { return method == RBX_METHOD_LINEAR ? 8 : 20; }

// ---- prep ---------------------------------------------------------------------------------------
// record (floats, original particle order): [0]=d  [1]=1/d  [2]=template row (int bits)  [3]=d - 1 (expm1)
// [4..]=interpolation weights * mass.  The cube kernels fetch records through the sorted index.
constexpr int kDminBias = 0x7f7fffff;   // ctrl[C_DMIN] holds kDminBias - bits(dmin): a zeroed ctrl means "no particle yet"

// With `coords` the spaxel assignment (rubix/telescope/utils.py:138-151, same searches as
// spaxel_assign_kernel: bit-exact) and the aperture filter (rubix/core/telescope.py:155-174, as pixel -1)
// happen here, so the particle arrays are read once; `pixel` then is an optional output.
// `vel` points at the Doppler component of particle 0 and advances by `vstride` floats per particle ((n, 3)
// arrays: vstride 3; a packed line-of-sight velocity array: 1); likewise cx / cy / cstride for the coordinates.
__global__ void __launch_bounds__(256, 4) prep_kernel(PlanView p, const float *__restrict__ vel, int vstride, const float *__restrict__ mass,
                            const float *__restrict__ met, const float *__restrict__ age,
                            int32_t *__restrict__ pixel, int n, int nseg, int cell_bits, int cell_shift,
                            uint32_t *__restrict__ keys, uint32_t *__restrict__ idx, int *__restrict__ counts,
                            int *__restrict__ ctrl, int smem_hist, float *__restrict__ rec, int stride,
                            const float *__restrict__ cx, const float *__restrict__ cy, int cstride,
                            const float *__restrict__ edges, int n_edges, int mark_outside, int edges_smem,
                            SortPlan sp, uint32_t *__restrict__ sort_state) {
  const bool coords = cx != nullptr;
  extern __shared__ int s_dyn[];
  int *s_hist = s_dyn;   // small cubes: per-block spaxel histogram; large cubes: count_runs_kernel after the sort
  float *s_axes = reinterpret_cast<float *>(s_dyn + (smem_hist ? nseg : 0));   // SSP metallicity and age axes
  float *s_edges = s_axes + p.nz + p.na;
  // digit histograms of every radix pass of the sort that follows (sort.cu), gathered while the keys are made
  int *s_dig = reinterpret_cast<int *>(s_edges + ((coords && edges_smem) ? n_edges : 0));
  uint16_t *s_alut = reinterpret_cast<uint16_t *>(s_dig + (sort_state ? sp.npass * 256 : 0));
  for (int s = threadIdx.x; s < p.alut_n; s += blockDim.x) s_alut[s] = p.alut[s];
  if (sort_state)
    for (int s = threadIdx.x; s < sp.npass * 256; s += blockDim.x) s_dig[s] = 0;
  if (smem_hist)
    for (int s = threadIdx.x; s < nseg; s += blockDim.x) s_hist[s] = 0;
  for (int s = threadIdx.x; s < p.nz; s += blockDim.x) s_axes[s] = p.zgrid[s];
  for (int s = threadIdx.x; s < p.na; s += blockDim.x) s_axes[p.nz + s] = p.agrid[s];
  const float *e = edges;
  if (coords && edges_smem) {
    for (int s = threadIdx.x; s < n_edges; s += blockDim.x) s_edges[s] = edges[s];
    e = s_edges;
  }
  __syncthreads();
  p.zgrid = s_axes;
  p.agrid = s_axes + p.nz;
  float elo = 0.f, ehi = 0.f;
  const int nb = n_edges - 1;
  if (coords) {   // the reference takes min() / max() of the edges (rubix/telescope/utils.py:170-174)
    elo = ehi = e[0];
    for (int s = 1; s < n_edges; ++s) { elo = fminf(elo, e[s]); ehi = fmaxf(ehi, e[s]); }
  }
  // The searches (two on the spaxel edges, one on each SSP axis) were 40 % of this kernel's instructions and of its
  // stalls as bisections.  They now start from a guess -- the uniform-grid index for the edges, a bucket table for the
  // age axis -- and walk to the exact searchsorted(side='right') answer on the GIVEN float32 values.
  const float e0 = coords ? e[0] : 0.f;
  const float einv = (coords && e[nb] > e[0]) ? (float)nb / (e[nb] - e[0]) : 0.f;
  const float zlo = p.zgrid[0], zhi = p.zgrid[p.nz - 1], alo = p.agrid[0], ahi = p.agrid[p.na - 1];
  const float abmax = (float)(p.alut_n - 1);
  auto guess_of = [](float t, float hi) { return (int)fminf(fmaxf(t, 0.f), hi); };   // NaN -> 0
  float dmin = 3.0e38f, dmax = 0.f;
  int nvalid = 0;
  // Latency bound (one particle's loads at a time leave the memory system idle): every thread first requests the
  // inputs of U particles, then processes them.
  constexpr int U = 4;
  const int gstride = gridDim.x * blockDim.x;
  for (int q0 = blockIdx.x * blockDim.x + threadIdx.x; q0 < n; q0 += U * gstride) {
    float in_x[U], in_y[U], in_v[U], in_m[U], in_z[U], in_a[U];
    int in_px[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int q = q0 + k * gstride;
      in_x[k] = in_y[k] = in_v[k] = in_m[k] = in_z[k] = in_a[k] = 0.f;
      in_px[k] = -1;
      if (q < n) {
        in_z[k] = met[q]; in_a[k] = age[q]; in_m[k] = mass[q];
        in_v[k] = vel[(size_t)vstride * q];
        if (coords) { in_x[k] = cx[(size_t)cstride * q]; in_y[k] = cy[(size_t)cstride * q]; }
        else in_px[k] = pixel[q];
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int q = q0 + k * gstride;
      if (q >= n) break;
      const float zq = in_z[k], aq = in_a[k], mq = in_m[k];
      // ssp_cell() with guided searches
      const bool inside = (zq >= zlo) && (zq <= zhi) && (aq >= alo) && (aq <= ahi);
      int zc = 0;
      if (p.nz <= 16) {
        for (int s = 0; s < p.nz; ++s) zc += !(p.zgrid[s] > zq) ? 1 : 0;   // sorted axis: the count of nodes <= zq (NaN: all)
      } else {
        zc = ss_right(p.zgrid, p.nz, zq);
      }
      const int i = min(max(zc, 1), p.nz - 1);
      const int ag = s_alut[guess_of((aq - alo) * p.alut_scale, abmax)];
      const int j = min(max(ss_right_from(p.agrid, p.na, aq, ag), 1), p.na - 1);
      int px;
      if (coords) {
        const float x = in_x[k], y = in_y[k];
        const int xi = min(max(ss_right_from(e, n_edges, x, guess_of((x - e0) * einv, (float)nb) + 1) - 1, 0), nb - 1);
        const int yi = min(max(ss_right_from(e, n_edges, y, guess_of((y - e0) * einv, (float)nb) + 1) - 1, 0), nb - 1);
        px = xi + nb * yi;
        if (mark_outside && !((x >= elo) && (x <= ehi) && (y >= elo) && (y <= ehi))) px = -1;
        if (pixel) pixel[q] = px;
      } else {
        px = in_px[k];
      }
      const float voc = in_v[k] / kSpeedOfLight;
      float d = expf(voc);
      bool valid = inside && (mq != 0.f) && px >= 0 && px < nseg && (d > 0.f) && (d < 3.0e38f);
      uint32_t key = (uint32_t)nseg << cell_bits;  // invalid: sorts behind every valid key
      if (valid) {
        uint32_t cell = (uint32_t)((i - 1) * (p.na - 1) + (j - 1));
        key = ((uint32_t)px << cell_bits) | (cell >> cell_shift);
        if (smem_hist) atomicAdd(s_hist + px, 1);
        dmin = fminf(dmin, d);
        dmax = fmaxf(dmax, d);
        ++nvalid;
        SspTerms tm;
        ssp_terms_at(p, zq, aq, mq, i, j, inside, tm);
        float4 *r = reinterpret_cast<float4 *>(rec + (size_t)q * stride);
        r[0] = make_float4(d, 1.f / d, __int_as_float(tm.row[0]), expm1f(voc));
        r[1] = make_float4(tm.w[0], tm.w[1], tm.w[2], tm.w[3]);
        if (p.method != RBX_METHOD_LINEAR) {
          r[2] = make_float4(tm.w[4], tm.w[5], tm.w[6], tm.w[7]);
          r[3] = make_float4(tm.w[8], tm.w[9], tm.w[10], tm.w[11]);
          r[4] = make_float4(tm.w[12], tm.w[13], tm.w[14], tm.w[15]);
        }
      }
      keys[q] = key;
      if (sort_state) {
  #pragma unroll
        for (int ps = 0; ps < kSortMaxPasses; ++ps)
          if (ps < sp.npass) atomicAdd(s_dig + ps * 256 + ((key >> sp.shift[ps]) & ((1u << sp.bits[ps]) - 1u)), 1);
      } else {
        idx[q] = (uint32_t)q;   // cub sorts explicit (key, index) pairs; the own sort's first pass knows the indices
      }
    
    }
  }
  // positive floats order like their bit patterns
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
  }
  if ((threadIdx.x & 31) == 0 && nvalid > 0) {
    atomicMax(ctrl + C_DMIN, kDminBias - __float_as_int(dmin));
    atomicMax(ctrl + C_DMAX, __float_as_int(dmax));
    atomicAdd(ctrl + C_NVALID, nvalid);
  }
  if (smem_hist || sort_state) __syncthreads();
  if (smem_hist) {
    for (int s = threadIdx.x; s < nseg; s += blockDim.x) {
      const int c = s_hist[s];
      if (c) atomicAdd(counts + s, c);
    }
  }
  if (sort_state) {
    const size_t per_pass = ((size_t)sp.ntiles + 2) * 256;
    for (int s = threadIdx.x; s < sp.npass * 256; s += blockDim.x) {
      const int c = s_dig[s];
      if (c) atomicAdd(sort_state + (size_t)(s >> 8) * per_pass + (s & 255), (uint32_t)c);
    }
  }
}

// Per-spaxel particle counts from the SORTED keys: a spaxel's particles are one run, counts = run length
// (no histogram atomics in prep_kernel).  counts is zeroed beforehand; invalid keys sort behind every spaxel.
__global__ void count_runs_kernel(const uint32_t *__restrict__ keys_sorted, int n, int nseg, int cell_bits,
                                  int *__restrict__ counts) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
    const int s = (int)(keys_sorted[q] >> cell_bits);
    if (s >= nseg) continue;
    const int sp = q > 0 ? (int)(keys_sorted[q - 1] >> cell_bits) : -1;
    const int sn = q + 1 < n ? (int)(keys_sorted[q + 1] >> cell_bits) : nseg;
    if (sp != s) atomicSub(counts + s, q);      // run start:  counts[s] -= first index
    if (sn != s) atomicAdd(counts + s, q + 1);  // run end:    counts[s] += one past the last index
  }
}

// number of channels below x, and t[k-1] (t[0] for k = 0).
// AFFINE grids: k = ceil((x - t0) / delta) in float32.  Within ~1e-3 A of a channel wavelength the
// rounded quotient may pick the neighbouring cell; the kink is then booked one channel off with the
// matching e, which changes that single channel by dm * (t_k - x) <= dm * 1e-3 A: p(t) is continuous.
template <bool AFFINE>
__device__ __forceinline__ int channel_of(float x, const PlanView &p, const unsigned char *s_lut,
                                          const float2 *s_tt, float &e) {
  if (AFFINE) {
    const float v = fminf(fmaxf((x - p.t0) * p.tinv, 0.f), (float)p.W);
    const int k = __float2int_ru(v);
    e = __fadd_rn(__fmul_rn((float)max(k - 1, 0), p.tdelta), p.t0);
    return k;
  }
  const int off = bucket_offset(x, p.tmin, p.trange, p.lut_scale);
  const int l = *reinterpret_cast<const uint16_t *>(s_lut + off);
  const float2 t2 = s_tt[l];
  const bool up = t2.y < x;
  e = up ? t2.y : t2.x;
  return l + (up ? 1 : 0);
}

// ---- warp kernel geometry shared by segment_kernel (which decides whether the warp kernel can run) and the
// kernel itself ------------------------------------------------------------------------------------------
constexpr int WK = 8;                 // knots per lane
constexpr int kWarpSlots = WK * 32;   // knot slots of one warp

// The warp kernel works in CHANNEL UNITS of the affine telescope grid t_w = t0 + w delta: a knot at lam_z sits at
//   u' = (lam_z d - t0) / delta - 1/2 = a eps + b,   a = lam_z / delta,  b = (lam_z - t0) / delta - 1/2,  eps = d - 1
// (a, b rounded once from double, eps = expm1(v / c): u' is good to ~2.5e-4 channels = 3e-4 A, better than the
// reference's own float32 lam_z * d).  The number of channels below the knot is ceil(u' + 1/2), obtained WITHOUT a
// float -> int conversion: fl(u' + 1.5 * 2^23) rounds u' to the nearest integer, which is ceil(u' + 1/2) - 1 (a
// knot exactly on a channel may land in the neighbouring cell with the matching offset: p(t) is continuous, the
// channel value does not change), and the integer sits in the mantissa.
constexpr float kMagic = 12582912.f;           // 1.5 * 2^23: ulp 1 on [2^23, 2^24)
constexpr int kMagicM1Bits = 0x4B3FFFFF;       // bits of kMagic - 1
constexpr int kMagicBits = 0x4B400000;         // bits of kMagic
struct KnotAB { float a, b; };
__device__ __forceinline__ KnotAB knot_ab(const PlanView &p, int j) {
  KnotAB k;
  if (j < 0) { k.a = 0.f; k.b = -1000.5f; return k; }                 // below every band
  if (j >= p.L) { k.a = 0.f; k.b = (float)p.W + 1000.f; return k; }   // above every band
  k.a = p.ka[j];   // (float)(lam_z / delta) and (float)((lam_z - t0) / delta - 1/2), evaluated in double by rbx_plan_create
  k.b = p.kb[j];
  return k;
}
// KA = kMagic - 1 + (number of channels below the knot, clamped to [0, W])
__device__ __forceinline__ float knot_ka(float u, float kahi) {
  return fminf(fmaxf(__fadd_rn(u, kMagic), kMagic - 1.f), kahi);
}
__device__ __forceinline__ int knot_cell(float ka) { return __float_as_int(ka) - kMagicM1Bits; }

// Chunk lines of the warp kernel are summed in registers: over the Doppler range present, the first chunk start
// inside lane `lane`'s channel span must be chunk cA or cA + 1, and the span must hold at most one chunk start.
// Returns false when that does not hold (Doppler range too wide for this chunk size).  eps_lo / eps_hi bound
// d - 1 of every particle (segment_kernel widens the observed range by a few ulps), and u' is monotone in eps, so
// every particle's first cell lies in [kmin0, kmax0].
struct LaneCells { int kmin0, kmax0, knext; };   // extreme first cells of a lane over the Doppler range, first cell of the next lane
__device__ __forceinline__ LaneCells warp_lane_cells(const PlanView &p, int jbase, int lane, float eps_lo, float eps_hi) {
  const int j0 = jbase + WK * lane, j1 = j0 + WK;
  const KnotAB k0 = knot_ab(p, j0), k1 = knot_ab(p, j1);
  const float kahi = kMagic + (float)(p.W - 1);
  LaneCells c;
  c.kmin0 = knot_cell(knot_ka(fmaf(k0.a, eps_lo, k0.b), kahi));
  c.kmax0 = knot_cell(knot_ka(fmaf(k0.a, eps_hi, k0.b), kahi));
  c.knext = lane == 31 ? c.kmax0 : knot_cell(knot_ka(fmaf(k1.a, eps_hi, k1.b), kahi));
  return c;
}
__device__ __forceinline__ bool lane_chunks_ok(const LaneCells &c, int chs, int &cA) {
  const int CH = 1 << chs;
  cA = (c.kmin0 + CH - 1) >> chs;
  return !((((c.kmax0 + CH - 1) >> chs) > cA + 1) || (c.knext - c.kmax0 + 2 >= CH));
}
__device__ __forceinline__ bool warp_lane_chunks(const PlanView &p, int jbase, int lane, float eps_lo, float eps_hi,
                                                 int chs, int &cA) {
  return lane_chunks_ok(warp_lane_cells(p, jbase, lane, eps_lo, eps_hi), chs, cA);
}

// d - 1 of every particle lies in [eps_lo_of(dmin), eps_hi_of(dmax)]: d = fl(exp(x)) and eps = expm1f(x) differ
// from the exact values by a few ulps of d
__device__ __forceinline__ float eps_lo_of(float dmin) { return (dmin - 1.f) - 1.0e-6f; }
__device__ __forceinline__ float eps_hi_of(float dmax) { return (dmax - 1.f) + 1.0e-6f; }

// ---- segments / items / knot window (one block) -----------------------------------------------
// exclusive block scan of NV ints per thread (1024 threads): warp shuffles + one shared hop
template <int NV>
__device__ __forceinline__ void block_scan(int (&v)[NV], int (&total)[NV], int *s_w /* [33 * NV] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) inc[q] = v[q];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      const int nb = __shfl_up_sync(0xffffffffu, inc[q], o);
      if (lane >= o) inc[q] += nb;
    }
  }
  if (lane == 31) {
#pragma unroll
    for (int q = 0; q < NV; ++q) s_w[q * 33 + warp] = inc[q];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      const int w = s_w[q * 33 + lane];
      int x = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int nb = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += nb;
      }
      __syncwarp();
      s_w[q * 33 + lane] = x - w;             // exclusive warp offsets
      if (lane == 31) s_w[q * 33 + 32] = x;   // total
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    v[q] = s_w[q * 33 + warp] + inc[q] - v[q];
    total[q] = s_w[q * 33 + 32];
  }
}

// How a spaxel's `c` sorted particles are cut into work items.  The bulk goes into items of `psub`
// particles; the last eighth goes into items a quarter of that size, and ALL small items sit behind ALL
// bulk items in the work queue, so that the warps that finish their last bulk item early fill up on small
// ones (the tail of the persistent kernel shrinks from half a bulk item to half a small one).
struct Cut { int bulk, tail, nb, ns, psmall; };
__device__ __forceinline__ Cut cut_spaxel(int c, int psub, int small_shift, int tail_shift) {
  Cut k;
  k.psmall = max(32, psub >> small_shift);
  if (c <= psub) {   // one item, no integer divisions (most spaxels of a large cube)
    k.tail = 0; k.bulk = c; k.nb = c > 0 ? 1 : 0; k.ns = 0;
    return k;
  }
  k.tail = 0;
  if (c > psub) k.tail = min(c, ((c >> tail_shift) + k.psmall - 1) / k.psmall * k.psmall);
  k.bulk = c - k.tail;
  k.nb = (k.bulk + psub - 1) / psub;
  k.ns = (k.tail + k.psmall - 1) / k.psmall;
  return k;
}

// item_start[s] .. item_start[s+1]: the spaxel's rows in `partials` (none when it is a single item)
__global__ void __launch_bounds__(1024)
segment_kernel(PlanView p, int nseg, int psub, int small_shift, int tail_shift, int max_items, int max_split,
                               const int *__restrict__ counts, int *__restrict__ seg_start,
                               int *__restrict__ item_start, Item *__restrict__ items, int *__restrict__ ctrl,
                               int warp_ok, int warp_chs, int group_chs, int lamz_smem, int counts_smem,
                               int *__restrict__ split_list, int tr_B) {
  // the knot-window searches and the chunk geometry below are chains of dependent reads of lam_z by single threads:
  // from shared memory they cost a few hundred cycles instead of ~10 us
  extern __shared__ float s_lamz[];
  if (lamz_smem) {
    for (int q = threadIdx.x; q < p.L; q += blockDim.x) s_lamz[q] = p.lamz[q];
    p.lamz = s_lamz;   // visible after the __syncthreads inside block_scan
  }
  // a thread owns a contiguous range of spaxels (22 of them at 150 x 150): the counts are fetched coalesced into
  // shared memory first, and the segment offsets leave through the same array
  int *s_cnt = reinterpret_cast<int *>(s_lamz + (lamz_smem ? p.L : 0));
  if (counts_smem) {
    for (int q = threadIdx.x; q < nseg; q += blockDim.x) s_cnt[q] = counts[q];
    __syncthreads();
  }
  const int *cnt = counts_smem ? s_cnt : counts;
  __shared__ int s_w[33 * 5];
  __shared__ int s_err, s_ja, s_jb;
  __shared__ int s_bad_w[16], s_bad_g[16];   // [chs]: the chunk geometry fails for 2^chs channels per chunk
  __shared__ int s_bad_tr;                   // ... and for the transposed layout's blocks of tr_B channels
  if (threadIdx.x < 16) { s_bad_w[threadIdx.x] = 0; s_bad_g[threadIdx.x] = 0; }
  if (threadIdx.x == 0) s_bad_tr = 0;
  const int T = blockDim.x, t = threadIdx.x;
  const int per = (nseg + T - 1) / T;
  const int lo = min(t * per, nseg), hi = min(lo + per, nseg);
  int v[5] = {0, 0, 0, 0, 0}, tot[5];  // particles, bulk items, small items, split rows, split spaxels
  for (int s = lo; s < hi; ++s) {
    const int c = cnt[s];
    const Cut k = cut_spaxel(c, psub, small_shift, tail_shift);
    const int ni = k.nb + k.ns;
    v[0] += c; v[1] += k.nb; v[2] += k.ns; v[3] += ni > 1 ? ni : 0; v[4] += ni > 1 ? 1 : 0;
  }
  block_scan<5>(v, tot, s_w);   // v: exclusive prefixes; tot: totals
  if (t == 0) {
    int err = 0;
    if (tot[1] + tot[2] > max_items || tot[3] > max_split) err = 1;
    // SSP knot window for the Doppler factors actually present
    int ja = 0, jb = 0;
    if (tot[0] > 0) {
      float dmin = __int_as_float(kDminBias - ctrl[C_DMIN]), dmax = __int_as_float(ctrl[C_DMAX]);
      float lo_l = p.tmin / dmax, hi_l = p.tmax / dmin;
      // first knot that can reach the band / one past the last knot that can be in it
      int a = 0, hi_a = p.L;
      while (a < hi_a) { int mid = (a + hi_a) >> 1; if (p.lamz[mid] < lo_l) a = mid + 1; else hi_a = mid; }
      int b = ss_right(p.lamz, p.L, hi_l);
      b = max(b, a);
      ja = max(0, a - 3);
      jb = min(p.L, b + 3);
      if (jb - ja > kMaxKnots) err = 2;
    }
    ctrl[C_NITEMS] = err ? 0 : tot[1] + tot[2];
    ctrl[C_JA] = ja;
    ctrl[C_JB] = jb;
    ctrl[C_ERROR] = err;
    ctrl[C_WORK] = 0;
    ctrl[C_NSPLIT] = tot[3];
    ctrl[C_NSPLITSEG] = err ? 0 : tot[4];
    seg_start[nseg] = tot[0];
    item_start[nseg] = tot[3];
    s_err = err;
    s_ja = ja; s_jb = jb;
  }
  // Which cube kernel, and with which chunk size: the warp kernel when the host found the plan eligible (warp_ok),
  // the knot window fits its 256 slots and some chunk size <= 2^10 channels makes every lane's chunk geometry hold
  // for the Doppler range present; else the group kernel (chunks <= 2^8 channels).  Benign races: threads only
  // ever store 1 into the flags.
  // The warp kernel's window is fixed per plan, so its check (warp 1) runs beside thread 0's knot-window search.
  if (tot[0] > 0 && t >= 32 && t < 64 && p.affine) {
    const float dmin = __int_as_float(kDminBias - ctrl[C_DMIN]), dmax = __int_as_float(ctrl[C_DMAX]);
    const int jbase = p.wt_jbase;   // the warp kernel's knot window is fixed per plan (window tables, plan.cu)
    const LaneCells lc = warp_lane_cells(p, jbase, t - 32, eps_lo_of(dmin), eps_hi_of(dmax));
    for (int chs = warp_chs; chs <= 10; ++chs) {
      int cA;
      if (!lane_chunks_ok(lc, chs, cA)) s_bad_w[chs] = 1;
    }
    if (tr_B > 0) {   // at most one block start inside a lane's span, and it is block cA or cA + 1 for every particle
      const int cA = (lc.kmin0 + tr_B - 1) / tr_B;
      if ((lc.kmax0 + tr_B - 1) / tr_B > cA + 1 || lc.knext - lc.kmax0 + 1 > tr_B) s_bad_tr = 1;
    }
  }
  __syncthreads();
  if (tot[0] > 0) {
    const float dmin = __int_as_float(kDminBias - ctrl[C_DMIN]), dmax = __int_as_float(ctrl[C_DMAX]);
    // group kernel: a lane's first knot is some even slot of the window; checked for every knot (conservative)
    const int jbg = (s_ja - 1) & ~1;
    for (int j = jbg + t; j <= s_jb; j += T) {
      auto lam = [&](int q) { return q < 0 ? -1.0e30f : (q >= p.L ? 1.0e30f : p.lamz[q]); };
      float e;
      int kmin0, kmax0, knext;
      if (p.affine) {
        kmin0 = channel_of<true>(__fmul_rn(lam(j), dmin), p, nullptr, nullptr, e);
        kmax0 = channel_of<true>(__fmul_rn(lam(j), dmax), p, nullptr, nullptr, e);
        knext = channel_of<true>(__fmul_rn(lam(j + KPL), dmax), p, nullptr, nullptr, e);
      } else {   // any monotone grid: searches on the channel wavelengths (as many channels below x as the kernel's lookup)
        auto below = [&](float x) { int a = 0, b = p.W; while (a < b) { int m = (a + b) >> 1; if (p.t[m] < x) a = m + 1; else b = m; } return a; };
        kmin0 = below(__fmul_rn(lam(j), dmin));
        kmax0 = below(__fmul_rn(lam(j), dmax));
        knext = below(__fmul_rn(lam(j + KPL), dmax));
      }
      for (int chs = group_chs; chs <= 8; ++chs) {
        const int CH = 1 << chs;
        const int cA = (kmin0 + CH - 1) >> chs;
        if ((((kmax0 + CH - 1) >> chs) > cA + 1) || (knext - kmax0 + 2 >= CH)) s_bad_g[chs] = 1;
      }
    }
  }
  __syncthreads();
  if (t == 0) {
    int wchs = -1, gchs = -1;
    for (int chs = 10; chs >= warp_chs; --chs) if (!s_bad_w[chs]) wchs = chs;
    for (int chs = 8; chs >= group_chs; --chs) if (!s_bad_g[chs]) gchs = chs;
    // slot 0 must lie below the band and the last slot beyond it for every Doppler factor present
    const int jbase = p.wt_jbase;
    const bool warp = warp_ok != 0 && wchs >= 0 && jbase <= s_ja - 1 && (s_jb - jbase + 1 <= kWarpSlots);
    ctrl[C_IMPL] = warp ? ((tr_B > 0 && !s_bad_tr) ? IMPL_WARP_TR : IMPL_WARP) : IMPL_GROUP;
    ctrl[C_CHS] = max(wchs, warp_chs);
    ctrl[C_GCHS] = max(gchs, group_chs);
    if (!warp && gchs < 0 && s_err == 0) {   // the Doppler range present is too wide for either kernel
      s_err = 3;
      ctrl[C_ERROR] = 3;
      ctrl[C_NITEMS] = 0;
    }
  }
  __syncthreads();
  const bool err = s_err != 0;
  int ra = v[0], rb = v[1], rs = tot[1] + v[2], rc = v[3], rl = v[4];
  for (int s = lo; s < hi; ++s) {
    const int c = cnt[s];
    const Cut k = cut_spaxel(c, psub, small_shift, tail_shift);
    const int ni = k.nb + k.ns;
    if (counts_smem) s_cnt[s] = ra; else seg_start[s] = ra;   // my own range: nobody else reads these counts
    item_start[s] = rc;
    if (ni > 1 && !err) split_list[rl++] = s;
    if (!err) {
      for (int q = 0; q < k.nb; ++q) {   // bulk items: front of the queue
        Item it;
        it.start = ra + q * psub;
        it.count = min(psub, k.bulk - q * psub);
        it.spaxel = s;
        it.slot = ni > 1 ? rc + q : -1;
        items[rb + q] = it;
      }
      for (int q = 0; q < k.ns; ++q) {   // small items: behind every bulk item
        Item it;
        it.start = ra + k.bulk + q * k.psmall;
        it.count = min(k.psmall, k.tail - q * k.psmall);
        it.spaxel = s;
        it.slot = rc + k.nb + q;         // ns > 0 implies ni > 1
        items[rs + q] = it;
      }
    }
    ra += c; rb += k.nb; rs += k.ns; rc += ni > 1 ? ni : 0;
  }
  if (counts_smem) {
    __syncthreads();
    for (int q = threadIdx.x; q < nseg; q += blockDim.x) seg_start[q] = s_cnt[q];
  }
}

// ---- the fused kernel -----------------------------------------------------------------------------
// Thread mapping.  A CTA (one per SM, 512 threads) holds the telescope lookup tables in shared memory
// (staged once with TMA bulk copies) and runs up to kMaxGroups independent *groups*; a group is
// NWG = ceil((KW + 2) / 60) warps that pull work items (<= psub particles of one spaxel) from a
// global queue.  Inside a group the SSP knot window is laid out along the lanes: lane l of warp wg
// owns the KPL = 2 consecutive knot slots s = 2 * (30 * wg + l - 1) + {0, 1} (lanes 0 and 31 are
// halos that only feed their neighbours through shuffles), slot s <-> SSP index jbase + s.  All
// per-knot constants (lam_z, 1/dlam_z, template offset) live in registers for the whole kernel.
//
// Per particle a lane evaluates, for its two knots: the mass-weighted spectrum S (template rows
// through the read-only path), the shifted position x = lam_z * d, the first telescope channel at or
// above x (k, via the bucket table -- no search, no F2I), the slope m of the segment to the next
// knot and the kink dm = m - m_prev.  The reference's two normalisation sums are evaluated per knot
// segment: total = sum S_j (x_j - x_{j-1}) [x_j in band] and
// new = sum_w p(t_w) dt_w = sum_j S_j D_j + m_j (Q_j - (x_j - tref) D_j), with D_j = t[k_{j+1}-1] -
// t[k_j-1] (dt telescopes exactly in float32) and Q from the double-float prefix table.  One
// transposed warp reduction + one named barrier per batch of NB = 4 particles gives the scales.
//
// Accumulation.  p(t) is piecewise linear, so the spaxel spectrum is kept as (a) per channel cell k the
// summed kinks (sum dm * (t[k-1] - x), sum dm) of all knots that fell into (t[k-1], t[k]] and (b) per
// chunk of CH channels the summed line (value at the chunk's first channel, slope) -- re-anchoring
// every chunk keeps float32 rounding from growing with wavelength distance.  Each warp adds into its
// own cell region (the channel range its knots can reach for the Doppler factors present), so the
// adds are plain shared-memory read-modify-writes in a fixed order: no atomics, bit-reproducible.
// Configurations where knots of one warp can share a channel (SSP grid finer than the telescope's)
// or where the regions do not fit use one shared region and 64-bit CAS adds instead.
// Once per work item the cells are expanded with two warp scans per 32 channels and stored coalesced.
struct FusedLayout {  // shared-memory layout (byte offsets), computed on the host
  int off_mbar, off_lut, off_tt, off_q, off_tc, off_group, group_stride;
  int g_step, g_base, g_rec, g_red, g_misc;  // offsets inside a group's block
  int lut_bytes, tt_bytes, q_bytes;          // multiples of 16 (TMA bulk copy sizes)
  int cap;         // step cells per group
  int nch, chs;    // chunks per row, log2(channels per chunk)
  int max_groups;
  int collide;     // two knots of one warp can fall into the same channel -> CAS adds
  int force_lut;   // use the lookup-table channel search even on an affine grid (RBX_FUSED_FORCE_LUT=1, tests)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void group_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <bool CAS>
__device__ __forceinline__ void cell_add(float2 *cell, float a, float b) {
  if (!CAS) {
    float2 v = *cell;
    v.x += a; v.y += b;
    *cell = v;
  } else {
    unsigned long long *addr = reinterpret_cast<unsigned long long *>(cell);
    unsigned long long old = *addr, assumed;
    do {
      assumed = old;
      float lo = __uint_as_float((unsigned)(assumed & 0xffffffffull)) + a;
      float hi = __uint_as_float((unsigned)(assumed >> 32)) + b;
      unsigned long long nv = ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
      old = atomicCAS(addr, assumed, nv);
    } while (old != assumed);
  }
}

// cell += sc * (a, b)
template <bool CAS>
__device__ __forceinline__ void cell_fma(float2 *cell, float sc, float a, float b) {
  if (!CAS) {
    float2 v = *cell;
    v.x = fmaf(sc, a, v.x); v.y = fmaf(sc, b, v.y);
    *cell = v;
  } else {
    cell_add<true>(cell, sc * a, sc * b);
  }
}

template <int METHOD, bool AFFINE>
__global__ void __launch_bounds__(kCtaThreads, 1)
fused_cube_kernel(PlanView p, const float *__restrict__ rec, const uint32_t *__restrict__ sidx,
                  const Item *__restrict__ items, int *__restrict__ ctrl,
                  float *__restrict__ cube, float *__restrict__ partials, int Wp, FusedLayout lay, int accumulate,
                  CubeLayout cl) {
  constexpr int NT = METHOD == RBX_METHOD_LINEAR ? 1 : 4;   // tables
  constexpr int RS = METHOD == RBX_METHOD_LINEAR ? 8 : 20;  // record stride (floats)
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (ctrl[C_IMPL] != IMPL_GROUP) return;   // segment_kernel selected the warp kernel
  const int chs = ctrl[C_GCHS];                  // >= chs: chosen by segment_kernel for the Doppler range present
  const int nch = (p.W + 1 + (1 << chs) - 1) >> chs;   // <= nch (the shared-memory layout)

  // ---- stage the lookup tables (TMA bulk copies, one elected thread) -----------------------------
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + lay.off_mbar);
  unsigned char *s_lut = smem + lay.off_lut;
  const float2 *s_tt = reinterpret_cast<const float2 *>(smem + lay.off_tt);
  const float2 *s_q = reinterpret_cast<const float2 *>(smem + lay.off_q);
  float *s_tc = reinterpret_cast<float *>(smem + lay.off_tc);   // [nch] wavelength of each chunk's first channel
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t total = AFFINE ? 0u : (uint32_t)(lay.q_bytes + lay.lut_bytes + lay.tt_bytes);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(total) : "memory");
    if (!AFFINE) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(s_q)), "l"(p.q), "r"((uint32_t)lay.q_bytes), "r"(smem_u32(mbar)) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(s_lut)), "l"(p.lut), "r"((uint32_t)lay.lut_bytes), "r"(smem_u32(mbar)) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(s_tt)), "l"(p.tt), "r"((uint32_t)lay.tt_bytes), "r"(smem_u32(mbar)) : "memory");
    }
  }

  // ---- group geometry (runtime: depends on the knot window found by segment_kernel) ---------------
  const int ja = ctrl[C_JA], jb = ctrl[C_JB];
  const int n_items = ctrl[C_NITEMS];
  const int jbase = (ja - 1) & ~1;                 // even; slot s <-> SSP index jbase + s
  const int need = jb - jbase + 1;                 // slots that must be owned
  const int NWG = max(1, (need + KPL * OWN - 1) / (KPL * OWN));
  const int ngroups = min(min(lay.max_groups, kCtaWarps / NWG), kMaxGroups);
  const int grp = warp / NWG, wg = warp - grp * NWG;
  const bool active = grp < ngroups;
  const int gthreads = NWG * 32;
  const int gt = wg * 32 + lane;                   // thread index inside the group

  unsigned char *gbase = smem + lay.off_group + (size_t)(active ? grp : 0) * lay.group_stride;
  float2 *s_step = reinterpret_cast<float2 *>(gbase + lay.g_step);   // [cap]
  float2 *s_base = reinterpret_cast<float2 *>(gbase + lay.g_base);   // [NWG][nch + 32]
  float *s_rec = reinterpret_cast<float *>(gbase + lay.g_rec);       // [2][NB][RS]
  float *s_red = reinterpret_cast<float *>(gbase + lay.g_red);       // [2][2*NB][kMaxGroupWarps]
  int *s_misc = reinterpret_cast<int *>(gbase + lay.g_misc);         // [0]=item, [2..10)=region klo, [12..20)=region khi

  if (active) {
    const int nbase = lay.nch + 32;  // per warp: one line per chunk + one dummy per lane
    for (int q = gt; q < lay.cap + kMaxGroupWarps * 32; q += gthreads) s_step[q] = make_float2(0.f, 0.f);
    for (int q = gt; q < NWG * nbase; q += gthreads) s_base[q] = make_float2(0.f, 0.f);
    for (int q = gt; q < 2 * 2 * NB * kRedStride; q += gthreads) s_red[q] = 0.f;
  }

  // ---- per-lane knot constants ---------------------------------------------------------------------
  const bool owner = lane >= 1 && lane <= OWN;
  const int s0 = KPL * (OWN * wg + lane - 1);      // first slot of this lane (halo lanes included)
  const int j0 = jbase + s0;                       // SSP index of slot 0 (even)
  // template offset of the pair load (clamped into the padded row) and which half each slot reads
  const int jpair = min(max(j0, 0), ((p.L - 1) & ~1));
  float lz[KPL], rdlv[KPL];
  bool take_hi[KPL];
#pragma unroll
  for (int r = 0; r < KPL; ++r) {
    const int j = j0 + r;
    const int jc = min(max(j, 0), p.L - 1);
    take_hi[r] = (jc - jpair) != 0;
    lz[r] = j < 0 ? -1.0e30f : (j >= p.L ? 1.0e30f : p.lamz[j]);
    rdlv[r] = (j < 0 || j >= p.L - 1) ? 0.f : p.rdl[j];
  }
  // previous knot of slot 0 for the "total" difference; diff0's first element is 0 (rubix/spectra/ifu.py:84-102)
  const float lzprev = (j0 - 1 >= 0 && j0 - 1 < p.L) ? p.lamz[j0 - 1] : lz[0];
  const bool pair_plain = !take_hi[0] && take_hi[1];
  const bool any_odd = __any_sync(0xffffffffu, !pair_plain);

  // wait for the tables
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}"
                   : "=r"(done) : "r"(smem_u32(mbar)), "r"(0u) : "memory");
    }
  }
  for (int c = tid; c < nch; c += kCtaThreads) s_tc[c] = p.t[min(c << chs, p.W - 1)];
  __syncthreads();
  if (!active) return;  // spare warps (no __syncthreads below this line)

  // ---- cell regions: the channel range each warp can reach for the Doppler factors present --------
  const float dmin = __int_as_float(kDminBias - ctrl[C_DMIN]), dmax = __int_as_float(ctrl[C_DMAX]);
  {
    float e;
    int klo = owner ? channel_of<AFFINE>(__fmul_rn(lz[0], dmin), p, s_lut, s_tt, e) : 0x7fffffff;
    int khi = owner ? channel_of<AFFINE>(__fmul_rn(lz[KPL - 1], dmax), p, s_lut, s_tt, e) : -1;
    // widest lane span at dmax: must hold at most one chunk start (see the base deposit below)
    float lznext = __shfl_down_sync(0xffffffffu, lz[0], 1);
    int kfirst = channel_of<AFFINE>(__fmul_rn(lz[0], dmax), p, s_lut, s_tt, e);
    int knext = channel_of<AFFINE>(__fmul_rn(lznext, dmax), p, s_lut, s_tt, e);
    int span = owner ? knext - kfirst + 2 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      klo = min(klo, __shfl_xor_sync(0xffffffffu, klo, o));
      khi = max(khi, __shfl_xor_sync(0xffffffffu, khi, o));
      span = max(span, __shfl_xor_sync(0xffffffffu, span, o));
    }
    if (lane == 0) {
      s_misc[2 + wg] = klo;
      s_misc[12 + wg] = khi;
      if (span >= (1 << chs)) atomicExch(ctrl + C_ERROR, 3);
    }
  }
  group_barrier(1 + grp, gthreads);
  // regions are packed back to back: offset of region w = sum of the sizes before it
  int total_cells = 0, my_off = 0;
  for (int w = 0; w < NWG; ++w) {
    if (w == wg) my_off = total_cells;
    total_cells += s_misc[12 + w] - s_misc[2 + w] + 1;
  }
  const bool shared_mode = lay.collide != 0 || total_cells > lay.cap;
  const int my_klo = shared_mode ? 0 : s_misc[2 + wg];
  const int nreg = shared_mode ? 1 : NWG;
  const int nbase = lay.nch + 32;
  // halo lanes add zeros-by-construction into a dummy cell of their own, so phase 2 needs no predicate
  const int own = owner ? 1 : 0;
  float2 *my_cells = s_step + (owner ? (shared_mode ? 0 : my_off - my_klo) : lay.cap + wg * 32 + lane);
  float2 *my_base = s_base + (size_t)wg * nbase;
  const int CH = 1 << chs;
  // chunk lines are summed in registers: over the Doppler range present a lane's first knot moves by
  // less than one chunk, so the chunk start inside its span is chunk cA or cA + 1
  int cA = 0;
  {
    float e;
    const int kmin0 = channel_of<AFFINE>(__fmul_rn(lz[0], dmin), p, s_lut, s_tt, e);
    const int kmax0 = channel_of<AFFINE>(__fmul_rn(lz[0], dmax), p, s_lut, s_tt, e);
    cA = (kmin0 + CH - 1) >> chs;
    if (owner && ((kmax0 + CH - 1) >> chs) > cA + 1) atomicExch(ctrl + C_ERROR, 3);
  }
  float accAv = 0.f, accAm = 0.f, accBv = 0.f, accBm = 0.f;

  const float *tab[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) tab[t] = p.tab[t] + jpair;
  const size_t rowB = (size_t)p.Lp, rowC = (size_t)p.na * p.Lp, rowD = (size_t)(p.na + 1) * p.Lp;

  while (true) {
    group_barrier(1 + grp, gthreads);
    if (gt == 0) s_misc[0] = atomicAdd(ctrl + C_WORK, 1);
    group_barrier(1 + grp, gthreads);
    const int item_id = s_misc[0];
    if (item_id >= n_items) break;
    const Item it = items[item_id];
    const int nbatch = (it.count + NB - 1) / NB;

    // records of batch 0
    for (int q = gt; q < NB * RS; q += gthreads) {
      int b = q / RS;
      const int col = q - b * RS;
      // padding slots repeat the item's first particle with zero weights: their knots stay inside the
      // warps' cell regions and they add exact zeros
      s_rec[q] = (b < it.count) ? rec[(size_t)sidx[it.start + b] * RS + col]
                                : (col < 4 ? rec[(size_t)sidx[it.start] * RS + col] : 0.f);
    }
    group_barrier(1 + grp, gthreads);

    for (int bt = 0; bt < nbatch; ++bt) {
      const int buf = bt & 1;
      const float *r_cur = s_rec + buf * NB * RS;
      if (bt + 1 < nbatch) {  // prefetch the next batch's records
        for (int q = gt; q < NB * RS; q += gthreads) {
          int b = q / RS;
          int pb = (bt + 1) * NB + b;
          const int col = q - b * RS;
          s_rec[(buf ^ 1) * NB * RS + q] = (pb < it.count) ? rec[(size_t)sidx[it.start + pb] * RS + col]
                                                           : (col < 4 ? rec[(size_t)sidx[it.start] * RS + col] : 0.f);
        }
      }

      // ---- phase 1: per-knot quantities and the two normalisation sums ----------------------------
      int ka[NB], kb[NB];              // cell of slot 0 / slot 1 (times `own`)
      float g0[NB], g1[NB], dm0[NB], dm1[NB];
      int cb[NB];                      // chunk line this lane adds: 0 none, 1 -> chunk cA, 2 -> chunk cA + 1
      float bv[NB], bm[NB];            // its line: value at the chunk's first channel, slope
      float red[2 * NB];               // tot[0..NB), new[0..NB)
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float *rb = r_cur + b * RS;
        const float4 r0 = *reinterpret_cast<const float4 *>(rb);  // d, 1/d, row, -
        const float d = r0.x, rd = r0.y;
        const size_t row = (size_t)__float_as_int(r0.z) * p.Lp;
        float Sa = 0.f, Sb = 0.f;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const float4 w = *reinterpret_cast<const float4 *>(rb + 4 + 4 * t);
          const float *f = tab[t] + row;
          const float2 fA = __ldg(reinterpret_cast<const float2 *>(f));
          const float2 fB = __ldg(reinterpret_cast<const float2 *>(f + (METHOD == RBX_METHOD_LINEAR ? rowB : rowC)));
          const float2 fC = __ldg(reinterpret_cast<const float2 *>(f + (METHOD == RBX_METHOD_LINEAR ? rowC : rowB)));
          const float2 fD = __ldg(reinterpret_cast<const float2 *>(f + rowD));
          // linear weights are ordered rows (+0, +1, +na, +na+1); cubic (jj, ii): rows (+0, +na, +1, +na+1)
          Sa = fmaf(w.x, fA.x, Sa); Sb = fmaf(w.x, fA.y, Sb);
          Sa = fmaf(w.y, fB.x, Sa); Sb = fmaf(w.y, fB.y, Sb);
          Sa = fmaf(w.z, fC.x, Sa); Sb = fmaf(w.z, fC.y, Sb);
          Sa = fmaf(w.w, fD.x, Sa); Sb = fmaf(w.w, fD.y, Sb);
        }
        float S0 = Sa, S1 = Sb;
        if (any_odd) {  // lanes at the true ends of the SSP grid: clamped copies (jnp.interp end values)
          S0 = take_hi[0] ? Sb : Sa;
          S1 = take_hi[1] ? Sb : Sa;
        }
        const float x0 = __fmul_rn(lz[0], d), x1 = __fmul_rn(lz[1], d);
        float e0, e1;
        const int k0 = channel_of<AFFINE>(x0, p, s_lut, s_tt, e0);
        const int k1 = channel_of<AFFINE>(x1, p, s_lut, s_tt, e1);
        // right neighbour's first knot, left neighbour's last slope
        const float S2 = __shfl_down_sync(0xffffffffu, S0, 1);
        const int k2 = __shfl_down_sync(0xffffffffu, k0, 1);
        const float e2 = __shfl_down_sync(0xffffffffu, e0, 1);
        const float m0 = (S1 - S0) * rdlv[0] * rd;
        const float m1 = (S2 - S1) * rdlv[1] * rd;
        const float mp = __shfl_up_sync(0xffffffffu, m1, 1);
        // total: sum S_j (x_j - x_{j-1}) over knots inside the band   (rubix/spectra/ifu.py:241-247)
        const float w0 = (x0 >= p.tmin && x0 <= p.tmax) ? x0 - __fmul_rn(lzprev, d) : 0.f;
        const float w1 = (x1 >= p.tmin && x1 <= p.tmax) ? x1 - x0 : 0.f;
        const float tot = fmaf(S1, w1, S0 * w0);
        // new: sum_w p(t_w) dt_w over the channels [k_j, k_{j+1}) of my two segments
        // (rubix/spectra/ifu.py:249-251): S_j D_j + m_j T_j with D_j = sum dt_w = e_{j+1} - e_j (dt
        // telescopes exactly in float32) and T_j = sum dt_w (t_w - x_j).
        const float D0 = e1 - e0, D1 = e2 - e1;
        const float gx0 = e0 - x0, gx1 = e1 - x1;   // t[k-1] - x, in (-dt, 0]
        float T0, T1;
        if (AFFINE) {
          // dt_w (t_w - x) = ((t_w - x)^2 - (t_{w-1} - x)^2) / 2 + dt_w^2 / 2: the squares telescope, and on
          // an arange grid sum dt_w^2 = 2 delta D - n delta^2 up to O(n ulp^2)
          const float n0 = (float)(max(k1, 1) - max(k0, 1)), n1 = (float)(max(k2, 1) - max(k1, 1));
          const float hd2 = 0.5f * p.tdelta * p.tdelta;
          T0 = fmaf(D0, fmaf(0.5f, D0, gx0 + p.tdelta), -hd2 * n0);
          T1 = fmaf(D1, fmaf(0.5f, D1, gx1 + p.tdelta), -hd2 * n1);
        } else {
          // general grid: T_j = Q[k_{j+1}] - Q[k_j] - (x_j - tref) D_j, Q = double-float prefix of dt_w (t_w - tref)
          const float2 q0 = s_q[k0], q1 = s_q[k1];
          const float q2x = __shfl_down_sync(0xffffffffu, q0.x, 1);
          const float q2y = __shfl_down_sync(0xffffffffu, q0.y, 1);
          T0 = fmaf(-(x0 - p.tref), D0, (q1.x - q0.x) + (q1.y - q0.y));
          T1 = fmaf(-(x1 - p.tref), D1, (q2x - q1.x) + (q2y - q1.y));
        }
        float nw = fmaf(m0, T0, S0 * D0);
        nw += fmaf(m1, T1, S1 * D1);
        red[b] = owner ? tot : 0.f;
        red[NB + b] = owner ? nw : 0.f;
        // kinks
        const float d0 = m0 - mp, d1 = m1 - m0;
        dm0[b] = d0; dm1[b] = d1;
        g0[b] = d0 * gx0;
        g1[b] = d1 * gx1;
        ka[b] = k0 * own; kb[b] = k1 * own;
        // chunk base: the line valid at the first chunk start inside [k0, k2)
        const int c = (k0 + CH - 1) >> chs;
        const int chan = c << chs;
        const bool has = owner && chan < k2 && chan < p.W;
        const bool second = chan >= k1;
        const float Sr = second ? S1 : S0, mr = second ? m1 : m0, xr = second ? x1 : x0;
        const float tch = s_tc[min(c, nch - 1)];
        cb[b] = has ? 1 + (c - cA) : 0;
        bv[b] = fmaf(mr, tch - xr, Sr);
        bm[b] = mr;
      }

      // ---- transposed warp reduction of the 2*NB sums ----------------------------------------------
      // after the three folding steps lane l holds value index ((l>>4)&1)*4 + ((l>>3)&1)*2 + ((l>>2)&1)
      float v1;
      int red_idx;
      bool red_writer;
      if (NB == 4) {
        float v4[4], v2[2];
        {
          const bool hi = lane & 16;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float keep = hi ? red[(i + 4) % (2 * NB)] : red[i], send = hi ? red[i] : red[(i + 4) % (2 * NB)];
            v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
        }
        {
          const bool hi = lane & 8;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float keep = hi ? v4[i + 2] : v4[i], send = hi ? v4[i] : v4[i + 2];
            v2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
        }
        {
          const bool hi = lane & 4;
          const float keep = hi ? v2[1] : v2[0], send = hi ? v2[0] : v2[1];
          v1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
        v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
        red_idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        red_writer = (lane & 3) == 0;
      } else {  // NB == 2: four sums
        float v2[2];
        {
          const bool hi = lane & 16;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float keep = hi ? red[(i + 2) % (2 * NB)] : red[i], send = hi ? red[i] : red[(i + 2) % (2 * NB)];
            v2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
        }
        {
          const bool hi = lane & 8;
          const float keep = hi ? v2[1] : v2[0], send = hi ? v2[0] : v2[1];
          v1 = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        v1 += __shfl_xor_sync(0xffffffffu, v1, 4);
        v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
        v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
        red_idx = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
        red_writer = (lane & 7) == 0;
      }
      float *red_cur = s_red + buf * (2 * NB * kRedStride);
      if (red_writer) red_cur[red_idx * kRedStride + wg] = v1;
      group_barrier(1 + grp, gthreads);
      // every warp: lane i < 2*NB sums value i over the group's warps in warp order
      float scale = 0.f;
      {
        // slots of warps beyond NWG stay zero, so the sum always runs over all kMaxGroupWarps entries
        const float4 *rp = reinterpret_cast<const float4 *>(red_cur + (lane & (2 * NB - 1)) * kRedStride);
        const float4 ra = rp[0], rb4 = rp[1];
        const float acc = ((ra.x + ra.y) + (ra.z + ra.w)) + ((rb4.x + rb4.y) + (rb4.z + rb4.w));
        const float nwv = __shfl_down_sync(0xffffffffu, acc, NB);
        scale = nan_to_num0(acc / nwv);  // lanes 0..NB-1: total / new   (rubix/spectra/ifu.py:252-255)
        if (bt * NB + lane >= it.count) scale = 0.f;  // padding slots of the last batch add nothing
      }

      // ---- phase 2: scaled kinks and chunk bases ----------------------------------------------------
      if (!shared_mode) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const float sc = __shfl_sync(0xffffffffu, scale, b);
          cell_fma<false>(my_cells + ka[b], sc, g0[b], dm0[b]);
          cell_fma<false>(my_cells + kb[b], sc, g1[b], dm1[b]);
          const float sA = cb[b] == 1 ? sc : 0.f, sB = cb[b] == 2 ? sc : 0.f;
          accAv = fmaf(sA, bv[b], accAv); accAm = fmaf(sA, bm[b], accAm);
          accBv = fmaf(sB, bv[b], accBv); accBm = fmaf(sB, bm[b], accBm);
          __syncwarp();
        }
      } else {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const float sc = __shfl_sync(0xffffffffu, scale, b);
          cell_fma<true>(my_cells + ka[b], sc, g0[b], dm0[b]);
          cell_fma<true>(my_cells + kb[b], sc, g1[b], dm1[b]);
          const float sA = cb[b] == 1 ? sc : 0.f, sB = cb[b] == 2 ? sc : 0.f;
          accAv = fmaf(sA, bv[b], accAv); accAm = fmaf(sA, bm[b], accAm);
          accBv = fmaf(sB, bv[b], accBv); accBm = fmaf(sB, bm[b], accBm);
          __syncwarp();
        }
      }
    }  // batches
    // flush the register chunk lines, one lane after the other (neighbouring lanes can share a chunk)
    for (int l = 1; l <= OWN; ++l) {
      if (lane == l) {
        if (cA < nch) cell_add<false>(my_base + cA, accAv, accAm);
        if (cA + 1 < nch) cell_add<false>(my_base + cA + 1, accBv, accBm);
      }
      __syncwarp();
    }
    accAv = accAm = accBv = accBm = 0.f;
    group_barrier(1 + grp, gthreads);

    // ---- expand the cells into the spaxel spectrum and store it -----------------------------------
    float *prow = partials + (size_t)max(it.slot, 0) * Wp;
    for (int c = wg; c < nch; c += NWG) {
      float vcar = 0.f, scar = 0.f;
      for (int w = 0; w < NWG; ++w) {
        const float2 bs = s_base[(size_t)w * nbase + c];
        vcar += bs.x; scar += bs.y;
      }
      __syncwarp();
      if (lane < NWG) s_base[(size_t)lane * nbase + c] = make_float2(0.f, 0.f);
      for (int h = 0; h < CH; h += 32) {
        const int ch = (c << chs) + h + lane;
        float A = 0.f, B = 0.f;
        if (ch <= p.W) {
          int off = 0;
          for (int w = 0; w < nreg; ++w) {
            const int klo_w = shared_mode ? 0 : s_misc[2 + w];
            const int size_w = shared_mode ? p.W + 1 : s_misc[12 + w] - klo_w + 1;
            const int idx = ch - klo_w;
            if ((unsigned)idx < (unsigned)size_w) {
              float2 *cellp = s_step + off + idx;
              const float2 cv = *cellp;
              A += cv.x; B += cv.y;
              *cellp = make_float2(0.f, 0.f);
            }
            off += size_w;
          }
        }
        const bool valid = ch < p.W;
        const bool start = (h + lane) == 0;   // the chunk's first channel takes the base line itself
        if (start || !valid) { A = 0.f; B = 0.f; }
        const float dtc = (start || !valid) ? 0.f : __ldg(p.dt + ch);
        // slope after the kinks of this channel, then the value increments
        float sB = B;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float n = __shfl_up_sync(0xffffffffu, sB, o);
          if (lane >= o) sB += n;
        }
        const float s = scar + sB;
        float inc = fmaf(s, dtc, A);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float n = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += n;
        }
        const float v = vcar + inc;
        if (valid) {
          if (it.slot < 0) cube_put(cube, cl, it.spaxel, ch, v, accumulate != 0);
          else prow[ch] = v;
        }
        scar = __shfl_sync(0xffffffffu, s, 31);
        vcar = __shfl_sync(0xffffffffu, v, 31);
      }
    }
  }
}


// ---- warp-per-particle variant ----------------------------------------------------------------------
// One WARP holds the whole knot window: lane l owns the WK = 8 consecutive SSP knots jbase + 8 l + {0..7}
// (256 slots), so the two normalisation sums of a particle are a lane-local sum plus one butterfly
// reduction -- no group barrier, no shared-memory exchange -- and the eight knots of a lane are eight
// independent instruction chains.  Every warp owns a private cell array (W + 1 cells) in shared memory
// and works on its own work item, so the warps of a CTA never synchronise after start-up.  A lane's
// knots fall into distinct cells (SSP spacing > channel spacing, checked on the host; the group kernel
// above with its CAS mode covers everything else), hence the eight read-modify-writes of a particle are
// issued as 8 loads, 16 FMAs, 8 stores: one shared-memory round trip per particle instead of one per
// knot.  Affine (arange) telescope grids only.

struct WarpLayout {
  int off_tc;        // [nch] wavelength of each chunk's first channel
  int off_warp;      // first warp block
  int warp_stride;   // bytes per warp: cells [(W + 2)] float2, then chunk lines [nch] float2
  int w_base;        // offset of the chunk lines inside a warp block
  int w_rec;         // offset of the record batch [32][RS] inside a warp block
  int nch, chs;      // chunks per row, log2(channels per chunk)
  int nwarps;        // warps per CTA
  int pair;          // two warps share one cell array (fused_cube_warp_kernel<.., true>)
  int rec_bytes;     // bytes of one warp's record batch
  int off_item;      // [arrays] work item handed from a pair's first warp to its second
  float skew;        // cell slot = k + floor(max(k - 1, 0) * skew): makes the lane stride an odd number of cells
  int ncells;        // cells per warp (W + 2 + skew of the last one)
  // transposed layout (fused_cube_warp_kernel<.., TR = true>): cells in blocks of tr_B, stored [row][32 columns], so a
  // cell's bank is its BLOCK and the 16 lanes of a half-warp (one lane stride <= tr_B cells apart) never share one
  int tr_B, tr_ncols;   // cells per block (= channels per chunk), blocks in use (<= 26)
  float tr_invB;        // (1 / tr_B) (1 + 2^-20)
  int w_stage;          // offset of the expansion's staging tile [32][8] floats inside a warp block
};

// Shared-memory slot of cell k (bank skew): k + floor(max(k - 1, 0) * alpha), computed in the float domain where
// the knot's cell already lives (KA = kMagic - 1 + k, kc = max(k - 1, 0)): one FFMA rounded down adds the skew to
// the integer in KA's mantissa.
__device__ __forceinline__ int cell_slot(float KA, float kc, float alpha) {
  return __float_as_int(__fmaf_rd(kc, alpha, KA)) - kMagicM1Bits;
}
__device__ __forceinline__ int cell_slot_of(int k, float alpha) {   // the same from the integer cell index (expansion)
  const float KA = (float)k + (kMagic - 1.f);
  return cell_slot(KA, fmaxf(KA, kMagic) - kMagic, alpha);
}

// PAIR: two warps share one cell array and one work item (14 warps per SM instead of 7: the kernel is latency bound
// and 8 private arrays do not fit).  Warp h of a pair takes the item's particles q = h, h + 2, ...; everything up
// to the scale factor is private to the warp, and the short read-modify-write of the cells is done in strictly
// alternating TURNS (w0 particle 0, w1 particle 1, w0 particle 2, ...) handed over with two named barriers per
// pair (bar.sync = wait for my turn, bar.arrive = pass the turn on): no atomics, the order of every addition is
// fixed, the result stays bit-reproducible and does not depend on which kernel variant ran only in rounding.
__device__ __forceinline__ void turn_wait(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void turn_pass(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

template <int METHOD, bool PAIR, int MAXT, bool TR = false>
__global__ void __launch_bounds__(MAXT, 1)
fused_cube_warp_kernel(PlanView p, const float *__restrict__ rec, const uint32_t *__restrict__ sidx,
                       const Item *__restrict__ items, int *__restrict__ ctrl,
                       float *__restrict__ cube, float *__restrict__ partials, int Wp, WarpLayout lay, int accumulate,
                       CubeLayout cl) {
  constexpr int NT = METHOD == RBX_METHOD_LINEAR ? 1 : 4;   // tables
  constexpr int RS = METHOD == RBX_METHOD_LINEAR ? 8 : 20;  // record stride (floats)
#ifndef RBX_PAIR_MAXFOL
#define RBX_PAIR_MAXFOL 1
#endif
  constexpr int MAXFOL = PAIR ? RBX_PAIR_MAXFOL : 2;        // cubic: followers of a run leader (registers: 8 each)
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  static_assert(!TR || PAIR, "the transposed layout is built for the pair kernel");
  if (ctrl[C_IMPL] != (TR ? IMPL_WARP_TR : IMPL_WARP)) return;   // segment_kernel selected another kernel
  const int chs = ctrl[C_CHS];                   // >= chs: chosen by segment_kernel for the Doppler range present
  const int nch = TR ? lay.tr_ncols : (p.W + 1 + (1 << chs) - 1) >> chs;   // <= nch (the shared-memory layout)
  const int TB = lay.tr_B;                       // TR: cells per block = channels per chunk
  // floats between parked records.  TR: a record sits in the unused columns 26.. of one cell row (linear: 32 bytes) or
  // of two rows (cubic: 48 + 32 bytes); voff(u) = float offset of its u-th 16-byte vector
  constexpr int RSS = TR ? (NT > 1 ? 128 : 64) : RS;
  auto voff = [](int u) { return (TR && u >= 3) ? 64 + 4 * (u - 3) : 4 * u; };
  const int arr = PAIR ? warp >> 1 : warp;       // my cell array
  const int half = PAIR ? warp & 1 : 0;          // which warp of the pair
  constexpr int PSTR = PAIR ? 2 : 1;             // particle stride inside an item
  const int t_mine = 1 + 2 * arr + half, t_other = 1 + 2 * arr + (1 - half);   // named barriers: my turn / the other warp's
  float *s_tc = reinterpret_cast<float *>(smem + lay.off_tc);
  volatile int *s_item = reinterpret_cast<volatile int *>(smem + lay.off_item) + arr;
  float2 *cells = reinterpret_cast<float2 *>(smem + lay.off_warp + (size_t)arr * lay.warp_stride);
  float2 *base = reinterpret_cast<float2 *>(smem + lay.off_warp + (size_t)arr * lay.warp_stride + lay.w_base);
  // TR: the cell array is [tr_B rows][32 columns] of which <= 26 columns hold blocks; a warp's 32 parked records sit in
  // the columns 26.. of rows 32 * half .. (linear; cubic: two rows each, rows 64 * half ..); the last row's columns
  // 26..31 are junk cells
  float *s_rec = TR ? reinterpret_cast<float *>(cells + 32 * (32 * (RSS / 64)) * half + 26)
                    : reinterpret_cast<float *>(smem + lay.off_warp + (size_t)arr * lay.warp_stride + lay.w_rec +
                                                (size_t)half * lay.rec_bytes);   // [32][RS]
  const int CH = TR ? TB : 1 << chs;

  // u' of every chunk's first channel, from the actual float32 channel wavelength
  for (int c = tid; c < nch; c += blockDim.x)
    s_tc[c] = (float)(((double)p.t[min(TR ? c * TB : c << chs, p.W - 1)] - (double)p.t0) / (double)p.tdelta - 0.5);
  for (int q = lane + 32 * half; q < lay.ncells; q += 32 * PSTR) cells[q] = make_float2(0.f, 0.f);
  for (int q = lane + 32 * half; q < nch; q += 32 * PSTR) base[q] = make_float2(0.f, 0.f);
  if (PAIR && half == 0 && lane == 0) *s_item = atomicAdd(ctrl + C_WORK, 1);   // the pair's first work item
  __syncthreads();

  // ---- per-lane knot constants -------------------------------------------------------------------
  const int n_items = ctrl[C_NITEMS];
  const int jbase = p.wt_jbase;                  // slot s <-> knot jbase + s: the plan's window tables (plan.cu)
  const int j0 = jbase + WK * lane;
  float au[WK], bu[WK];                           // u' = au * (d - 1) + bu: knot position in channel units (see knot_ab)
  float rdlu[WK];                                 // delta / (lam_z[j+1] - lam_z[j]): slope per channel = dS * rdlu / d
  float dl[WK];                                   // lam_z[j] - lam_z[j-1]: total_lum = d * sum S_j dl_j over the knots in band
#pragma unroll
  for (int r = 0; r < WK; ++r) {
    const int j = j0 + r;
    const KnotAB ab = knot_ab(p, j);
    au[r] = ab.a; bu[r] = ab.b;
    rdlu[r] = (j < 0 || j >= p.L - 1) ? 0.f : p.rdl[j] * p.tdelta;
    // diff0's first element is 0 (rubix/spectra/ifu.py:84-102)
    dl[r] = (j >= 1 && j < p.L) ? p.lamz[j] - p.lamz[j - 1] : 0.f;
  }
  if (lane == 31) rdlu[WK - 1] = 0.f;             // the segment leaving the window (always beyond the band)
  const float kahi = kMagic + (float)(p.W - 1);
  const float kain = kMagic + (float)(p.W - 2);   // kMagic <= KA <= kain: the knot lies in the band (t_0, t_{W-1}]

  const float dmin = __int_as_float(kDminBias - ctrl[C_DMIN]), dmax = __int_as_float(ctrl[C_DMAX]);
  // chunk lines are summed in registers: a lane's span [k_0, k_8) holds at most one chunk start, and over
  // the Doppler range present that start is chunk cA or cA + 1
  int cA = 0;
  if (TR) cA = (warp_lane_cells(p, jbase, lane, eps_lo_of(dmin), eps_hi_of(dmax)).kmin0 + TB - 1) / TB;
  else warp_lane_chunks(p, jbase, lane, eps_lo_of(dmin), eps_hi_of(dmax), chs, cA);   // segment_kernel checked that this holds
  // TR: slot of cell k = 32 k - blk (32 B - 1), blk = floor(k / B) from one rounded FMA
  const float trHalf = -(kMagic - 1.f) - 0.5f * (float)TB;   // KA + trHalf = k - B / 2 (exact)
  // shared-memory byte address of cell k: 256 k - blk (256 B - 8) + cells = 256 bits(KA) - (256 B - 8) bits(blkf) + trC
  const unsigned trM = 256u * (unsigned)TB - 8u;
  const unsigned trC = smem_u32(cells) - 256u * (unsigned)kMagicM1Bits + trM * (unsigned)kMagicBits;
  // Lanes whose knots lie above the band for every particle would all hit cell W, which shares its bank with the
  // last block's real cells (a 2-way conflict in every instruction): they get junk slots of their own in the unused
  // columns 26.. of the last row.  (Lanes below the band share the junk slot of block -1: column 31, a free bank.)
  unsigned trAnd = 0xffffffffu, trOr = 0u;
  if (TR) {
    const LaneCells lc0 = warp_lane_cells(p, jbase, lane, eps_lo_of(dmin), eps_hi_of(dmax));
    const bool dead_hi = lc0.kmin0 >= p.W;
    const unsigned dm = __ballot_sync(0xffffffffu, dead_hi);
    if (dead_hi) {
      const int idx = min(lane - (__ffs(dm) - 1), 4);
      trAnd = 0u;
      trOr = smem_u32(cells) + 8u * (unsigned)(32 * (TB - 1) + 26 + idx);
    }
  }
  float accAv = 0.f, accAm = 0.f, accBv = 0.f, accBm = 0.f;

  const float *tab[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) tab[t] = p.wt[t] + 4 * lane;
  // window tables: a row is 256 floats (knots clamped to the SSP grid: jnp.interp's end values); my knots 0..3 sit at
  // float 4 * lane, my knots 4..7 at 128 + 4 * lane, so every warp-wide 16-byte load is one contiguous 512-byte block
  const size_t rowB = (size_t)kWarpSlots, rowC = (size_t)p.na * kWarpSlots, rowD = (size_t)(p.na + 1) * kWarpSlots;

  // The next work item is popped, and its first record batch requested, BEFORE the current item's cells are
  // expanded: the queue atomic and two dependent global loads hide behind the expansion.
  float4 nrec[RS / 4];   // my record of the NEXT batch, in flight during the current one
  // PAIR: my particles of an item are q = half, half + 2, ...
  auto my_count = [&](const Item &w) { return PAIR ? (w.count + 1 - half) >> 1 : w.count; };
  auto fetch_rec = [&](const Item &w, int b0) {
    if (b0 + lane < my_count(w)) {
      const float4 *src =
          reinterpret_cast<const float4 *>(rec + (size_t)__ldg(sidx + w.start + half + PSTR * (b0 + lane)) * RS);
#pragma unroll
      for (int u = 0; u < RS / 4; ++u) nrec[u] = __ldg(src + u);
    }
  };
  float4 nf[8];          // template vectors of the next (particle, table) group
  int nf_row = -1;       // linear: the template row they were loaded for
  Item it_next;
  it_next.start = 0; it_next.count = 0; it_next.spaxel = 0; it_next.slot = -1;
  auto take_item = [&](int id) -> bool {
    if (id >= n_items) return false;
    it_next = items[id];
    fetch_rec(it_next, 0);
    return true;
  };
  auto pop_item = [&]() -> bool {   // !PAIR
    int id = 0;
    if (lane == 0) id = atomicAdd(ctrl + C_WORK, 1);
    id = __shfl_sync(0xffffffffu, id, 0);
    return take_item(id);
  };
  bool have = PAIR ? take_item(*s_item) : pop_item();
  if (PAIR && half == 1) turn_pass(t_other);   // the first turn is warp 0's

  while (have) {
    const Item it = it_next;

    // Software pipeline.  Records: a batch of 32 sorted records is loaded coalesced (lane l loads record l)
    // and parked in the warp's shared-memory slot, so every particle's record is a broadcast read away.
    // Template rows: the eight 16-byte vectors of the NEXT (particle, table) group are in flight while the
    // current group is folded in and, for the last table, during the whole knot arithmetic of the particle.
    const int cnt = my_count(it);
    for (int b0 = 0; b0 < cnt; b0 += 32) {
      const int nb = min(32, cnt - b0);
      __syncwarp();
      if (lane < nb) {
#pragma unroll
        for (int u = 0; u < RS / 4; ++u) *reinterpret_cast<float4 *>(s_rec + lane * RSS + voff(u)) = nrec[u];
      }
      __syncwarp();
      if (b0 + 32 < cnt) fetch_rec(it, b0 + 32);
      auto issue_rows = [&](int i, int t) {   // rows of particle i (in this batch), table t
        const int rw = __float_as_int(s_rec[i * RSS + 2]);
        // linear: particles are sorted by (spaxel, template cell), so runs of particles share their four rows;
        // the vectors already in registers are kept (at 10^7 particles ~90 % of the row loads go away)
        if (NT == 1 && rw == nf_row) return;
        nf_row = rw;
        const float *f = tab[t] + (size_t)rw * kWarpSlots;
        // linear weights are ordered rows (+0, +1, +na, +na+1); cubic (jj, ii): rows (+0, +na, +1, +na+1)
        const size_t o1 = METHOD == RBX_METHOD_LINEAR ? rowB : rowC, o2 = METHOD == RBX_METHOD_LINEAR ? rowC : rowB;
        nf[0] = __ldg(reinterpret_cast<const float4 *>(f));
        nf[1] = __ldg(reinterpret_cast<const float4 *>(f + 128));
        nf[2] = __ldg(reinterpret_cast<const float4 *>(f + o1));
        nf[3] = __ldg(reinterpret_cast<const float4 *>(f + o1 + 128));
        nf[4] = __ldg(reinterpret_cast<const float4 *>(f + o2));
        nf[5] = __ldg(reinterpret_cast<const float4 *>(f + o2 + 128));
        nf[6] = __ldg(reinterpret_cast<const float4 *>(f + rowD));
        nf[7] = __ldg(reinterpret_cast<const float4 *>(f + rowD + 128));
      };
      issue_rows(0, 0);
      float S2[WK], S3[WK];      // cubic: spectra of the followers in a run of particles of one template cell
      int run_pending = 0, run_len = 1;
#pragma unroll
      for (int r = 0; r < WK; ++r) { S2[r] = 0.f; S3[r] = 0.f; }
    for (int qi = 0; qi < nb; ++qi) {
      const float *rb = s_rec + qi * RSS;
      const float4 r0 = *reinterpret_cast<const float4 *>(rb);   // d, 1/d, row, d - 1
      const float d = r0.x, rd = r0.y, eps = r0.w;

      // ---- mass-weighted spectrum at my eight knots ------------------------------------------------
      // Cubic (16 rows per particle, no room to keep them in registers): when the next one or two particles read
      // the same template cell, their spectra (S2, S3) are accumulated from the same row vectors and they skip
      // their own loads.
      float S[WK + 1];
      const int follower = NT > 1 ? run_pending : 0;   // 0: leads a run, 1 / 2: first / second follower
      int run_next = 0;
      if (follower == 1) {
#pragma unroll
        for (int r = 0; r < WK; ++r) S[r] = S2[r];
        run_next = run_len > 2 ? 2 : 0;
      } else if (follower == 2) {
#pragma unroll
        for (int r = 0; r < WK; ++r) S[r] = S3[r];
      } else {
        int nfol = 0;   // followers of this particle
        if (NT > 1 && qi + 1 < nb && __float_as_int(s_rec[(qi + 1) * RSS + 2]) == __float_as_int(r0.z)) {
          nfol = 1;
          if (MAXFOL >= 2 && qi + 2 < nb && __float_as_int(s_rec[(qi + 2) * RSS + 2]) == __float_as_int(r0.z)) nfol = 2;
        }
        run_len = 1 + nfol;
        run_next = nfol ? 1 : 0;
        const float *rb2 = rb + RSS, *rb3 = rb + 2 * RSS;
#pragma unroll
        for (int r = 0; r < WK; ++r) { S[r] = 0.f; S2[r] = 0.f; S3[r] = 0.f; }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const float4 w = *reinterpret_cast<const float4 *>(rb + voff(1 + t));
          const float wv[4] = {w.x, w.y, w.z, w.w};
          float4 w2 = make_float4(0.f, 0.f, 0.f, 0.f), w3 = w2;
          if (nfol >= 1) w2 = *reinterpret_cast<const float4 *>(rb2 + voff(1 + t));
          if (MAXFOL >= 2 && nfol >= 2) w3 = *reinterpret_cast<const float4 *>(rb3 + voff(1 + t));
          const float wv2[4] = {w2.x, w2.y, w2.z, w2.w}, wv3[4] = {w3.x, w3.y, w3.z, w3.w};
          {
            // cubic pairs (168 registers): the vectors are consumed in place and the next table's are requested after
            // the FMAs (the other warps of the SM cover the latency; a second register set would spill)
#ifndef RBX_CUBIC_INPLACE
#define RBX_CUBIC_INPLACE 1
#endif
            constexpr bool INPLACE = NT > 1 && PAIR && RBX_CUBIC_INPLACE;
            float4 cur[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) cur[u] = nf[u];
            auto issue_next = [&]() {
              if (t + 1 < NT) {
                issue_rows(qi, t + 1);
              } else {
                const int nq = qi + 1 + nfol;
                if (nq < nb) issue_rows(nq, 0);
              }
            };
            if (!INPLACE) issue_next();
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              const float4 lo = cur[2 * a], hi = cur[2 * a + 1];
              ffma2s(S[0], S[1], wv[a], wv[a], lo.x, lo.y);
              ffma2s(S[2], S[3], wv[a], wv[a], lo.z, lo.w);
              ffma2s(S[4], S[5], wv[a], wv[a], hi.x, hi.y);
              ffma2s(S[6], S[7], wv[a], wv[a], hi.z, hi.w);
            }
            if (NT > 1 && nfol >= 1) {
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                const float4 lo = cur[2 * a], hi = cur[2 * a + 1];
                ffma2s(S2[0], S2[1], wv2[a], wv2[a], lo.x, lo.y);
                ffma2s(S2[2], S2[3], wv2[a], wv2[a], lo.z, lo.w);
                ffma2s(S2[4], S2[5], wv2[a], wv2[a], hi.x, hi.y);
                ffma2s(S2[6], S2[7], wv2[a], wv2[a], hi.z, hi.w);
              }
            }
            if (NT > 1 && MAXFOL >= 2 && nfol >= 2) {
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                const float4 lo = cur[2 * a], hi = cur[2 * a + 1];
                ffma2s(S3[0], S3[1], wv3[a], wv3[a], lo.x, lo.y);
                ffma2s(S3[2], S3[3], wv3[a], wv3[a], lo.z, lo.w);
                ffma2s(S3[4], S3[5], wv3[a], wv3[a], hi.x, hi.y);
                ffma2s(S3[6], S3[7], wv3[a], wv3[a], hi.z, hi.w);
              }
            }
            if (INPLACE) issue_next();
          }
        }
      }
      run_pending = run_next;

      // ---- knot positions in channel units, first channel at or above each knot ---------------------------
      // KA = kMagic - 1 + (channels below the knot): the knot's cell, kept as a float (compares, skew and the
      // shared-memory address all work on it; no float <-> int conversion anywhere)
      // The float arithmetic of two neighbouring knots goes through the packed f32x2 forms (one issue slot for two
      // IEEE operations; each half is an ordinary round-to-nearest op, so the results are those of the scalar form).
      float u[WK], KA[WK + 1], kc[WK + 1], g[WK];
#pragma unroll
      for (int r = 0; r < WK; r += 2) {
        ffma2o(u[r], u[r + 1], au[r], au[r + 1], eps, eps, bu[r], bu[r + 1]);
        float kb0, kb1;
        fadd2s(kb0, kb1, u[r], u[r + 1], kMagic, kMagic);
        KA[r] = fminf(fmaxf(kb0, kMagic - 1.f), kahi);
        KA[r + 1] = fminf(fmaxf(kb1, kMagic - 1.f), kahi);
      }
      S[WK] = __shfl_down_sync(0xffffffffu, S[0], 1);
      KA[WK] = __shfl_down_sync(0xffffffffu, KA[0], 1);
      if (lane == 31) { S[WK] = S[WK - 1]; KA[WK] = KA[WK - 1]; }
#pragma unroll
      for (int r = 0; r < WK; r += 2)   // last channel below the knot, 0 below the band
        fadd2s(kc[r], kc[r + 1], fmaxf(KA[r], kMagic), fmaxf(KA[r + 1], kMagic), -kMagic, -kMagic);
      kc[WK] = fmaxf(KA[WK], kMagic) - kMagic;

      // ---- slopes, the two normalisation sums (rubix/spectra/ifu.py:241-251) -------------------------
      //   total = sum S_j (x_j - x_{j-1}) [x_j in band] = d * sum S_j dl_j
      //   new = sum_w p(t_w) dt_w = delta * sum_j D_j (S_j + mu_j (D_j / 2 + g_j)), over the channels [k_j, k_{j+1}):
      //   D_j channels, mu_j the slope per channel and g_j = kc_j - u'_j = (t[k_j - 1] - x_j) / delta + 1/2
      float mu[WK];
      float tot = 0.f, nw = 0.f, nw1 = 0.f;
#pragma unroll
      for (int r = 0; r < WK; r += 2) {
        const float dS0 = S[r + 1] - S[r], dS1 = S[r + 2] - S[r + 1];
        float t0, t1;
        fmul2s(t0, t1, dS0, dS1, rdlu[r], rdlu[r + 1]);
        fmul2s(mu[r], mu[r + 1], t0, t1, rd, rd);
        ffma2o(g[r], g[r + 1], u[r], u[r + 1], -1.f, -1.f, kc[r], kc[r + 1]);
        // knots in the band [t_0, t_{W-1}] (1 <= cell <= W - 1): a predicated FFMA
#pragma unroll
        for (int q = r; q < r + 2; ++q)
          asm("{\n.reg .pred p, q;\nsetp.ge.f32 p, %1, %2;\nsetp.le.and.f32 q, %1, %3, p;\n@q fma.rn.f32 %0, %4, %5, %0;\n}"
              : "+f"(tot) : "f"(KA[q]), "f"(kMagic), "f"(kain), "f"(S[q]), "f"(dl[q]));
        const float D0 = kc[r + 1] - kc[r], D1 = kc[r + 2] - kc[r + 1];
        float Tp0, Tp1, in0, in1;
        ffma2o(Tp0, Tp1, D0, D1, 0.5f, 0.5f, g[r], g[r + 1]);
        ffma2o(in0, in1, mu[r], mu[r + 1], Tp0, Tp1, S[r], S[r + 1]);
        ffma2s(nw, nw1, D0, D1, in0, in1);       // even and odd knots in two accumulators
      }
      nw += nw1;
      float mp = __shfl_up_sync(0xffffffffu, mu[WK - 1], 1);
      if (lane == 0) mp = mu[0];   // slot 0 of the window lies below the band for every Doppler factor present

      // Everything that does not need the scale is issued BEFORE the reduction, so that the cell loads and
      // the chunk-line selection overlap the shuffle latency.
      // ---- kinks and their cells: 8 loads now, 16 FMAs + 8 stores after the scale ------------------------
      float2 cv[WK];
      float dmv[WK];
      int ka[WK];
      if (TR) {
        // slot = 32 (k mod B) + floor(k / B): the bank is the block.  floor(k / B) = RN((k - B / 2) / B) (the factor
        // 1 + 2^-20 in tr_invB breaks the ties upwards); k = 0 (a knot below the band) gives block -1 = the junk slot
#pragma unroll
        for (int r = 0; r < WK; r += 2) {
          float h0, h1, b0, b1;
          fadd2s(h0, h1, KA[r], KA[r + 1], trHalf, trHalf);
          ffma2o(b0, b1, h0, h1, lay.tr_invB, lay.tr_invB, kMagic, kMagic);
          ka[r] = (int)(((256u * (unsigned)__float_as_int(KA[r]) - trM * (unsigned)__float_as_int(b0) + trC) & trAnd) | trOr);
          ka[r + 1] = (int)(((256u * (unsigned)__float_as_int(KA[r + 1]) - trM * (unsigned)__float_as_int(b1) + trC) & trAnd) | trOr);
        }
      } else {
#pragma unroll
        for (int r = 0; r < WK; ++r) ka[r] = cell_slot(KA[r], kc[r], lay.skew);
      }
#ifdef RBX_FAKE_BANKS
      // timing experiment only (wrong results): what would conflict-free cell updates be worth?
#pragma unroll
      for (int r = 0; r < WK; ++r) if (!TR) ka[r] = (ka[r] & ~31) + lane;
#endif
#ifdef RBX_RACECHECK
      // Knots outside the band (k = 0 or k = W) of several lanes land in the two junk cells 0 and W, which are
      // never read back: a benign write-write overlap that compute-sanitizer racecheck reports.  This build
      // gives every lane its own junk cell so that racecheck can show there is no other hazard.
#pragma unroll
      for (int r = 0; r < WK; ++r)
        if (!TR && (KA[r] < kMagic || KA[r] >= kahi)) ka[r] = lay.ncells + lane;
#endif
      if (!PAIR) {
#pragma unroll
        for (int r = 0; r < WK; ++r) cv[r] = cells[ka[r]];
      }
#pragma unroll
      for (int r = 0; r < WK; ++r) dmv[r] = mu[r] - (r == 0 ? mp : mu[r - 1]);
      // ---- chunk line: the segment valid at the first chunk start inside [k_0, k_8) ---------------------
      float bv, mr;
      int sel;
      {
        int c;
        float chanKA;
        bool has;
        if (TR) {   // c = ceil(k_0 / B) = RN((k_0 + B - 1 - B / 2) / B), in the float domain like the cell index
          const float cf = fmaf(KA[0] + (trHalf + (float)(TB - 1)), lay.tr_invB, kMagic);
          c = __float_as_int(cf) - kMagicBits;
          chanKA = fmaf(cf - kMagic, (float)TB, kMagic - 1.f);
          has = chanKA < KA[WK] && chanKA < (kMagic - 1.f) + (float)p.W;
        } else {
          c = (knot_cell(KA[0]) + CH - 1) >> chs;
          const int chan = c << chs;
          chanKA = (float)chan + (kMagic - 1.f);
          has = chanKA < KA[WK] && chan < p.W;
        }
        float Sr = S[0], ur = u[0];
        mr = mu[0];
#pragma unroll
        for (int r = 1; r < WK; ++r) {
          const bool take = KA[r] <= chanKA;
          Sr = take ? S[r] : Sr; mr = take ? mu[r] : mr; ur = take ? u[r] : ur;
        }
        const float tch = s_tc[min(c, nch - 1)];
        bv = fmaf(mr, tch - ur, Sr);
        sel = has ? 1 + (c - cA) : 0;
      }

      // ---- total / new   (rubix/spectra/ifu.py:252-255) -----------------------------------------------
      // both sums in one butterfly: after the first exchange the lower half-warp carries `total`, the upper `new`
      {
        const bool lowh = lane < 16;
        float v = lowh ? tot : nw;
        const float w = lowh ? nw : tot;
        v += __shfl_xor_sync(0xffffffffu, w, 16);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        tot = __shfl_sync(0xffffffffu, v, 0);
        nw = __shfl_sync(0xffffffffu, v, 16);
      }
      const float sc = nan_to_num0((d * tot) * p.tinv / nw);

      {
        const float sA = sel == 1 ? sc : 0.f, sB = sel == 2 ? sc : 0.f;
        accAv = fmaf(sA, bv, accAv); accAm = fmaf(sA, mr, accAm);
        accBv = fmaf(sB, bv, accBv); accBm = fmaf(sB, mr, accBm);
      }
      // cell k: (sum dmu (g + 1/2), sum dmu) in channel units; the expansion takes the 1/2 out again
      float ag[WK];
#pragma unroll
      for (int r = 0; r < WK; r += 2) fmul2s(ag[r], ag[r + 1], dmv[r], dmv[r + 1], g[r], g[r + 1]);
      if (PAIR) {
        turn_wait(t_mine);   // the cells are mine until turn_pass
        if (TR) {   // ka holds shared-memory byte addresses
#pragma unroll
          for (int r = 0; r < WK; ++r)
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(cv[r].x), "=f"(cv[r].y) : "r"(ka[r]) : "memory");
        } else {
#pragma unroll
          for (int r = 0; r < WK; ++r) cv[r] = cells[ka[r]];
        }
      }
#pragma unroll
      for (int r = 0; r < WK; ++r) ffma2s(cv[r].x, cv[r].y, sc, sc, ag[r], dmv[r]);
      if (TR) {
#pragma unroll
        for (int r = 0; r < WK; ++r)
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(ka[r]), "f"(cv[r].x), "f"(cv[r].y) : "memory");
      } else {
#pragma unroll
        for (int r = 0; r < WK; ++r) cells[ka[r]] = cv[r];
      }
      if (PAIR) turn_pass(t_other); else __syncwarp();
    }  // particles
    }  // record batches
    if (PAIR && cnt < ((it.count + 1) >> 1)) {   // odd item: the second warp's empty last turn keeps the order
      turn_wait(t_mine);
      turn_pass(t_other);
    }

    // flush the register chunk lines: cA is non-decreasing along the lanes, so lanes that share a chunk form
    // runs; a segmented shuffle scan sums each run in a fixed order and its last lane adds the total
    {
      auto flush = [&](int key, float v, float m) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int kk = __shfl_up_sync(0xffffffffu, key, o);
          const float vv = __shfl_up_sync(0xffffffffu, v, o), mm = __shfl_up_sync(0xffffffffu, m, o);
          if (lane >= o && kk == key) { v += vv; m += mm; }
        }
        const int knext = __shfl_down_sync(0xffffffffu, key, 1);
        if ((lane == 31 || knext != key) && key < nch) cell_add<false>(base + key, v, m);
        __syncwarp();
      };
      if (PAIR) turn_wait(t_mine);   // one more turn each: the chunk lines go into the pair's shared array
      flush(cA, accAv, accAm);
      flush(cA + 1, accBv, accBm);
      if (PAIR) {
        // warp 0 pops the pair's next item inside its turn; warp 1 reads it inside its own (after warp 0's)
        if (half == 0 && lane == 0) *s_item = atomicAdd(ctrl + C_WORK, 1);
        __syncwarp();
        const int id = half == 1 ? *s_item : 0;
        turn_pass(t_other);
        if (half == 0) {
          turn_wait(t_mine);         // warp 1's flush is done: every turn of this item is complete
          turn_pass(t_other);        // acknowledge: warp 1 may go on (no barrier ever collects two arrivals of one warp)
          have = take_item(*s_item);
        } else {
          turn_wait(t_mine);         // warp 0 has seen the end of my flush
          have = take_item(id);
        }
      }
    }
    accAv = accAm = accBv = accBm = 0.f;
    if (!PAIR) have = pop_item();

    // ---- expand the cells into the spaxel spectrum and store it ---------------------------------------
    // Per chunk: slope_k = scar + sum_{j<=k} B_j and value_k = vcar + sum_{j<=k} (slope_j dt_j + A_j)
    //   = vcar + scar (t_k - t_ref) + sum_{j<=k} (sB_j dt_j + A_j)      (sum dt_j telescopes exactly)
    // so the two warp scans of a 32-channel group do not depend on the carries: four groups are scanned
    // at once (four independent shuffle chains) and the carries are applied afterwards.
    float *prow = partials + (size_t)max(it.slot, 0) * Wp;
    if (TR) {
      // Transposed layout: lane b walks block b = chunk b (channels b B .. b B + B - 1) row by row -- at every step the
      // 32 lanes read 32 different banks -- starting from the chunk's line, so there are no carries between lanes and
      // no scans: slope_k = slope_{k-1} + B_k, value_k = value_{k-1} + slope_k dt_k + A_k.  Eight rows at a time go
      // through a staging tile so that the stores are runs of eight consecutive channels.  The pair's first warp does
      // this alone (the second one has nothing to wait for but its next turn).
      if (half == 0) {
        float *stage = reinterpret_cast<float *>(smem + lay.off_warp + (size_t)arr * lay.warp_stride + lay.w_stage);
        const int blk = lane;
        const bool active = blk < nch;
        float v = 0.f, sl = 0.f;
        if (active) { const float2 bs = base[blk]; v = bs.x; sl = bs.y; base[blk] = make_float2(0.f, 0.f); }
        const int ch0 = blk * TB;
        float tprev = 0.f;
        for (int i0 = 0; i0 < TB; i0 += 8) {
          float out[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int row = i0 + r, ch = ch0 + row;
            float2 cv = make_float2(0.f, 0.f);
            if (active) { cv = cells[32 * row + blk]; cells[32 * row + blk] = make_float2(0.f, 0.f); }
            // t_w = fl(fl(w delta) + t0): what rbx_plan_create verified for this grid (PlanView::affine)
            const float tk = __fadd_rn(__fmul_rn((float)ch, p.tdelta), p.t0);
            const bool use = row > 0 && ch < p.W;   // the chunk's first channel takes the line itself
            const float dtc = use ? __fmul_rn(__fsub_rn(tk, tprev), p.tinv) : 0.f;
            const float A = use ? fmaf(-0.5f, cv.y, cv.x) : 0.f;   // kinks: offset in channels, slope per channel
            sl += use ? cv.y : 0.f;
            v = fmaf(sl, dtc, v + A);
            tprev = tk;
            out[r] = v;
          }
          __syncwarp();
          float4 *st = reinterpret_cast<float4 *>(stage + lane * 8);
          st[0] = make_float4(out[0], out[1], out[2], out[3]);
          st[1] = make_float4(out[4], out[5], out[6], out[7]);
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int run = q * 4 + (lane >> 3), e = lane & 7;
            const int ch = run * TB + i0 + e;
            const float val = stage[run * 8 + e];
            if (run < nch && ch < p.W) {
              if (it.slot < 0) cube_put(cube, cl, it.spaxel, ch, val, accumulate != 0);
              else prow[ch] = val;
            }
          }
        }
      }
    } else
    for (int c = half; c < nch; c += PSTR) {   // PAIR: the two warps expand alternate chunks
      const float2 bs = base[c];
      float vcar = bs.x, scar = bs.y;
      __syncwarp();
      if (lane == 0) base[c] = make_float2(0.f, 0.f);
      const int cstart = c << chs;
      for (int h = 0; h < CH; h += 128) {
        if (cstart + h > p.W) break;
        float A[4], sB[4], dtc[4], tk[4], tref[4];
        bool valid[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int gstart = cstart + h + 32 * g;
          const int ch = gstart + lane;
          A[g] = 0.f; sB[g] = 0.f;
          if (ch <= p.W) {
            const int ca = cell_slot_of(ch, lay.skew);
            const float2 cvv = cells[ca];
            A[g] = fmaf(-0.5f, cvv.y, cvv.x); sB[g] = cvv.y;   // kinks (offset in channels, slope per channel)
            cells[ca] = make_float2(0.f, 0.f);
          }
          valid[g] = ch < p.W;
          const bool start = ch == cstart;      // the chunk's first channel takes the base line itself
          if (start || !valid[g]) { A[g] = 0.f; sB[g] = 0.f; }
          // channel widths / distances in units of delta, from the actual float32 channel wavelengths
          dtc[g] = (start || !valid[g]) ? 0.f : __ldg(p.dt + ch) * p.tinv;
          tk[g] = __ldg(p.t + min(ch, p.W - 1));
          tref[g] = __ldg(p.t + min(max(gstart - 1, cstart), p.W - 1));
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float n = __shfl_up_sync(0xffffffffu, sB[g], o);
            if (lane >= o) sB[g] += n;
          }
        }
        float inc[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) inc[g] = fmaf(sB[g], dtc[g], A[g]);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float n = __shfl_up_sync(0xffffffffu, inc[g], o);
            if (lane >= o) inc[g] += n;
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int ch = cstart + h + 32 * g + lane;
          const float v = vcar + fmaf(scar, (tk[g] - tref[g]) * p.tinv, inc[g]);
          if (valid[g]) {
            if (it.slot < 0) cube_put(cube, cl, it.spaxel, ch, v, accumulate != 0);
            else prow[ch] = v;
          }
          scar += __shfl_sync(0xffffffffu, sB[g], 31);
          vcar = __shfl_sync(0xffffffffu, v, 31);
        }
      }
    }
    __syncwarp();
    // PAIR: my half of the cells is clear again; that is warp 0's token for the first turn of the next item
    // (warp 1's first turn follows warp 0's, which follows warp 0's own expansion)
    if (PAIR && half == 1 && have) turn_pass(t_other);
  }
}

// cube[s] = sum over the spaxel's partial rows, in sub-range order (deterministic two-level reduction).
// A configuration the kernels cannot hold (knot window too large) poisons the cube with NaN here rather than
// returning a silently wrong result.
__global__ void reduce_partials_kernel(const int *__restrict__ slot_start, const int *__restrict__ split_list,
                                       const float *__restrict__ partials, int Wp, int W, int nseg,
                                       const int *__restrict__ ctrl, float *__restrict__ cube, int accumulate, CubeLayout cl) {
  if (ctrl[C_ERROR] != 0) {   // poison
    for (int s = blockIdx.y; s < nseg; s += gridDim.y)
      for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < W; w += gridDim.x * blockDim.x)
        cube_put(cube, cl, s, w, __int_as_float(0x7fc00000), false);
    return;
  }
  // only the spaxels that were cut into several items have rows to add (segment_kernel lists them: at 150 x 150
  // a few hundred of 22500)
  const int nsplit = ctrl[C_NSPLITSEG];
  for (int e = blockIdx.y; e < nsplit; e += gridDim.y) {
    const int s = split_list[e];
    const int i0 = slot_start[s], i1 = slot_start[s + 1];
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < W; w += gridDim.x * blockDim.x) {
      float acc = accumulate ? cube_get(cube, cl, s, w) : 0.f;
      // eight rows requested at once, added in row order (the kernel was bound by one dependent load per row)
      for (int k = i0; k < i1; k += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = k + u < i1 ? __ldg(partials + (size_t)(k + u) * Wp + w) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (k + u < i1) acc += v[u];
      }
      cube_put(cube, cl, s, w, acc, false);
    }
  }
}

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Particles per bulk work item.  Measured on B200 with the 12-warp kernel (tools/gpu_psub.sh, profiles/r02_item_sweep.txt):
// every item costs a fixed expansion (3721 channels, two warps), so few large items win as long as the queue still
// holds several items per warp pair and the quarter-size tail items keep the end of the persistent kernel short:
// 512 up to 2.5 10^6 particles, 1024 at 5 10^6, 2048 at 10^7 (about 4000 .. 5000 bulk items).
static int choose_psub(int64_t n) {
  if (opt(OPT_PSUB) > 0) return (int)std::max<int64_t>(32, opt(OPT_PSUB));
  const int64_t t = n / 3500;
  int ps = 256;
  while (2 * ps <= t && ps < 2048) ps <<= 1;
  if (n >= 600000) ps = std::max(ps, 512);
  return ps;
}

static int layout_workspace(const rbx_plan *plan, int64_t n, int nseg, void *base, FusedWs &ws, size_t &total) {
  const PlanView &v = plan->v;
  // Sort key: spaxel in the high bits (that is all correctness needs: the sort is stable, so the order
  // inside a spaxel is the input order of each key class), the template cell in the low bits so that
  // consecutive particles read neighbouring template rows.  Measured at 10^6 particles: coarsening the
  // cell bins to make the key 16 bits (two radix passes instead of three) saves 16 us in the sort and
  // costs 58 us in fused_cube_kernel (L1 misses), so the full cell index is kept (RBX_SORT_BITS caps it).
  ws.ncell = (v.nz - 1) * (v.na - 1);
  int seg_bits = 1, ncell_bits = 1;
  while ((1ull << seg_bits) <= (uint64_t)nseg) ++seg_bits;
  while ((1 << ncell_bits) < ws.ncell) ++ncell_bits;
  int want = 31;
  if (opt(OPT_SORT_BITS) > 0) want = (int)opt(OPT_SORT_BITS);
  ws.cell_bits = std::min(ncell_bits, std::max(2, want - seg_bits));
  if (seg_bits + ws.cell_bits > 31) ws.cell_bits = std::max(0, 31 - seg_bits);
  ws.cell_shift = ncell_bits - ws.cell_bits;
  ws.end_bit = seg_bits + ws.cell_bits;
  ws.rec_stride = v.method == RBX_METHOD_LINEAR ? 8 : 20;
  ws.psub = choose_psub(n);
  // a spaxel is cut only when it holds more than psub particles (at most n / psub of them), into at most
  // 2 c / psub + 2 items (cut_spaxel)
  ws.max_split = (int)(4 * (n / ws.psub) + 4);
  ws.max_items = nseg + (int)(4 * (n / ws.psub)) + 4;
  ws.Wp = (v.W + 3) & ~3;
  ws.cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, ws.cub_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                  (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, 0, ws.end_bit);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return (char *)base + o; };
  ws.keys_in = (uint32_t *)take(sizeof(uint32_t) * n);
  ws.keys_out = (uint32_t *)take(sizeof(uint32_t) * n);
  ws.idx_in = (uint32_t *)take(sizeof(uint32_t) * n);
  ws.idx_out = (uint32_t *)take(sizeof(uint32_t) * n);
  ws.rec = (float *)take(sizeof(float) * n * ws.rec_stride);
  // counts, ctrl and the sort's state are adjacent: zeroed by one memset
  ws.sp = make_sort_plan(n, ws.end_bit);
  const size_t cc_bytes = align_up(sizeof(int) * (nseg + 1 + C_COUNT));
  ws.zero_bytes = cc_bytes + sizeof(uint32_t) * sort_state_words(ws.sp);
  ws.counts = (int *)take(ws.zero_bytes);
  ws.ctrl = ws.counts + (nseg + 1);
  ws.sort_state = (uint32_t *)((char *)ws.counts + cc_bytes);
  ws.seg_start = (int *)take(sizeof(int) * (nseg + 1));
  ws.item_start = (int *)take(sizeof(int) * (nseg + 1));
  ws.split_list = (int *)take(sizeof(int) * (nseg + 1));
  ws.items = (Item *)take(sizeof(Item) * ws.max_items);
  ws.partials = (float *)take(sizeof(float) * (size_t)ws.max_split * ws.Wp);
  ws.cub_temp = take(ws.cub_bytes);
  total = off;
  return RBX_OK;
}

}  // namespace rbx

using namespace rbx;

// ---- optional timing of the dominant kernel (bench.py's roofline) ---------------------------------
static std::atomic<int> g_profile{0};
static cudaEvent_t g_ev[2] = {nullptr, nullptr};
static double g_fused_ms_sum = 0.0;
static int64_t g_fused_n = 0;
static bool g_ev_pending = false;

static void profile_collect() {
  if (g_ev_pending && cudaEventSynchronize(g_ev[1]) == cudaSuccess) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_ev[0], g_ev[1]) == cudaSuccess) { g_fused_ms_sum += ms; ++g_fused_n; }
  }
  g_ev_pending = false;
}

extern "C" int rbx_profile_enable(int on) {
  if (on && !g_ev[0]) {
    RBX_CUDA_OK(cudaEventCreate(&g_ev[0]));
    RBX_CUDA_OK(cudaEventCreate(&g_ev[1]));
  }
  if (!on) profile_collect();
  g_profile.store(on);
  return RBX_OK;
}

// mean duration (ms) and number of fused_cube_kernel launches timed since the last reset
extern "C" int rbx_profile_fused(double *mean_ms, int64_t *launches, int reset) {
  profile_collect();
  if (mean_ms) *mean_ms = g_fused_n ? g_fused_ms_sum / (double)g_fused_n : 0.0;
  if (launches) *launches = g_fused_n;
  if (reset) { g_fused_ms_sum = 0.0; g_fused_n = 0; }
  return RBX_OK;
}

// Shared-memory layout and static limits of fused_cube_kernel for this plan.
static int fused_layout(const rbx_plan *plan, FusedLayout &lay, size_t &smem_bytes) {
  const PlanView &v = plan->v;
  if ((!v.affine && (!plan->lut_ok || v.nb <= 0)) || v.W + 1 >= 65535) {
    set_error("rbx_build_cube: telescope wavelength grid not supported by the fused kernel (too long, or too "
              "uneven for the channel lookup table); use the stage calls");
    return RBX_ERR_UNSUPPORTED;
  }
  for (int w = 1; w < v.W; ++w)
    if (!(plan->h_t[w] > plan->h_t[w - 1])) {
      set_error("rbx_build_cube: telescope wavelength grid must be strictly increasing");
      return RBX_ERR_UNSUPPORTED;
    }
  for (int l = 1; l < v.L; ++l)
    if (!(plan->h_lamz[l] >= plan->h_lamz[l - 1])) {
      set_error("rbx_build_cube: SSP wavelength grid must be non-decreasing");
      return RBX_ERR_UNSUPPORTED;
    }
  // knots that can reach the band for |v| up to ~0.03 c
  const double lo = (double)v.tmin / 1.03, hi = (double)v.tmax * 1.03;
  int inband = 0;
  double min2 = 1e300, max2 = 0.0;  // extremes of lam_z[j+2] - lam_z[j] among those knots
  for (int l = 0; l < v.L; ++l) {
    const double x = plan->h_lamz[l];
    if (x < lo || x > hi) continue;
    ++inband;
    if (l + 2 < v.L) {
      const double d2 = (double)plan->h_lamz[l + 2] - x;
      min2 = std::fmin(min2, d2);
      max2 = std::fmax(max2, d2);
    }
  }
  if (inband + 8 > kMaxKnots) {
    set_error("rbx_build_cube: SSP grid too fine for the fused kernel's knot window; use the stage calls");
    return RBX_ERR_UNSUPPORTED;
  }
  // channels per chunk: a lane's two segments must hold at most one chunk start
  const double span = max2 > 0.0 ? max2 * 1.03 / (double)plan->min_dt + 4.0 : 4.0;
  int chs = 6;
  while (chs <= 8 && span >= (double)(1 << chs)) ++chs;
  if (chs > 8) {
    set_error("rbx_build_cube: SSP grid too coarse relative to the telescope channels for the fused kernel; "
              "use the stage calls");
    return RBX_ERR_UNSUPPORTED;
  }
  lay.chs = chs;
  lay.nch = (v.W + 1 + (1 << chs) - 1) >> chs;
  lay.collide = (min2 * 0.97 <= (double)plan->max_dt) ? 1 : 0;
  lay.force_lut = opt_on(OPT_FUSED_FORCE_LUT) ? 1 : 0;
  if (opt_on(OPT_FUSED_FORCE_CAS)) lay.collide = 1;
  lay.cap = v.W + 1 + kMaxGroupWarps * kRegionSlack;
  auto a16 = [](int x) { return (x + 15) & ~15; };
  auto a128 = [](int x) { return (x + 127) & ~127; };
  lay.lut_bytes = a16(2 * std::max(v.nb, 0));
  lay.tt_bytes = a16(8 * v.W);
  lay.q_bytes = a16(8 * (v.W + 1));
  lay.off_mbar = 0;
  lay.off_lut = 16;
  lay.off_tt = lay.off_lut + lay.lut_bytes;
  lay.off_q = lay.off_tt + lay.tt_bytes;
  lay.off_tc = lay.off_q + lay.q_bytes;
  lay.off_group = a128(lay.off_tc + a16(4 * lay.nch));
  const int rs = v.method == RBX_METHOD_LINEAR ? 8 : 20;
  lay.g_step = 0;
  lay.g_base = a16((lay.cap + kMaxGroupWarps * 32) * 8);  // + one dummy cell per lane (halo lanes)
  lay.g_rec = lay.g_base + a16(kMaxGroupWarps * (lay.nch + 32) * 8);
  lay.g_red = lay.g_rec + a16(2 * NB * rs * 4);
  lay.g_misc = lay.g_red + a16(2 * 2 * NB * kRedStride * 4);
  lay.group_stride = a128(lay.g_misc + 32 * 4);
  const int budget = 227 * 1024;
  lay.max_groups = std::min(kMaxGroups, (budget - lay.off_group) / lay.group_stride);
  if (lay.max_groups < 1) {
    set_error("rbx_build_cube: telescope wavelength grid too long for the fused kernel's shared memory; use the "
              "stage calls");
    return RBX_ERR_UNSUPPORTED;
  }
  smem_bytes = (size_t)lay.off_group + (size_t)lay.max_groups * lay.group_stride;
  return RBX_OK;
}

// Static limits and shared-memory layout of fused_cube_warp_kernel.  false: the plan needs the group kernel.
static bool warp_layout(const rbx_plan *plan, WarpLayout &lay, size_t &smem_bytes, bool pair, int max_arrays) {
  const PlanView &v = plan->v;
  if (!v.affine || !v.wt[0] || v.W + 2 >= (1 << 20)) return false;
  if (opt_on(OPT_FUSED_IMPL) || opt_on(OPT_FUSED_FORCE_LUT) || opt_on(OPT_FUSED_FORCE_CAS)) return false;
  // knots that can reach the band for |v| up to ~0.03 c
  const double lo = (double)v.tmin / 1.03, hi = (double)v.tmax * 1.03;
  int inband = 0;
  double min1 = 1e300, max8 = 0.0;   // extremes of lam_z[j+1] - lam_z[j] and lam_z[j+8] - lam_z[j] among those knots
  for (int l = 0; l < v.L; ++l) {
    const double x = plan->h_lamz[l];
    if (x < lo || x > hi) continue;
    ++inband;
    if (l + 1 < v.L) min1 = std::fmin(min1, (double)plan->h_lamz[l + 1] - x);
    max8 = std::fmax(max8, (double)plan->h_lamz[std::min(l + WK, v.L - 1)] - x);
  }
  if (inband + 16 > kWarpSlots) return false;
  if (!(min1 * 0.97 > (double)plan->max_dt)) return false;   // two knots of a lane could share a cell
  const double span = max8 * 1.03 / (double)plan->min_dt + 4.0;
  int chs = 7;   // the expansion walks a chunk 128 channels at a time
  while (chs <= 10 && span >= (double)(1 << chs)) ++chs;
  if (chs > 10) return false;
  if (opt(OPT_FUSED_CHS) > 0) chs = std::max(chs, (int)opt(OPT_FUSED_CHS));
  auto a128 = [](int x) { return (x + 127) & ~127; };
  lay.chs = chs;
  lay.nch = (v.W + 1 + (1 << chs) - 1) >> chs;
  lay.off_tc = 0;
  lay.off_item = a128(4 * lay.nch);
  lay.off_warp = lay.off_item + a128(4 * 8);
  // bank skew: the lane stride in cells (8 knots) becomes the next odd integer, so the 16 lanes of a
  // 64-bit shared-memory wavefront hit 16 different banks
  {
    double sum = 0.0;
    int cnt = 0;
    for (int l = 0; l + 1 < v.L; ++l) {
      const double x = plan->h_lamz[l];
      if (x < lo || x > hi) continue;
      sum += (double)plan->h_lamz[l + 1] - x;
      ++cnt;
    }
    const double mean_dt = (double)v.trange / std::max(1, v.W - 1);
    const double stride = cnt ? WK * (sum / cnt) / mean_dt : 0.0;
    double target = std::ceil(stride);
    if (((long long)target & 1) == 0) target += 1.0;
    double alpha = stride > 1.0 ? (target - stride) / stride : 0.0;
    if (opt_on(OPT_FUSED_NO_SKEW)) alpha = 0.0;
    lay.skew = (float)std::fmin(alpha, 0.25);
    lay.ncells = v.W + 2 + (int)std::floor((double)(v.W + 2) * (double)lay.skew) + 2;
  }
  lay.w_base = a128(8 * (lay.ncells + 32));   // + one junk cell per lane (RBX_RACECHECK builds)
  lay.w_rec = lay.w_base + a128(8 * lay.nch);
  lay.pair = pair ? 1 : 0;
  lay.rec_bytes = a128(4 * 32 * (v.method == RBX_METHOD_LINEAR ? 8 : 20));
  lay.warp_stride = lay.w_rec + (pair ? 2 : 1) * lay.rec_bytes;
  int na = std::min(std::min(pair ? 7 : 8, max_arrays), (227 * 1024 - lay.off_warp) / lay.warp_stride);   // cell arrays
  if (opt(OPT_FUSED_WARPS) > 0) na = std::min(na, (int)opt(OPT_FUSED_WARPS));
  if (na < 4) return false;
  lay.nwarps = na * (pair ? 2 : 1);
  smem_bytes = (size_t)lay.off_warp + (size_t)na * lay.warp_stride;
  return true;
}

// The transposed-cell variant of the linear pair kernel: blocks of B cells, B the first multiple
// of 8 that holds a lane's span of 8 knots (+ 0.4 % Doppler stretch + 2), stored [B rows][32 columns]; at most 26
// columns hold blocks (columns 26..29 of rows 0..63 park the two warps' records, column 31 of the last row is the junk
// cell).  segment_kernel verifies the geometry for the Doppler range present and falls back to the linear layout.
static bool warp_layout_tr(const rbx_plan *plan, const WarpLayout &base, WarpLayout &lay, size_t &smem_bytes, int max_arrays) {
  const PlanView &v = plan->v;
  if (opt(OPT_FUSED_TR) == 0) return false;
  const double lo = (double)v.tmin / 1.03, hi = (double)v.tmax * 1.03;
  double max8 = 0.0;
  for (int l = 0; l < v.L; ++l) {
    const double x = plan->h_lamz[l];
    if (x < lo || x > hi) continue;
    max8 = std::fmax(max8, (double)plan->h_lamz[std::min(l + WK, v.L - 1)] - x);
  }
  const int B = ((int)std::ceil(max8 * 1.004 / (double)plan->min_dt) + 2 + 7) & ~7;
  const int ncols = (v.W + 2 + B - 1) / B;
  // rows 0..63 (linear) / 0..127 (cubic) park the records; the last row's spare columns are the junk cells
  if (B < (v.method == RBX_METHOD_LINEAR ? 72 : 136) || ncols > 26 || 32 * B > (1 << 15)) return false;
  auto a128 = [](int x) { return (x + 127) & ~127; };
  lay = base;
  lay.tr_B = B;
  lay.tr_ncols = ncols;
  lay.tr_invB = (float)((1.0 / (double)B) * (1.0 + 1.0 / 1048576.0));
  lay.nch = ncols;
  lay.ncells = 32 * B;
  lay.off_item = a128(4 * 32);
  lay.off_warp = lay.off_item + a128(4 * 8);
  lay.w_base = a128(8 * lay.ncells);
  lay.w_stage = lay.w_base + a128(8 * ncols);
  lay.w_rec = lay.w_stage;   // unused: the records are parked inside the cell array
  lay.rec_bytes = 0;
  lay.warp_stride = lay.w_stage + 1024;
  int na = std::min(std::min(7, max_arrays), (227 * 1024 - lay.off_warp) / lay.warp_stride);
  if (opt(OPT_FUSED_WARPS) > 0) na = std::min(na, (int)opt(OPT_FUSED_WARPS));
  if (na < max_arrays) return false;   // fewer arrays than the linear layout holds: not worth it
  lay.nwarps = na * 2;
  smem_bytes = (size_t)lay.off_warp + (size_t)na * lay.warp_stride;
  return true;
}

static int check_fused_config(const rbx_plan *plan, int num_spaxels) {
  FusedLayout lay;
  size_t smem = 0;
  int rc = fused_layout(plan, lay, smem);
  if (rc != RBX_OK) return rc;
  if (num_spaxels < 1 || (int64_t)num_spaxels * num_spaxels > (1 << 24)) {
    set_error("rbx_build_cube: num_spaxels out of range");
    return RBX_ERR_INVALID_ARGUMENT;
  }
  return RBX_OK;
}

extern "C" size_t rbx_build_cube_workspace_bytes(const rbx_plan *plan, int64_t n, int num_spaxels) {
  if (!plan || n < 0 || num_spaxels < 1) return 0;
  FusedWs ws;
  size_t total = 0;
  layout_workspace(plan, n > 0 ? n : 1, num_spaxels * num_spaxels, nullptr, ws, total);
  return total + 256;
}

extern "C" int rbx_build_cube(const rbx_plan *plan, const float *d_vel, const float *d_mass, const float *d_met,
                              const float *d_age, const int32_t *d_pixel, int64_t n, int num_spaxels,
                              float *d_cube, void *d_ws, size_t ws_bytes, void *stream_) {
  RBX_REQUIRE(plan && (n == 0 || d_pixel), "rbx_build_cube: null pointer");
  CubeBuild b;
  b.vel = d_vel ? d_vel + plan->v.vel_comp : nullptr;
  b.mass = d_mass; b.met = d_met; b.age = d_age;
  b.pixel = const_cast<int32_t *>(d_pixel);
  return rbx::build_cube_impl(plan, b, n, num_spaxels, d_cube, d_ws, ws_bytes, stream_);
}

extern "C" int rbx_assign_build_cube(const rbx_plan *plan, const float *d_coords, const float *d_edges, int n_edges,
                                     int apply_filter, const float *d_vel, const float *d_mass, const float *d_met,
                                     const float *d_age, int64_t n, int num_spaxels, int32_t *d_pixel_out,
                                     float *d_cube, void *d_ws, size_t ws_bytes, void *stream_) {
  RBX_REQUIRE(plan && n_edges >= 2 && (n == 0 || (d_coords && d_edges)), "rbx_assign_build_cube: bad argument");
  CubeBuild b;
  b.vel = d_vel ? d_vel + plan->v.vel_comp : nullptr;
  b.mass = d_mass; b.met = d_met; b.age = d_age;
  b.pixel = d_pixel_out;
  b.cx = d_coords; b.cy = d_coords ? d_coords + 1 : nullptr;
  b.edges = d_edges; b.n_edges = n_edges; b.mark_outside = apply_filter ? 1 : 0;
  return rbx::build_cube_impl(plan, b, n, num_spaxels, d_cube, d_ws, ws_bytes, stream_);
}

// Structure-of-arrays particles: x, y and the line-of-sight velocity as arrays of their own (24 bytes per particle
// instead of the 40 of the reference's (n, 3) coords / velocity arrays, of which this path never reads 16).
extern "C" int rbx_assign_build_cube_packed(const rbx_plan *plan, const float *d_x, const float *d_y, const float *d_edges,
                                            int n_edges, int apply_filter, const float *d_vlos, const float *d_mass,
                                            const float *d_met, const float *d_age, int64_t n, int num_spaxels,
                                            int32_t *d_pixel_out, float *d_cube, void *d_ws, size_t ws_bytes,
                                            void *stream_) {
  RBX_REQUIRE(plan && n_edges >= 2 && (n == 0 || (d_x && d_y && d_edges)), "rbx_assign_build_cube_packed: bad argument");
  CubeBuild b;
  b.vel = d_vlos; b.vstride = 1;
  b.mass = d_mass; b.met = d_met; b.age = d_age;
  b.pixel = d_pixel_out;
  b.cx = d_x; b.cy = d_y; b.cstride = 1;
  b.edges = d_edges; b.n_edges = n_edges; b.mark_outside = apply_filter ? 1 : 0;
  return rbx::build_cube_impl(plan, b, n, num_spaxels, d_cube, d_ws, ws_bytes, stream_);
}

// Slab-major partial cube for the multi-GPU exchange: nslab wavelength slabs, each an (S*S, ws) block
// with `halo` channels of its neighbours on both sides, so one ncclReduceScatter hands rank r its summed slab
// ready for the PSF + LSF pass.
extern "C" int rbx_slab_geometry(int W, int nslab, int halo, int *wslab, int *ws) {
  RBX_REQUIRE(W >= 1 && nslab >= 1 && halo >= 0, "rbx_slab_geometry: bad argument");
  const int w = (W + nslab - 1) / nslab;
  if (wslab) *wslab = w;
  if (ws) *ws = w + 2 * halo;
  return RBX_OK;
}

extern "C" int rbx_assign_build_cube_slabs(const rbx_plan *plan, const float *d_coords, const float *d_edges, int n_edges,
                                           int apply_filter, const float *d_vel, const float *d_mass, const float *d_met,
                                           const float *d_age, int64_t n, int num_spaxels, int nslab, int halo,
                                           float *d_slabs, void *d_ws, size_t ws_bytes, void *stream_) {
  RBX_REQUIRE(plan && n_edges >= 2 && (n == 0 || (d_coords && d_edges)), "rbx_assign_build_cube_slabs: bad argument");
  RBX_REQUIRE(nslab >= 1 && halo >= 0, "rbx_assign_build_cube_slabs: bad slab geometry");
  const int wslab = (plan->v.W + nslab - 1) / nslab;
  RBX_REQUIRE(nslab == 1 || halo <= wslab, "rbx_assign_build_cube_slabs: halo wider than a slab");
  CubeBuild b;
  b.vel = d_vel ? d_vel + plan->v.vel_comp : nullptr;
  b.mass = d_mass; b.met = d_met; b.age = d_age;
  b.cx = d_coords; b.cy = d_coords ? d_coords + 1 : nullptr;
  b.edges = d_edges; b.n_edges = n_edges; b.mark_outside = apply_filter ? 1 : 0;
  b.nslab = nslab; b.halo = halo;
  return rbx::build_cube_impl(plan, b, n, num_spaxels, d_slabs, d_ws, ws_bytes, stream_);
}

// What the last build on this workspace did: *h_error = ctrl[C_ERROR] (0 ok; != 0: the cube was poisoned with NaN),
// *h_impl = the kernel segment_kernel selected (0 warp kernel, 1 group kernel).  Synchronises `stream`.
extern "C" int rbx_build_cube_status(const rbx_plan *plan, int64_t n, int num_spaxels, const void *d_ws, int *h_error,
                                     int *h_impl, void *stream_) {
  RBX_REQUIRE(plan && d_ws && n >= 0 && num_spaxels >= 1, "rbx_build_cube_status: bad argument");
  if (h_error) *h_error = 0;
  if (h_impl) *h_impl = 0;
  if (n == 0) return RBX_OK;
  FusedWs ws;
  size_t need = 0;
  const uintptr_t base = ((uintptr_t)d_ws + 255) & ~(uintptr_t)255;
  layout_workspace(plan, n, num_spaxels * num_spaxels, (void *)base, ws, need);
  int h[C_COUNT];
  RBX_CUDA_OK(cudaMemcpyAsync(h, ws.ctrl, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
  RBX_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream_));
  if (h_error) *h_error = h[C_ERROR];
  if (h_impl) *h_impl = h[C_IMPL] == IMPL_WARP_TR ? IMPL_WARP : h[C_IMPL];   // both are the warp kernel
  return RBX_OK;
}

extern "C" int rbx_build_cube_cell_layout(const rbx_plan *plan, int64_t n, int num_spaxels, const void *d_ws,
                                          int *h_transposed, void *stream_) {
  RBX_REQUIRE(plan && d_ws && n >= 0 && num_spaxels >= 1, "rbx_build_cube_cell_layout: bad argument");
  if (h_transposed) *h_transposed = 0;
  if (n == 0) return RBX_OK;
  FusedWs ws;
  size_t need = 0;
  const uintptr_t base = ((uintptr_t)d_ws + 255) & ~(uintptr_t)255;
  layout_workspace(plan, n, num_spaxels * num_spaxels, (void *)base, ws, need);
  int h[C_COUNT];
  RBX_CUDA_OK(cudaMemcpyAsync(h, ws.ctrl, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
  RBX_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream_));
  if (h_transposed) *h_transposed = h[C_IMPL] == IMPL_WARP_TR ? 1 : 0;
  return RBX_OK;
}

// b.accumulate != 0: d_cube += the cube of these particles (rbx_pipeline_host bins a galaxy in ranges so that the
// host-to-device copy of one range overlaps the kernels of the previous one).
// b.cx != NULL: spaxel assignment (+ aperture filter as pixel -1) inside prep_kernel; b.pixel is then an optional
// output.
int rbx::build_cube_impl(const rbx_plan *plan, const CubeBuild &b, int64_t n, int num_spaxels, float *d_cube,
                         void *d_ws, size_t ws_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RBX_REQUIRE(plan && d_cube, "rbx_build_cube: null plan or cube");
  RBX_REQUIRE(n >= 0 && n < (1ll << 31) - 1, "rbx_build_cube: n out of range");
  int rc = check_fused_config(plan, num_spaxels);
  if (rc != RBX_OK) return rc;
  const PlanView &v = plan->v;
  const int nseg = num_spaxels * num_spaxels;
  CubeLayout cl;
  cl.nslab = std::max(1, b.nslab);
  cl.halo = cl.nslab > 1 ? b.halo : 0;
  cl.wslab = cl.nslab > 1 ? (v.W + cl.nslab - 1) / cl.nslab : v.W;
  cl.ws = cl.wslab + 2 * cl.halo;
  cl.slab_stride = (long long)nseg * cl.ws;
  if (!b.accumulate)
    RBX_CUDA_OK(cudaMemsetAsync(d_cube, 0, sizeof(float) * (size_t)cl.nslab * (size_t)cl.slab_stride, stream));
  if (n == 0) return RBX_OK;
  RBX_REQUIRE(b.vel && b.mass && b.met && b.age && (b.pixel || b.cx) && d_ws, "rbx_build_cube: null pointer");
  FusedWs ws;
  size_t need = 0;
  uintptr_t base = ((uintptr_t)d_ws + 255) & ~(uintptr_t)255;
  layout_workspace(plan, n, nseg, (void *)base, ws, need);
  if (need + (base - (uintptr_t)d_ws) > ws_bytes) {
    set_error("rbx_build_cube: workspace too small (see rbx_build_cube_workspace_bytes)");
    return RBX_ERR_WORKSPACE_TOO_SMALL;
  }
  // counts and ctrl are adjacent; all-zero is their initial state (dmin is stored biased, dmax as bits)
  // the library's own radix sort (sort.cu); option sort_impl = 1: cub::DeviceRadixSort (kept for A/B runs)
  const bool own_sort = opt(OPT_SORT_IMPL) != 1 && n < (1ll << 30);
  RBX_CUDA_OK(cudaMemsetAsync(ws.counts, 0, own_sort ? ws.zero_bytes : sizeof(int) * (nseg + 1 + C_COUNT), stream));

  const int threads = 256;
  int blocks = (int)std::min<int64_t>((n + threads - 1) / threads, 148 * 16);
  {
    const int edges_smem = (b.cx && b.n_edges <= 4096) ? 1 : 0;
    // spaxel counts: a per-block shared-memory histogram for MUSE-size cubes (one launch less); for large cubes
    // (150 x 150: 90 KB of histogram per block would halve the occupancy and cost 22500 flush atomics per block)
    // the run lengths of the sorted keys
    const int smem_hist = nseg <= 4096 ? 1 : 0;
    const size_t dyn = (smem_hist ? sizeof(int) * (size_t)nseg : 0) +
                       sizeof(float) * (size_t)(v.nz + v.na + (edges_smem ? b.n_edges : 0)) +
                       (own_sort ? sizeof(int) * 256 * (size_t)ws.sp.npass : 0) + sizeof(uint16_t) * (size_t)v.alut_n;
    int pcap = 148 * 4;   // measured (B200, 10^6 particles): 148 blocks 98 us, 296: 55, 592: 38, 1184: 44, 2368: 54 -- every
                          // block flushes its histogram with one atomic per non-empty spaxel
    // large inputs: the flush is amortised over many particles per block, so fill the SMs (8 blocks each) instead
    pcap = (int)std::min<int64_t>(148 * 16, std::max<int64_t>(pcap, n / 4096));
    if (opt(OPT_PREP_BLOCKS) > 0) pcap = (int)opt(OPT_PREP_BLOCKS);
    const int pblocks = smem_hist ? (int)std::min<int64_t>(blocks, pcap) : blocks;
    if (dyn > 48 * 1024)
      RBX_CUDA_OK(cudaFuncSetAttribute(prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    prep_kernel<<<pblocks, threads, dyn, stream>>>(v, b.vel, b.vstride, b.mass, b.met, b.age, b.pixel, (int)n, nseg,
                                                   ws.cell_bits, ws.cell_shift, ws.keys_in, ws.idx_in, ws.counts, ws.ctrl,
                                                   smem_hist, ws.rec, ws.rec_stride, b.cx, b.cy, b.cstride, b.edges,
                                                   b.n_edges, b.mark_outside, edges_smem, ws.sp,
                                                   own_sort ? ws.sort_state : nullptr);
  }
  count_launch();
  RBX_LAUNCH_OK();
  if (own_sort) {
    uint32_t *ks = nullptr, *vs = nullptr;
    rc = radix_sort_pairs(ws.sp, ws.keys_in, ws.idx_in, ws.keys_out, ws.idx_out, n, ws.sort_state, &ks, &vs, stream);
    if (rc != RBX_OK) return rc;
    ws.keys_out = ks;
    ws.idx_out = vs;
  } else {
    size_t cb = ws.cub_bytes;
    RBX_CUDA_OK(cub::DeviceRadixSort::SortPairs(ws.cub_temp, cb, ws.keys_in, ws.keys_out, ws.idx_in, ws.idx_out,
                                                (int)n, 0, ws.end_bit, stream));
    // cub's kernels are library code: not counted in rbx_launch_count
  }
  if (nseg > 4096) {
    count_runs_kernel<<<blocks, threads, 0, stream>>>(ws.keys_out, (int)n, nseg, ws.cell_bits, ws.counts);
    count_launch();
    RBX_LAUNCH_OK();
  }
  int small_shift = 2, tail_shift = 3;
  if (opt(OPT_SMALL_SHIFT) >= 0) small_shift = (int)std::max<int64_t>(0, std::min<int64_t>(5, opt(OPT_SMALL_SHIFT)));
  if (opt(OPT_TAIL_SHIFT) >= 0) tail_shift = (int)std::max<int64_t>(1, std::min<int64_t>(6, opt(OPT_TAIL_SHIFT)));
  FusedLayout lay;
  size_t smem = 0;
  rc = fused_layout(plan, lay, smem);
  if (rc != RBX_OK) return rc;
  WarpLayout wlay;
  size_t wsmem = 0;
  // The warp kernel runs when the plan allows it (host: grid geometry) AND the particles do (device, segment_kernel:
  // knot window and chunk geometry for the Doppler range present); otherwise the group kernel takes the same work
  // queue.  Both are launched; the one not selected returns at once.
  // Cube-kernel variant (option fused_variant; A/B numbers in DESIGN.md section 5):
  //   1  one warp per cell array (7 warps per SM)
  //   2  two warps per array, 6 arrays (12 warps, 168 registers)      <- default for ssp.method linear
  //   3  two warps per array, 7 arrays (14 warps, 128 registers: spills)
  //   4  two warps per array, 5 arrays (10 warps, 168 registers, more L1)
  // cubic: with the plain template rows (32-byte lane stride, 8-9 cache lines per warp-wide load) the kernel was bound
  // by the L1 data pipe and pairs gained nothing (10^6 particles: 1.2127 against 1.2145 ms); with the window tables
  // it is latency bound and the pair variant wins: 10^6 1.10 -> 0.95 ms, 10^7 8.74 -> 7.40 ms.
  int variant = (int)opt(OPT_FUSED_VARIANT);
  if (variant < 1 || variant > 4) variant = 2;
  const bool pair = variant != 1;
  const int max_arrays = variant == 1 ? 8 : variant == 2 ? 6 : variant == 3 ? 7 : 5;
  const bool warp_ok = warp_layout(plan, wlay, wsmem, pair, max_arrays);
  WarpLayout tlay;
  size_t tsmem = 0;
  const bool tr_ok = warp_ok && variant == 2 && warp_layout_tr(plan, wlay, tlay, tsmem, max_arrays);
  const int lamz_smem = v.L <= 10000 ? 1 : 0;
  const int counts_smem = nseg > 1024 && nseg <= 40000 ? 1 : 0;   // one spaxel per thread needs no staging
  const size_t seg_dyn = (lamz_smem ? sizeof(float) * v.L : 0) + (counts_smem ? sizeof(int) * (size_t)nseg : 0);
  if (seg_dyn > 48 * 1024)
    RBX_CUDA_OK(cudaFuncSetAttribute(segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seg_dyn));
  segment_kernel<<<1, 1024, seg_dyn, stream>>>(
      v, nseg, ws.psub, small_shift, tail_shift, ws.max_items, ws.max_split, ws.counts, ws.seg_start, ws.item_start,
      ws.items, ws.ctrl, warp_ok ? 1 : 0, warp_ok ? wlay.chs : 7, lay.chs, lamz_smem, counts_smem, ws.split_list,
      tr_ok ? tlay.tr_B : 0);
  count_launch();
  RBX_LAUNCH_OK();
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const bool prof = g_profile.load() != 0;
  if (prof) { profile_collect(); cudaEventRecord(g_ev[0], stream); }
  if (tr_ok) {   // the transposed-cell variant; returns at once unless segment_kernel selected it
    auto kernel = v.method == RBX_METHOD_LINEAR ? fused_cube_warp_kernel<RBX_METHOD_LINEAR, true, 384, true>
                                                : fused_cube_warp_kernel<RBX_METHOD_CUBIC, true, 384, true>;
    RBX_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
    kernel<<<nsm, tlay.nwarps * 32, tsmem, stream>>>(v, ws.rec, ws.idx_out, ws.items, ws.ctrl, d_cube, ws.partials, ws.Wp,
                                                     tlay, b.accumulate, cl);
    count_launch();
    RBX_LAUNCH_OK();
  }
  if (warp_ok) {
    auto wlaunch = [&](auto kernel) -> int {
      RBX_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
      kernel<<<nsm, wlay.nwarps * 32, wsmem, stream>>>(v, ws.rec, ws.idx_out, ws.items, ws.ctrl, d_cube, ws.partials, ws.Wp,
                                                       wlay, b.accumulate, cl);
      return RBX_OK;
    };
    if (v.method == RBX_METHOD_LINEAR)
      rc = variant == 1   ? wlaunch(fused_cube_warp_kernel<RBX_METHOD_LINEAR, false, 256>)
           : variant == 2 ? wlaunch(fused_cube_warp_kernel<RBX_METHOD_LINEAR, true, 384>)
           : variant == 3 ? wlaunch(fused_cube_warp_kernel<RBX_METHOD_LINEAR, true, 448>)
                          : wlaunch(fused_cube_warp_kernel<RBX_METHOD_LINEAR, true, 320>);
    else
      rc = variant == 1   ? wlaunch(fused_cube_warp_kernel<RBX_METHOD_CUBIC, false, 256>)
           : variant == 2 ? wlaunch(fused_cube_warp_kernel<RBX_METHOD_CUBIC, true, 384>)
           : variant == 3 ? wlaunch(fused_cube_warp_kernel<RBX_METHOD_CUBIC, true, 448>)
                          : wlaunch(fused_cube_warp_kernel<RBX_METHOD_CUBIC, true, 320>);
    if (rc != RBX_OK) return rc;
    count_launch();
    RBX_LAUNCH_OK();
  }
  {
    auto launch = [&](auto kernel) -> int {
      RBX_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kernel<<<nsm, kCtaThreads, smem, stream>>>(v, ws.rec, ws.idx_out, ws.items, ws.ctrl, d_cube, ws.partials, ws.Wp, lay,
                                                 b.accumulate, cl);
      return RBX_OK;
    };
    const bool affine = v.affine != 0 && !lay.force_lut;
    if (v.method == RBX_METHOD_LINEAR)
      rc = affine ? launch(fused_cube_kernel<RBX_METHOD_LINEAR, true>) : launch(fused_cube_kernel<RBX_METHOD_LINEAR, false>);
    else
      rc = affine ? launch(fused_cube_kernel<RBX_METHOD_CUBIC, true>) : launch(fused_cube_kernel<RBX_METHOD_CUBIC, false>);
    if (rc != RBX_OK) return rc;
    count_launch();
    RBX_LAUNCH_OK();
  }
  if (prof) { cudaEventRecord(g_ev[1], stream); g_ev_pending = true; }
  // blocks loop over spaxels (most have nothing to add); large cubes hold few split spaxels: fewer, longer blocks
  dim3 rgrid((v.W + 255) / 256, std::min(nseg, 592));
  reduce_partials_kernel<<<rgrid, 256, 0, stream>>>(ws.item_start, ws.split_list, ws.partials, ws.Wp, v.W, nseg, ws.ctrl,
                                                    d_cube, b.accumulate, cl);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}
