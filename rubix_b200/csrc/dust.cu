// Dust extinction, the calc_dusty_ifu variant of the path: rubix/core/dust.py:15-65 ->
// rubix/spectra/dust/dust_extinction.py:169-358.
//
//   per gas cell   log_OH = 12 + log10(metals[4] / (16 metals[0]))                       (dust_extinction.py:283-285)
//                  D/G    = 1 / 10^(a + alpha (8.69 - log_OH)), Remy-Ruyer 2014 table 1    (:13-93)
//                  A_V,i  = mass * D/G * 3 pi (Msun/kpc^2 -> g/cm^2) / (0.4 ln10 lambda_V rho_grain) / spaxel area
//                                                                                         (:96-166, :291-297)
//   per spaxel     gas cells lexsorted by (pixel, z); cumulative A_V along z               (:240-249, :318)
//   per star       A_V = jnp.interp(z_star, z_gas, cumulative, left="extrapolate")         (:327-336)
//   per spectrum   spectra * 10^(-0.4 * axav(lambda) * A_V)                                (:341-356,
//                  dust_baseclasses.py:126-164); axav(lambda) depends on the configuration only and is
//                  evaluated on the host (rubix_b200/dust.py), like the PSF / LSF taps.
//
// The reference evaluates one spaxel at a time over ALL gas cells: the cells outside the spaxel are moved
// to z * 1e30 with value 0 and the table is re-sorted (dust_extinction.py:313-325).  With every gas cell
// at a |z| that 1e30 carries past all particles (|z| > ~1e-20 kpc) the sorted table of
// spaxel s is  [cells outside s with z < 0] [the cells of s by z] [cells outside s with z > 0]  and only
// the innermost outside cell on either side (value 0) can take part in an interpolation.  dust_av_kernel
// evaluates exactly that table: a binary search in the spaxel's own cells plus the two "far" neighbours,
// which are found with two 64-bit atomic max / min passes (best cell overall, best cell of another spaxel).
//
// Kernels:  dust_cell_kernel (cell A_V, sort key, far-cell pass 1)  ->  dust_far2_kernel (far-cell pass 2)
//        -> cub radix sort of (pixel, z) keys  ->  dust_bounds_kernel (spaxel boundaries, gather)
//        -> dust_scan_kernel (one block per spaxel, cumulative A_V)  ->  dust_av_kernel (one thread per star)
//        apply_extinction_kernel: the per-channel factor (stage form);  resample_dusty_cube_kernel: Doppler
//        shift + resampling + extinction + per-spaxel sum in one pass, without the (n, W) intermediate.
#include <cub/block/block_scan.cuh>
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace rbx {

__device__ __forceinline__ uint32_t f32_orderable(float f) {
  const uint32_t b = __float_as_uint(f);
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float f32_from_orderable(uint32_t u) {
  return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}

struct DustParams {
  float a_high, alpha_high, a_low, alpha_low, x_transition;  // D/G = 1 / 10^(a + alpha (8.69 - x)); high branch for x > x_transition
  float ext_const;                                           // 3 pi conv / (0.4 ln10 lambda_V rho_grain)
  float spaxel_area;
};

// far[0] = best far-left cell overall  (max of z*1e30 < 0), far[1] = best far-left cell of another spaxel,
// far[2] / far[3] the same on the right (min of z*1e30 > 0).  Encoded (orderable(xp) << 32) | pixel bits.
constexpr unsigned long long kNoLeft = 0ull, kNoRight = ~0ull;

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  return v;
}

__device__ __forceinline__ void far_candidates(float z, int32_t pixel, unsigned long long &l, unsigned long long &r) {
  const float xp = __fmul_rn(z, 1e30f);   // gas_mask2 = where(mask, 1, 1e30), float32        (dust_extinction.py:316-321)
  l = kNoLeft;
  r = kNoRight;
  const unsigned long long enc = ((unsigned long long)f32_orderable(xp) << 32) | (uint32_t)pixel;
  if (xp < 0.f) l = enc;
  if (xp > 0.f) r = enc;
}

__global__ void dust_cell_kernel(const float *__restrict__ coords, const int32_t *__restrict__ pixel,
                                 const float *__restrict__ mass, const float *__restrict__ metals, int n_metals,
                                 int64_t n, int n_spaxels, DustParams p, float *__restrict__ cell_av,
                                 unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals,
                                 unsigned long long *__restrict__ far) {
  unsigned long long bl = kNoLeft, br = kNoRight;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float m0 = metals[i * n_metals], m4 = metals[i * n_metals + 4];
    const float log_oh = 12.f + log10f(__fdiv_rn(m4, __fmul_rn(16.f, m0)));
    const bool hi = log_oh > p.x_transition;
    const float a = hi ? p.a_high : p.a_low, al = hi ? p.alpha_high : p.alpha_low;
    const float dtg = __fdiv_rn(1.f, exp10f(a + al * (8.69f - log_oh)));
    const float dust_mass = __fmul_rn(mass[i], dtg);
    cell_av[i] = __fdiv_rn(__fmul_rn(dust_mass, p.ext_const), p.spaxel_area);
    const float z = coords[3 * i + 2];
    const int32_t px = pixel[i];
    // lexsort((z, pixel)); ids below 0 / at or above n_spaxels (the assignment never produces them) collapse to one
    // slot in front of / behind every spaxel, which keeps the key at 32 + log2(n_spaxels + 2) bits
    const uint32_t slot = (uint32_t)(min(max(px, -1), n_spaxels) + 1);
    keys[i] = ((unsigned long long)slot << 32) | f32_orderable(z);
    vals[i] = (uint32_t)i;
    unsigned long long l, r;
    far_candidates(z, px, l, r);
    bl = l > bl ? l : bl;
    br = r < br ? r : br;
  }
  bl = warp_max_u64(bl);
  br = warp_min_u64(br);
  if ((threadIdx.x & 31) == 0) {
    if (bl != kNoLeft) atomicMax(far + 0, bl);
    if (br != kNoRight) atomicMin(far + 2, br);
  }
}

__global__ void dust_far2_kernel(const float *__restrict__ coords, const int32_t *__restrict__ pixel, int64_t n,
                                 unsigned long long *__restrict__ far) {
  const unsigned long long f0 = far[0], f2 = far[2];
  const int32_t pl = (int32_t)(uint32_t)f0, pr = (int32_t)(uint32_t)f2;
  unsigned long long bl = kNoLeft, br = kNoRight;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long l, r;
    const int32_t px = pixel[i];
    far_candidates(coords[3 * i + 2], px, l, r);
    if (f0 != kNoLeft && px != pl && l > bl) bl = l;
    if (f2 != kNoRight && px != pr && r < br) br = r;
  }
  bl = warp_max_u64(bl);
  br = warp_min_u64(br);
  if ((threadIdx.x & 31) == 0) {
    if (bl != kNoLeft) atomicMax(far + 1, bl);
    if (br != kNoRight) atomicMin(far + 3, br);
  }
}

// spaxel boundaries in the sorted gas order (searchsorted(pixel_sorted, arange(S), 'left') + [n], :249-270) and
// the sorted z / cell A_V arrays
__global__ void dust_bounds_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals,
                                   const float *__restrict__ cell_av, int64_t n, int n_spaxels,
                                   int32_t *__restrict__ bounds, float *__restrict__ zs, float *__restrict__ es) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t <= n_spaxels) {
    int64_t lo = 0, hi = n;
    if (t == n_spaxels) lo = n;   // jnp.array([len(sorted)])
    const unsigned long long want = (unsigned long long)((uint32_t)t + 1u) << 32;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (keys[mid] < want) lo = mid + 1; else hi = mid;
    }
    bounds[t] = (int32_t)lo;
  }
  for (int64_t i = t; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    zs[i] = f32_from_orderable((uint32_t)keys[i]);
    es[i] = cell_av[vals[i]];
  }
}

// cumsum(extinction * gas_mask) * gas_mask (:318): the running sum of the spaxel's cells along z.  Double
// accumulation (the reference's float32 scan order is XLA's choice; double is closer to exact than any of them).
__global__ void dust_scan_kernel(const int32_t *__restrict__ bounds, const float *__restrict__ es, float *__restrict__ cs) {
  using Scan = cub::BlockScan<double, 256>;
  __shared__ typename Scan::TempStorage tmp;
  const int g0 = bounds[blockIdx.x], g1 = bounds[blockIdx.x + 1];
  double carry = 0.0;
  for (int base = g0; base < g1; base += 256) {
    const int i = base + threadIdx.x;
    const double v = i < g1 ? (double)es[i] : 0.0;
    double incl, total;
    Scan(tmp).InclusiveSum(v, incl, total);
    if (i < g1) cs[i] = (float)(carry + incl);
    carry += total;
    __syncthreads();
  }
}

// jnp.interp's two-point formula (fp[i-1] + (delta / dx) * df, a zero-width pair returns fp[i-1])
__device__ __forceinline__ float interp_pair(float x, float x0, float f0, float x1, float f1) {
  const float dx = x1 - x0, df = f1 - f0, delta = x - x0;
  return (fabsf(dx) <= 1.4210855e-14f) ? f0 : f0 + __fdiv_rn(delta, dx) * df;
}

__global__ void dust_av_kernel(const float *__restrict__ star_coords, const int32_t *__restrict__ star_pixel, int64_t n,
                               int n_spaxels, const int32_t *__restrict__ bounds, const float *__restrict__ zs,
                               const float *__restrict__ cs, const unsigned long long *__restrict__ far,
                               float *__restrict__ av) {
  const unsigned long long fl0 = far[0], fl1 = far[1], fr0 = far[2], fr1 = far[3];
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int32_t s = star_pixel[q];
    if (s < 0 || s >= n_spaxels) { av[q] = 0.f; continue; }   // star_mask of no spaxel: Av stays 0
    const float x = star_coords[3 * q + 2];
    const int g0 = bounds[s], g1 = bounds[s + 1], m = g1 - g0;
    // innermost cells of other spaxels on either side
    unsigned long long el = (fl0 != kNoLeft && (int32_t)(uint32_t)fl0 == s) ? fl1 : fl0;
    unsigned long long er = (fr0 != kNoRight && (int32_t)(uint32_t)fr0 == s) ? fr1 : fr0;
    const bool has_l = el != kNoLeft, has_r = er != kNoRight;
    const float xl = f32_from_orderable((uint32_t)(el >> 32)), xr = f32_from_orderable((uint32_t)(er >> 32));
    int lo = g0, hi = g1;   // j = number of the spaxel's cells with z <= x (searchsorted 'right')
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (zs[mid] <= x) lo = mid + 1; else hi = mid;
    }
    const int j = lo - g0;
    const bool have_left = j >= 1 || has_l, have_right = j < m || has_r;
    float f;
    if (m == 0) {
      f = 0.f;   // every table entry is 0
    } else if (have_left && have_right) {
      const float x0 = j >= 1 ? zs[g0 + j - 1] : xl, f0 = j >= 1 ? cs[g0 + j - 1] : 0.f;
      const float x1 = j < m ? zs[g0 + j] : xr, f1 = j < m ? cs[g0 + j] : 0.f;
      f = interp_pair(x, x0, f0, x1, f1);
    } else if (!have_left) {
      // in front of the first table entry: left="extrapolate" through entries 0 and 1
      if (m >= 2) f = interp_pair(x, zs[g0], cs[g0], zs[g0 + 1], cs[g0 + 1]);
      else if (has_r) f = interp_pair(x, zs[g0], cs[g0], xr, 0.f);
      else f = cs[g0];
    } else {
      // behind the last table entry: index clipped to the last pair, then the right end value
      if (m >= 2) f = interp_pair(x, zs[g1 - 2], cs[g1 - 2], zs[g1 - 1], cs[g1 - 1]);
      else if (has_l) f = interp_pair(x, xl, 0.f, zs[g0], cs[g0]);
      else f = cs[g0];
    }
    if (m >= 1 && !has_r && x > zs[g1 - 1]) f = cs[g1 - 1];   // where(x > xp[-1], fp[-1], f)
    av[q] = f;
  }
}

// spectra * 10^(-0.4 * axav * Av)   (dust_baseclasses.py:163, dust_extinction.py:356)
__global__ void apply_extinction_kernel(const float *__restrict__ spec, const float *__restrict__ av,
                                        const float *__restrict__ axav, int64_t n, int W, float *__restrict__ out) {
  for (int64_t q = blockIdx.x; q < n; q += gridDim.x) {
    const float a = av[q];
    const float *s = spec + q * W;
    float *o = out + q * W;
    for (int w = threadIdx.x; w < W; w += blockDim.x)
      o[w] = s[w] * exp10f(__fmul_rn(__fmul_rn(-0.4f, axav[w]), a));
  }
}

// ---- doppler_shift_and_resampling -> calculate_extinction -> calculate_datacube in one pass --------------------
// The reference materialises (n, W) twice between these stages (rubix/spectra/ifu.py:224-266 -> dust_extinction.py:356
// -> ifu.py:270-288).  Here a block walks a chunk of the spaxel-sorted particle list; every thread owns kk =
// ceil(W / 256) CONSECUTIVE channels, so the interval search of jnp.interp is one binary search for the first channel
// and a forward walk for the rest; the slope of every knot interval is computed once per particle (one IEEE division
// per interval, not per channel: f = fp[i-1] + delta * (df / dx) instead of fp[i-1] + (delta / dx) * df, an ulp-level
// reordering); the particle's resampled values stay in registers until total / new is known, are scaled, multiplied
// by 10^(-0.4 axav Av) = 2^(e_w Av) and added to the thread's per-channel accumulators, which are flushed into the cube
// (RED.ADD.F32) when the spaxel changes or the chunk ends.  t, dt and e_w are staged in shared memory once per block;
// (lambda', spectrum, slopes) of the current particle too.  Traffic: 4 L bytes per particle instead of 16 W.
template <int KMAX>
__global__ void __launch_bounds__(256, KMAX <= 16 ? 4 : 1) resample_dusty_cube_kernel(
    PlanView p, const float *__restrict__ spec, const float *__restrict__ vel, const int32_t *__restrict__ spaxel_sorted,
    const uint32_t *__restrict__ order, const float *__restrict__ av, const float *__restrict__ axav, int64_t n, int nseg,
    int chunk, float *__restrict__ cube) {
  extern __shared__ float sm[];
  // t, dt, -0.4 log2(10) axav per channel (once per block); lambda', spectrum and the interval slopes of the particle
  float *t_s = sm, *dt_s = sm + p.W, *e_s = sm + 2 * p.W, *lam = sm + 3 * p.W, *s = lam + p.L, *slope = s + p.L;
  __shared__ float red[16];
  __shared__ float bc;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int kk = (p.W + 255) / 256, w0 = tid * kk;
  const int cnt = max(0, min(kk, p.W - w0));   // this thread's channels: w0 .. w0 + cnt - 1
  for (int w = tid; w < p.W; w += 256) {
    t_s[w] = p.t[w];
    dt_s[w] = p.dt[w];
    e_s[w] = axav ? __fmul_rn(-0.4f, axav[w]) * 3.3219281f : 0.f;   // 10^(-0.4 k A) = 2^(e A)
  }
  float acc[KMAX];
  const int64_t nchunks = (n + chunk - 1) / chunk;
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    int cur = -1;
    const int64_t pend = (c + 1) * chunk < n ? (c + 1) * chunk : n;
    for (int64_t pos = c * chunk; pos < pend; ++pos) {
      const int spx = spaxel_sorted[pos];
      if (spx < 0 || spx >= nseg) continue;   // segment_sum drops ids outside the range
      if (spx != cur) {
        if (cur >= 0) {
#pragma unroll
          for (int k = 0; k < KMAX; ++k)
            if (k < cnt && acc[k] != 0.f) atomicAdd(cube + (size_t)cur * p.W + w0 + k, acc[k]);
        }
        cur = spx;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) acc[k] = 0.f;
      }
      const int64_t q = order[pos];
      const float d = expf(vel[3 * q + p.vel_comp] / kSpeedOfLight);
      __syncthreads();   // the previous particle's readers are done with lam / s / slope / red
      const float *sp = spec + q * p.L;
      for (int l = tid; l < p.L; l += 256) {
        lam[l] = __fmul_rn(p.lamz[l], d);
        s[l] = sp[l];
      }
      __syncthreads();
      float tot = 0.f, nws = 0.f;
      // per knot interval (l-1, l): the in-band flux sum (ifu.py:241-247) and the slope df / dx of jnp.interp's
      // fp[i-1] + (delta / dx) * df, one IEEE division per interval instead of one per channel; a zero-width
      // interval returns fp[i-1] (slope 0)
      for (int l = tid + 1; l < p.L; l += 256) {
        const float x = lam[l], dx = x - lam[l - 1];
        if (x >= p.tmin && x <= p.tmax) tot += s[l] * dx;
        slope[l] = (fabsf(dx) <= 1.4210855e-14f) ? 0.f : __fdiv_rn(s[l] - s[l - 1], dx);
      }
      __syncthreads();
      float pv[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) pv[k] = 0.f;
      if (cnt > 0) {
        int i = min(max(ss_right(lam, p.L, t_s[w0]), 1), p.L - 1);
        float x0 = lam[i - 1], x1 = lam[i], f0 = s[i - 1], m = slope[i];
        const float lam_lo = lam[0], lam_hi = lam[p.L - 1], s_lo = s[0], s_hi = s[p.L - 1];
        const bool inside = t_s[w0] >= lam_lo && t_s[w0 + cnt - 1] <= lam_hi;   // no end-value clamp for this thread
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          if (k < cnt) {
            const float x = t_s[w0 + k];
            while (i < p.L - 1 && !(x1 > x)) {   // i = clip(searchsorted(lam, x, 'right'), 1, L-1)
              ++i;
              x0 = x1; f0 = s[i - 1];
              x1 = lam[i]; m = slope[i];
            }
            float f = __fmaf_rn(x - x0, m, f0);
            if (!inside) {
              if (x < lam_lo) f = s_lo;
              if (x > lam_hi) f = s_hi;
            }
            pv[k] = f;
            nws = __fmaf_rn(f, dt_s[w0 + k], nws);
          }
        }
      }
      tot = warp_sum(tot);
      nws = warp_sum(nws);
      if (lane == 0) { red[wid] = tot; red[8 + wid] = nws; }
      __syncthreads();
      if (tid == 0) {
        float a = 0.f, b = 0.f;
        for (int k = 0; k < 8; ++k) { a += red[k]; b += red[8 + k]; }
        bc = nan_to_num0(a / b);   // rubix/spectra/ifu.py:252-255
      }
      __syncthreads();
      const float scale = bc;
      const float a_v = av ? av[q] : 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < cnt) {
          float v = __fmul_rn(pv[k], scale);                             // the resampled spectrum (ifu.py:257)
          if (av) v = __fmul_rn(v, exp2f(e_s[w0 + k] * a_v));            // * extinction (dust_extinction.py:356)
          acc[k] += v;                                                   // segment_sum (ifu.py:286)
        }
    }
    if (cur >= 0) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < cnt && acc[k] != 0.f) atomicAdd(cube + (size_t)cur * p.W + w0 + k, acc[k]);
    }
  }
}

__global__ void dusty_keys_kernel(const int32_t *__restrict__ pixel, int64_t n, int nseg, uint32_t *__restrict__ keys,
                                  uint32_t *__restrict__ vals) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t px = pixel[i];
    keys[i] = (px < 0 || px >= nseg) ? (uint32_t)nseg : (uint32_t)px;   // dropped ids sort behind every spaxel
    vals[i] = (uint32_t)i;
  }
}

// ---- the dusty cube through the knot-based cube kernel (fused.cu) ------------------------------------------------
// The per-star factor 10^(-0.4 k_w A_V) is smooth in A_V, so stars are binned by A_V (bin centre A_b, width chosen
// so that x = ln10 * 0.4 * k_w * (A_V - A_b) stays below 0.05 in magnitude) and the factor is expanded inside a bin:
//   10^(-0.4 k_w A_V) = 10^(-0.4 k_w A_b) * (1 - x + x^2/2 - x^3/6 + ...),   truncation x^4/24 < 2.6e-7 relative.
// The cube of every (spaxel, bin) pair is then sum_m (a_w^m / m!) C_m with a_w = -0.4 ln10 k_w and C_m the ordinary
// (dust-free) cube of the bin's stars weighted by mass * (A_V - A_b)^m.  dusty_bins_kernel writes every star four
// times (one virtual particle per moment, weight mass * (A_V - A_b)^m, virtual spaxel (s * n_bins + b) * 4 + m), ONE run
// of rbx_build_cube bins the 4 n virtual particles onto the S^2 * n_bins * 4 virtual spaxels, and
// dusty_combine_kernel folds bins and moments into the (S^2, W) cube.
constexpr int kDustMoments = 4;

__global__ void dusty_bins_kernel(const float *__restrict__ av, const int32_t *__restrict__ pixel,
                                  const float *__restrict__ mass, const float *__restrict__ vel,
                                  const float *__restrict__ met, const float *__restrict__ age, int64_t n, int nseg,
                                  int n_bins, float av0, float step, int32_t *__restrict__ vpixel, float *__restrict__ wmass,
                                  float *__restrict__ vel_out, float *__restrict__ met_out, float *__restrict__ age_out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float a = av[i], m = mass[i];
    const int32_t px = pixel[i];
    int b = (int)floorf((a - av0) / step);
    b = b < 0 ? 0 : (b >= n_bins ? n_bins - 1 : b);
    const float da = a - (av0 + ((float)b + 0.5f) * step);
    const int32_t base = (px < 0 || px >= nseg) ? -1 : (px * n_bins + b) * kDustMoments;
    const float v0 = vel[3 * i], v1 = vel[3 * i + 1], v2 = vel[3 * i + 2], z = met[i], t = age[i];
    float w = m;
#pragma unroll
    for (int k = 0; k < kDustMoments; ++k) {   // one virtual particle per moment, in its own virtual spaxel
      const size_t o = (size_t)k * n + i;
      vpixel[o] = base < 0 ? -1 : base + k;
      wmass[o] = w;
      vel_out[3 * o] = v0; vel_out[3 * o + 1] = v1; vel_out[3 * o + 2] = v2;
      met_out[o] = z;
      age_out[o] = t;
      w *= da;
    }
  }
}

__global__ void dusty_combine_kernel(const float *__restrict__ vcube, int n_bins, float av0, float step,
                                     const float *__restrict__ axav, int W, float *__restrict__ cube) {
  const int s = blockIdx.x;
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    const float m04 = __fmul_rn(-0.4f, axav[w]);
    const float a = m04 * 2.302585093f;
    float acc = 0.f;
    for (int b = 0; b < n_bins; ++b) {
      const float *r = vcube + ((size_t)s * n_bins + b) * kDustMoments * W + w;
      const float c0 = r[0], c1 = r[W], c2 = r[2 * (size_t)W], c3 = r[3 * (size_t)W];
      const float poly = c0 + a * (c1 + a * (0.5f * c2 + a * (0.16666667f * c3)));
      if (poly != 0.f) acc += exp10f(m04 * (av0 + ((float)b + 0.5f) * step)) * poly;
    }
    cube[(size_t)s * W + w] = acc;
  }
}

}  // namespace rbx

using namespace rbx;

namespace {
struct DustWs {
  float *cell_av, *zs, *es, *cs;
  unsigned long long *keys, *keys_out, *far;
  uint32_t *vals, *vals_out;
  int32_t *bounds;
  void *cub_temp;
  size_t cub_bytes, total;
};

inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

DustWs carve_dust(void *base, int64_t n, int n_spaxels) {
  DustWs w{};
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                  (const uint32_t *)nullptr, (uint32_t *)nullptr, (int64_t)(n > 0 ? n : 1), 0, 64);
  uintptr_t p = up256((uintptr_t)base);
  auto take = [&](size_t bytes) { void *r = (void *)p; p += up256(bytes); return r; };
  const size_t nn = (size_t)(n > 0 ? n : 1);
  w.far = (unsigned long long *)take(4 * sizeof(unsigned long long));
  w.cell_av = (float *)take(4 * nn);
  w.zs = (float *)take(4 * nn);
  w.es = (float *)take(4 * nn);
  w.cs = (float *)take(4 * nn);
  w.keys = (unsigned long long *)take(8 * nn);
  w.keys_out = (unsigned long long *)take(8 * nn);
  w.vals = (uint32_t *)take(4 * nn);
  w.vals_out = (uint32_t *)take(4 * nn);
  w.bounds = (int32_t *)take(4 * ((size_t)n_spaxels + 1));
  w.cub_temp = take(cub_bytes);
  w.cub_bytes = cub_bytes;
  w.total = (size_t)(p - (uintptr_t)base) + 256;
  return w;
}

int grid1d(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (int)b;
}
}  // namespace

extern "C" size_t rbx_dust_av_workspace_bytes(int64_t n_gas, int n_spaxels) {
  return carve_dust(nullptr, n_gas, n_spaxels).total;
}

extern "C" int rbx_dust_av(const float *d_gas_coords, const int32_t *d_gas_pixel, const float *d_gas_mass,
                           const float *d_gas_metals, int n_metals, int64_t n_gas, const float *d_star_coords,
                           const int32_t *d_star_pixel, int64_t n_star, int n_spaxels, const float *h_dust_to_gas,
                           float ext_const, float spaxel_area, float *d_av, float *d_cell_av_out, void *d_workspace,
                           size_t workspace_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RBX_REQUIRE(n_gas >= 0 && n_star >= 0 && n_spaxels > 0, "rbx_dust_av: bad sizes");
  RBX_REQUIRE(n_gas < (int64_t)1 << 31, "rbx_dust_av: more than 2^31 gas cells");
  RBX_REQUIRE(h_dust_to_gas, "rbx_dust_av: dust-to-gas parameters missing");
  RBX_REQUIRE(n_metals >= 5 || n_gas == 0, "rbx_dust_av: metals needs at least 5 species (H = 0, O = 4)");
  if (n_star == 0) return RBX_OK;
  RBX_REQUIRE(d_star_coords && d_star_pixel && d_av, "rbx_dust_av: null star pointer");
  if (n_gas == 0) {   // no gas: the table is empty, every star keeps Av = 0
    RBX_CUDA_OK(cudaMemsetAsync(d_av, 0, sizeof(float) * (size_t)n_star, stream));
    return RBX_OK;
  }
  RBX_REQUIRE(d_gas_coords && d_gas_pixel && d_gas_mass && d_gas_metals && d_workspace, "rbx_dust_av: null gas pointer");
  RBX_REQUIRE(workspace_bytes >= rbx_dust_av_workspace_bytes(n_gas, n_spaxels), "rbx_dust_av: workspace too small");
  DustWs w = carve_dust(d_workspace, n_gas, n_spaxels);
  DustParams p{h_dust_to_gas[0], h_dust_to_gas[1], h_dust_to_gas[2], h_dust_to_gas[3], h_dust_to_gas[4], ext_const, spaxel_area};
  RBX_CUDA_OK(cudaMemsetAsync(w.far, 0x00, 2 * sizeof(unsigned long long), stream));        // kNoLeft
  RBX_CUDA_OK(cudaMemsetAsync(w.far + 2, 0xFF, 2 * sizeof(unsigned long long), stream));    // kNoRight
  dust_cell_kernel<<<grid1d(n_gas, 256), 256, 0, stream>>>(d_gas_coords, d_gas_pixel, d_gas_mass, d_gas_metals, n_metals,
                                                            n_gas, n_spaxels, p, w.cell_av, w.keys, w.vals, w.far);
  count_launch();
  RBX_LAUNCH_OK();
  dust_far2_kernel<<<grid1d(n_gas, 256), 256, 0, stream>>>(d_gas_coords, d_gas_pixel, n_gas, w.far);
  count_launch();
  RBX_LAUNCH_OK();
  size_t cub_bytes = w.cub_bytes;
  int key_bits = 33;
  while ((1ll << (key_bits - 32)) < (long long)n_spaxels + 2) ++key_bits;
  RBX_CUDA_OK(cub::DeviceRadixSort::SortPairs(w.cub_temp, cub_bytes, (const unsigned long long *)w.keys, w.keys_out,
                                              (const uint32_t *)w.vals, w.vals_out, n_gas, 0, key_bits, stream));
  count_launch(3);
  int bgrid = grid1d(n_gas, 256);
  if (bgrid < (n_spaxels + 1 + 255) / 256) bgrid = (n_spaxels + 1 + 255) / 256;
  dust_bounds_kernel<<<bgrid, 256, 0, stream>>>(w.keys_out, w.vals_out, w.cell_av, n_gas, n_spaxels, w.bounds, w.zs, w.es);
  count_launch();
  RBX_LAUNCH_OK();
  dust_scan_kernel<<<n_spaxels, 256, 0, stream>>>(w.bounds, w.es, w.cs);
  count_launch();
  RBX_LAUNCH_OK();
  dust_av_kernel<<<grid1d(n_star, 256), 256, 0, stream>>>(d_star_coords, d_star_pixel, n_star, n_spaxels, w.bounds, w.zs,
                                                          w.cs, w.far, d_av);
  count_launch();
  RBX_LAUNCH_OK();
  if (d_cell_av_out)
    RBX_CUDA_OK(cudaMemcpyAsync(d_cell_av_out, w.cell_av, sizeof(float) * (size_t)n_gas, cudaMemcpyDeviceToDevice, stream));
  return RBX_OK;
}

extern "C" int rbx_apply_extinction(const float *d_spectra, const float *d_av, const float *d_axav, int64_t n, int W,
                                    float *d_out, void *stream) {
  RBX_REQUIRE(n >= 0 && W > 0, "rbx_apply_extinction: bad shape");
  if (n == 0) return RBX_OK;
  RBX_REQUIRE(d_spectra && d_av && d_axav && d_out, "rbx_apply_extinction: null pointer");
  int64_t b = n > 148 * 64 ? 148 * 64 : n;
  apply_extinction_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(d_spectra, d_av, d_axav, n, W, d_out);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}

namespace {
struct DustyWs {
  uint32_t *keys, *keys_out, *vals, *vals_out;
  void *cub_temp;
  size_t cub_bytes, total;
};
DustyWs carve_dusty(void *base, int64_t n) {
  DustyWs w{};
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                  (uint32_t *)nullptr, (int64_t)(n > 0 ? n : 1), 0, 32);
  uintptr_t p = up256((uintptr_t)base);
  auto take = [&](size_t bytes) { void *r = (void *)p; p += up256(bytes); return r; };
  const size_t nn = (size_t)(n > 0 ? n : 1);
  w.keys = (uint32_t *)take(4 * nn);
  w.keys_out = (uint32_t *)take(4 * nn);
  w.vals = (uint32_t *)take(4 * nn);
  w.vals_out = (uint32_t *)take(4 * nn);
  w.cub_temp = take(cub_bytes);
  w.cub_bytes = cub_bytes;
  w.total = (size_t)(p - (uintptr_t)base) + 256;
  return w;
}
}  // namespace

extern "C" size_t rbx_build_cube_dusty_workspace_bytes(int64_t n) { return carve_dusty(nullptr, n).total; }

extern "C" int rbx_build_cube_dusty(const rbx_plan *plan, const float *d_spectra, const float *d_velocity,
                                    const int32_t *d_pixel, const float *d_av, const float *d_axav, int64_t n,
                                    int num_spaxels, float *d_cube, void *d_workspace, size_t workspace_bytes,
                                    void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RBX_REQUIRE(plan && d_cube && num_spaxels > 0 && n >= 0, "rbx_build_cube_dusty: bad argument");
  RBX_REQUIRE(n < (int64_t)1 << 31, "rbx_build_cube_dusty: more than 2^31 particles");
  const PlanView &v = plan->v;
  const int nseg = num_spaxels * num_spaxels;
  RBX_CUDA_OK(cudaMemsetAsync(d_cube, 0, sizeof(float) * (size_t)nseg * v.W, stream));
  if (n == 0) return RBX_OK;
  RBX_REQUIRE(d_spectra && d_velocity && d_pixel && d_workspace, "rbx_build_cube_dusty: null pointer");
  RBX_REQUIRE((d_av == nullptr) == (d_axav == nullptr), "rbx_build_cube_dusty: d_av and d_axav go together");
  RBX_REQUIRE(workspace_bytes >= rbx_build_cube_dusty_workspace_bytes(n), "rbx_build_cube_dusty: workspace too small");
  const int kk = (v.W + 255) / 256;
  const size_t smem = sizeof(float) * (3 * (size_t)v.W + 3 * (size_t)v.L);
  if (kk > 32 || smem > 200 * 1024) {
    set_error("rbx_build_cube_dusty: telescope grid beyond 8192 channels (or SSP grid too long for shared memory)");
    return RBX_ERR_UNSUPPORTED;
  }
  DustyWs w = carve_dusty(d_workspace, n);
  dusty_keys_kernel<<<grid1d(n, 256), 256, 0, stream>>>(d_pixel, n, nseg, w.keys, w.vals);
  count_launch();
  RBX_LAUNCH_OK();
  int bits = 1;
  while ((1ll << bits) <= nseg) ++bits;
  size_t cub_bytes = w.cub_bytes;
  RBX_CUDA_OK(cub::DeviceRadixSort::SortPairs(w.cub_temp, cub_bytes, (const uint32_t *)w.keys, w.keys_out,
                                              (const uint32_t *)w.vals, w.vals_out, n, 0, bits, stream));
  count_launch(2);
  // chunks: enough blocks to fill the machine several times over, at most 64 particles per chunk
  int chunk = (int)((n + 148 * 32 - 1) / (148 * 32));
  chunk = chunk < 8 ? 8 : (chunk > 64 ? 64 : chunk);
  const int64_t nchunks = (n + chunk - 1) / chunk;
  const int grid = (int)(nchunks < 148 * 8 ? nchunks : 148 * 8);
  auto kern = kk <= 16 ? resample_dusty_cube_kernel<16> : resample_dusty_cube_kernel<32>;
  if (smem > 48 * 1024) RBX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, 256, smem, stream>>>(v, d_spectra, d_velocity, (const int32_t *)w.keys_out, w.vals_out, d_av, d_axav, n, nseg,
                                    chunk, d_cube);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}

extern "C" int rbx_dusty_moments(void) { return kDustMoments; }

extern "C" int rbx_dusty_bins(const float *d_av, const int32_t *d_pixel, const float *d_mass, const float *d_velocity,
                              const float *d_metallicity, const float *d_age, int64_t n, int nseg, int n_bins, float av0,
                              float step, int32_t *d_vpixel, float *d_wmass, float *d_vvelocity, float *d_vmetallicity,
                              float *d_vage, void *stream) {
  RBX_REQUIRE(n >= 0 && nseg > 0 && n_bins > 0 && step > 0.f, "rbx_dusty_bins: bad argument");
  RBX_REQUIRE((int64_t)nseg * n_bins * kDustMoments < (int64_t)1 << 31, "rbx_dusty_bins: spaxels x bins beyond int32");
  if (n == 0) return RBX_OK;
  RBX_REQUIRE(d_av && d_pixel && d_mass && d_velocity && d_metallicity && d_age && d_vpixel && d_wmass && d_vvelocity &&
                  d_vmetallicity && d_vage, "rbx_dusty_bins: null pointer");
  dusty_bins_kernel<<<grid1d(n, 256), 256, 0, (cudaStream_t)stream>>>(d_av, d_pixel, d_mass, d_velocity, d_metallicity, d_age,
                                                                      n, nseg, n_bins, av0, step, d_vpixel, d_wmass,
                                                                      d_vvelocity, d_vmetallicity, d_vage);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}

extern "C" int rbx_dusty_combine(const float *d_vcube, int nseg, int n_bins, float av0, float step, const float *d_axav, int W,
                                 float *d_cube, void *stream) {
  RBX_REQUIRE(d_vcube && d_axav && d_cube, "rbx_dusty_combine: null pointer");
  RBX_REQUIRE(nseg > 0 && n_bins > 0 && W > 0, "rbx_dusty_combine: bad shape");
  dusty_combine_kernel<<<nseg, 256, 0, (cudaStream_t)stream>>>(d_vcube, n_bins, av0, step, d_axav, W, d_cube);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}
