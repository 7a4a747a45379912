// Stage kernels: one per reference stage, materialising the same intermediates as the reference.
// They serve the stepwise (notebook) path, stage-level parity tests, and configurations outside the
// fused kernel's limits.  All are bandwidth-bound streaming kernels: coalesced along the wavelength
// axis, one block per particle for the (n, L) / (n, W) producers.
#include "common.cuh"

namespace rbx {

constexpr int kMaxEdgesSmem = 4096;

// ---- a0: spaxel assignment + aperture mask + filter ----------------------------------------
// rubix/telescope/utils.py:138-151 (digitize = searchsorted 'right', -1, clip, x + nb*y) and
// :170-174 (mask, inclusive), rubix/core/telescope.py:155-174 (where(mask, x, 0)).
__global__ void spaxel_assign_kernel(const float *__restrict__ coords, int64_t n,
                                     const float *__restrict__ edges, int n_edges,
                                     int32_t *__restrict__ pixel, uint8_t *__restrict__ mask,
                                     float *__restrict__ mass, float *__restrict__ met,
                                     float *__restrict__ age, int mark_outside) {
  extern __shared__ float s_edges[];
  const float *e = edges;
  if (n_edges <= kMaxEdgesSmem) {
    for (int i = threadIdx.x; i < n_edges; i += blockDim.x) s_edges[i] = edges[i];
    __syncthreads();
    e = s_edges;
  }
  float lo = e[0], hi = e[0];
  // edges come from arange (increasing) but the reference takes min()/max(); honour that
  for (int i = 1; i < n_edges; ++i) { lo = fminf(lo, e[i]); hi = fmaxf(hi, e[i]); }
  const int nb = n_edges - 1;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    float x = coords[3 * p], y = coords[3 * p + 1];
    if (pixel) {
      int xi = min(max(ss_right(e, n_edges, x) - 1, 0), nb - 1);
      int yi = min(max(ss_right(e, n_edges, y) - 1, 0), nb - 1);
      pixel[p] = xi + nb * yi;
    }
    bool in = (x >= lo) && (x <= hi) && (y >= lo) && (y <= hi);
    if (mask) mask[p] = in ? 1 : 0;
    if (mark_outside && !in && pixel) pixel[p] = -1;  // dropped downstream like a zero-mass particle
    if (!in) {
      if (mass) mass[p] = 0.f;
      if (met) met[p] = 0.f;
      if (age) age[p] = 0.f;
    }
  }
}

// ---- a1: SSP lookup -> (n, L) ----------------------------------------------------------------
// rubix/core/ifu.py:95-118 (the 250k chunking there only bounds XLA memory; the values are per
// particle).  One block per particle; each thread evaluates the (<=16) interpolation terms once
// and streams its share of the L wavelength bins.
__global__ void ssp_lookup_kernel(PlanView p, const float *__restrict__ met, const float *__restrict__ age,
                                  int64_t n, float *__restrict__ out) {
  for (int64_t q = blockIdx.x; q < n; q += gridDim.x) {
    SspTerms tm;
    ssp_terms(p, met[q], age[q], 1.0f, tm);
    float *o = out + q * p.L;
    if (tm.n == 0) {
      for (int l = threadIdx.x; l < p.L; l += blockDim.x) o[l] = 0.f;
      continue;
    }
    for (int l = threadIdx.x; l < p.L; l += blockDim.x) {
      float acc = 0.f;
      for (int k = 0; k < tm.n; ++k) acc = fmaf(tm.w[k], p.tab[tm.tabid[k]][(size_t)tm.row[k] * p.Lp + l], acc);
      o[l] = acc;
    }
  }
}

// ---- a2: mass scaling -------------------------------------------------------------------------
// rubix/core/ifu.py:152-154: exactly one float32 multiply per element (test_core_ifu.py:277 uses
// array_equal on it).
__global__ void scale_by_mass_kernel(const float *__restrict__ spec, const float *__restrict__ mass,
                                     int64_t n, int L, float *__restrict__ out) {
  size_t total = (size_t)n * L;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    out[i] = __fmul_rn(spec[i], mass[i / L]);
  }
}

// ---- a3 + a4: Doppler shift and flux-conserving resample -> (n, W) --------------------------
// rubix/spectra/ifu.py:190 (lam' = lam_z * exp(v/c)) and :241-260 (mask, total, jnp.interp, new
// total, nan_to_num(total/new), scale).  One block per particle; lam' and the spectrum live in
// shared memory; each thread binary-searches its own targets (jnp.interp semantics:
// i = clip(searchsorted(xp, x, 'right'), 1, L-1); fp[i-1] + (delta/dx)*df; end values outside).
__global__ void doppler_resample_kernel(PlanView p, const float *__restrict__ spec,
                                        const float *__restrict__ vel, int64_t n, float *__restrict__ out) {
  extern __shared__ float sm[];
  float *lam = sm;
  float *s = sm + p.L;
  __shared__ float red[32];
  __shared__ float bc[2];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int64_t q = blockIdx.x; q < n; q += gridDim.x) {
    const float d = expf(vel[3 * q + p.vel_comp] / kSpeedOfLight);
    const float *sp = spec + q * p.L;
    for (int l = threadIdx.x; l < p.L; l += blockDim.x) {
      lam[l] = __fmul_rn(p.lamz[l], d);
      s[l] = sp[l];
    }
    __syncthreads();
    float tot = 0.f;
    for (int l = threadIdx.x + 1; l < p.L; l += blockDim.x) {
      float x = lam[l];
      if (x >= p.tmin && x <= p.tmax) tot += s[l] * (x - lam[l - 1]);
    }
    float nw_sum = 0.f;
    float *o = out + q * p.W;
    const float eps = 1.4210855e-14f;  // np.spacing(np.finfo(float32).eps)
    for (int w = threadIdx.x; w < p.W; w += blockDim.x) {
      float x = p.t[w];
      int i = min(max(ss_right(lam, p.L, x), 1), p.L - 1);
      float df = s[i] - s[i - 1], dx = lam[i] - lam[i - 1], delta = x - lam[i - 1];
      float f = (fabsf(dx) <= eps) ? s[i - 1] : s[i - 1] + __fdiv_rn(delta, dx) * df;
      if (x < lam[0]) f = s[0];
      if (x > lam[p.L - 1]) f = s[p.L - 1];
      o[w] = f;
      nw_sum += f * p.dt[w];
    }
    tot = warp_sum(tot);
    nw_sum = warp_sum(nw_sum);
    if (lane == 0) { red[wid] = tot; red[16 + wid] = nw_sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, b = 0.f;
      for (int k = 0; k < nw; ++k) { a += red[k]; b += red[16 + k]; }
      bc[0] = nan_to_num0(a / b);
    }
    __syncthreads();
    const float scale = bc[0];
    for (int w = threadIdx.x; w < p.W; w += blockDim.x) o[w] = o[w] * scale;
    __syncthreads();
  }
}

// ---- a5: segment sum --------------------------------------------------------------------------
// rubix/spectra/ifu.py:286: jax.ops.segment_sum(spectra, idx, num_segments): ids outside the range
// are dropped.  Stage version: float atomics (RED.ADD.F32) straight into the cube.
__global__ void segment_sum_kernel(const float *__restrict__ spec, const int32_t *__restrict__ pixel,
                                   int64_t n, int W, int nseg, float *__restrict__ cube) {
  for (int64_t q = blockIdx.x; q < n; q += gridDim.x) {
    int id = pixel[q];
    if (id < 0 || id >= nseg) continue;
    const float *s = spec + q * W;
    float *c = cube + (size_t)id * W;
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
      float v = s[w];
      if (v != 0.f) atomicAdd(c + w, v);
    }
  }
}

// a5, bit-reproducible form: the particles of a spaxel are added ONE AFTER THE OTHER in particle order (the order of
// the stable sort's run), a thread per channel -- the float32 sum the reference's CPU segment_sum forms (SURVEY 8a a5:
// "order = particle order on CPU"), with no atomics.
__global__ void __launch_bounds__(128)
segment_sum_sorted_kernel(const float *__restrict__ spec, const int32_t *__restrict__ order,
                          const int32_t *__restrict__ offsets, int W, int nseg, float *__restrict__ cube) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  for (int s = blockIdx.y; s < nseg; s += gridDim.y) {
    const int lo = offsets[s], hi = offsets[s + 1];
    float acc = 0.f;
#pragma unroll 4
    for (int i = lo; i < hi; ++i) acc += spec[(size_t)order[i] * W + w];
    cube[(size_t)s * W + w] = acc;
  }
}

static int grid_for(int64_t n, int per_block) {
  int64_t b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > 148 * 64) b = 148 * 64;
  return (int)b;
}

}  // namespace rbx

using namespace rbx;

static int spaxel_common(const float *d_coords, int64_t n, const float *d_edges, int n_edges, int32_t *d_pixel,
                         uint8_t *d_mask, float *mass, float *met, float *age, cudaStream_t stream,
                         int mark_outside = 0) {
  RBX_REQUIRE(n >= 0 && n_edges >= 2, "spaxel assignment: need n >= 0 and at least 2 bin edges");
  if (n == 0) return RBX_OK;
  RBX_REQUIRE(d_coords && d_edges, "spaxel assignment: null pointer");
  size_t smem = n_edges <= kMaxEdgesSmem ? sizeof(float) * n_edges : 0;
  spaxel_assign_kernel<<<grid_for(n, 256), 256, smem, stream>>>(d_coords, n, d_edges, n_edges, d_pixel, d_mask,
                                                                 mass, met, age, mark_outside);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}

extern "C" int rbx_spaxel_assign(const float *d_coords, int64_t n, const float *d_edges, int n_edges,
                                 int32_t *d_pixel, uint8_t *d_mask, void *stream) {
  RBX_REQUIRE(d_pixel || n == 0, "rbx_spaxel_assign: null output");
  return spaxel_common(d_coords, n, d_edges, n_edges, d_pixel, d_mask, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int rbx_filter_particles(const float *d_coords, int64_t n, const float *d_edges, int n_edges,
                                    float *d_mass, float *d_met, float *d_age, uint8_t *d_mask, void *stream) {
  return spaxel_common(d_coords, n, d_edges, n_edges, nullptr, d_mask, d_mass, d_met, d_age, (cudaStream_t)stream);
}

extern "C" int rbx_filter_and_assign(const float *d_coords, int64_t n, const float *d_edges, int n_edges,
                                     float *d_mass, float *d_met, float *d_age, int32_t *d_pixel, uint8_t *d_mask,
                                     void *stream) {
  RBX_REQUIRE(d_pixel || n == 0, "rbx_filter_and_assign: null output");
  const int mark = (!d_mass && !d_met && !d_age) ? 1 : 0;
  return spaxel_common(d_coords, n, d_edges, n_edges, d_pixel, d_mask, d_mass, d_met, d_age, (cudaStream_t)stream, mark);
}

extern "C" int rbx_ssp_lookup(const rbx_plan *plan, const float *d_met, const float *d_age, int64_t n,
                              float *d_spectra, void *stream) {
  RBX_REQUIRE(plan, "rbx_ssp_lookup: null plan");
  if (n == 0) return RBX_OK;
  RBX_REQUIRE(d_met && d_age && d_spectra && n > 0, "rbx_ssp_lookup: null pointer");
  ssp_lookup_kernel<<<grid_for(n, 1), 256, 0, (cudaStream_t)stream>>>(plan->v, d_met, d_age, n, d_spectra);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}

extern "C" int rbx_scale_by_mass(const float *d_spectra, const float *d_mass, int64_t n, int L, float *d_out,
                                 void *stream) {
  if (n == 0) return RBX_OK;
  RBX_REQUIRE(d_spectra && d_mass && d_out && n > 0 && L > 0, "rbx_scale_by_mass: bad argument");
  scale_by_mass_kernel<<<grid_for(n * (int64_t)L, 256 * 8), 256, 0, (cudaStream_t)stream>>>(d_spectra, d_mass, n, L, d_out);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}

extern "C" int rbx_doppler_resample(const rbx_plan *plan, const float *d_spectra, const float *d_velocity,
                                    int64_t n, float *d_out, void *stream) {
  RBX_REQUIRE(plan, "rbx_doppler_resample: null plan");
  if (n == 0) return RBX_OK;
  RBX_REQUIRE(d_spectra && d_velocity && d_out && n > 0, "rbx_doppler_resample: null pointer");
  size_t smem = sizeof(float) * 2 * plan->v.L;
  if (smem > 200 * 1024) {
    set_error("rbx_doppler_resample: SSP wavelength grid too long for shared memory");
    return RBX_ERR_UNSUPPORTED;
  }
  if (smem > 48 * 1024)
    RBX_CUDA_OK(cudaFuncSetAttribute(doppler_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  doppler_resample_kernel<<<grid_for(n, 1), 256, smem, (cudaStream_t)stream>>>(plan->v, d_spectra, d_velocity, n, d_out);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}

extern "C" int rbx_segment_sum(const float *d_spectra, const int32_t *d_pixel, int64_t n, int W, int nseg,
                               float *d_cube, int zero_first, void *stream) {
  RBX_REQUIRE(d_cube && W > 0 && nseg > 0, "rbx_segment_sum: bad argument");
  if (zero_first) RBX_CUDA_OK(cudaMemsetAsync(d_cube, 0, sizeof(float) * (size_t)nseg * W, (cudaStream_t)stream));
  if (n == 0) return RBX_OK;
  RBX_REQUIRE(d_spectra && d_pixel && n > 0, "rbx_segment_sum: null pointer");
  segment_sum_kernel<<<grid_for(n, 1), 256, 0, (cudaStream_t)stream>>>(d_spectra, d_pixel, n, W, nseg, d_cube);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}

extern "C" int rbx_segment_sum_sorted(const float *d_spectra, const int32_t *d_order, const int32_t *d_offsets,
                                      int W, int nseg, float *d_cube, void *stream) {
  RBX_REQUIRE(d_cube && W > 0 && nseg > 0, "rbx_segment_sum_sorted: bad argument");
  // d_spectra / d_order may be NULL for an empty particle set (all offsets equal: nothing is dereferenced)
  RBX_REQUIRE(d_offsets, "rbx_segment_sum_sorted: null offsets");
  const dim3 grid((unsigned)((W + 127) / 128), (unsigned)std::min(nseg, 32768));
  segment_sum_sorted_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(d_spectra, d_order, d_offsets, W, nseg, d_cube);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}
