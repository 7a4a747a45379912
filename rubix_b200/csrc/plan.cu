// Plan creation: per-configuration device tables (SSP template + derivative tables, redshifted
// SSP wavelengths, telescope grid and its chunk-local helper tables).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace rbx {

static thread_local std::string g_error;
std::atomic<int64_t> g_launches{0};
void set_error(const std::string &msg) { g_error = msg; }

// interpax approx_df(method="cubic"): node derivative = one-sided secant at the ends, plain mean of
// the two adjacent secants inside.  One thread per output element of an (nz, na, Lp) table.
__global__ void approx_df_kernel(const float *__restrict__ x, const float *__restrict__ f,
                                 float *__restrict__ out, int nz, int na, int L, int Lp, int axis) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)nz * na * Lp;
  if (tid >= total) return;
  int l = (int)(tid % Lp);
  int a = (int)((tid / Lp) % na);
  int z = (int)(tid / ((size_t)Lp * na));
  if (l >= L) { out[tid] = 0.f; return; }
  int n = axis == 0 ? nz : na;
  int i = axis == 0 ? z : a;
  size_t stride = axis == 0 ? (size_t)na * Lp : (size_t)Lp;
  const float *base = f + tid - (size_t)i * stride;
  auto secant = [&](int k) {  // slope between node k and k+1
    float dx = x[k + 1] - x[k];
    float dxi = dx == 0.f ? 0.f : 1.f / dx;
    return dxi * (base[(size_t)(k + 1) * stride] - base[(size_t)k * stride]);
  };
  float v;
  if (n == 1) v = 0.f;
  else if (i == 0) v = secant(0);
  else if (i == n - 1) v = secant(n - 2);
  else v = 0.5f * (secant(i - 1) + secant(i));
  out[tid] = v;
}

// Channel lookup table of the fused kernel.  bucket[w] = bucket of channel w under bucket_offset();
// lut[b] = number of channels whose bucket is < b.  With at most one channel per bucket,
// #channels below x is lut[b(x)] + (t[lut[b(x)]] < x) for every x, because b() is monotone.
__global__ void lut_bucket_kernel(const float *__restrict__ t, int W, float tmin, float trange, float scale,
                                  int *__restrict__ bucket, int *__restrict__ ok) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  int b = bucket_offset(t[w], tmin, trange, scale) >> 1;
  bucket[w] = b;
  if (w > 0) {
    int bp = bucket_offset(t[w - 1], tmin, trange, scale) >> 1;
    if (bp >= b) atomicExch(ok, 0);
  }
}

// Is the telescope grid exactly t0 + w * delta evaluated in float32 (multiply, then add)?  That is how
// numpy / jnp.arange fill a float32 range (rubix/telescope/utils.py:53), so it holds for every
// telescope of the reference; the fused kernel then finds channels arithmetically.
__global__ void affine_check_kernel(const float *__restrict__ t, int W, float t0, float delta, int *__restrict__ ok) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  if (t[w] != __fadd_rn(__fmul_rn((float)w, delta), t0)) atomicExch(ok, 0);
}

__global__ void lut_fill_kernel(const int *__restrict__ bucket, int W, int nb, uint16_t *__restrict__ lut) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  int lo = 0, hi = W;  // first w with bucket[w] >= b
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (bucket[mid] < b) lo = mid + 1; else hi = mid;
  }
  lut[b] = (uint16_t)lo;
}

// Window table of the warp cube kernel: out[row][perm(s)] = tab[row][clamp(jbase + s, 0, L - 1)] for the 256 knot
// slots s of the kernel's window.  Lane l owns the slots 8l .. 8l+7 and fetches them with two 16-byte loads; perm puts
// the first halves of all 32 lanes into one contiguous 512-byte block and the second halves into the next, so each
// warp-wide load touches 4 cache lines instead of 8-9 (rows of the plain table: 32-byte lane stride).
__global__ void window_table_kernel(const float *__restrict__ tab, float *__restrict__ out, int rows, int L, int Lp,
                                    int jbase) {
  const int row = blockIdx.x, s = threadIdx.x;   // 256 threads
  if (row >= rows) return;
  const int l = s >> 3, r = s & 7;
  const int dst = r < 4 ? 4 * l + r : 128 + 4 * l + (r - 4);
  out[(size_t)row * 256 + dst] = tab[(size_t)row * Lp + min(max(jbase + s, 0), L - 1)];
}

}  // namespace rbx

using namespace rbx;

extern "C" const char *rbx_last_error(void) { return g_error.c_str(); }
extern "C" int rbx_version(void) { return 100; }
extern "C" int64_t rbx_launch_count(void) { return g_launches.load(); }

template <typename T>
static int upload(rbx_plan *pl, const T *h, size_t n, const T **d, cudaStream_t s) {
  void *p = nullptr;
  RBX_CUDA_OK(cudaMalloc(&p, n * sizeof(T) + 16));
  pl->allocs.push_back(p);
  RBX_CUDA_OK(cudaMemcpyAsync(p, h, n * sizeof(T), cudaMemcpyHostToDevice, s));
  *d = (const T *)p;
  return RBX_OK;
}

extern "C" int rbx_plan_create(rbx_plan **out, const float *h_met, int nz, const float *h_age, int na,
                               const float *h_wave, int L, const float *h_flux, const float *h_t, int W,
                               double redshift, int method, int vel_component, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RBX_REQUIRE(out && h_met && h_age && h_wave && h_flux && h_t, "rbx_plan_create: null argument");
  RBX_REQUIRE(nz >= 2 && na >= 2 && L >= 2 && W >= 1, "rbx_plan_create: need nz,na,L >= 2 and W >= 1");
  RBX_REQUIRE(method == RBX_METHOD_LINEAR || method == RBX_METHOD_CUBIC, "rbx_plan_create: unknown method");
  RBX_REQUIRE(vel_component >= 0 && vel_component <= 2, "rbx_plan_create: vel_component must be 0, 1 or 2");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("rbx_plan_create: no CUDA device");
    return RBX_ERR_NO_DEVICE;
  }
  rbx_plan *pl = new rbx_plan();
  cudaGetDevice(&pl->device);
  PlanView &v = pl->v;
  std::memset(&v, 0, sizeof(v));
  v.nz = nz; v.na = na; v.L = L; v.W = W; v.method = method; v.vel_comp = vel_component;
  v.Lp = (L + 3) & ~3;
  int rc;
#define TRY(x) do { rc = (x); if (rc != RBX_OK) { rbx_plan_destroy(pl); return rc; } } while (0)

  TRY(upload(pl, h_met, nz, &v.zgrid, stream));
  TRY(upload(pl, h_age, na, &v.agrid, stream));
  {  // prep_kernel's age search starts from a bucket table instead of bisecting 221 nodes (8 dependent steps)
    const int nbk = 1024;
    std::vector<uint16_t> lut(nbk, 0);
    const double a0 = h_age[0], a1 = h_age[na - 1];
    v.alut_scale = a1 > a0 ? (float)((double)nbk / (a1 - a0)) : 0.f;
    v.alut_n = nbk;
    for (int b = 0; b < nbk; ++b) {
      const double left = a0 + (a1 - a0) * (double)b / (double)nbk;
      int k = 0;
      while (k < na && !((double)h_age[k] > left)) ++k;   // elements <= the bucket's left edge
      lut[b] = (uint16_t)std::min(k, 65535);
    }
    TRY(upload(pl, lut.data(), lut.size(), &v.alut, stream));
  }

  // template, rows padded to a multiple of 4 floats so every row starts 16-byte aligned
  size_t rows = (size_t)nz * na;
  std::vector<float> padded(rows * v.Lp, 0.f);
  for (size_t r = 0; r < rows; ++r) std::memcpy(&padded[r * v.Lp], h_flux + r * L, sizeof(float) * L);
  TRY(upload(pl, padded.data(), padded.size(), &v.tab[0], stream));
  if (method == RBX_METHOD_CUBIC) {
    float *d[3];
    for (int k = 0; k < 3; ++k) {
      void *p = nullptr;
      if (cudaMalloc(&p, padded.size() * sizeof(float)) != cudaSuccess) {
        set_error("rbx_plan_create: cudaMalloc failed");
        rbx_plan_destroy(pl);
        return RBX_ERR_CUDA;
      }
      pl->allocs.push_back(p);
      d[k] = (float *)p;
      v.tab[k + 1] = d[k];
    }
    size_t total = padded.size();
    int blocks = (int)((total + 255) / 256);
    approx_df_kernel<<<blocks, 256, 0, stream>>>(v.zgrid, v.tab[0], d[0], nz, na, L, v.Lp, 0);  // fx
    approx_df_kernel<<<blocks, 256, 0, stream>>>(v.agrid, v.tab[0], d[1], nz, na, L, v.Lp, 1);  // fy
    approx_df_kernel<<<blocks, 256, 0, stream>>>(v.agrid, d[0], d[2], nz, na, L, v.Lp, 1);      // fxy = d/dy fx
    count_launch(3);
  }

  // (1+z) * wavelength in float32, exactly as rubix/spectra/ifu.py:80 evaluates it
  pl->h_lamz.resize(L);
  const float onepz = (float)(1.0 + redshift);
  for (int l = 0; l < L; ++l) pl->h_lamz[l] = onepz * h_wave[l];
  std::vector<float> rdl(L, 0.f);
  for (int l = 0; l + 1 < L; ++l) {
    float d = pl->h_lamz[l + 1] - pl->h_lamz[l];
    rdl[l] = d > 0.f ? 1.f / d : 0.f;
  }
  TRY(upload(pl, pl->h_lamz.data(), (size_t)L, &v.lamz, stream));
  TRY(upload(pl, rdl.data(), (size_t)L, &v.rdl, stream));

  // telescope grid
  pl->h_t.assign(h_t, h_t + W);
  std::vector<float> dt(W, 0.f);
  std::vector<float2> tt(W), q(W + 1);
  float tmin = h_t[0], tmax = h_t[0], min_dt = 3.0e38f, max_dt = 0.f;
  for (int w = 0; w < W; ++w) {
    if (w > 0) {
      dt[w] = h_t[w] - h_t[w - 1];
      min_dt = std::fmin(min_dt, dt[w]);
      max_dt = std::fmax(max_dt, dt[w]);
    }
    tmin = std::fmin(tmin, h_t[w]);
    tmax = std::fmax(tmax, h_t[w]);
    tt[w] = make_float2(h_t[w > 0 ? w - 1 : 0], h_t[w]);
  }
  v.tmin = tmin; v.tmax = tmax; v.trange = tmax - tmin;
  v.tref = h_t[W / 2];
  pl->min_dt = min_dt; pl->max_dt = max_dt;
  {  // q[k] = sum_{w<k} dt[w] * (t[w] - tref) as an unevaluated float pair (hi, lo)
    double acc = 0.0;
    for (int k = 0; k <= W; ++k) {
      float hi = (float)acc;
      q[k] = make_float2(hi, (float)(acc - (double)hi));
      if (k < W) acc += (double)dt[k] * ((double)h_t[k] - (double)v.tref);
    }
  }
  TRY(upload(pl, h_t, (size_t)W, &v.t, stream));
  TRY(upload(pl, dt.data(), (size_t)W, &v.dt, stream));
  TRY(upload(pl, tt.data(), (size_t)W, &v.tt, stream));
  TRY(upload(pl, q.data(), (size_t)W + 1, &v.q, stream));
  // affine grid test (on the device: same float32 operations as the kernel)
  v.affine = 0; v.t0 = h_t[0]; v.tdelta = W > 1 ? h_t[1] - h_t[0] : 1.f; v.tinv = 1.f / v.tdelta;
  if (W >= 2 && v.tdelta > 0.f) {
    void *d_ok = nullptr;
    if (cudaMalloc(&d_ok, sizeof(int)) != cudaSuccess) {
      set_error("rbx_plan_create: cudaMalloc failed");
      rbx_plan_destroy(pl);
      return RBX_ERR_CUDA;
    }
    pl->allocs.push_back(d_ok);
    int one = 1, ok = 0;
    cudaMemcpyAsync(d_ok, &one, sizeof(int), cudaMemcpyHostToDevice, stream);
    affine_check_kernel<<<(W + 255) / 256, 256, 0, stream>>>(v.t, W, v.t0, v.tdelta, (int *)d_ok);
    count_launch();
    cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (cudaStreamSynchronize(stream) == cudaSuccess) v.affine = ok;
  }
  // channel lookup table, built on the device with the kernel's own bucket function
  pl->lut_ok = 0;
  v.nb = 0; v.lut = nullptr; v.lut_scale = 0.f;
  if (W >= 2 && min_dt > 0.f && v.trange > 0.f) {
    double want = (double)v.trange / (double)min_dt * 1.03 + 4.0;
    if (want <= (double)kMaxLutBuckets) {
      v.nb = (int)want;
      v.lut_scale = (float)((2.0 * v.nb - 1.0) / (double)v.trange);
      void *d_bucket = nullptr, *d_ok = nullptr, *d_lut = nullptr;
      if (cudaMalloc(&d_bucket, sizeof(int) * W) != cudaSuccess || cudaMalloc(&d_ok, sizeof(int)) != cudaSuccess ||
          cudaMalloc(&d_lut, sizeof(uint16_t) * v.nb + 16) != cudaSuccess) {
        set_error("rbx_plan_create: cudaMalloc failed");
        rbx_plan_destroy(pl);
        return RBX_ERR_CUDA;
      }
      pl->allocs.push_back(d_bucket); pl->allocs.push_back(d_ok); pl->allocs.push_back(d_lut);
      int one = 1;
      cudaMemcpyAsync(d_ok, &one, sizeof(int), cudaMemcpyHostToDevice, stream);
      lut_bucket_kernel<<<(W + 255) / 256, 256, 0, stream>>>(v.t, W, v.tmin, v.trange, v.lut_scale, (int *)d_bucket, (int *)d_ok);
      lut_fill_kernel<<<(v.nb + 255) / 256, 256, 0, stream>>>((const int *)d_bucket, W, v.nb, (uint16_t *)d_lut);
      count_launch(2);
      int ok = 0;
      cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, stream);
      if (cudaStreamSynchronize(stream) == cudaSuccess) pl->lut_ok = ok;
      v.lut = (const uint16_t *)d_lut;
    }
  }
  {  // knot positions in channel units of the affine grid, rounded once from double (fused.cu: knot_ab)
    std::vector<float> ka(L, 0.f), kb(L, 0.f);
    if (v.affine) {
      const double t0 = (double)v.t0, dl = (double)v.tdelta;
      for (int l = 0; l < L; ++l) {
        const double lz = (double)pl->h_lamz[l];
        ka[l] = (float)(lz / dl);
        kb[l] = (float)((lz - t0) / dl - 0.5);
      }
    }
    TRY(upload(pl, ka.data(), (size_t)L, &v.ka, stream));
    TRY(upload(pl, kb.data(), (size_t)L, &v.kb, stream));
  }
  // Window tables for the warp cube kernel (affine grids): the knot window is fixed per plan -- the knots that reach
  // the band at rest (+3 either side, as segment_kernel counts them) with the spare slots split 1 : 2 between the blue
  // and the red end (the red end moves twice as many knots per unit of Doppler shift).  segment_kernel checks that
  // the Doppler range actually present stays inside it (BC03 on MUSE: |v| up to ~0.045 c) and hands wider ranges to the
  // group kernel.
  v.wt_jbase = 0;
  for (int k = 0; k < 4; ++k) v.wt[k] = nullptr;
  if (v.affine) {
    int a0 = 0;
    while (a0 < L && pl->h_lamz[a0] < tmin) ++a0;
    int b0 = a0;
    while (b0 < L && pl->h_lamz[b0] <= tmax) ++b0;
    const int ja0 = std::max(0, a0 - 3), jb0 = std::min(L, b0 + 3);
    const int need = jb0 - ja0 + 2;   // slots ja0 - 1 .. jb0
    if (need <= 256) {
      const int slack = 256 - need;
      v.wt_jbase = ((ja0 - 1) - slack / 3) & ~3;
      const int ntab = method == RBX_METHOD_CUBIC ? 4 : 1;
      for (int k = 0; k < ntab; ++k) {
        void *p = nullptr;
        if (cudaMalloc(&p, rows * 256 * sizeof(float)) != cudaSuccess) {
          set_error("rbx_plan_create: cudaMalloc failed");
          rbx_plan_destroy(pl);
          return RBX_ERR_CUDA;
        }
        pl->allocs.push_back(p);
        window_table_kernel<<<(int)rows, 256, 0, stream>>>(v.tab[k], (float *)p, (int)rows, L, v.Lp, v.wt_jbase);
        count_launch();
        v.wt[k] = (const float *)p;
      }
    }
  }
#undef TRY
  // the staging vectors above die at return: make sure the async copies have consumed them
  if (cudaStreamSynchronize(stream) != cudaSuccess) {
    set_error("rbx_plan_create: stream sync failed");
    rbx_plan_destroy(pl);
    return RBX_ERR_CUDA;
  }
  *out = pl;
  return RBX_OK;
}

extern "C" int rbx_plan_destroy(rbx_plan *pl) {
  if (!pl) return RBX_OK;
  for (void *p : pl->allocs) cudaFree(p);
  delete pl;
  return RBX_OK;
}

extern "C" int rbx_plan_dims(const rbx_plan *pl, int *nz, int *na, int *L, int *W) {
  RBX_REQUIRE(pl, "rbx_plan_dims: null plan");
  if (nz) *nz = pl->v.nz;
  if (na) *na = pl->v.na;
  if (L) *L = pl->v.L;
  if (W) *W = pl->v.W;
  return RBX_OK;
}
