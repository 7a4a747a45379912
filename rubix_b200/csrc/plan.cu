// Plan creation: per-configuration device tables (SSP template + derivative tables, redshifted
// SSP wavelengths, telescope grid and its chunk-local helper tables).
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
  v.nchunks = (W + kChunk - 1) / kChunk;
  int rc;
#define TRY(x) do { rc = (x); if (rc != RBX_OK) { rbx_plan_destroy(pl); return rc; } } while (0)

  TRY(upload(pl, h_met, nz, &v.zgrid, stream));
  TRY(upload(pl, h_age, na, &v.agrid, stream));

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
  std::vector<float> dt(W, 0.f), tau(W), tc(v.nchunks);
  std::vector<float2> suf(W);
  float tmin = h_t[0], tmax = h_t[0];
  for (int w = 0; w < W; ++w) {
    if (w > 0) dt[w] = h_t[w] - h_t[w - 1];
    tmin = std::fmin(tmin, h_t[w]);
    tmax = std::fmax(tmax, h_t[w]);
  }
  v.tmin = tmin; v.tmax = tmax;
  for (int c = 0; c < v.nchunks; ++c) {
    int w0 = c * kChunk, w1 = std::min(W, w0 + kChunk);
    tc[c] = h_t[std::min(W - 1, w0 + kChunk / 2)];
    double sd = 0, st = 0;
    for (int w = w1 - 1; w >= w0; --w) {
      tau[w] = h_t[w] - tc[c];
      sd += (double)dt[w];
      st += (double)tau[w] * (double)dt[w];
      suf[w] = make_float2((float)sd, (float)st);
    }
  }
  TRY(upload(pl, h_t, (size_t)W, &v.t, stream));
  TRY(upload(pl, dt.data(), (size_t)W, &v.dt, stream));
  TRY(upload(pl, tau.data(), (size_t)W, &v.tau, stream));
  TRY(upload(pl, suf.data(), (size_t)W, &v.suf, stream));
  TRY(upload(pl, tc.data(), (size_t)v.nchunks, &v.tc, stream));
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
