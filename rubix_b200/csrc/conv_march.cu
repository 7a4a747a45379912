// a6 + a7 in one pass for HOST-known taps: rbx_psf_lsf_taps.
//
// The reference builds both kernels on the host from the config (rubix/telescope/psf/kernels.py:5-31,
// rubix/telescope/lsf/lsf.py:12-26), so the taps are known when the launch is issued.  That buys three
// things the device-tap entry point (rbx_psf_lsf, conv.cu) cannot have:
//   * the taps travel as kernel parameters: every FFMA reads its tap from the constant bank, no
//     registers and no loads are spent on them;
//   * a PSF that is an outer product (every Gaussian is) runs as two 1-D passes: P + P instead of P * P
//     FMAs per voxel;
//   * LSF taps that cannot change a float32 sum (|k| < 1e-14 max|k|: for the MUSE config sigma = 0.5 A on
//     a 1.25 A grid that is every tap beyond +-3 of the 25) are dropped, which also shrinks the spectral
//     halo a block has to load from 24 to 6 channels.
// With 17 instead of 50 FMAs per voxel the kernel is bound by HBM (8 bytes per voxel).
//
// psf_lsf_march_kernel: a block owns TX spaxel columns x TL channels and MARCHES along y.  A thread is one
// channel of the tile (lanes along lambda: every global access is a coalesced 128-byte row); it keeps
// the P partially summed output rows of its TX columns in registers, so each input voxel is loaded
// exactly once per block, x-convolved in registers and folded into the P rows it contributes to.  A
// finished row goes through a double-buffered shared-memory tile for the LSF along lambda (the only
// cross-thread exchange, one __syncthreads per row) and is stored coalesced.
#include <cmath>

#include "common.cuh"

namespace rbx {

constexpr int kMarchMaxP = 7;
constexpr int kMarchMaxK = 25;

struct MarchTaps {
  float kx[kMarchMaxP];   // PSF factor along x (columns of the reference's kernel)
  float ky[kMarchMaxP];   // PSF factor along y (rows)
  float kl[kMarchMaxK];   // effective LSF window, reversed: out[w] = sum_u kl[u] * mid[w - he + u]
};

template <int P, int KE, int TX, int NT>
__global__ void __launch_bounds__(NT)
psf_lsf_march_kernel(const float *__restrict__ in, float *__restrict__ out, int ny, int nx, int W,
                     int rows_per_seg, const __grid_constant__ MarchTaps tp) {
  constexpr int TL = NT - (KE - 1);      // output channels per block
  constexpr int HE = (KE - 1) / 2;       // spectral halo on each side
  constexpr int C = (P - 1) / 2;         // jax "same": out[y] = sum_m K[m] in[y - m + C]
  constexpr int H = P - 1 - C;           // rows / columns before the tile
  constexpr int IX = TX + P - 1;
  __shared__ float s_mid[2][TX][NT];

  const int tiles_x = (nx + TX - 1) / TX;
  const int x0 = (blockIdx.x % tiles_x) * TX;
  const int w0 = (blockIdx.x / tiles_x) * TL;
  const int ya = blockIdx.y * rows_per_seg;
  const int yb = min(ya + rows_per_seg, ny);
  const int c = threadIdx.x;
  const int q = w0 - HE + c;             // input channel of this thread
  const bool qok = q >= 0 && q < W;
  const bool xin = x0 - H >= 0 && x0 + TX + C <= nx;
  const size_t rowstride = (size_t)nx * W;
  const float *colbase = in + ((ptrdiff_t)(x0 - H)) * W + q;   // + yy * rowstride
  const int nsteps = (yb - ya) + P - 1;
  const bool lsf_thread = c < TL && w0 + c < W;

  float A[P][TX];
#pragma unroll
  for (int a = 0; a < P; ++a)
#pragma unroll
    for (int b = 0; b < TX; ++b) A[a][b] = 0.f;

  int buf = 0;
  for (int i0 = 0; i0 < nsteps; i0 += P) {
#pragma unroll
    for (int u = 0; u < P; ++u) {
      const int i = i0 + u;
      if (i >= nsteps) break;            // block-uniform
      const int yy = ya - H + i;         // input row of this step
      float v[IX];
      const bool rowok = qok && yy >= 0 && yy < ny;
      const float *src = colbase + (ptrdiff_t)yy * (ptrdiff_t)rowstride;
      if (rowok && xin) {
#pragma unroll
        for (int ix = 0; ix < IX; ++ix) v[ix] = __ldg(src + (size_t)ix * W);
      } else {
#pragma unroll
        for (int ix = 0; ix < IX; ++ix) {
          const int xx = x0 - H + ix;
          v[ix] = (rowok && xx >= 0 && xx < nx) ? __ldg(src + (ptrdiff_t)ix * W) : 0.f;
        }
      }
      // x pass: h[ox] = sum_n kx[n] * in[x - n + C]
      float h[TX];
#pragma unroll
      for (int ox = 0; ox < TX; ++ox) {
        float acc = tp.kx[0] * v[ox + P - 1];
#pragma unroll
        for (int n = 1; n < P; ++n) acc = fmaf(tp.kx[n], v[ox + P - 1 - n], acc);
        h[ox] = acc;
      }
      // y pass: this input row feeds output rows yy - C + m through tap m; slot (u + m) % P is static.
      // Tap P-1 is a row's first contribution (it initialises the slot), tap 0 its last.
#pragma unroll
      for (int m = 0; m < P - 1; ++m)
#pragma unroll
        for (int ox = 0; ox < TX; ++ox) A[(u + m) % P][ox] = fmaf(tp.ky[m], h[ox], A[(u + m) % P][ox]);
#pragma unroll
      for (int ox = 0; ox < TX; ++ox)
        A[(u + P - 1) % P][ox] = (P == 1) ? tp.ky[0] * h[ox] : tp.ky[P - 1] * h[ox];
      if (P == 1 || i >= P - 1) {
        const int oy = yy - C;           // finished output row
        // for P == 1 slot 0 was just written; for P > 1 slot u got its last tap (m = 0) above
#pragma unroll
        for (int ox = 0; ox < TX; ++ox) s_mid[buf][ox][c] = A[u % P][ox];
        __syncthreads();
        if (lsf_thread) {
          float *dst = out + ((size_t)oy * nx + x0) * W + (w0 + c);
#pragma unroll
          for (int ox = 0; ox < TX; ++ox) {
            if (x0 + ox < nx) {
              float acc = tp.kl[0] * s_mid[buf][ox][c];
#pragma unroll
              for (int t = 1; t < KE; ++t) acc = fmaf(tp.kl[t], s_mid[buf][ox][c + t], acc);
              dst[(size_t)ox * W] = acc;
            }
          }
        }
        buf ^= 1;
      }
    }
  }
}

}  // namespace rbx

using namespace rbx;

namespace {

struct TapPlan {
  MarchTaps taps;
  int P, KE;
};

// Rank-1 test of the PSF in double: K ~= u (x) v with u = the pivot column, v = the pivot row / pivot.
// Accepts when no tap is off by more than 3e-7 of the largest tap (the float32 rounding of the taps
// themselves is 6e-8 relative).
bool separable_factors(const float *K, int M, int N, float *ky, float *kx) {
  int im = 0, jm = 0;
  double best = 0.0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j)
      if (std::fabs((double)K[i * N + j]) > best) { best = std::fabs((double)K[i * N + j]); im = i; jm = j; }
  if (!(best > 0.0)) return false;
  const double piv = K[im * N + jm];
  for (int i = 0; i < M; ++i) ky[i] = K[i * N + jm];
  for (int j = 0; j < N; ++j) kx[j] = (float)((double)K[im * N + j] / piv);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j)
      if (std::fabs((double)ky[i] * (double)kx[j] - (double)K[i * N + j]) > 3e-7 * best) return false;
  return true;
}

template <int P, int KE, int TX, int NT>
int launch_march(const float *d_in, float *d_out, int ny, int nx, int W, const MarchTaps &taps, cudaStream_t stream) {
  auto kernel = psf_lsf_march_kernel<P, KE, TX, NT>;
  constexpr int TL = NT - (KE - 1);
  static int slots_cached[64] = {};
  int dev = 0;
  RBX_CUDA_OK(cudaGetDevice(&dev));
  int slots = (dev >= 0 && dev < 64) ? slots_cached[dev] : 0;
  if (!slots) {
    int per_sm = 0, sms = 0;
    RBX_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NT, 0));
    RBX_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    slots = std::max(1, per_sm * sms);
    if (dev >= 0 && dev < 64) slots_cached[dev] = slots;
  }
  const int bx = ((nx + TX - 1) / TX) * ((W + TL - 1) / TL);
  // y segments: enough blocks to fill the machine, as few re-read halo rows as possible, full waves
  int best_seg = 1;
  double best_cost = 1e300;
  for (int nseg = 1; nseg <= std::max(1, ny / 4) && nseg <= 65535; ++nseg) {
    const int rows = (ny + nseg - 1) / nseg;
    const int segs = (ny + rows - 1) / rows;
    const double blocks = (double)bx * segs;
    const double waves = std::ceil(blocks / slots);
    const double cost = waves * (rows + P - 1);   // time ~ waves x input rows marched per block
    if (cost < best_cost * 0.999) { best_cost = cost; best_seg = segs; }
  }
  const int rows = (ny + best_seg - 1) / best_seg;
  dim3 grid(bx, (ny + rows - 1) / rows);
  kernel<<<grid, NT, 0, stream>>>(d_in, d_out, ny, nx, W, rows, taps);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}

template <int P, int KE>
int launch_march_tx(const float *d_in, float *d_out, int ny, int nx, int W, const MarchTaps &taps, cudaStream_t stream) {
  if (nx >= 40) return launch_march<P, KE, 10, 128>(d_in, d_out, ny, nx, W, taps, stream);
  return launch_march<P, KE, 5, 128>(d_in, d_out, ny, nx, W, taps, stream);
}

template <int P>
int launch_march_ke(int KE, const float *d_in, float *d_out, int ny, int nx, int W, const MarchTaps &taps,
                    cudaStream_t stream) {
  switch (KE) {
    case 1: return launch_march_tx<P, 1>(d_in, d_out, ny, nx, W, taps, stream);
    case 7: return launch_march_tx<P, 7>(d_in, d_out, ny, nx, W, taps, stream);
    case 13: return launch_march_tx<P, 13>(d_in, d_out, ny, nx, W, taps, stream);
    default: return launch_march_tx<P, 25>(d_in, d_out, ny, nx, W, taps, stream);
  }
}

}  // namespace

// Host-tap PSF + LSF.  h_psf (M, N) or NULL (identity); h_lsf (K = 2 ext + 1) or NULL (identity).
// Returns RBX_ERR_UNSUPPORTED (no error text) when the taps do not fit the marching kernel (PSF not an
// outer product, not square 1/3/5/7, or more than 25 significant LSF taps): the caller then uses the
// device-tap kernels of conv.cu.
extern "C" int rbx_psf_lsf_taps(const float *d_in, float *d_out, int ny, int nx, int W, const float *h_psf, int M,
                                int N, const float *h_lsf, int K, int ext, void *stream) {
  RBX_REQUIRE(d_in && d_out && d_in != d_out, "rbx_psf_lsf_taps: bad pointers (no aliasing)");
  RBX_REQUIRE(ny > 0 && nx > 0 && W > 0, "rbx_psf_lsf_taps: bad shape");
  RBX_REQUIRE(h_psf || h_lsf, "rbx_psf_lsf_taps: at least one kernel is needed");
  TapPlan tp = {};
  tp.P = 1; tp.KE = 1;
  tp.taps.kx[0] = tp.taps.ky[0] = 1.f;
  tp.taps.kl[0] = 1.f;
  if (h_psf) {
    RBX_REQUIRE(M > 0 && N > 0, "rbx_psf_lsf_taps: bad PSF shape");
    // jax.scipy.signal.convolve2d: "One input must be smaller than the other in every dimension."
    RBX_REQUIRE((M <= ny && N <= nx) || (M >= ny && N >= nx),
                "One input must be smaller than the other in every dimension.");
    if (!(M == N && (M == 1 || M == 3 || M == 5 || M == 7))) return RBX_ERR_UNSUPPORTED;
    if (!separable_factors(h_psf, M, N, tp.taps.ky, tp.taps.kx)) return RBX_ERR_UNSUPPORTED;
    tp.P = M;
  }
  if (h_lsf) {
    RBX_REQUIRE(K > 0 && K == 2 * ext + 1, "rbx_psf_lsf_taps: LSF kernel length must be 2*extend_factor+1");
    double mx = 0.0;
    for (int m = 0; m < K; ++m) mx = std::fmax(mx, std::fabs((double)h_lsf[m]));
    int he = 0;
    for (int m = 0; m < K; ++m)
      if (std::fabs((double)h_lsf[m]) > 1e-14 * mx) he = std::max(he, std::abs(m - ext));
    int KE = 2 * he + 1;
    KE = KE <= 1 ? 1 : (KE <= 7 ? 7 : (KE <= 13 ? 13 : 25));
    if (2 * he + 1 > kMarchMaxK) return RBX_ERR_UNSUPPORTED;
    const int hw = (KE - 1) / 2;
    // out[w] = sum_m k[m] in[w + ext - m]; window u = 0..KE-1 <-> m = ext + hw - u
    for (int u = 0; u < KE; ++u) {
      const int m = ext + hw - u;
      tp.taps.kl[u] = (m >= 0 && m < K) ? h_lsf[m] : 0.f;
    }
    tp.KE = KE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  switch (tp.P) {
    case 1: return launch_march_ke<1>(tp.KE, d_in, d_out, ny, nx, W, tp.taps, s);
    case 3: return launch_march_ke<3>(tp.KE, d_in, d_out, ny, nx, W, tp.taps, s);
    case 5: return launch_march_ke<5>(tp.KE, d_in, d_out, ny, nx, W, tp.taps, s);
    default: return launch_march_ke<7>(tp.KE, d_in, d_out, ny, nx, W, tp.taps, s);
  }
}
