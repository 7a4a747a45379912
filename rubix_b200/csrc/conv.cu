// a6 / a7: spatial PSF and spectral LSF convolutions of the (ny, nx, W) cube (lambda fastest).
//   rbx_convolve_psf  per-slice zero-padded true 2-D convolution (jax.scipy.signal.convolve2d "same")
//   rbx_convolve_lsf  zero-padded 1-D convolution along lambda  (convolve "full" + slice == "same")
//   rbx_psf_lsf       both in one pass over the cube: a shared-memory tile with spatial and spectral
//                     halos; the PSF result never goes to HBM.
// HBM-bound (read once + write once = 8 bytes / voxel); lambda is the coalesced axis everywhere.
#include "common.cuh"

namespace rbx {

constexpr int kMaxTaps = 1024;  // PSF M*N and LSF K held in shared memory

// ---- separate kernels -------------------------------------------------------------------------
// out[y,x,w] = sum_{m,n} K[m,n] in[y-m+cm, x-n+cn, w],  cm=(M-1)/2, cn=(N-1)/2
// (rubix/telescope/psf/psf.py:9-11,56-57; jax _convolve_nd "same" padding)
__global__ void psf_kernel(const float *__restrict__ in, float *__restrict__ out, int ny, int nx, int W,
                           const float *__restrict__ K, int M, int N) {
  __shared__ float sk[kMaxTaps];
  for (int i = threadIdx.x; i < M * N; i += blockDim.x) sk[i] = K[i];
  __syncthreads();
  const int cm = (M - 1) / 2, cn = (N - 1) / 2;
  const int pos = blockIdx.y;  // y*nx + x
  const int y = pos / nx, x = pos % nx;
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < W; w += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int m = 0; m < M; ++m) {
      int yy = y - m + cm;
      if (yy < 0 || yy >= ny) continue;
      for (int n = 0; n < N; ++n) {
        int xx = x - n + cn;
        if (xx < 0 || xx >= nx) continue;
        acc = fmaf(sk[m * N + n], in[((size_t)yy * nx + xx) * W + w], acc);
      }
    }
    out[(size_t)pos * W + w] = acc;
  }
}

// out[r,w] = sum_k k[m] in[r, w+ext-m]   (rubix/telescope/lsf/lsf.py:59-65)
__global__ void lsf_kernel(const float *__restrict__ in, float *__restrict__ out, int64_t rows, int W,
                           const float *__restrict__ k, int K, int ext) {
  __shared__ float sk[kMaxTaps];
  for (int i = threadIdx.x; i < K; i += blockDim.x) sk[i] = k[i];
  __syncthreads();
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const float *s = in + r * W;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < W; w += gridDim.x * blockDim.x) {
      float acc = 0.f;
      for (int m = 0; m < K; ++m) {
        int q = w + ext - m;
        if (q >= 0 && q < W) acc = fmaf(sk[m], s[q], acc);
      }
      out[r * W + w] = acc;
    }
  }
}

// ---- fused PSF + LSF --------------------------------------------------------------------------
// Block = TY x TX spaxels x TL output channels.  Shared memory:
//   s_in  [(TY+M-1)*(TX+N-1)][TLH]   input tile incl. spatial halo and spectral halo (TLH = TL+K-1)
//   s_mid [TY*TX][TLH]               PSF-convolved tile (still with the spectral halo)
// Phase 1: coalesced zero-padded load.  Phase 2: PSF, one thread per (row y, channel) computing TX
// outputs from registers.  Phase 3: LSF, lanes along lambda, coalesced store.
template <int TY, int TX>
__global__ void psf_lsf_kernel(const float *__restrict__ in, float *__restrict__ out, int ny, int nx, int W,
                               const float *__restrict__ Kp, int M, int N, const float *__restrict__ kl,
                               int K, int ext, int TL) {
  extern __shared__ float smf[];
  const int TLH = TL + K - 1;
  const int IY = TY + M - 1, IX = TX + N - 1;
  float *s_in = smf;
  float *s_mid = s_in + (size_t)IY * IX * TLH;
  float *s_kp = s_mid + (size_t)TY * TX * TLH;
  float *s_kl = s_kp + M * N;
  for (int i = threadIdx.x; i < M * N; i += blockDim.x) s_kp[i] = Kp[i];
  for (int i = threadIdx.x; i < K; i += blockDim.x) s_kl[i] = kl[i];
  const int cm = (M - 1) / 2, cn = (N - 1) / 2;
  const int tiles_x = (nx + TX - 1) / TX;
  const int ty0 = (blockIdx.y / tiles_x) * TY, tx0 = (blockIdx.y % tiles_x) * TX;
  const int w0 = blockIdx.x * TL;  // first output channel of this block
  // input channel q maps to tile column q - (w0 + ext - (K-1)); LSF output w needs inputs w+ext-m
  const int q0 = w0 + ext - (K - 1);
  // spatial input origin: output (y,x) needs in[y-m+cm][x-n+cn]  => rows from ty0+cm-(M-1)
  const int iy0 = ty0 + cm - (M - 1), ix0 = tx0 + cn - (N - 1);

  // phase 1
  for (int r = 0; r < IY * IX; ++r) {
    int yy = iy0 + r / IX, xx = ix0 + r % IX;
    bool ok = yy >= 0 && yy < ny && xx >= 0 && xx < nx;
    const float *src = in + ((size_t)yy * nx + xx) * W;
    for (int c = threadIdx.x; c < TLH; c += blockDim.x) {
      int q = q0 + c;
      s_in[(size_t)r * TLH + c] = (ok && q >= 0 && q < W) ? src[q] : 0.f;
    }
  }
  __syncthreads();

  // phase 2: mid[y][x][c] = sum_{m,n} Kp[m][n] * s_in[(y + (M-1) - m)][(x + (N-1) - n)][c]
  for (int item = threadIdx.x; item < TY * TLH; item += blockDim.x) {
    const int y = item / TLH, c = item % TLH;
    float acc[TX];
#pragma unroll
    for (int x = 0; x < TX; ++x) acc[x] = 0.f;
    for (int m = 0; m < M; ++m) {
      const float *rowp = s_in + ((size_t)(y + (M - 1) - m) * IX) * TLH + c;
      for (int n = 0; n < N; ++n) {
        const float kv = s_kp[m * N + n];
#pragma unroll
        for (int x = 0; x < TX; ++x) acc[x] = fmaf(kv, rowp[(size_t)(x + (N - 1) - n) * TLH], acc[x]);
      }
    }
#pragma unroll
    for (int x = 0; x < TX; ++x) s_mid[((size_t)(y * TX + x)) * TLH + c] = acc[x];
  }
  __syncthreads();

  // phase 3: out[y][x][w0 + j] = sum_m kl[m] * mid[..][j + (K-1) - m]
  for (int item = threadIdx.x; item < TY * TX * TL; item += blockDim.x) {
    const int pos = item / TL, j = item % TL;
    const int y = ty0 + pos / TX, x = tx0 + pos % TX, w = w0 + j;
    if (y >= ny || x >= nx || w >= W) continue;
    const float *mp = s_mid + (size_t)pos * TLH + j + (K - 1);
    float acc = 0.f;
    for (int m = 0; m < K; ++m) acc = fmaf(s_kl[m], mp[-m], acc);
    out[((size_t)y * nx + x) * W + w] = acc;
  }
}

// ---- fused PSF + LSF, register tiled ---------------------------------------------------------------
// Block = TY x TX spaxels x TL output channels; one thread per channel of the tile *including* the
// spectral halo (TL + K - 1 channels).  PSF: every input voxel of the (TY+P-1) x (TX+P-1) spatial
// neighbourhood is loaded once (coalesced along lambda) into a register and used for up to P*P
// FMAs into TY*TX register accumulators; the taps sit in registers too.  The PSF result goes to
// shared memory only, the LSF then runs along lambda with 4 outputs per thread (7 LDS.128 for 100
// FMAs) and stores coalesced.  HBM traffic is the algorithmic 8 bytes / voxel as long as the
// (TY+P-1) spaxel rows being worked on stay in L2 (blocks walk x, then lambda, then y).
template <int TY, int TX, int P, int KMAX>
__global__ void __launch_bounds__(160)
psf_lsf_reg_kernel(const float *__restrict__ in, float *__restrict__ out, int ny, int nx, int W,
                   const float *__restrict__ Kp, const float *__restrict__ kl, int K, int ext) {
  constexpr int TL = 128;                          // output channels per block
  extern __shared__ __align__(16) float s_mid[];   // [TY*TX][pitch]
  const int TLH = TL + K - 1;
  const int pitch = (TLH + 3 + 4) & ~3;            // room for the 4-wide LSF reads of the last group
  const int tiles_x = (nx + TX - 1) / TX;
  const int x0 = (blockIdx.x % tiles_x) * TX;
  const int w0 = (blockIdx.x / tiles_x) * TL;
  const int y0 = blockIdx.y * TY;
  constexpr int C = (P - 1) / 2;                   // jax "same": out[y] = sum_m K[m] in[y - m + C]
  constexpr int H = P - 1 - C;                     // rows / columns of halo before the tile
  float kp[P * P];                                 // PSF taps in registers (static indices below)
#pragma unroll
  for (int i = 0; i < P * P; ++i) kp[i] = Kp ? __ldg(Kp + i) : 1.f;   // Kp == NULL: identity (P == 1)
  const bool interior = y0 - H >= 0 && y0 + TY + C <= ny && x0 - H >= 0 && x0 + TX + C <= nx;
  const size_t rowstride = (size_t)nx * W;

  // ---- PSF into shared memory -------------------------------------------------------------------
  for (int c = threadIdx.x; c < TLH; c += blockDim.x) {
    const int q = w0 + ext - (K - 1) + c;          // input channel of tile column c
    float acc[TY][TX];
#pragma unroll
    for (int a = 0; a < TY; ++a)
#pragma unroll
      for (int b = 0; b < TX; ++b) acc[a][b] = 0.f;
    if (q >= 0 && q < W) {
      const float *base = in + ((ptrdiff_t)(y0 - H) * nx + (x0 - H)) * W + q;   // voxel (iy = 0, ix = 0)
#pragma unroll
      for (int iy = 0; iy < TY + P - 1; ++iy) {
        const int yy = y0 + iy - H;
        float v[TX + P - 1];
        if (interior) {
#pragma unroll
          for (int ix = 0; ix < TX + P - 1; ++ix) v[ix] = __ldg(base + iy * rowstride + (size_t)ix * W);
        } else {
          const bool rowok = yy >= 0 && yy < ny;
#pragma unroll
          for (int ix = 0; ix < TX + P - 1; ++ix) {
            const int xx = x0 + ix - H;
            v[ix] = (rowok && xx >= 0 && xx < nx) ? __ldg(base + iy * rowstride + (ptrdiff_t)ix * W) : 0.f;
          }
        }
#pragma unroll
        for (int oy = 0; oy < TY; ++oy) {
          const int m = oy + P - 1 - iy;           // tap row feeding output row oy (compile time)
          if (m < 0 || m >= P) continue;
#pragma unroll
          for (int n = 0; n < P; ++n)
#pragma unroll
            for (int ox = 0; ox < TX; ++ox) acc[oy][ox] = fmaf(kp[m * P + n], v[ox + P - 1 - n], acc[oy][ox]);
        }
        // keep the loads of the next input row behind this row's FMAs: hoisting them all costs ~250 registers
        asm volatile("" ::: "memory");
      }
    }
#pragma unroll
    for (int a = 0; a < TY; ++a)
#pragma unroll
      for (int b = 0; b < TX; ++b) s_mid[(a * TX + b) * pitch + c] = acc[a][b];
  }
  // zero the pad columns read by the last 4-wide group
  for (int i = threadIdx.x; i < TY * TX * (pitch - TLH); i += blockDim.x) {
    const int pix = i / (pitch - TLH), c = TLH + i % (pitch - TLH);
    s_mid[pix * pitch + c] = 0.f;
  }
  __syncthreads();

  // ---- LSF: out[w0 + j] = sum_m kl[m] mid[j + K - 1 - m] -----------------------------------------
  float krev[KMAX];  // taps reversed and zero padded: krev[u] = kl[K - 1 - u]
#pragma unroll
  for (int u = 0; u < KMAX; ++u) krev[u] = u < K ? (kl ? __ldg(kl + (K - 1 - u)) : 1.f) : 0.f;  // kl == NULL: identity (K == 1)
  constexpr int G = TL / 4;                        // 4-channel groups per spaxel
  for (int item = threadIdx.x; item < TY * TX * G; item += blockDim.x) {
    const int pix = item / G, j0 = (item % G) * 4;
    const int y = y0 + pix / TX, x = x0 + pix % TX;
    if (y >= ny || x >= nx) continue;
    const float4 *mp = reinterpret_cast<const float4 *>(s_mid + pix * pitch + j0);
    float win[KMAX + 3 + 1];
#pragma unroll
    for (int i = 0; i < (KMAX + 3 + 3) / 4; ++i) {
      const float4 t4 = (4 * i < K + 3) ? mp[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      win[4 * i] = t4.x; win[4 * i + 1] = t4.y; win[4 * i + 2] = t4.z; win[4 * i + 3] = t4.w;
    }
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    // out[j0 + r] = sum_m kl[m] * mid[j0 + r + K - 1 - m];  with u = K - 1 - m: kl[K-1-u] * win[r + u]
#pragma unroll
    for (int u = 0; u < KMAX; ++u)
#pragma unroll
      for (int r = 0; r < 4; ++r) o[r] = fmaf(krev[u], win[r + u], o[r]);
    float *dst = out + ((size_t)y * nx + x) * W + w0 + j0;
    if (w0 + j0 + 3 < W) {
      dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2]; dst[3] = o[3];
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (w0 + j0 + r < W) dst[r] = o[r];
    }
  }
}

// rubix/telescope/psf/kernels.py:26-31 in float32; single block
__global__ void gaussian_psf_kernel(int m, int n, float sigma, float *__restrict__ out) {
  __shared__ float ssum;
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < n; ++j) {
        float x = -((m - 1) / 2.f) + i, y = -((n - 1) / 2.f) + j;
        s += expf(-(x * x + y * y) / (2.f * sigma * sigma));
      }
    ssum = s;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < m * n; q += blockDim.x) {
    int i = q / n, j = q % n;
    float x = -((m - 1) / 2.f) + i, y = -((n - 1) / 2.f) + j;
    out[q] = expf(-(x * x + y * y) / (2.f * sigma * sigma)) / ssum;
  }
}

// rubix/telescope/lsf/lsf.py:12-26: x = arange(-f*wr, f*wr + wr, wr); exp(-0.5 x^2 / sigma^2) / sum
__global__ void gaussian_lsf_kernel(float sigma, float wr, int factor, int K, float *__restrict__ out) {
  __shared__ float ssum;
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < K; ++i) {
      float x = -factor * wr + i * wr;
      s += expf(-0.5f * (x * x) / (sigma * sigma));
    }
    ssum = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    float x = -factor * wr + i * wr;
    out[i] = expf(-0.5f * (x * x) / (sigma * sigma)) / ssum;
  }
}

}  // namespace rbx

using namespace rbx;

// Launch the register-tiled kernel when it covers the configuration (square PSF of 1/3/5/7 taps per side,
// LSF of <= 25 taps, or identity for either); RBX_ERR_UNSUPPORTED otherwise (no error text set).
static int psf_lsf_reg_dispatch(const float *d_in, float *d_out, int ny, int nx, int W, const float *d_psf, int M,
                                int N, const float *d_lsf, int K, int ext, cudaStream_t stream) {
  if (!(M == N && (M == 1 || M == 3 || M == 5 || M == 7) && (K <= 25))) return RBX_ERR_UNSUPPORTED;
  const int TLr = 128;
  auto run = [&](auto kernel, int ty, int tx) -> int {
    const int pitch = (TLr + K - 1 + 3 + 4) & ~3;
    const size_t smem = sizeof(float) * (size_t)ty * tx * pitch;
    if (smem > 48 * 1024) RBX_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(((nx + tx - 1) / tx) * ((W + TLr - 1) / TLr), (ny + ty - 1) / ty);
    kernel<<<grid, 160, smem, stream>>>(d_in, d_out, ny, nx, W, d_psf, d_lsf, K, ext);
    count_launch();
    RBX_LAUNCH_OK();
    return RBX_OK;
  };
  const bool five = (nx % 5 == 0) && (ny % 5 == 0);
  if (K == 1) {  // PSF only
    if (M == 3) return run(psf_lsf_reg_kernel<4, 8, 3, 1>, 4, 8);
    if (M == 5) return five ? run(psf_lsf_reg_kernel<5, 5, 5, 1>, 5, 5) : run(psf_lsf_reg_kernel<4, 8, 5, 1>, 4, 8);
    if (M == 7) return run(psf_lsf_reg_kernel<4, 8, 7, 1>, 4, 8);
    return RBX_ERR_UNSUPPORTED;
  }
  if (M == 1) return five ? run(psf_lsf_reg_kernel<5, 5, 1, 25>, 5, 5) : run(psf_lsf_reg_kernel<4, 8, 1, 25>, 4, 8);
  if (M == 3) return run(psf_lsf_reg_kernel<4, 8, 3, 25>, 4, 8);
  if (M == 5) {
    if (five && nx % 10 == 0 && nx >= 50) return run(psf_lsf_reg_kernel<5, 10, 5, 25>, 5, 10);
    return five ? run(psf_lsf_reg_kernel<5, 5, 5, 25>, 5, 5) : run(psf_lsf_reg_kernel<4, 8, 5, 25>, 4, 8);
  }
  return run(psf_lsf_reg_kernel<4, 8, 7, 25>, 4, 8);
}

extern "C" int rbx_convolve_psf(const float *d_in, float *d_out, int ny, int nx, int W, const float *d_kernel,
                                int M, int N, void *stream) {
  RBX_REQUIRE(d_in && d_out && d_kernel && d_in != d_out, "rbx_convolve_psf: bad pointers (no aliasing)");
  RBX_REQUIRE(ny > 0 && nx > 0 && W > 0 && M > 0 && N > 0 && M * N <= kMaxTaps, "rbx_convolve_psf: bad shape");
  // jax.scipy.signal.convolve2d: "One input must be smaller than the other in every dimension."
  RBX_REQUIRE((M <= ny && N <= nx) || (M >= ny && N >= nx),
              "One input must be smaller than the other in every dimension.");
  {  // register-tiled kernel with an identity LSF
    int rc = psf_lsf_reg_dispatch(d_in, d_out, ny, nx, W, d_kernel, M, N, nullptr, 1, 0, (cudaStream_t)stream);
    if (rc != RBX_ERR_UNSUPPORTED) return rc;
  }
  dim3 grid((W + 255) / 256, ny * nx);
  psf_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_in, d_out, ny, nx, W, d_kernel, M, N);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}

extern "C" int rbx_convolve_lsf(const float *d_in, float *d_out, int64_t rows, int W, const float *d_kernel, int K,
                                int ext, void *stream) {
  RBX_REQUIRE(d_in && d_out && d_kernel && d_in != d_out, "rbx_convolve_lsf: bad pointers (no aliasing)");
  RBX_REQUIRE(rows > 0 && W > 0 && K > 0 && K <= kMaxTaps, "rbx_convolve_lsf: bad shape");
  RBX_REQUIRE(K == 2 * ext + 1, "rbx_convolve_lsf: kernel length must be 2*extend_factor+1");
  if (rows < (1 << 30)) {  // register-tiled kernel with an identity PSF
    int fx = 1;  // any factorisation rows = fy * fx is a valid "image" for an identity PSF: pick one that fills tiles
    for (int d : {40, 25, 20, 10, 8, 5, 4, 2})
      if (rows % d == 0) { fx = d; break; }
    int rc = psf_lsf_reg_dispatch(d_in, d_out, (int)(rows / fx), fx, W, nullptr, 1, 1, d_kernel, K, ext, (cudaStream_t)stream);
    if (rc != RBX_ERR_UNSUPPORTED) return rc;
  }
  dim3 grid((W + 255) / 256, (unsigned)(rows < 65535 ? rows : 65535));
  lsf_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_in, d_out, rows, W, d_kernel, K, ext);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}

extern "C" int rbx_psf_lsf(const float *d_in, float *d_out, int ny, int nx, int W, const float *d_psf, int M,
                           int N, const float *d_lsf, int K, int ext, void *stream) {
  RBX_REQUIRE(d_in && d_out && d_psf && d_lsf && d_in != d_out, "rbx_psf_lsf: bad pointers (no aliasing)");
  RBX_REQUIRE(ny > 0 && nx > 0 && W > 0 && M > 0 && N > 0 && K > 0, "rbx_psf_lsf: bad shape");
  RBX_REQUIRE(K == 2 * ext + 1, "rbx_psf_lsf: LSF kernel length must be 2*extend_factor+1");
  RBX_REQUIRE((M <= ny && N <= nx) || (M >= ny && N >= nx),
              "One input must be smaller than the other in every dimension.");
  {
    int rc = psf_lsf_reg_dispatch(d_in, d_out, ny, nx, W, d_psf, M, N, d_lsf, K, ext, (cudaStream_t)stream);
    if (rc != RBX_ERR_UNSUPPORTED) return rc;
  }
  constexpr int TY = 5, TX = 5;
  int TL = 128;
  auto smem_for = [&](int tl) {
    size_t tlh = tl + K - 1;
    return sizeof(float) * ((size_t)(TY + M - 1) * (TX + N - 1) * tlh + (size_t)TY * TX * tlh + M * N + K);
  };
  while (TL > 32 && smem_for(TL) > 200 * 1024) TL /= 2;
  if (smem_for(TL) > 200 * 1024 || M * N > kMaxTaps || K > kMaxTaps) {
    set_error("rbx_psf_lsf: kernels too large for the fused tile; call rbx_convolve_psf + rbx_convolve_lsf");
    return RBX_ERR_UNSUPPORTED;
  }
  size_t smem = smem_for(TL);
  RBX_CUDA_OK(cudaFuncSetAttribute(psf_lsf_kernel<TY, TX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((W + TL - 1) / TL, ((ny + TY - 1) / TY) * ((nx + TX - 1) / TX));
  psf_lsf_kernel<TY, TX><<<grid, 256, smem, (cudaStream_t)stream>>>(d_in, d_out, ny, nx, W, d_psf, M, N, d_lsf, K, ext, TL);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}

extern "C" int rbx_gaussian_psf_kernel(int m, int n, float sigma, float *d_kernel, void *stream) {
  RBX_REQUIRE(d_kernel && m > 0 && n > 0 && m * n <= kMaxTaps, "rbx_gaussian_psf_kernel: bad argument");
  gaussian_psf_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(m, n, sigma, d_kernel);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}

extern "C" int rbx_gaussian_lsf_kernel(float sigma, float wave_res, int factor, float *d_kernel, void *stream) {
  RBX_REQUIRE(d_kernel && factor >= 0 && 2 * factor + 1 <= kMaxTaps, "rbx_gaussian_lsf_kernel: bad argument");
  gaussian_lsf_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(sigma, wave_res, factor, 2 * factor + 1, d_kernel);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}
