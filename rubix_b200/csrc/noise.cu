// apply_noise, the stage right after the path: rubix/core/noise.py:15-78 ->
// rubix/telescope/noise/noise.py:8-115.
//   flux image   = sum over lambda                                   (noise.py:62)
//   median flux  = jnp.median(where(flux > 0, flux, nan)), nan -> 0   (noise.py:66-69): jnp.median propagates
//                  NaN, so a single spaxel without flux makes the median 0 and the noise vanish.  Reproduced.
//   S2N[y, x]    = where(flux > 0, (sqrt(median) / signal_to_noise) / sqrt(flux), 0)        (noise.py:72-78)
//   cube        += cube * N * S2N[y, x, None],  N = jax.random.normal(PRNGKey(0), cube.shape)   (noise.py:104-113,
//                  core/noise.py:73) or jax.random.uniform for "uniform"
//
// The random numbers restate JAX's counter-based generator: threefry2x32 (20 rounds; pinned to the Random123
// known-answer vectors in the tests), one block per element with the element's row-major index as the
// counter and bits = x0 ^ x1 (JAX's partitionable threefry, the default since jax 0.5), uniform from the top 23
// bits, normal = sqrt(2) * erfinv(u) with XLA's single-precision erfinv polynomial (Giles 2010).  jax is not
// pinned by the reference and not installable here: the integer stream is exact by construction, the float
// mapping is "parity unpinned" (SURVEY 8c).
#include "common.cuh"

namespace rbx {

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t &x0, uint32_t &x1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  const int R0[4] = {13, 15, 26, 6}, R1[4] = {17, 29, 16, 24};
  x0 += ks[0]; x1 += ks[1];
#pragma unroll
  for (int g = 0; g < 5; ++g) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      x0 += x1;
      x1 = rotl32(x1, (g & 1) ? R1[r] : R0[r]);
      x1 ^= x0;
    }
    x0 += ks[(g + 1) % 3];
    x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
  }
}

// XLA ErfInv for float32 (Giles, "Approximating the erfinv function", single precision)
__device__ __forceinline__ float erfinv_f32(float x) {
  float w = -log1pf(-x * x);
  float p;
  if (w < 5.0f) {
    w = w - 2.5f;
    p = 2.81022636e-08f;
    p = fmaf(p, w, 3.43273939e-07f);
    p = fmaf(p, w, -3.5233877e-06f);
    p = fmaf(p, w, -4.39150654e-06f);
    p = fmaf(p, w, 0.00021858087f);
    p = fmaf(p, w, -0.00125372503f);
    p = fmaf(p, w, -0.00417768164f);
    p = fmaf(p, w, 0.246640727f);
    p = fmaf(p, w, 1.50140941f);
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    p = fmaf(p, w, 0.000100950558f);
    p = fmaf(p, w, 0.00134934322f);
    p = fmaf(p, w, -0.00367342844f);
    p = fmaf(p, w, 0.00573950773f);
    p = fmaf(p, w, -0.0076224613f);
    p = fmaf(p, w, 0.00943887047f);
    p = fmaf(p, w, 1.00167406f);
    p = fmaf(p, w, 2.83297682f);
  }
  return fabsf(x) == 1.0f ? copysignf(INFINITY, x) : p * x;
}

__device__ __forceinline__ float noise_sample(uint32_t k0, uint32_t k1, uint64_t index, int uniform) {
  uint32_t x0 = (uint32_t)(index >> 32), x1 = (uint32_t)index;
  threefry2x32(k0, k1, x0, x1);
  const uint32_t bits = x0 ^ x1;
  const float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;   // [0, 1)
  if (uniform) return f;
  const float lo = -0.99999994f;   // nextafter(-1, 0)
  const float u = fmaxf(lo, fmaf(f, 1.0f - lo, lo));
  return 1.41421356f * erfinv_f32(u);
}

// flux[s] = sum_w cube[s, w], fixed summation order (thread-strided partials, shuffle tree, warp order)
__global__ void __launch_bounds__(128)
flux_image_kernel(const float *__restrict__ cube, int W, float *__restrict__ flux) {
  const float *row = cube + (size_t)blockIdx.x * W;
  float acc = 0.f;
  for (int w = threadIdx.x; w < W; w += 128) acc += row[w];
  __shared__ float sh[4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) flux[blockIdx.x] = (sh[0] + sh[1]) + (sh[2] + sh[3]);
}

// one block: median of the flux image by a bitonic sort in shared memory, then the S2N map
__global__ void __launch_bounds__(1024)
noise_s2n_kernel(const float *__restrict__ flux, int n, int npow2, float signal_to_noise, float *__restrict__ s2n) {
  extern __shared__ float s_v[];
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
    const float v = i < n ? flux[i] : INFINITY;
    if (i < n && !(v > 0.f)) s_bad = 1;   // where(mask, flux, nan) holds a NaN: jnp.median returns NaN -> 0
    s_v[i] = v;
  }
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const float a = s_v[i], b = s_v[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { s_v[i] = b; s_v[l] = a; }
        }
      }
      __syncthreads();
    }
  float median = 0.f;
  if (!s_bad && n > 0) {
    const float lo = s_v[(n - 1) / 2], hi = s_v[n / 2];
    median = lo + (hi - lo) * 0.5f;   // jnp.quantile(0.5), linear interpolation
  }
  const float factor = sqrtf(median) / signal_to_noise;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float f = flux[i];
    s2n[i] = f > 0.f ? factor / sqrtf(f) : 0.f;
  }
}

__global__ void apply_noise_kernel(const float *__restrict__ in, float *__restrict__ out, size_t total, int W,
                                   const float *__restrict__ s2n, uint32_t k0, uint32_t k1, int uniform) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const float c = in[i];
    const float nz = noise_sample(k0, k1, (uint64_t)i, uniform) * s2n[i / W];
    out[i] = c + c * nz;    // datacube += cube * noise   (noise.py:113, core/noise.py:76)
  }
}

__global__ void noise_samples_kernel(float *__restrict__ out, size_t total, uint32_t k0, uint32_t k1, int uniform,
                                     uint32_t *__restrict__ bits_out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    out[i] = noise_sample(k0, k1, (uint64_t)i, uniform);
    if (bits_out) {
      uint32_t x0 = (uint32_t)((uint64_t)i >> 32), x1 = (uint32_t)i;
      threefry2x32(k0, k1, x0, x1);
      bits_out[i] = x0 ^ x1;
    }
  }
}

}  // namespace rbx

using namespace rbx;

extern "C" size_t rbx_apply_noise_workspace_bytes(int ny, int nx) { return sizeof(float) * 2 * (size_t)ny * nx + 512; }

extern "C" int rbx_apply_noise(const float *d_in, float *d_out, int ny, int nx, int W, float signal_to_noise,
                               int distribution, uint32_t key0, uint32_t key1, void *d_workspace, size_t workspace_bytes,
                               void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RBX_REQUIRE(d_in && d_out && d_workspace, "rbx_apply_noise: null pointer");
  RBX_REQUIRE(ny > 0 && nx > 0 && W > 0, "rbx_apply_noise: bad shape");
  RBX_REQUIRE(distribution == 0 || distribution == 1, "rbx_apply_noise: distribution must be 0 (normal) or 1 (uniform)");
  RBX_REQUIRE(workspace_bytes >= rbx_apply_noise_workspace_bytes(ny, nx), "rbx_apply_noise: workspace too small");
  const int n = ny * nx;
  int npow2 = 1;
  while (npow2 < n) npow2 <<= 1;
  if (npow2 > 32768) {
    set_error("rbx_apply_noise: more than 32768 spaxels (the median runs in one block's shared memory)");
    return RBX_ERR_UNSUPPORTED;
  }
  float *flux = reinterpret_cast<float *>(((uintptr_t)d_workspace + 255) & ~(uintptr_t)255);
  float *s2n = flux + n;
  flux_image_kernel<<<n, 128, 0, stream>>>(d_in, W, flux);
  count_launch();
  RBX_LAUNCH_OK();
  const size_t smem = sizeof(float) * (size_t)npow2;
  if (smem > 48 * 1024)
    RBX_CUDA_OK(cudaFuncSetAttribute(noise_s2n_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  noise_s2n_kernel<<<1, 1024, smem, stream>>>(flux, n, npow2, signal_to_noise, s2n);
  count_launch();
  RBX_LAUNCH_OK();
  const size_t total = (size_t)n * W;
  apply_noise_kernel<<<148 * 8, 256, 0, stream>>>(d_in, d_out, total, W, s2n, key0, key1, distribution);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}

// the raw sample stream (tests): d_out[i] = normal / uniform sample i, d_bits[i] (may be NULL) = its 32 random bits
extern "C" int rbx_noise_samples(float *d_out, uint32_t *d_bits, int64_t n, int distribution, uint32_t key0, uint32_t key1,
                                 void *stream_) {
  RBX_REQUIRE(d_out && n >= 0, "rbx_noise_samples: bad argument");
  if (n == 0) return RBX_OK;
  noise_samples_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream_>>>(d_out, (size_t)n, key0, key1, distribution, d_bits);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}
