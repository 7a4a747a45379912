// rotate_galaxy (the stage right before the path: rubix/core/rotation.py:76-115 -> rubix/galaxy/alignment.py).
//   moment_of_inertia_tensor  alignment.py:67-125   particles within the half-mass radius (inclusive)
//   rotation_matrix_from_inertia_tensor  :128-146   eigh, eigenvectors ordered by ascending eigenvalue
//   apply_init_rotation / apply_rotation :149-229   pos @ R, then pos @ E (E = R_z R_y R_x from Euler angles)
//   rotate_galaxy             :233-265              the same two products for the velocities
//
// Three launches, no host synchronisation:
//   inertia_partial_kernel   one pass over coords + mass: six second moments per block (double accumulators,
//                            fixed block-tree order -> bit-reproducible)
//   inertia_finish_kernel    one block: fixed-order sum of the block partials, the reference's padding term
//                            (see below), cyclic Jacobi eigen-decomposition of the 3 x 3 tensor in double, R
//   rotate_apply_kernel      per particle: (p @ R) @ E for coords and velocity, float32, in the reference's order
//
// Reference quirk reproduced: alignment.py:103-106 selects the particles with
// jnp.where(mask, size=N)[0], which PADS the index list with 0 up to N entries, so particle 0 is added
// (N - n_inside) more times.  The same term is added here.
//
// Eigenvector signs: eigh's signs are backend-dependent in the reference itself (LAPACK on CPU, cuSOLVER on
// GPU).  Here every eigenvector is normalised so that its component of largest magnitude is positive; the
// matrix is returned so callers / tests can see which signs were used.
#include "common.cuh"

namespace rbx {

constexpr int kRotBlocks = 592;
constexpr int kRotThreads = 256;

// partial layout: [block][8] doubles: Sxx, Syy, Szz, Sxy, Sxz, Syz (mass weighted), n_inside, unused
__global__ void __launch_bounds__(kRotThreads)
inertia_partial_kernel(const float *__restrict__ coords, const float *__restrict__ mass, int64_t n, float radius,
                       double *__restrict__ partial) {
  double s[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const float x = coords[3 * q], y = coords[3 * q + 1], z = coords[3 * q + 2];
    // distances = sqrt(sum(pos**2)) <= radius, in float32 like the reference (alignment.py:98-101)
    const float d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    if (d <= radius) {
      const double m = mass[q];
      s[0] += m * x * x; s[1] += m * y * y; s[2] += m * z * z;
      s[3] += m * x * y; s[4] += m * x * z; s[5] += m * y * z;
      s[6] += 1.0;
    }
  }
  __shared__ double sh[kRotThreads / 32][7];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    double v = s[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    double v = 0.0;
    for (int w = 0; w < kRotThreads / 32; ++w) v += sh[w][threadIdx.x];
    partial[(size_t)blockIdx.x * 8 + threadIdx.x] = v;
  }
}

// cyclic Jacobi for a symmetric 3 x 3 matrix; V's columns are the eigenvectors
__device__ void jacobi3(double A[3][3], double V[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    const double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
    if (off <= 1e-300 || off <= 1e-17 * diag) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {   // A <- A J
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {   // A <- J^T A
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
}

// moments[12]: the six second moments, n_inside, n, and (x, y, z, m) of the galaxy's particle 0 -- every entry a plain
// sum over particle shards (only the shard holding particle 0 contributes the last four), so a sharded galaxy
// all-reduces this vector and every rank derives the same rotation
__global__ void inertia_moments_kernel(const double *__restrict__ partial, int nblocks, const float *__restrict__ coords,
                                       const float *__restrict__ mass, int64_t n, int first_shard,
                                       double *__restrict__ moments) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int b = 0; b < nblocks; ++b)
    for (int k = 0; k < 7; ++k) s[k] += partial[(size_t)b * 8 + k];
  for (int k = 0; k < 7; ++k) moments[k] = s[k];
  moments[7] = (double)n;
  const bool has0 = first_shard && n > 0;
  moments[8] = has0 ? coords[0] : 0.0; moments[9] = has0 ? coords[1] : 0.0; moments[10] = has0 ? coords[2] : 0.0;
  moments[11] = has0 ? mass[0] : 0.0;
}

__global__ void inertia_finish_kernel(const double *__restrict__ moments, float *__restrict__ R,
                                      double *__restrict__ tensor_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s[7];
  for (int k = 0; k < 7; ++k) s[k] = moments[k];
  const double n = moments[7];
  // jnp.where(mask, size=N) pads with index 0: particle 0 enters (N - n_inside) more times
  if (n > 0) {
    const double pad = n - s[6];
    const double x = moments[8], y = moments[9], z = moments[10], m = moments[11];
    s[0] += pad * m * x * x; s[1] += pad * m * y * y; s[2] += pad * m * z * z;
    s[3] += pad * m * x * y; s[4] += pad * m * x * z; s[5] += pad * m * y * z;
  }
  // I_ii = sum m (r^2 - x_i^2), I_ij = -sum m x_i x_j   (alignment.py:109-124)
  double A[3][3], V[3][3];
  A[0][0] = s[1] + s[2]; A[1][1] = s[0] + s[2]; A[2][2] = s[0] + s[1];
  A[0][1] = A[1][0] = -s[3]; A[0][2] = A[2][0] = -s[4]; A[1][2] = A[2][1] = -s[5];
  if (tensor_out)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) tensor_out[i * 3 + j] = A[i][j];
  jacobi3(A, V);
  // ascending eigenvalues (stable for ties, like argsort)
  int order[3] = {0, 1, 2};
  double ev[3] = {A[0][0], A[1][1], A[2][2]};
  for (int i = 1; i < 3; ++i)
    for (int j = i; j > 0 && ev[order[j]] < ev[order[j - 1]]; --j) { int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
  for (int c = 0; c < 3; ++c) {
    const int src = order[c];
    int big = 0;
    for (int r = 1; r < 3; ++r)
      if (fabs(V[r][src]) > fabs(V[big][src])) big = r;
    const double sign = V[big][src] < 0 ? -1.0 : 1.0;
    for (int r = 0; r < 3; ++r) R[r * 3 + c] = (float)(sign * V[r][src]);
  }
}

// out = (p @ R) @ E with float32 products in the reference's order (jnp.dot twice, alignment.py:161,228)
__device__ __forceinline__ void rot2(const float p[3], const float R[9], const float E[9], float out[3]) {
  float t[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) t[c] = fmaf(p[2], R[6 + c], fmaf(p[1], R[3 + c], __fmul_rn(p[0], R[c])));
#pragma unroll
  for (int c = 0; c < 3; ++c) out[c] = fmaf(t[2], E[6 + c], fmaf(t[1], E[3 + c], __fmul_rn(t[0], E[c])));
}

struct Mat3 { float m[9]; };

__global__ void rotate_apply_kernel(const float *__restrict__ coords, const float *__restrict__ vel, int64_t n,
                                    const float *__restrict__ Rdev, Mat3 E, float *__restrict__ coords_out,
                                    float *__restrict__ vel_out) {
  float R[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = Rdev[i];
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    float p[3] = {coords[3 * q], coords[3 * q + 1], coords[3 * q + 2]}, o[3];
    rot2(p, R, E.m, o);
    coords_out[3 * q] = o[0]; coords_out[3 * q + 1] = o[1]; coords_out[3 * q + 2] = o[2];
    if (vel) {
      float v[3] = {vel[3 * q], vel[3 * q + 1], vel[3 * q + 2]};
      rot2(v, R, E.m, o);
      vel_out[3 * q] = o[0]; vel_out[3 * q + 1] = o[1]; vel_out[3 * q + 2] = o[2];
    }
  }
}

}  // namespace rbx

using namespace rbx;

extern "C" size_t rbx_rotate_galaxy_workspace_bytes(void) { return sizeof(double) * (8 * kRotBlocks + 16) + 256; }

// Step 1 of rotate_galaxy for a galaxy whose particles are sharded over ranks: this shard's contribution to the
// inertia sums, d_moments[12] doubles (see inertia_moments_kernel).  first_shard != 0 on the rank that holds the
// galaxy's particle 0.  The host sums d_moments over the ranks (rbx_allreduce_f64) before rbx_rotate_apply.
extern "C" int rbx_rotate_moments(const float *d_coords, const float *d_mass, int64_t n, float halfmass_radius,
                                  int first_shard, double *d_moments, void *d_workspace, size_t workspace_bytes,
                                  void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RBX_REQUIRE(n >= 0 && d_moments && d_workspace, "rbx_rotate_moments: bad argument");
  RBX_REQUIRE(n == 0 || (d_coords && d_mass), "rbx_rotate_moments: null pointer");
  RBX_REQUIRE(workspace_bytes >= rbx_rotate_galaxy_workspace_bytes(), "rbx_rotate_moments: workspace too small");
  double *partial = reinterpret_cast<double *>(((uintptr_t)d_workspace + 255) & ~(uintptr_t)255);
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(kRotBlocks, (n + kRotThreads - 1) / kRotThreads));
  inertia_partial_kernel<<<blocks, kRotThreads, 0, stream>>>(d_coords, d_mass, n, halfmass_radius, partial);
  count_launch();
  RBX_LAUNCH_OK();
  inertia_moments_kernel<<<1, 32, 0, stream>>>(partial, blocks, d_coords, d_mass, n, first_shard, d_moments);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}

// Step 2: rotation from the (summed) moments, then (p @ R) @ E on this shard's particles.
extern "C" int rbx_rotate_apply(const float *d_coords, const float *d_velocity, int64_t n, const double *d_moments,
                                const float *h_euler, float *d_coords_out, float *d_velocity_out, float *d_rotation,
                                void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RBX_REQUIRE(n >= 0 && h_euler && d_rotation && d_moments, "rbx_rotate_apply: bad argument");
  RBX_REQUIRE(n == 0 || (d_coords && d_coords_out), "rbx_rotate_apply: null pointer");
  RBX_REQUIRE((d_velocity == nullptr) == (d_velocity_out == nullptr), "rbx_rotate_apply: velocity in/out must both be given or both NULL");
  inertia_finish_kernel<<<1, 32, 0, stream>>>(d_moments, d_rotation, nullptr);
  count_launch();
  RBX_LAUNCH_OK();
  if (n > 0) {
    Mat3 E;
    for (int i = 0; i < 9; ++i) E.m[i] = h_euler[i];
    const int ablocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
    rotate_apply_kernel<<<ablocks, 256, 0, stream>>>(d_coords, d_velocity, n, d_rotation, E, d_coords_out, d_velocity_out);
    count_launch();
    RBX_LAUNCH_OK();
  }
  return RBX_OK;
}

extern "C" int rbx_rotate_galaxy(const float *d_coords, const float *d_velocity, const float *d_mass, int64_t n,
                                 float halfmass_radius, const float *h_euler, float *d_coords_out,
                                 float *d_velocity_out, float *d_rotation, void *d_workspace, size_t workspace_bytes,
                                 void *stream_) {
  RBX_REQUIRE(n >= 0 && h_euler && d_rotation, "rbx_rotate_galaxy: bad argument");
  RBX_REQUIRE(n == 0 || (d_coords && d_mass && d_coords_out && d_workspace), "rbx_rotate_galaxy: null pointer");
  RBX_REQUIRE((d_velocity == nullptr) == (d_velocity_out == nullptr), "rbx_rotate_galaxy: velocity in/out must both be given or both NULL");
  RBX_REQUIRE(workspace_bytes >= rbx_rotate_galaxy_workspace_bytes(), "rbx_rotate_galaxy: workspace too small");
  double *moments = reinterpret_cast<double *>(((uintptr_t)d_workspace + 255) & ~(uintptr_t)255) + 8 * kRotBlocks;
  int rc = rbx_rotate_moments(d_coords, d_mass, n, halfmass_radius, 1, moments, d_workspace, workspace_bytes, stream_);
  if (rc != RBX_OK) return rc;
  return rbx_rotate_apply(d_coords, d_velocity, n, moments, h_euler, d_coords_out, d_velocity_out, d_rotation, stream_);
}
