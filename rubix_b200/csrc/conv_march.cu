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
#include <cstdlib>

#include "common.cuh"

namespace rbx {

#ifndef RBX_MARCH_MINB
#define RBX_MARCH_MINB 4
#endif
constexpr int kMarchMaxP = 7;
constexpr int kMarchMaxK = 25;

struct MarchTaps {
  float kx[kMarchMaxP];   // PSF factor along x (columns of the reference's kernel)
  float ky[kMarchMaxP];   // PSF factor along y (rows)
  float kl[kMarchMaxK];   // effective LSF window, reversed: out[w] = sum_u kl[u] * mid[w - he + u]
  float2 ky2[kMarchMaxP]; // (ky, ky) and (kl, kl): operands of the packed FFMA2 (two columns per instruction)
  float2 kl2[kMarchMaxK];
};

// d = a * b + c on two packed float32 lanes (sm_100 FFMA2: one issue slot for two FMAs; each lane is an
// ordinary round-to-nearest fma, so results are bit-identical to the scalar form)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

// base + K * stride_bytes with a 32-bit product (the column stride W is a runtime value; the compiler's own
// 64-bit element indexing costs four integer instructions per access, this form two).
template <int K>
__device__ __forceinline__ float *col_at(const float *base, unsigned stride_bytes) {
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(stride_bytes), "n"(K), "l"(base));
  return reinterpret_cast<float *>(r);
}

template <int I, int N>
struct Cols {
  static __device__ __forceinline__ void load(const float *base, unsigned W4, float (&v)[N]) {
    v[I] = __ldg(col_at<I>(base, W4));
    if constexpr (I + 1 < N) Cols<I + 1, N>::load(base, W4, v);
  }
  // columns whose bit is clear in `mask` are outside the cube: zero, and never dereferenced
  static __device__ __forceinline__ void load_masked(const float *base, unsigned W4, unsigned mask, float (&v)[N]) {
    v[I] = ((mask >> I) & 1u) ? __ldg(col_at<I>(base, W4)) : 0.f;
    if constexpr (I + 1 < N) Cols<I + 1, N>::load_masked(base, W4, mask, v);
  }
  static __device__ __forceinline__ void store(float *base, unsigned W4, const float (&o)[N]) {
    *col_at<I>(base, W4) = o[I];
    if constexpr (I + 1 < N) Cols<I + 1, N>::store(base, W4, o);
  }
  static __device__ __forceinline__ void store_n(float *base, unsigned W4, const float (&o)[N], int nvalid) {
    if (I < nvalid) *col_at<I>(base, W4) = o[I];
    if constexpr (I + 1 < N) Cols<I + 1, N>::store_n(base, W4, o, nvalid);
  }
};

// EDGE = false: every column, channel and store of the block is inside the cube (no predicates at all).
template <int P, int KE, int TX, int NT, bool EDGE>
__device__ __forceinline__ void march_body(const float *__restrict__ in, float *__restrict__ out, int ny, int nx, int W,
                                           int pin, int pout, int x0, int w0, int ya, int yb, const MarchTaps &tp,
                                           float (&s_mid)[4][TX][NT]) {
  constexpr int TL = NT - (KE - 1);      // output channels per block
  constexpr int HE = (KE - 1) / 2;       // spectral halo on each side
  constexpr int C = (P - 1) / 2;         // jax "same": out[y] = sum_m K[m] in[y - m + C]
  constexpr int H = P - 1 - C;           // rows / columns before the tile
  constexpr int IX = TX + P - 1;
  const int c = threadIdx.x;
  const int q = w0 - HE + c;             // input channel of this thread
  const bool qok = q >= 0 && q < W;
  const int nsteps = (yb - ya) + P - 1;
  const bool lsf_thread = c < TL && w0 + c < W;
  const int nvalid = lsf_thread ? min(TX, nx - x0) : 0;   // output columns this thread stores
  const unsigned W4 = 4u * (unsigned)pin, W4o = 4u * (unsigned)pout;   // column strides in bytes (in, out)
  const size_t rowstride = (size_t)nx * pin, rowstride_o = (size_t)nx * pout;
  unsigned colmask = 0;
  if (EDGE) {
#pragma unroll
    for (int ix = 0; ix < IX; ++ix)
      if (qok && x0 - H + ix >= 0 && x0 - H + ix < nx) colmask |= 1u << ix;
  }
  // input row `yy` of this thread's channel: IX columns starting at x0 - H (only called for rows inside the cube)
  const float *in0 = in + ((ptrdiff_t)(x0 - H) * pin + q);
  auto load_row = [&](int yy, float (&v)[IX]) {
    const float *rowp = in0 + (ptrdiff_t)yy * (ptrdiff_t)rowstride;
    if (EDGE) Cols<0, IX>::load_masked(rowp, W4, colmask, v);
    else Cols<0, IX>::load(rowp, W4, v);
  };

  float A[P][TX];
#pragma unroll
  for (int a = 0; a < P; ++a)
#pragma unroll
    for (int b = 0; b < TX; ++b) A[a][b] = 0.f;

  float vn[IX];                          // next input row, in flight while the current one is folded in
#pragma unroll
  for (int ix = 0; ix < IX; ++ix) vn[ix] = 0.f;
  if (ya - H >= 0) load_row(ya - H, vn);
  int buf = 0;
  for (int i0 = 0; i0 < nsteps; i0 += P) {
#pragma unroll
    for (int u = 0; u < P; ++u) {
      const int i = i0 + u;
      if (i >= nsteps) break;            // block-uniform
      const int yy = ya - H + i;         // input row of this step
      const bool row_in = yy >= 0 && yy < ny;   // rows outside the cube are zero padding: nothing to add
      if (row_in) {
        float v[IX];
#pragma unroll
        for (int ix = 0; ix < IX; ++ix) v[ix] = vn[ix];
        if (i + 1 < nsteps && yy + 1 < ny) load_row(yy + 1, vn);
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
        for (int ox = 0; ox < TX; ++ox) A[(u + P - 1) % P][ox] = tp.ky[P - 1] * h[ox];
      } else {
        if (i + 1 < nsteps && yy + 1 >= 0 && yy + 1 < ny) load_row(yy + 1, vn);
#pragma unroll
        for (int ox = 0; ox < TX; ++ox) A[(u + P - 1) % P][ox] = 0.f;
      }
      if (P == 1 || i >= P - 1) {
        const int oy = yy - C;           // finished output row: slot u got its last tap (m = 0) above
#pragma unroll
        for (int ox = 0; ox < TX; ++ox) s_mid[buf][ox][c] = A[u % P][ox];
        __syncthreads();
        if (!EDGE || nvalid > 0) {
          float o[TX];
#pragma unroll
          for (int ox = 0; ox < TX; ++ox) {
            float acc = tp.kl[0] * s_mid[buf][ox][c];
#pragma unroll
            for (int t = 1; t < KE; ++t) acc = fmaf(tp.kl[t], s_mid[buf][ox][c + t], acc);
            o[ox] = acc;
          }
          float *dst = out + (size_t)oy * rowstride_o + (size_t)x0 * pout + (w0 + c);
          if (EDGE) Cols<0, TX>::store_n(dst, W4o, o, nvalid);
          else if (c < TL) Cols<0, TX>::store(dst, W4o, o);
        }
        buf ^= 1;
      }
    }
  }
}

// ---- bulk-copy (TMA) staged variant -----------------------------------------------------------------
// The input rows are staged into shared memory by cp.async.bulk copies issued by one thread, NST rows
// ahead of the row being folded in: the bytes in flight no longer cost registers or per-thread address
// arithmetic.  W is odd for MUSE (3721), so a row segment starts at an arbitrary 4-byte offset; the copy
// fetches the enclosing 16-byte aligned 528 bytes and the readers add the (warp-uniform) shift.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}

constexpr int kMarchStages = 4;
constexpr int kMarchRowFloats = 132;   // 128 channels + up to 3 floats of alignment shift, 16-byte multiple

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// y pass of step u (u = step mod P, a compile-time value so that the P partial rows stay in registers):
// the x-convolved input row h feeds output rows through taps 0..P-1; tap P-1 is a row's first
// contribution (it initialises the slot), tap 0 its last -- that row is finished and returned in e.
template <int P, int TX, int NT, int U>
__device__ __forceinline__ void y_pass(float (&A)[P][TX], const float (&h)[TX], const MarchTaps &tp, float *dst, int c) {
  if constexpr (TX % 2 == 0) {
#pragma unroll
    for (int m = 0; m < P - 1; ++m)
#pragma unroll
      for (int ox = 0; ox < TX; ox += 2) {
        const float2 r = ffma2(tp.ky2[m], make_float2(h[ox], h[ox + 1]),
                               make_float2(A[(U + m) % P][ox], A[(U + m) % P][ox + 1]));
        A[(U + m) % P][ox] = r.x; A[(U + m) % P][ox + 1] = r.y;
      }
#pragma unroll
    for (int ox = 0; ox < TX; ox += 2) {
      const float2 r = fmul2(tp.ky2[P - 1], make_float2(h[ox], h[ox + 1]));
      A[(U + P - 1) % P][ox] = r.x; A[(U + P - 1) % P][ox + 1] = r.y;
    }
  } else {
#pragma unroll
    for (int m = 0; m < P - 1; ++m)
#pragma unroll
      for (int ox = 0; ox < TX; ++ox) A[(U + m) % P][ox] = fmaf(tp.ky[m], h[ox], A[(U + m) % P][ox]);
#pragma unroll
    for (int ox = 0; ox < TX; ++ox) A[(U + P - 1) % P][ox] = tp.ky[P - 1] * h[ox];
  }
  // the finished row (slot U % P) goes straight to the LSF tile `dst` (NULL while the first rows fill up):
  // even TX channel-major [NT][TX / 2] pairs, odd TX column-major [TX][NT]
  if (dst) {
    if constexpr (TX % 2 == 0) {
      float2 *tile = reinterpret_cast<float2 *>(dst);
#pragma unroll
      for (int ox = 0; ox < TX; ox += 2) tile[c * (TX / 2) + ox / 2] = make_float2(A[U % P][ox], A[U % P][ox + 1]);
    } else {
#pragma unroll
      for (int ox = 0; ox < TX; ++ox) dst[ox * NT + c] = A[U % P][ox];
    }
  }
}

template <int P, int TX, int NT>
__device__ __forceinline__ void y_pass_dyn(int u, float (&A)[P][TX], const float (&h)[TX], const MarchTaps &tp,
                                           float *dst, int c) {
  switch (u) {
    case 0: y_pass<P, TX, NT, 0>(A, h, tp, dst, c); break;
    case 1: if constexpr (P > 1) y_pass<P, TX, NT, 1>(A, h, tp, dst, c); break;
    case 2: if constexpr (P > 2) y_pass<P, TX, NT, 2>(A, h, tp, dst, c); break;
    case 3: if constexpr (P > 3) y_pass<P, TX, NT, 3>(A, h, tp, dst, c); break;
    case 4: if constexpr (P > 4) y_pass<P, TX, NT, 4>(A, h, tp, dst, c); break;
    case 5: if constexpr (P > 5) y_pass<P, TX, NT, 5>(A, h, tp, dst, c); break;
    default: if constexpr (P > 6) y_pass<P, TX, NT, 6>(A, h, tp, dst, c); break;
  }
}

template <int P, int KE, int TX, int NT>
__device__ __forceinline__ void march_body_bulk(const float *__restrict__ in, float *__restrict__ out, int ny, int nx,
                                                int W, int pin, int pout, int x0, int w0, int ya, int yb, const MarchTaps &tp,
                                                float (&s_mid)[4][TX][NT],
                                                float (&s_in)[kMarchStages][TX + P - 1][kMarchRowFloats],
                                                uint64_t (&s_bar)[kMarchStages + 4]) {
  static_assert(NT == 128 && kMarchStages == 4, "row staging is sized for 128 channels, 4 warps, 4 stages");
  constexpr int NST = kMarchStages;
  constexpr int TL = NT - (KE - 1);
  constexpr int HE = (KE - 1) / 2;
  constexpr int C = (P - 1) / 2;
  constexpr int H = P - 1 - C;
  constexpr int IX = TX + P - 1;
  constexpr uint32_t kRowBytes = kMarchRowFloats * 4;
  const int c = threadIdx.x;
  const int q0 = w0 - HE;                // first input channel of the tile (the caller checked the window fits)
  const int nsteps = (yb - ya) + P - 1;
  const unsigned W4o = 4u * (unsigned)pout;
  const size_t rowstride = (size_t)nx * pin, rowstride_o = (size_t)nx * pout;
  const int nvalid = min(TX, nx - x0);
  const int yfirst = max(ya - H, 0), ylast = min(yb + C, ny);   // valid input rows [yfirst, ylast)
  unsigned colmask = 0;
#pragma unroll
  for (int ix = 0; ix < IX; ++ix)
    if (x0 - H + ix >= 0 && x0 - H + ix < nx) colmask |= 1u << ix;
  const int ncols = __popc(colmask);

  if (c == 0) {
#pragma unroll
    for (int st = 0; st < NST; ++st) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&s_bar[st])));
#pragma unroll
    for (int b = 0; b < 4; ++b)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&s_bar[NST + b])), "r"(NT));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // columns outside the cube are never copied: zero them once, the readers need no predicate
  if (ncols != IX) {
    for (int t = c; t < NST * IX * kMarchRowFloats; t += NT) {
      const int ix = (t / kMarchRowFloats) % IX;
      if (!((colmask >> ix) & 1u)) (&s_in[0][0][0])[t] = 0.f;
    }
  }
  __syncthreads();

  // element index (in floats, relative to a 16-byte aligned origin) of (row 0, column `lane`, channel q0)
  const uint64_t in_words = (uint64_t)(uintptr_t)in >> 2;
  const int lane = c & 31, warp = c >> 5;
  const bool copy_lane = lane < IX && ((colmask >> lane) & 1u);
  const uint64_t e_lane = in_words + (uint64_t)(copy_lane ? x0 - H + lane : 0) * (uint64_t)pin + (uint64_t)q0;
  // Row k is issued by warp k % 4 (the copy work rotates over the warps, i.e. over the SM's four
  // schedulers): lane 0 arms the stage's barrier, lane ix copies column ix.
  auto issue_row = [&](int yy) {
    const int k = yy - yfirst;
    if ((k & 3) != warp) return;
    const int st = k % NST;
    const uint32_t bar = smem_addr(&s_bar[st]);
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kRowBytes * (uint32_t)ncols) : "memory");
    __syncwarp();
    if (copy_lane) {
      const uint64_t e = e_lane + (uint64_t)yy * (uint64_t)rowstride;
      const uint64_t src = (e & ~(uint64_t)3) << 2;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_addr(&s_in[st][lane][0])), "l"(src), "r"(kRowBytes), "r"(bar) : "memory");
    }
  };
  for (int r = 0; r < NST && yfirst + r < ylast; ++r) issue_row(yfirst + r);

  // LSF of the finished row `oy` held in s_mid[b], then the coalesced store
  auto lsf_store = [&](int oy, int b) {
    if (c < TL) {
      float o[TX];
      if constexpr (TX % 2 == 0) {
        // even TX: the tile is stored channel-major, [channel][column], so a column PAIR is one 64-bit load
        // (a half-warp's 16 addresses, TX words apart, fall into 16 different bank pairs) feeding one FFMA2
        const float2 *tile = reinterpret_cast<const float2 *>(&s_mid[b][0][0]);   // [NT][TX / 2]
#pragma unroll
        for (int ox = 0; ox < TX; ox += 2) {
          float2 acc = fmul2(tp.kl2[0], tile[c * (TX / 2) + ox / 2]);
#pragma unroll
          for (int t = 1; t < KE; ++t) acc = ffma2(tp.kl2[t], tile[(c + t) * (TX / 2) + ox / 2], acc);
          o[ox] = acc.x; o[ox + 1] = acc.y;
        }
      } else {
#pragma unroll
        for (int ox = 0; ox < TX; ++ox) {
          float acc = tp.kl[0] * s_mid[b][ox][c];
#pragma unroll
          for (int t = 1; t < KE; ++t) acc = fmaf(tp.kl[t], s_mid[b][ox][c + t], acc);
          o[ox] = acc;
        }
      }
      float *dst = out + (size_t)oy * rowstride_o + (size_t)x0 * pout + (w0 + c);
      if (nvalid == TX) Cols<0, TX>::store(dst, W4o, o);
      else Cols<0, TX>::store_n(dst, W4o, o, nvalid);
    }
  };

  float A[P][TX];
#pragma unroll
  for (int a = 0; a < P; ++a)
#pragma unroll
    for (int b = 0; b < TX; ++b) A[a][b] = 0.f;

  // Step i folds input row yy = ya - H + i in and finishes output row yy - C (emit index j = i - (P-1)).
  // Order inside a step: [row i: stage wait, x pass, y pass with the finished row j stored to s_mid, arrive j]
  // -> [wait for row j-1 in s_mid, LSF, store].  A whole step of arithmetic sits between a thread's arrive
  // for a row and its wait for it, so the block never stalls on its slowest warp.  Four s_mid buffers: a
  // thread that stores row j has passed the wait for row j-2, i.e. every thread has finished step j-3 and
  // with it the LSF of row j-4, the previous user of the buffer.
  int u = 0;
  for (int i = 0; i < nsteps; ++i) {
    const int yy = ya - H + i;
    const bool row_in = yy >= yfirst && yy < ylast;   // rows outside the cube are zero padding
    float h[TX];
    if (row_in) {
      const int k = yy - yfirst;
      const int st = k % NST;
      mbar_wait(smem_addr(&s_bar[st]), (uint32_t)((k / NST) & 1));
      // alignment shift of column ix: ((row, column, q0) element index) mod 4, warp-uniform
      const uint32_t e0 = (uint32_t)(in_words & 3u) + (uint32_t)(q0 & 3) +
                          (uint32_t)((((uint64_t)yy * nx + (uint64_t)(x0 - H + IX)) * (uint64_t)pin) & 3u);
      float v[IX];
#pragma unroll
      for (int ix = 0; ix < IX; ++ix) {
        // (x0 - H + ix) * pitch = (x0 - H + IX) * pitch - (IX - ix) * pitch, and -x == 3x (mod 4)
        const uint32_t sh = (e0 + (uint32_t)((IX - ix) * 3) * (uint32_t)(pin & 3)) & 3u;
        v[ix] = s_in[st][ix][sh + c];
      }
#pragma unroll
      for (int ox = 0; ox < TX; ++ox) {   // x pass: h[ox] = sum_n kx[n] * in[x - n + C]
        float acc = tp.kx[0] * v[ox + P - 1];
#pragma unroll
        for (int n = 1; n < P; ++n) acc = fmaf(tp.kx[n], v[ox + P - 1 - n], acc);
        h[ox] = acc;
      }
    } else {
#pragma unroll
      for (int ox = 0; ox < TX; ++ox) h[ox] = 0.f;
    }
    const int j = i - (P - 1);            // emit index of this step's finished row (valid when >= 0)
    y_pass_dyn<P, TX, NT>(u, A, h, tp, j >= 0 ? &s_mid[j & 3][0][0] : nullptr, c);
    u = (u + 1 == P) ? 0 : u + 1;
    if (j >= 0) mbar_arrive(smem_addr(&s_bar[NST + (j & 3)]));
    if (j >= 1) {
      mbar_wait(smem_addr(&s_bar[NST + ((j - 1) & 3)]), (uint32_t)(((j - 1) >> 2) & 1));
      // every thread has read the stage of input row yy - 1: refill it NST rows ahead
      if (yy - 1 >= yfirst && yy - 1 + NST < ylast) issue_row(yy - 1 + NST);
      lsf_store(ya + j - 1, (j - 1) & 3);
    } else if (i >= 1) {
      // no finished row yet: the stage hand-over still needs every thread past its reads of row yy - 1
      __syncthreads();
      if (yy - 1 >= yfirst && yy - 1 + NST < ylast) issue_row(yy - 1 + NST);
    }
  }
  {
    const int j = nsteps - P;             // last finished row
    mbar_wait(smem_addr(&s_bar[NST + (j & 3)]), (uint32_t)((j >> 2) & 1));
    lsf_store(ya + j, j & 3);
  }
}

template <int P, int KE, int TX, int NT>
__global__ void __launch_bounds__(NT, RBX_MARCH_MINB * 128 / NT)
psf_lsf_march_kernel(const float *__restrict__ in, float *__restrict__ out, int ny, int nx, int W, int pin, int pout,
                     int rows_per_seg, int use_bulk, const __grid_constant__ MarchTaps tp) {
  constexpr int TL = NT - (KE - 1);
  constexpr int HE = (KE - 1) / 2;
  // dynamic shared memory (more than the 48 KB static limit): [s_in | s_mid | s_bar]
  extern __shared__ __align__(128) unsigned char march_smem[];
  using InT = float[kMarchStages][TX + P - 1][kMarchRowFloats];
  using MidT = float[4][TX][NT];
  using BarT = uint64_t[kMarchStages + 4];
  InT &s_in = *reinterpret_cast<InT *>(march_smem);
  MidT &s_mid = *reinterpret_cast<MidT *>(march_smem + sizeof(InT));
  BarT &s_bar = *reinterpret_cast<BarT *>(march_smem + sizeof(InT) + sizeof(MidT));
  const int tiles_x = (nx + TX - 1) / TX;
  const int x0 = (blockIdx.x % tiles_x) * TX;
  const int w0 = (blockIdx.x / tiles_x) * TL;
  const int ya = blockIdx.y * rows_per_seg;
  const int yb = min(ya + rows_per_seg, ny);
  // bulk path: the whole 132-float staging window of every column lies inside its spaxel's spectrum
  const bool lam_in = w0 - HE >= 0 && w0 - HE + kMarchRowFloats <= W;
  if (use_bulk && lam_in) march_body_bulk<P, KE, TX, NT>(in, out, ny, nx, W, pin, pout, x0, w0, ya, yb, tp, s_mid, s_in, s_bar);
  else march_body<P, KE, TX, NT, true>(in, out, ny, nx, W, pin, pout, x0, w0, ya, yb, tp, s_mid);
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
int launch_march(const float *d_in, float *d_out, int ny, int nx, int W, int pin, int pout, const MarchTaps &taps,
                 cudaStream_t stream) {
  auto kernel = psf_lsf_march_kernel<P, KE, TX, NT>;
  constexpr int TL = NT - (KE - 1);
  constexpr size_t kSmem = sizeof(float) * ((size_t)kMarchStages * (TX + P - 1) * kMarchRowFloats + 4 * TX * NT) +
                           sizeof(uint64_t) * (kMarchStages + 4);
  static int slots_cached[64] = {};
  int dev = 0;
  RBX_CUDA_OK(cudaGetDevice(&dev));
  int slots = (dev >= 0 && dev < 64) ? slots_cached[dev] : 0;
  if (!slots) {
    int per_sm = 0, sms = 0;
    RBX_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    RBX_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NT, kSmem));
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
  // bulk copies need 16-byte aligned global addresses: the staging aligns down relative to d_in
  const bool no_bulk = opt_on(OPT_MARCH_NO_BULK);
  // (an unaligned slab start inside a wider cube is fine: the bytes before it belong to the same cube)
  const int use_bulk = ((((uintptr_t)d_in & 15u) == 0 || pin > W) && !no_bulk) ? 1 : 0;
  kernel<<<grid, NT, kSmem, stream>>>(d_in, d_out, ny, nx, W, pin, pout, rows, use_bulk, taps);
  count_launch();
  RBX_LAUNCH_OK();
  return RBX_OK;
}

template <int P, int KE>
int launch_march_tx(const float *d_in, float *d_out, int ny, int nx, int W, int pin, int pout, const MarchTaps &taps,
                    cudaStream_t stream) {
  if (nx >= 40) return launch_march<P, KE, 10, 128>(d_in, d_out, ny, nx, W, pin, pout, taps, stream);
  return launch_march<P, KE, 5, 128>(d_in, d_out, ny, nx, W, pin, pout, taps, stream);
}

template <int P>
int launch_march_ke(int KE, const float *d_in, float *d_out, int ny, int nx, int W, int pin, int pout,
                    const MarchTaps &taps, cudaStream_t stream) {
  switch (KE) {
    case 1: return launch_march_tx<P, 1>(d_in, d_out, ny, nx, W, pin, pout, taps, stream);
    case 7: return launch_march_tx<P, 7>(d_in, d_out, ny, nx, W, pin, pout, taps, stream);
    case 13: return launch_march_tx<P, 13>(d_in, d_out, ny, nx, W, pin, pout, taps, stream);
    default: return launch_march_tx<P, 25>(d_in, d_out, ny, nx, W, pin, pout, taps, stream);
  }
}

}  // namespace

// Host-tap PSF + LSF.  h_psf (M, N) or NULL (identity); h_lsf (K = 2 ext + 1) or NULL (identity).
// Returns RBX_ERR_UNSUPPORTED (no error text) when the taps do not fit the marching kernel (PSF not an
// outer product, not square 1/3/5/7, or more than 25 significant LSF taps): the caller then uses the
// device-tap kernels of conv.cu.
extern "C" int rbx_psf_lsf_taps(const float *d_in, float *d_out, int ny, int nx, int W, const float *h_psf, int M,
                                int N, const float *h_lsf, int K, int ext, void *stream) {
  return rbx_psf_lsf_taps_pitched(d_in, W, d_out, W, ny, nx, W, h_psf, M, N, h_lsf, K, ext, stream);
}

// The same on a wavelength SLAB of a cube: W channels starting at d_in / d_out, consecutive spaxels
// in_pitch / out_pitch floats apart (>= W).  Channels outside [0, W) of the slab count as zero.
extern "C" int rbx_psf_lsf_taps_pitched(const float *d_in, int in_pitch, float *d_out, int out_pitch, int ny, int nx,
                                        int W, const float *h_psf, int M, int N, const float *h_lsf, int K, int ext,
                                        void *stream) {
  const int pin = in_pitch, pout = out_pitch;
  RBX_REQUIRE(d_in && d_out && d_in != d_out, "rbx_psf_lsf_taps: bad pointers (no aliasing)");
  RBX_REQUIRE(ny > 0 && nx > 0 && W > 0 && pin >= W && pout >= W, "rbx_psf_lsf_taps: bad shape");
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
  for (int i = 0; i < kMarchMaxP; ++i) tp.taps.ky2[i] = make_float2(tp.taps.ky[i], tp.taps.ky[i]);
  for (int i = 0; i < kMarchMaxK; ++i) tp.taps.kl2[i] = make_float2(tp.taps.kl[i], tp.taps.kl[i]);
  cudaStream_t s = (cudaStream_t)stream;
  switch (tp.P) {
    case 1: return launch_march_ke<1>(tp.KE, d_in, d_out, ny, nx, W, pin, pout, tp.taps, s);
    case 3: return launch_march_ke<3>(tp.KE, d_in, d_out, ny, nx, W, pin, pout, tp.taps, s);
    case 5: return launch_march_ke<5>(tp.KE, d_in, d_out, ny, nx, W, pin, pout, tp.taps, s);
    default: return launch_march_ke<7>(tp.KE, d_in, d_out, ny, nx, W, pin, pout, tp.taps, s);
  }
}
