// Stable LSD radix sort of (key, particle index) pairs for the cube build -- the "device radix sort of particles by
// spaxel index" of the north star, written for this key layout: key = spaxel << cell_bits | template cell, at most
// ~26 significant bits, so 3 (MUSE) or 4 (150 x 150) passes of <= 8 bits.
//
// One kernel per pass, each key read once per pass (single-pass chained scan with decoupled look-back):
//   * the digit histograms of ALL passes come for free from prep_kernel, which writes the keys (a shared-memory
//     histogram per block, flushed with atomics): no histogram kernel, no scan kernel;
//   * a block takes the next tile of 4096 keys (atomic ticket: a tile only ever waits for tiles that are already
//     running), ranks its keys per digit -- inside a warp with __match_any_sync in index order, across warps by a
//     scan of the per-warp counts, so equal digits keep their input order: the sort is STABLE and the result
//     bit-reproducible -- publishes its per-digit counts, looks back over the preceding tiles' counts / inclusive
//     prefixes, puts the tile in digit order in shared memory and writes every digit's run to its final position
//     (consecutive threads write consecutive elements: coalesced runs instead of 4-byte scatters).
//   * pass 0 takes the values to be the particle indices 0 .. n-1 (nothing is read for them).
#include "common.cuh"

namespace rbx {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;   // keys per tile
constexpr uint32_t kFlagAgg = 0x40000000u, kFlagPrefix = 0x80000000u, kCountMask = 0x3fffffffu;

SortPlan make_sort_plan(int64_t n, int end_bit) {
  SortPlan sp;
  sp.npass = std::max(1, (end_bit + kSortMaxBits - 1) / kSortMaxBits);
  const int per = (end_bit + sp.npass - 1) / sp.npass;
  int shift = 0;
  for (int i = 0; i < kSortMaxPasses; ++i) {
    sp.shift[i] = shift;
    sp.bits[i] = i < sp.npass ? std::max(1, std::min(per, end_bit - shift)) : 0;
    shift += sp.bits[i];
  }
  sp.ntiles = (int)((n + kSortTile - 1) / kSortTile);
  return sp;
}

size_t sort_state_words(const SortPlan &sp) {
  // per pass: global digit histogram [256], ticket [1 (+ padding to 256)], tile states [ntiles][256]
  return (size_t)sp.npass * ((size_t)sp.ntiles + 2) * 256;
}

__global__ void __launch_bounds__(kSortThreads, 3)
radix_pass_kernel(const uint32_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint32_t *__restrict__ kout,
                  uint32_t *__restrict__ vout, int n, int shift, int bits, uint32_t *__restrict__ pass_state) {
  // pass_state: [0, 256) global digit histogram (from prep_kernel), [256] ticket, [512 + 256 tile, +256) tile states
  __shared__ uint32_t s_whist[kSortThreads / 32][256];
  __shared__ uint32_t s_base[256];   // global position of the tile's first element of each digit
  __shared__ uint32_t s_scan[256];
  __shared__ uint32_t s_loc[256];    // position of each digit's run inside the tile
  __shared__ uint32_t s_keys[kSortTile], s_vals[kSortTile];
  __shared__ int s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = 1 << bits;
  const uint32_t dmask = (uint32_t)nb - 1u;
  if (tid == 0) s_tile = (int)atomicAdd(pass_state + 256, 1u);
  for (int q = tid; q < (kSortThreads / 32) * 256; q += kSortThreads) (&s_whist[0][0])[q] = 0u;
  __syncthreads();
  const int tile = s_tile;
  volatile uint32_t *state = pass_state + 512;
  const int tbase = tile * kSortTile + warp * (32 * kSortItems);

  // ---- rank my keys: inside the warp in index order (match-any groups), item by item --------------------------
  // Three sweeps so that nothing waits on its predecessor: (1) the 16 key loads and match-any votes are independent;
  // (2) the leader of every group bumps the warp's digit counter with ONE shared-memory atomic -- atomics of a warp on
  // one address are applied in issue order, so item i + 1 sees item i's count without a __syncwarp, and the 16
  // atomics are in flight together (the former load / add / store per item was a chain of 16 shared-memory round
  // trips); (3) the counter values are broadcast from the leaders.
  uint32_t key[kSortItems];
  int rank[kSortItems];
  unsigned grp[kSortItems];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const int idx = tbase + i * 32 + lane;
    key[i] = idx < n ? kin[idx] : 0u;
  }
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const bool valid = tbase + i * 32 + lane < n;
    const uint32_t d = (key[i] >> shift) & dmask;
    grp[i] = __match_any_sync(0xffffffffu, valid ? d : 0x10000u + (uint32_t)lane);
  }
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const bool valid = tbase + i * 32 + lane < n;
    const uint32_t d = (key[i] >> shift) & dmask;
    rank[i] = 0;
    if (lane == __ffs(grp[i]) - 1 && valid) rank[i] = (int)atomicAdd(&s_whist[warp][d], (uint32_t)__popc(grp[i]));
    __syncwarp();   // orders the atomics of different lanes on one counter (it does not wait for their results)
  }
  // the values travel with the keys: requested here, they arrive while the tile waits for its look-back
  uint32_t val[kSortItems];
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const int idx = tbase + i * 32 + lane;
    val[i] = (vin && idx < n) ? vin[idx] : (uint32_t)idx;
  }
#pragma unroll
  for (int i = 0; i < kSortItems; ++i)
    rank[i] = __shfl_sync(0xffffffffu, rank[i], __ffs(grp[i]) - 1) + __popc(grp[i] & lt);
  __syncthreads();

  // ---- per digit: exclusive prefix over the warps, tile total, look-back over the preceding tiles ----------------
  if (tid < nb) {
    uint32_t tot = 0;
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w) {
      const uint32_t t = s_whist[w][tid];
      s_whist[w][tid] = tot;
      tot += t;
    }
    uint32_t excl = 0;
    if (tile == 0) {
      state[tid] = tot | kFlagPrefix;
    } else {
      state[(size_t)tile * 256 + tid] = tot | kFlagAgg;
      // look back eight tiles at a time (independent loads: the chain of dependent L2 round trips is what a
      // tile waits for); tiles before tile 0 count as an empty inclusive prefix
      int t = tile - 1;
      bool done = false;
      while (!done) {
        uint32_t v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = t - q >= 0 ? (uint32_t)state[(size_t)(t - q) * 256 + tid] : 0x80000000u;
        int used = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (done || used != q) continue;            // stop at the first prefix / the first tile that is not ready
          if (v[q] & (kFlagPrefix | kFlagAgg)) {
            excl += v[q] & kCountMask;
            ++used;
            if (v[q] & kFlagPrefix) done = true;
          }
        }
        t -= used;
      }
      state[(size_t)tile * 256 + tid] = (excl + tot) | kFlagPrefix;
    }
    s_base[tid] = excl;
    s_scan[tid] = pass_state[tid];   // global count of this digit
    s_loc[tid] = tot;                // count of this digit in the tile
  }
  __syncthreads();
  // exclusive scans over the digits (<= 256 bins: one warp each, 8 bins per lane): warp 0 the global digit histogram
  // (-> global base), warp 1 the tile's digit counts (-> run start inside the tile)
  if (warp < 2) {
    uint32_t *arr = warp == 0 ? s_scan : s_loc;
    uint32_t v[8], sum = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { const int b = lane * 8 + q; v[q] = b < nb ? arr[b] : 0u; sum += v[q]; }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    uint32_t run = inc - sum;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int b = lane * 8 + q;
      if (b < nb) { if (warp == 0) s_base[b] += run; else s_loc[b] = run; }
      run += v[q];
    }
  }
  __syncthreads();

  // ---- the tile in digit order in shared memory, then every run to its final position ---------------------------
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const int idx = tbase + i * 32 + lane;
    if (idx < n) {
      const uint32_t d = (key[i] >> shift) & dmask;
      const uint32_t lp = s_loc[d] + s_whist[warp][d] + (uint32_t)rank[i];
      s_keys[lp] = key[i];
      s_vals[lp] = val[i];
    }
  }
  __syncthreads();
  const int cnt = min(kSortTile, n - tile * kSortTile);
  for (int j = tid; j < cnt; j += kSortThreads) {
    const uint32_t k = s_keys[j];
    const uint32_t d = (k >> shift) & dmask;
    const uint32_t pos = s_base[d] + ((uint32_t)j - s_loc[d]);
    kout[pos] = k;
    vout[pos] = s_vals[j];
  }
}

// Sorts n (key, index) pairs by the low end_bit bits of the key; the values of the first pass are 0 .. n-1.
// buf_a holds the keys on entry; the result lands in *keys_sorted / *vals_sorted (one of the two buffer pairs).
// d_state (sort_state_words() words) must hold the digit histograms written by prep_kernel and be zero elsewhere.
int radix_sort_pairs(const SortPlan &sp, uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n,
                     uint32_t *d_state, uint32_t **keys_sorted, uint32_t **vals_sorted, cudaStream_t stream) {
  uint32_t *kin = keys_a, *vin = nullptr, *kout = keys_b, *vout = vals_b;
  const size_t per_pass = ((size_t)sp.ntiles + 2) * 256;
  for (int p = 0; p < sp.npass; ++p) {
    radix_pass_kernel<<<sp.ntiles, kSortThreads, 0, stream>>>(kin, vin, kout, vout, (int)n, sp.shift[p], sp.bits[p],
                                                               d_state + (size_t)p * per_pass);
    count_launch();
    RBX_LAUNCH_OK();
    // ping-pong: the pair just written becomes the input (pass 0 wrote (keys_b, vals_b) without reading values)
    uint32_t *nk = kout, *nv = vout;
    kout = (nk == keys_b) ? keys_a : keys_b;
    vout = (nv == vals_b) ? vals_a : vals_b;
    kin = nk;
    vin = nv;
  }
  *keys_sorted = kin;
  *vals_sorted = vin;
  return RBX_OK;
}

// ---- stand-alone form (rbx_sort_by_spaxel): keys and digit histograms from the spaxel ids ------------------------
// Inside the cube build prep_kernel writes the keys and their digit histograms; here a small kernel does it for a
// plain array of spaxel ids.  Ids outside [0, nseg) -- the ones segment_sum drops -- become the key nseg: they sort
// behind every segment.
__global__ void __launch_bounds__(256)
spaxel_keys_kernel(const int32_t *__restrict__ pixel, int n, int nseg, uint32_t *__restrict__ keys,
                   uint32_t *__restrict__ state, SortPlan sp, size_t per_pass) {
  __shared__ uint32_t s_hist[kSortMaxPasses][256];
  for (int q = threadIdx.x; q < kSortMaxPasses * 256; q += blockDim.x) (&s_hist[0][0])[q] = 0u;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t p = (uint32_t)pixel[i];           // a negative id is a huge unsigned one
    const uint32_t k = p < (uint32_t)nseg ? p : (uint32_t)nseg;
    keys[i] = k;
#pragma unroll
    for (int s = 0; s < kSortMaxPasses; ++s)
      if (s < sp.npass) atomicAdd(&s_hist[s][(k >> sp.shift[s]) & ((1u << sp.bits[s]) - 1u)], 1u);
  }
  __syncthreads();
  for (int q = threadIdx.x; q < sp.npass * 256; q += blockDim.x) {
    const uint32_t v = (&s_hist[0][0])[q];
    if (v) atomicAdd(state + (size_t)(q >> 8) * per_pass + (q & 255), v);
  }
}

// offsets[s] = first position of the sorted keys with key >= s, s = 0 .. nseg (searchsorted 'left'): every run
// boundary fills the entries of the (possibly empty) segments it steps over.
__global__ void segment_offsets_kernel(const uint32_t *__restrict__ sorted, int n, int nseg,
                                       int32_t *__restrict__ offsets) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)sorted[i];
    const int prev = i > 0 ? (int)sorted[i - 1] : -1;
    for (int s = prev + 1; s <= k; ++s) offsets[s] = (int32_t)i;
    if (i == n - 1)
      for (int s = k + 1; s <= nseg; ++s) offsets[s] = n;
  }
}

static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

static int spaxel_end_bit(int nseg) {
  int b = 1;
  while (b < 31 && (1u << b) <= (uint32_t)nseg) ++b;   // bits of the largest key, nseg
  return b;
}

}  // namespace rbx

using namespace rbx;

extern "C" size_t rbx_sort_by_spaxel_workspace_bytes(int64_t n, int num_segments) {
  if (n <= 0 || num_segments <= 0) return 256;
  const SortPlan sp = make_sort_plan(n, spaxel_end_bit(num_segments));
  return align256(sizeof(uint32_t) * sort_state_words(sp)) + 3 * align256(sizeof(uint32_t) * (size_t)n) + 256;
}

extern "C" int rbx_sort_by_spaxel(const int32_t *d_pixel, int64_t n, int num_segments, int32_t *d_order,
                                  int32_t *d_sorted, int32_t *d_offsets, void *d_workspace, size_t workspace_bytes,
                                  void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RBX_REQUIRE(n >= 0 && n <= (int64_t(1) << 30), "rbx_sort_by_spaxel: n outside [0, 2^30]");
  RBX_REQUIRE(num_segments >= 1 && num_segments < (1 << 30), "rbx_sort_by_spaxel: num_segments outside [1, 2^30)");
  if (n == 0) {
    if (d_offsets) RBX_CUDA_OK(cudaMemsetAsync(d_offsets, 0, sizeof(int32_t) * ((size_t)num_segments + 1), stream));
    return RBX_OK;
  }
  RBX_REQUIRE(d_pixel && d_order, "rbx_sort_by_spaxel: null pointer");
  RBX_REQUIRE(d_workspace, "rbx_sort_by_spaxel: null workspace");
  if (workspace_bytes < rbx_sort_by_spaxel_workspace_bytes(n, num_segments)) {
    set_error("rbx_sort_by_spaxel: workspace too small (see rbx_sort_by_spaxel_workspace_bytes)");
    return RBX_ERR_WORKSPACE_TOO_SMALL;
  }
  const SortPlan sp = make_sort_plan(n, spaxel_end_bit(num_segments));
  const size_t state_bytes = align256(sizeof(uint32_t) * sort_state_words(sp));
  const size_t arr_bytes = align256(sizeof(uint32_t) * (size_t)n);
  char *ws = (char *)(((uintptr_t)d_workspace + 255) & ~(uintptr_t)255);
  uint32_t *state = (uint32_t *)ws;
  uint32_t *w0 = (uint32_t *)(ws + state_bytes), *w1 = (uint32_t *)(ws + state_bytes + arr_bytes),
           *w2 = (uint32_t *)(ws + state_bytes + 2 * arr_bytes);
  // pass p writes the (b) pair when p is even and the (a) pair when it is odd: the pair the LAST pass writes is
  // the caller's output arrays, the other pair lives in the workspace
  uint32_t *out_keys = d_sorted ? (uint32_t *)d_sorted : w2, *out_vals = (uint32_t *)d_order;
  const bool last_is_b = (sp.npass & 1) != 0;
  uint32_t *keys_a = last_is_b ? w0 : out_keys, *vals_a = last_is_b ? w1 : out_vals;
  uint32_t *keys_b = last_is_b ? out_keys : w0, *vals_b = last_is_b ? out_vals : w1;
  RBX_CUDA_OK(cudaMemsetAsync(state, 0, sizeof(uint32_t) * sort_state_words(sp), stream));
  const int blocks = (int)std::min<int64_t>((n + 256 * 16 - 1) / (256 * 16), 148 * 8);
  spaxel_keys_kernel<<<blocks, 256, 0, stream>>>(d_pixel, (int)n, num_segments, keys_a, state, sp,
                                                 ((size_t)sp.ntiles + 2) * 256);
  count_launch();
  RBX_LAUNCH_OK();
  uint32_t *ks = nullptr, *vs = nullptr;
  const int rc = radix_sort_pairs(sp, keys_a, vals_a, keys_b, vals_b, n, state, &ks, &vs, stream);
  if (rc != RBX_OK) return rc;
  RBX_REQUIRE(ks == out_keys && vs == out_vals, "rbx_sort_by_spaxel: internal buffer order");
  if (d_offsets) {
    segment_offsets_kernel<<<blocks * 4, 256, 0, stream>>>(ks, (int)n, num_segments, d_offsets);
    count_launch();
    RBX_LAUNCH_OK();
  }
  return RBX_OK;
}
