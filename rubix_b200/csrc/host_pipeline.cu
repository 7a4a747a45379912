// rbx_pipeline_host: the whole path for callers that hold host (numpy) arrays.
// filter_particles -> spaxel_assignment -> fused cube -> PSF + LSF, with the H2D copies of the
// particle arrays and the D2H copy of the cube inside the call.  Device scratch comes from the
// stream-ordered allocator (cudaMallocAsync), so repeated calls reuse the pool without cudaMalloc.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

using namespace rbx;

namespace {
// One stream-ordered pool per device that never trims: with the default release threshold (0) the
// driver hands the scratch back to the OS at every synchronisation and the next call pays for
// mapping hundreds of MB again (measured: 6 ms .. 1.1 s per call at 10^6 .. 10^7 particles).
std::mutex g_pool_mutex;
cudaMemPool_t g_pools[64] = {};

int scratch_pool(cudaMemPool_t *out) {
  int dev = 0;
  RBX_CUDA_OK(cudaGetDevice(&dev));
  RBX_REQUIRE(dev >= 0 && dev < 64, "rbx_pipeline_host: device ordinal out of range");
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  if (!g_pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    RBX_CUDA_OK(cudaMemPoolCreate(&g_pools[dev], &props));
    uint64_t keep = UINT64_MAX;
    RBX_CUDA_OK(cudaMemPoolSetAttribute(g_pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
  }
  *out = g_pools[dev];
  return RBX_OK;
}

// second stream + events for the chunked host-to-device overlap, one set per device
struct CopyLane {
  cudaStream_t stream = nullptr;
  cudaEvent_t ready = nullptr;
  cudaEvent_t copied[16] = {};
  std::mutex busy;   // one chunked call at a time per device (the events are shared)
};
CopyLane g_lanes[64];

int copy_lane(CopyLane **out) {
  int dev = 0;
  RBX_CUDA_OK(cudaGetDevice(&dev));
  RBX_REQUIRE(dev >= 0 && dev < 64, "rbx_pipeline_host: device ordinal out of range");
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  CopyLane &l = g_lanes[dev];
  if (!l.stream) {
    RBX_CUDA_OK(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
    RBX_CUDA_OK(cudaEventCreateWithFlags(&l.ready, cudaEventDisableTiming));
    for (auto &e : l.copied) RBX_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  *out = &l;
  return RBX_OK;
}

struct Scratch {
  std::vector<void *> ptrs;
  cudaStream_t s;
  cudaMemPool_t pool = nullptr;
  explicit Scratch(cudaStream_t st) : s(st) {}
  ~Scratch() {
    for (void *p : ptrs) cudaFreeAsync(p, s);
  }
  template <typename T>
  int get(T **out, size_t count) {
    void *p = nullptr;
    if (!pool) {
      int rc = scratch_pool(&pool);
      if (rc != RBX_OK) return rc;
    }
    RBX_CUDA_OK(cudaMallocFromPoolAsync(&p, count * sizeof(T) + 256, pool, s));
    ptrs.push_back(p);
    *out = (T *)p;
    return RBX_OK;
  }
};
}  // namespace

#define TRY(x) do { int _rc = (x); if (_rc != RBX_OK) return _rc; } while (0)

namespace {
// The particle arrays of one call: the reference's layout ((n, 3) coords and velocity: 40 bytes per particle) or
// structure-of-arrays x, y, line-of-sight velocity (24 bytes per particle: what the path reads).
struct HostParticles {
  const float *coords = nullptr, *velocity = nullptr;        // (n, 3)
  const float *x = nullptr, *y = nullptr, *vlos = nullptr;   // (n,) each
  const float *mass = nullptr, *metallicity = nullptr, *age = nullptr;
  bool packed = false;
};

// d_cube_ext != nullptr: rbx_build_cube_host -- the cube (slab-major when nslab > 1) is left on the device in the
// caller's buffer, no PSF / LSF, no device-to-host copy, no synchronisation (stream-ordered like cudaMemcpyAsync: the
// host arrays must stay valid until `stream` has passed this call).
int pipeline_host_impl(const rbx_plan *plan, const HostParticles &hp, int64_t n, const float *h_edges, int n_edges,
                       int num_spaxels, int apply_filter, const float *h_psf, int M, int N, const float *h_lsf, int K,
                       int ext, float *h_cube, void *stream_, float *d_cube_ext = nullptr, int nslab = 1, int halo = 0) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RBX_REQUIRE(plan && (h_cube || d_cube_ext) && h_edges, "rbx_pipeline_host: null argument");
  RBX_REQUIRE(n >= 0 && n_edges >= 2 && num_spaxels >= 1, "rbx_pipeline_host: bad sizes");
  RBX_REQUIRE(n == 0 || (hp.mass && hp.metallicity && hp.age &&
                         (hp.packed ? (hp.x && hp.y && hp.vlos) : (hp.coords && hp.velocity))),
              "rbx_pipeline_host: null particle array");
  const float *h_mass = hp.mass, *h_metallicity = hp.metallicity, *h_age = hp.age;
  const int nc = hp.packed ? 1 : 3;   // floats per particle in the coordinate / velocity arrays
  const int W = plan->v.W;
  const size_t cube_elems = (size_t)num_spaxels * num_spaxels * W;
  Scratch sc(stream);
  float *d_coords, *d_y = nullptr, *d_vel, *d_mass, *d_met, *d_age, *d_edges, *d_cube, *d_cube2 = nullptr, *d_psf = nullptr,
        *d_lsf = nullptr;
  int32_t *d_pixel;
  void *d_ws;
  const size_t np = n > 0 ? (size_t)n : 1;
  TRY(sc.get(&d_coords, nc * np));
  if (hp.packed) TRY(sc.get(&d_y, np));
  TRY(sc.get(&d_vel, nc * np));
  TRY(sc.get(&d_mass, np));
  TRY(sc.get(&d_met, np));
  TRY(sc.get(&d_age, np));
  TRY(sc.get(&d_pixel, np));
  TRY(sc.get(&d_edges, (size_t)n_edges));
  if (d_cube_ext) d_cube = d_cube_ext;
  else TRY(sc.get(&d_cube, cube_elems));
  const size_t ws_bytes = rbx_build_cube_workspace_bytes(plan, n, num_spaxels);
  TRY(sc.get((char **)&d_ws, ws_bytes));
  RBX_CUDA_OK(cudaMemcpyAsync(d_edges, h_edges, sizeof(float) * n_edges, cudaMemcpyHostToDevice, stream));
  // The galaxy is binned in `chunks` contiguous particle ranges: the host-to-device copies of range c+1 run
  // on a second stream while the kernels of range c execute, and every range adds into the same cube
  // (fixed ranges, fixed order: the result stays deterministic).
  // (measured on B200, profiles/r01_e2e_chunks.txt: 10^7 particles 16.4 -> 12.1 ms with 4-6 ranges; at 10^6 the
  // per-range cost -- every range expands all spaxels and runs the whole launch sequence -- eats the overlap.
  // Round 2 tried 8 geometrically shrinking ranges (ratio 0.7 / 0.8 / 0.9, so that only a small last range's kernels
  // stay exposed): 10.5 / 10.2 / 9.7 ms against 8.7 ms with 5 equal ranges on the same workload -- the per-range
  // cost outweighs the shorter tail, so the ranges stay equal; option host_ratio keeps the experiment reachable.)
  int chunks = n >= 3000000 ? 5 : (n >= 1500000 ? 2 : 1);
  if (d_cube_ext && n >= 700000 && n < 3000000) chunks = n >= 1500000 ? 3 : 2;   // a rank's shard: no PSF / copy-out tail
  double ratio = 1.0;
  if (opt(OPT_HOST_CHUNKS) > 0) { chunks = (int)std::min<int64_t>(16, opt(OPT_HOST_CHUNKS)); ratio = 1.0; }
  if (opt(OPT_HOST_RATIO) > 0) ratio = std::min(1.0, std::max(0.3, (double)opt(OPT_HOST_RATIO) / 100.0));
  if (n == 0) chunks = 1;
  int64_t bnd[17];
  {
    double wsum = 0.0, w = 1.0, acc = 0.0;
    for (int c = 0; c < chunks; ++c) { wsum += w; w *= ratio; }
    w = 1.0;
    bnd[0] = 0;
    for (int c = 0; c < chunks; ++c) {
      acc += w / wsum;
      w *= ratio;
      bnd[c + 1] = c + 1 == chunks ? n : std::min<int64_t>(n, (int64_t)(acc * (double)n));
    }
  }
  CopyLane *lane = nullptr;
  std::unique_lock<std::mutex> lane_lock;
  if (chunks > 1) {
    TRY(copy_lane(&lane));
    lane_lock = std::unique_lock<std::mutex>(lane->busy);
    // the scratch was allocated in `stream` order: the copy stream may touch it only after that point
    RBX_CUDA_OK(cudaEventRecord(lane->ready, stream));
    RBX_CUDA_OK(cudaStreamWaitEvent(lane->stream, lane->ready, 0));
  }
  if (n == 0) {
    CubeBuild b0;
    b0.nslab = nslab; b0.halo = halo;
    TRY(build_cube_impl(plan, b0, 0, num_spaxels, d_cube, d_ws, ws_bytes, stream));
  }
  bool first_range = true;
  for (int c = 0; c < chunks && n > 0; ++c) {
    const int64_t lo = bnd[c], hi = bnd[c + 1], m = hi - lo;
    if (m <= 0) continue;
    cudaStream_t cs = chunks > 1 ? lane->stream : stream;
    if (hp.packed) {
      RBX_CUDA_OK(cudaMemcpyAsync(d_coords + lo, hp.x + lo, sizeof(float) * m, cudaMemcpyHostToDevice, cs));
      RBX_CUDA_OK(cudaMemcpyAsync(d_y + lo, hp.y + lo, sizeof(float) * m, cudaMemcpyHostToDevice, cs));
      RBX_CUDA_OK(cudaMemcpyAsync(d_vel + lo, hp.vlos + lo, sizeof(float) * m, cudaMemcpyHostToDevice, cs));
    } else {
      RBX_CUDA_OK(cudaMemcpyAsync(d_coords + 3 * lo, hp.coords + 3 * lo, sizeof(float) * 3 * m, cudaMemcpyHostToDevice, cs));
      RBX_CUDA_OK(cudaMemcpyAsync(d_vel + 3 * lo, hp.velocity + 3 * lo, sizeof(float) * 3 * m, cudaMemcpyHostToDevice, cs));
    }
    RBX_CUDA_OK(cudaMemcpyAsync(d_mass + lo, h_mass + lo, sizeof(float) * m, cudaMemcpyHostToDevice, cs));
    RBX_CUDA_OK(cudaMemcpyAsync(d_met + lo, h_metallicity + lo, sizeof(float) * m, cudaMemcpyHostToDevice, cs));
    RBX_CUDA_OK(cudaMemcpyAsync(d_age + lo, h_age + lo, sizeof(float) * m, cudaMemcpyHostToDevice, cs));
    if (chunks > 1) {
      RBX_CUDA_OK(cudaEventRecord(lane->copied[c], cs));
      RBX_CUDA_OK(cudaStreamWaitEvent(stream, lane->copied[c], 0));
    }
    // spaxel assignment + aperture filter (pixel -1: the same cube as zeroing the mass) inside the cube build's
    // first kernel: the particle arrays are read once
    CubeBuild b;
    b.vel = hp.packed ? d_vel + lo : d_vel + 3 * lo + plan->v.vel_comp;
    b.vstride = nc;
    b.mass = d_mass + lo; b.met = d_met + lo; b.age = d_age + lo;
    b.cx = d_coords + nc * lo;
    b.cy = hp.packed ? d_y + lo : d_coords + 3 * lo + 1;
    b.cstride = nc;
    b.edges = d_edges; b.n_edges = n_edges; b.mark_outside = apply_filter ? 1 : 0;
    b.accumulate = first_range ? 0 : 1;
    b.nslab = nslab; b.halo = halo;
    first_range = false;
    TRY(build_cube_impl(plan, b, m, num_spaxels, d_cube, d_ws, ws_bytes, stream));
  }
  if (d_cube_ext) return RBX_OK;   // the scratch is released in stream order (Scratch's destructor)
  float *result = d_cube;
  if (h_psf || h_lsf) {
    TRY(sc.get(&d_cube2, cube_elems));
    // host-known taps: separable PSF + pruned LSF in one marching pass (conv_march.cu)
    int rc = rbx_psf_lsf_taps(d_cube, d_cube2, num_spaxels, num_spaxels, W, h_psf, M, N, h_lsf, K, ext, stream);
    if (rc == RBX_OK) {
      result = d_cube2;
    } else if (rc != RBX_ERR_UNSUPPORTED) {
      return rc;
    } else {  // general taps: device-tap kernels
      if (h_psf) {
        TRY(sc.get(&d_psf, (size_t)M * N));
        RBX_CUDA_OK(cudaMemcpyAsync(d_psf, h_psf, sizeof(float) * M * N, cudaMemcpyHostToDevice, stream));
      }
      if (h_lsf) {
        TRY(sc.get(&d_lsf, (size_t)K));
        RBX_CUDA_OK(cudaMemcpyAsync(d_lsf, h_lsf, sizeof(float) * K, cudaMemcpyHostToDevice, stream));
      }
      if (h_psf && h_lsf) {
        rc = rbx_psf_lsf(d_cube, d_cube2, num_spaxels, num_spaxels, W, d_psf, M, N, d_lsf, K, ext, stream);
        if (rc == RBX_ERR_UNSUPPORTED) {  // taps too large for the fused tile: two passes
          TRY(rbx_convolve_psf(d_cube, d_cube2, num_spaxels, num_spaxels, W, d_psf, M, N, stream));
          TRY(rbx_convolve_lsf(d_cube2, d_cube, (int64_t)num_spaxels * num_spaxels, W, d_lsf, K, ext, stream));
          result = d_cube;
        } else {
          TRY(rc);
          result = d_cube2;
        }
      } else if (h_psf) {
        TRY(rbx_convolve_psf(d_cube, d_cube2, num_spaxels, num_spaxels, W, d_psf, M, N, stream));
        result = d_cube2;
      } else {
        TRY(rbx_convolve_lsf(d_cube, d_cube2, (int64_t)num_spaxels * num_spaxels, W, d_lsf, K, ext, stream));
        result = d_cube2;
      }
    }
  }
  RBX_CUDA_OK(cudaMemcpyAsync(h_cube, result, sizeof(float) * cube_elems, cudaMemcpyDeviceToHost, stream));
  RBX_CUDA_OK(cudaStreamSynchronize(stream));  // the caller reads h_cube right after this returns
  return RBX_OK;
}
}  // namespace

extern "C" int rbx_pipeline_host(const rbx_plan *plan, const float *h_coords, const float *h_velocity,
                                 const float *h_mass, const float *h_metallicity, const float *h_age, int64_t n,
                                 const float *h_edges, int n_edges, int num_spaxels, int apply_filter,
                                 const float *h_psf, int M, int N, const float *h_lsf, int K, int ext,
                                 float *h_cube, void *stream_) {
  HostParticles hp;
  hp.coords = h_coords; hp.velocity = h_velocity;
  hp.mass = h_mass; hp.metallicity = h_metallicity; hp.age = h_age;
  return pipeline_host_impl(plan, hp, n, h_edges, n_edges, num_spaxels, apply_filter, h_psf, M, N, h_lsf, K, ext, h_cube,
                            stream_);
}

extern "C" int rbx_pipeline_host_packed(const rbx_plan *plan, const float *h_x, const float *h_y, const float *h_vlos,
                                        const float *h_mass, const float *h_metallicity, const float *h_age, int64_t n,
                                        const float *h_edges, int n_edges, int num_spaxels, int apply_filter,
                                        const float *h_psf, int M, int N, const float *h_lsf, int K, int ext,
                                        float *h_cube, void *stream_) {
  HostParticles hp;
  hp.x = h_x; hp.y = h_y; hp.vlos = h_vlos; hp.packed = true;
  hp.mass = h_mass; hp.metallicity = h_metallicity; hp.age = h_age;
  return pipeline_host_impl(plan, hp, n, h_edges, n_edges, num_spaxels, apply_filter, h_psf, M, N, h_lsf, K, ext, h_cube,
                            stream_);
}

// A rank's host shard -> its partial cube on the device (the first half of rbx_pipeline_host, for the multi-GPU
// path: the exchange rbx_reduce_cube / rbx_reduce_scatter_cube and the PSF + LSF follow on the device).  The shard
// is copied in contiguous ranges on a second stream while the kernels of the previous range run, like
// rbx_pipeline_host does.  nslab > 1: slab-major cube (rbx_assign_build_cube_slabs).  Stream-ordered: returns
// without synchronising; the host arrays must stay valid until `stream` has passed this call.
extern "C" int rbx_build_cube_host(const rbx_plan *plan, const float *h_coords, const float *h_velocity,
                                   const float *h_mass, const float *h_metallicity, const float *h_age, int64_t n,
                                   const float *h_edges, int n_edges, int num_spaxels, int apply_filter, int nslab,
                                   int halo, float *d_cube, void *stream_) {
  RBX_REQUIRE(d_cube, "rbx_build_cube_host: null cube");
  RBX_REQUIRE(nslab >= 1 && halo >= 0, "rbx_build_cube_host: bad slab geometry");
  HostParticles hp;
  hp.coords = h_coords; hp.velocity = h_velocity;
  hp.mass = h_mass; hp.metallicity = h_metallicity; hp.age = h_age;
  return pipeline_host_impl(plan, hp, n, h_edges, n_edges, num_spaxels, apply_filter, nullptr, 0, 0, nullptr, 0, 0, nullptr,
                            stream_, d_cube, nslab, nslab > 1 ? halo : 0);
}
