/*
 * CPU oracle (C restatement) for the rubix particle -> IFU datacube path.
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py.  Nothing under rubix_b200/ may use it.
 *
 * It follows the reference (AstroAI-Lab/rubix @ dbb4487) function by function, keeping the
 * reference's per-particle dataflow (a full 842-bin SSP spectrum and a full W-bin resampled spectrum
 * are materialised for every particle, then scatter-added), so that it can be timed as "the
 * reference's algorithm on host cores".  Compiled twice: REAL=float (rbxo32_*) mirrors the JAX
 * float32 arithmetic, REAL=double (rbxo64_*) evaluates the same formulas in double on the same
 * float32 inputs.  Validated against oracle/rubix_oracle.py (numpy) in tests/test_oracle_c.py and against the
 * cube the reference's own source files give (tests/golden/ref_numpy_cube.npz, tests/test_oracle_vs_reference_source.py).
 *
 * interp2d is interpax.interp2d (PyPI, unpinned by the reference): off-node values are
 * "parity unpinned" -- see the header of rubix_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef REAL
#define REAL float
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#ifndef PREFIX
#define PREFIX rbxo32_
#endif
#define FN(name) CAT(PREFIX, name)

/* searchsorted(a, v, side='right'): number of elements <= v */
static int ss_right_f32(const float *a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
static int ss_right_real(const REAL *a, int n, REAL v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* rubix/telescope/utils.py:138-151 + :170-174 + core/telescope.py:155-174 */
void FN(spaxel_assign)(const float *coords, int64_t n, const float *edges, int n_edges,
                       int32_t *idx, uint8_t *mask) {
  int nb = n_edges - 1;
  float lo = edges[0], hi = edges[0];
  for (int i = 1; i < n_edges; ++i) { if (edges[i] < lo) lo = edges[i]; if (edges[i] > hi) hi = edges[i]; }
  for (int64_t p = 0; p < n; ++p) {
    float x = coords[3 * p], y = coords[3 * p + 1];
    int xi = clampi(ss_right_f32(edges, n_edges, x) - 1, 0, nb - 1);
    int yi = clampi(ss_right_f32(edges, n_edges, y) - 1, 0, nb - 1);
    idx[p] = xi + nb * yi;
    if (mask) mask[p] = (x >= lo && x <= hi && y >= lo && y <= hi) ? 1 : 0;
  }
}

/* interpax approx_df(method="cubic") along one axis of a (nz, na, L) table */
static void approx_df(const REAL *x, int nx, const REAL *f, REAL *out, int nz, int na, int L, int axis) {
  int n = axis == 0 ? nz : na;
  (void)nx;
  size_t stride = axis == 0 ? (size_t)na * L : (size_t)L;
  REAL *dxi = (REAL *)malloc(sizeof(REAL) * (n > 1 ? n - 1 : 1));
  for (int i = 0; i + 1 < n; ++i) { REAL dx = x[i + 1] - x[i]; dxi[i] = dx == 0 ? 0 : (REAL)1 / dx; }
  int outer = axis == 0 ? 1 : nz;
  int inner = axis == 0 ? na * L : L;
  for (int o = 0; o < outer; ++o) {
    const REAL *fb = f + (size_t)o * na * L;
    REAL *ob = out + (size_t)o * na * L;
    for (int q = 0; q < inner; ++q) {
      for (int i = 0; i < n; ++i) {
        REAL v;
        if (n == 1) v = 0;
        else if (i == 0) v = dxi[0] * (fb[stride + q] - fb[q]);
        else if (i == n - 1) v = dxi[n - 2] * (fb[(size_t)(n - 1) * stride + q] - fb[(size_t)(n - 2) * stride + q]);
        else {
          REAL a = dxi[i - 1] * (fb[(size_t)i * stride + q] - fb[(size_t)(i - 1) * stride + q]);
          REAL b = dxi[i] * (fb[(size_t)(i + 1) * stride + q] - fb[(size_t)i * stride + q]);
          v = (REAL)0.5 * (a + b);
        }
        ob[(size_t)i * stride + q] = v;
      }
    }
  }
  free(dxi);
}

typedef struct {
  int nz, na, L;
  REAL *zg, *ag, *f, *fx, *fy, *fxy;
} ssp_t;

static void ssp_init(ssp_t *s, const float *zg, int nz, const float *ag, int na, const float *flux, int L, int cubic) {
  s->nz = nz; s->na = na; s->L = L;
  size_t tot = (size_t)nz * na * L;
  s->zg = (REAL *)malloc(sizeof(REAL) * nz);
  s->ag = (REAL *)malloc(sizeof(REAL) * na);
  s->f = (REAL *)malloc(sizeof(REAL) * tot);
  for (int i = 0; i < nz; ++i) s->zg[i] = zg[i];
  for (int i = 0; i < na; ++i) s->ag[i] = ag[i];
  for (size_t i = 0; i < tot; ++i) s->f[i] = flux[i];
  s->fx = s->fy = s->fxy = NULL;
  if (cubic) {
    s->fx = (REAL *)malloc(sizeof(REAL) * tot);
    s->fy = (REAL *)malloc(sizeof(REAL) * tot);
    s->fxy = (REAL *)malloc(sizeof(REAL) * tot);
    approx_df(s->zg, nz, s->f, s->fx, nz, na, L, 0);
    approx_df(s->ag, na, s->f, s->fy, nz, na, L, 1);
    approx_df(s->ag, na, s->fx, s->fxy, nz, na, L, 1);
  }
}
static void ssp_free(ssp_t *s) { free(s->zg); free(s->ag); free(s->f); free(s->fx); free(s->fy); free(s->fxy); }

/* interpax.interp2d(xq, yq, x, y, f, method, extrap=0) for one query point -> out[L]
 * (call site rubix/spectra/ssp/grid.py:113-120, rubix/core/ifu.py:107-110) */
static void interp2d_one(const ssp_t *s, REAL xq, REAL yq, int cubic, REAL *out) {
  int L = s->L, nz = s->nz, na = s->na;
  if (xq < s->zg[0] || xq > s->zg[nz - 1] || yq < s->ag[0] || yq > s->ag[na - 1] || xq != xq || yq != yq) {
    for (int l = 0; l < L; ++l) out[l] = 0;
    return;
  }
  int i = clampi(ss_right_real(s->zg, nz, xq), 1, nz - 1);
  int j = clampi(ss_right_real(s->ag, na, yq), 1, na - 1);
  REAL x0 = s->zg[i - 1], x1 = s->zg[i], y0 = s->ag[j - 1], y1 = s->ag[j];
  REAL dx = x1 - x0, dy = y1 - y0;
  REAL dxi = dx == 0 ? 0 : (REAL)1 / dx, dyi = dy == 0 ? 0 : (REAL)1 / dy;
  const REAL *f00 = s->f + ((size_t)(i - 1) * na + (j - 1)) * L;
  const REAL *f01 = s->f + ((size_t)(i - 1) * na + j) * L;
  const REAL *f10 = s->f + ((size_t)i * na + (j - 1)) * L;
  const REAL *f11 = s->f + ((size_t)i * na + j) * L;
  if (!cubic) {
    REAL tx0 = x1 - xq, tx1 = xq - x0, ty0 = y1 - yq, ty1 = yq - y0, sc = dxi * dyi;
    for (int l = 0; l < L; ++l) {
      REAL acc = f00[l] * tx0 * ty0 + f01[l] * tx0 * ty1 + f10[l] * tx1 * ty0 + f11[l] * tx1 * ty1;
      out[l] = sc * acc;
    }
    return;
  }
  /* bicubic Hermite patch (== interpax A_BICUBIC form; checked in tests/test_oracle_golden.py) */
  REAL tx = (xq - x0) * dxi, ty = (yq - y0) * dyi;
  REAL tx2 = tx * tx, tx3 = tx2 * tx, ty2 = ty * ty, ty3 = ty2 * ty;
  REAL hx[4] = {2 * tx3 - 3 * tx2 + 1, -2 * tx3 + 3 * tx2, tx3 - 2 * tx2 + tx, tx3 - tx2};
  REAL hy[4] = {2 * ty3 - 3 * ty2 + 1, -2 * ty3 + 3 * ty2, ty3 - 2 * ty2 + ty, ty3 - ty2};
  const REAL *tabs[4] = {s->f, s->fx, s->fy, s->fxy};
  REAL w[16];
  const REAL *rows[16];
  int k = 0;
  for (int t = 0; t < 4; ++t) {
    int sx = (t & 1) ? 2 : 0, sy = (t & 2) ? 2 : 0;
    REAL scale = ((t & 1) ? dx : 1) * ((t & 2) ? dy : 1);
    for (int jj = 0; jj < 2; ++jj)
      for (int ii = 0; ii < 2; ++ii) {
        w[k] = hx[sx + ii] * hy[sy + jj] * scale;
        rows[k] = tabs[t] + ((size_t)(i - 1 + ii) * na + (j - 1 + jj)) * L;
        ++k;
      }
  }
  for (int l = 0; l < L; ++l) {
    REAL acc = 0;
    for (int q = 0; q < 16; ++q) acc += w[q] * rows[q][l];
    out[l] = acc;
  }
}

/* jnp.interp + flux conservation, rubix/spectra/ifu.py:241-260 */
static void resample_one(const REAL *s, const REAL *lam, int L, const REAL *t, const REAL *dt, int W,
                         REAL tmin, REAL tmax, REAL *out) {
  REAL total = 0;
  for (int l = 1; l < L; ++l) {
    if (lam[l] >= tmin && lam[l] <= tmax) total += s[l] * (lam[l] - lam[l - 1]);
  }
  REAL newtot = 0;
  int i = 1; /* targets are increasing: march instead of a fresh binary search (same result) */
  const REAL eps = sizeof(REAL) == 4 ? (REAL)1.4210855e-14f : (REAL)4.930380657631324e-32;
  for (int w = 0; w < W; ++w) {
    REAL x = t[w];
    while (i < L - 1 && lam[i] <= x) ++i;
    /* i == clip(searchsorted(lam, x, 'right'), 1, L-1) */
    REAL df = s[i] - s[i - 1], dx = lam[i] - lam[i - 1], delta = x - lam[i - 1];
    REAL f = (fabs((double)dx) <= (double)eps) ? s[i - 1] : s[i - 1] + (delta / dx) * df;
    if (x < lam[0]) f = s[0];
    if (x > lam[L - 1]) f = s[L - 1];
    out[w] = f;
    newtot += f * dt[w];
  }
  REAL scale = total / newtot;
  if (scale != scale) scale = 0;
  else if (isinf((double)scale)) scale = scale > 0 ? (sizeof(REAL) == 4 ? (REAL)3.4028235e38f : (REAL)1.7976931348623157e308)
                                                    : (sizeof(REAL) == 4 ? (REAL)-3.4028235e38f : (REAL)-1.7976931348623157e308);
  for (int w = 0; w < W; ++w) out[w] *= scale;
}

/*
 * filter_particles -> spaxel_assignment -> calculate_spectra -> scale_spectrum_by_mass ->
 * doppler_shift_and_resampling -> calculate_datacube   (rubix/config/pipeline_config.yml:1-45)
 * cube is (S*S, W) in REAL, zeroed here.  Particle order is kept inside each thread's contiguous
 * range (the reference's segment_sum adds in particle order); per-thread partial cubes are then
 * added in thread order.  Returns 0.
 */
int FN(particles_to_cube)(const float *coords, const float *velocity, const float *mass,
                          const float *metallicity, const float *age, int64_t n,
                          const float *edges, int n_edges, int num_spaxels,
                          const float *ssp_z, int nz, const float *ssp_age, int na,
                          const float *ssp_wave, int L, const float *ssp_flux,
                          const float *target_wave, int W, double redshift, int cubic,
                          int vel_component, int apply_filter, int n_threads, REAL *cube) {
  ssp_t ssp;
  ssp_init(&ssp, ssp_z, nz, ssp_age, na, ssp_flux, L, cubic);
  int nb = n_edges - 1;
  float elo = edges[0], ehi = edges[0];
  for (int i = 1; i < n_edges; ++i) { if (edges[i] < elo) elo = edges[i]; if (edges[i] > ehi) ehi = edges[i]; }
  REAL *lamz = (REAL *)malloc(sizeof(REAL) * L);
  REAL *t = (REAL *)malloc(sizeof(REAL) * W);
  REAL *dt = (REAL *)malloc(sizeof(REAL) * W);
  REAL onepz = (REAL)(1.0 + redshift);
  for (int l = 0; l < L; ++l) lamz[l] = onepz * (REAL)ssp_wave[l];
  REAL tmin = target_wave[0], tmax = target_wave[0];
  for (int w = 0; w < W; ++w) {
    t[w] = target_wave[w];
    dt[w] = w == 0 ? 0 : (REAL)target_wave[w] - (REAL)target_wave[w - 1];
    if (t[w] < tmin) tmin = t[w];
    if (t[w] > tmax) tmax = t[w];
  }
  size_t csz = (size_t)num_spaxels * num_spaxels * W;
  memset(cube, 0, sizeof(REAL) * csz);
  if (n_threads < 1) n_threads = 1;
  REAL **partial = (REAL **)calloc(n_threads, sizeof(REAL *));
  partial[0] = cube;
  for (int th = 1; th < n_threads; ++th) partial[th] = (REAL *)calloc(csz, sizeof(REAL));
  const REAL c_light = (REAL)299792.458;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads)
#endif
  {
#ifdef _OPENMP
    int th = omp_get_thread_num();
#else
    int th = 0;
#endif
    if (th < n_threads) {
      int64_t per = (n + n_threads - 1) / n_threads;
      int64_t p0 = th * per, p1 = p0 + per > n ? n : p0 + per;
      REAL *spec = (REAL *)malloc(sizeof(REAL) * L);
      REAL *lam = (REAL *)malloc(sizeof(REAL) * L);
      REAL *res = (REAL *)malloc(sizeof(REAL) * W);
      REAL *my = partial[th];
      for (int64_t p = p0; p < p1; ++p) {
        float x = coords[3 * p], y = coords[3 * p + 1];
        int xi = clampi(ss_right_f32(edges, n_edges, x) - 1, 0, nb - 1);
        int yi = clampi(ss_right_f32(edges, n_edges, y) - 1, 0, nb - 1);
        int idx = xi + nb * yi;
        REAL m = mass[p], zq = metallicity[p], aq = age[p];
        if (apply_filter && !(x >= elo && x <= ehi && y >= elo && y <= ehi)) { m = 0; zq = 0; aq = 0; }
        interp2d_one(&ssp, zq, aq, cubic, spec);
        for (int l = 0; l < L; ++l) spec[l] = spec[l] * m;
        REAL v = velocity[3 * p + vel_component];
        REAL d = sizeof(REAL) == 4 ? (REAL)expf((float)(v / c_light)) : (REAL)exp((double)(v / c_light));
        for (int l = 0; l < L; ++l) lam[l] = lamz[l] * d;
        resample_one(spec, lam, L, t, dt, W, tmin, tmax, res);
        if (idx >= 0 && idx < num_spaxels * num_spaxels) {
          REAL *row = my + (size_t)idx * W;
          for (int w = 0; w < W; ++w) row[w] += res[w];
        }
      }
      free(spec); free(lam); free(res);
    }
  }
  for (int th = 1; th < n_threads; ++th) {
    for (size_t q = 0; q < csz; ++q) cube[q] += partial[th][q];
    free(partial[th]);
  }
  free(partial); free(lamz); free(t); free(dt);
  ssp_free(&ssp);
  return 0;
}

/* rubix/telescope/psf/psf.py:56-57: zero-padded true 2-D convolution per wavelength slice,
 * out[i,j,w] = sum_mn K[m,n] in[i-m+(M-1)/2, j-n+(N-1)/2, w] */
void FN(apply_psf)(const REAL *in, REAL *out, int H, int Wd, int Lw, const REAL *K, int M, int N) {
  int cm = (M - 1) / 2, cn = (N - 1) / 2;
#ifdef _OPENMP
#pragma omp parallel for collapse(2)
#endif
  for (int i = 0; i < H; ++i)
    for (int j = 0; j < Wd; ++j) {
      REAL *o = out + ((size_t)i * Wd + j) * Lw;
      for (int w = 0; w < Lw; ++w) o[w] = 0;
      for (int m = 0; m < M; ++m) {
        int ii = i - m + cm;
        if (ii < 0 || ii >= H) continue;
        for (int nn = 0; nn < N; ++nn) {
          int jj = j - nn + cn;
          if (jj < 0 || jj >= Wd) continue;
          const REAL *src = in + ((size_t)ii * Wd + jj) * Lw;
          REAL k = K[m * N + nn];
          for (int w = 0; w < Lw; ++w) o[w] += k * src[w];
        }
      }
    }
}

/* rubix/telescope/lsf/lsf.py:59-65: convolve(full)[:, ext : W+K-1-ext], K taps, ext = (K-1)/2 */
void FN(apply_lsf)(const REAL *in, REAL *out, int64_t rows, int Lw, const REAL *k, int K, int ext) {
#ifdef _OPENMP
#pragma omp parallel for
#endif
  for (int64_t r = 0; r < rows; ++r) {
    const REAL *s = in + r * Lw;
    REAL *o = out + r * Lw;
    for (int w = 0; w < Lw; ++w) {
      REAL acc = 0;
      for (int m = 0; m < K; ++m) {
        int q = w + ext - m;
        if (q >= 0 && q < Lw) acc += k[m] * s[q];
      }
      o[w] = acc;
    }
  }
}
