/* A C host of the whole path: no Python, no torch, no CUDA headers -- only include/rubix_b200.h and host buffers.
 *
 *   c_host_pipeline <dir>
 *
 * reads raw little-endian arrays from <dir> (written by tests/helpers.py: write_c_host_inputs):
 *   dims.i32        nz na L W n n_edges S M N K method
 *   metallicity.f32 age.f32 wavelength.f32 flux.f32 wave.f32        the SSP template and the telescope grid
 *   coords.f32 velocity.f32 mass.f32 met.f32 age_p.f32 edges.f32    particles ((n, 3) row-major) and spatial bin edges
 *   psf.f32 lsf.f32                                                 PSF (M, N) and LSF (K,) taps
 * runs rbx_plan_create + rbx_pipeline_host (filter -> spaxel assignment -> fused cube -> PSF -> LSF, copies included)
 * and writes <dir>/cube.f32 (S, S, W).  Exit status 0 on success; the library's error text goes to stderr otherwise. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rubix_b200.h"

static void *slurp(const char *dir, const char *name, size_t bytes) {
  char path[4096];
  FILE *f;
  void *buf = malloc(bytes ? bytes : 1);
  snprintf(path, sizeof path, "%s/%s", dir, name);
  f = fopen(path, "rb");
  if (!f || !buf || fread(buf, 1, bytes, f) != bytes) {
    fprintf(stderr, "c_host_pipeline: cannot read %zu bytes from %s\n", bytes, path);
    exit(2);
  }
  fclose(f);
  return buf;
}

int main(int argc, char **argv) {
  const char *dir = argc > 1 ? argv[1] : ".";
  int32_t *d = (int32_t *)slurp(dir, "dims.i32", 11 * sizeof(int32_t));
  const int nz = d[0], na = d[1], L = d[2], W = d[3], n = d[4], n_edges = d[5], S = d[6], M = d[7], N = d[8], K = d[9],
            method = d[10];
  const size_t f = sizeof(float);
  float *met_grid = (float *)slurp(dir, "metallicity.f32", f * nz), *age_grid = (float *)slurp(dir, "age.f32", f * na);
  float *lam = (float *)slurp(dir, "wavelength.f32", f * L);
  float *flux = (float *)slurp(dir, "flux.f32", f * (size_t)nz * na * L), *wave = (float *)slurp(dir, "wave.f32", f * W);
  float *coords = (float *)slurp(dir, "coords.f32", f * 3 * (size_t)n);
  float *vel = (float *)slurp(dir, "velocity.f32", f * 3 * (size_t)n);
  float *mass = (float *)slurp(dir, "mass.f32", f * n), *met = (float *)slurp(dir, "met.f32", f * n);
  float *age = (float *)slurp(dir, "age_p.f32", f * n), *edges = (float *)slurp(dir, "edges.f32", f * n_edges);
  float *psf = (float *)slurp(dir, "psf.f32", f * (size_t)M * N), *lsf = (float *)slurp(dir, "lsf.f32", f * K);
  float *cube = (float *)malloc(f * (size_t)S * S * W);
  rbx_plan *plan = NULL;
  char path[4096];
  FILE *out;
  int rc;

  rc = rbx_plan_create(&plan, met_grid, nz, age_grid, na, lam, L, flux, wave, W, 0.1, method, /* z direction */ 2, NULL);
  if (rc != RBX_OK) { fprintf(stderr, "rbx_plan_create: %d %s\n", rc, rbx_last_error()); return 1; }
  rc = rbx_pipeline_host(plan, coords, vel, mass, met, age, n, edges, n_edges, S, /* apply_filter */ 1, psf, M, N, lsf, K,
                         (K - 1) / 2, cube, NULL);
  if (rc != RBX_OK) { fprintf(stderr, "rbx_pipeline_host: %d %s\n", rc, rbx_last_error()); return 1; }
  snprintf(path, sizeof path, "%s/cube.f32", dir);
  out = fopen(path, "wb");
  if (!out || fwrite(cube, f, (size_t)S * S * W, out) != (size_t)S * S * W) { fprintf(stderr, "cannot write %s\n", path); return 2; }
  fclose(out);
  printf("c_host_pipeline: %d particles -> %d x %d x %d cube, %lld kernels launched\n", n, S, S, W,
         (long long)rbx_launch_count());
  rbx_plan_destroy(plan);
  return 0;
}
