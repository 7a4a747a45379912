/*
 * rubix_b200 -- C ABI of the B200-native particle -> IFU-datacube path.
 *
 * This is the drop-in boundary for the hot path of AstroAI-Lab/rubix (reference @ dbb4487; all
 * file:line citations below are relative to the reference tree).  rubix itself is pure Python/JAX
 * and has no FFI; the entry points below are what a `jax.ffi.ffi_call` (or ctypes / cffi) binding
 * for this path binds -- see INTEGRATION.md for the reference-side stubs.
 *
 * Conventions
 *   - Every function returns RBX_OK (0) or a negative rbx_status; rbx_last_error() gives the text
 *     (thread-local).
 *   - Pointers named d_* are DEVICE pointers owned by the caller (JAX / torch allocator); h_* are
 *     HOST pointers.  Outputs are pre-allocated by the caller.  The only hidden allocations are
 *     inside rbx_plan_create (per-config tables, a few MB) and the rbx_*_host convenience calls.
 *   - All floating-point data is float32 and all indices are int32, like the reference (JAX x64 off:
 *     rubix/spectra/ssp/grid.py:327, rubix/core/data.py:545, rubix/telescope/utils.py:151).
 *   - Launches are asynchronous on the given stream; nothing here calls cudaDeviceSynchronize.
 *     Functions are re-entrant per (plan, stream, workspace).
 *   - `stream` is a cudaStream_t passed as void* so that this header needs no CUDA include.
 */
#ifndef RUBIX_B200_H
#define RUBIX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  RBX_OK = 0,
  RBX_ERR_INVALID_ARGUMENT = -1,
  RBX_ERR_CUDA = -2,
  RBX_ERR_WORKSPACE_TOO_SMALL = -3,
  RBX_ERR_UNSUPPORTED = -4, /* configuration outside the fused kernel's limits: use the stage calls */
  RBX_ERR_NO_DEVICE = -5,
  RBX_ERR_NCCL = -6
} rbx_status;

/* rubix/core/ssp.py:57-62: `ssp.method`; rubix's default when the key is absent is "cubic". */
typedef enum { RBX_METHOD_LINEAR = 0, RBX_METHOD_CUBIC = 1 } rbx_method;

typedef struct rbx_plan rbx_plan; /* opaque, device-resident per-configuration tables */

const char *rbx_last_error(void);
int rbx_version(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
int64_t rbx_launch_count(void);

/* Tuning / test switches: "psub", "sort_bits", "fused_force_lut", "fused_force_cas", "fused_impl" (1 = group kernel),
 * "fused_chs", "fused_no_skew", "fused_warps", "prep_blocks", "small_shift", "tail_shift", "host_chunks",
 * "march_no_bulk", "sort_impl" (1 = cub), "fused_variant", "host_ratio" (percent).  value < 0 restores the library's own choice.  The
 * environment (RBX_<NAME>) seeds them ONCE when the library is loaded; no launch path calls getenv(). */
int rbx_set_option(const char *name, int64_t value);
int rbx_get_option(const char *name, int64_t *value);

/* Optional CUDA-event timing of the dominant kernel (fused_cube_kernel) on its launch stream, for
 * bench.py's roofline line.  Not thread-safe; off by default. */
int rbx_profile_enable(int on);
int rbx_profile_fused(double *mean_ms, int64_t *launches, int reset);

/* ---------------------------------------------------------------------------------------------
 * Plan: everything that depends only on the configuration (SSP template, telescope wavelength
 * grid, redshift, interpolation method, Doppler direction).
 * Replaces the host-side setup in get_lookup_interpolation (rubix/core/ssp.py:38-65),
 * SSPGrid.get_lookup_interpolation (rubix/spectra/ssp/grid.py:61-124: binds x=metallicity, y=age,
 * f=flux, extrap=0) and get_doppler_shift_and_resampling's closure state
 * (rubix/core/ifu.py:248-266: (1+z)*ssp.wavelength, telescope.wave_seq, velocity direction).
 *   h_flux is (nz, na, L) C-order; h_target_wave is the telescope wave_seq (W,), increasing.
 *   vel_component: 0/1/2 for rubix_config ifu.doppler.velocity_direction "x"/"y"/"z".
 * ------------------------------------------------------------------------------------------- */
int rbx_plan_create(rbx_plan **plan, const float *h_metallicity, int nz, const float *h_age, int na,
                    const float *h_wavelength, int L, const float *h_flux,
                    const float *h_target_wave, int W, double redshift, int method,
                    int vel_component, void *stream);
int rbx_plan_destroy(rbx_plan *plan);
int rbx_plan_dims(const rbx_plan *plan, int *nz, int *na, int *L, int *W);

/* ---------------------------------------------------------------------------------------------
 * a0  square_spaxel_assignment (rubix/telescope/utils.py:138-151) and
 *     mask_particles_outside_aperture (:170-174).  d_coords is (n, 3) row-major; d_edges (n_edges,).
 *     d_pixel: int32 (n,), x + nbins*y, bit-exact with the reference.  d_mask: uint8 (n,) or NULL.
 * ------------------------------------------------------------------------------------------- */
int rbx_spaxel_assign(const float *d_coords, int64_t n, const float *d_edges, int n_edges,
                      int32_t *d_pixel, uint8_t *d_mask, void *stream);
/* get_filter_particles (rubix/core/telescope.py:155-174): where(mask, x, 0) on mass, metallicity,
 * age in place (any may be NULL); coords and velocity are left alone, as in the reference. */
int rbx_filter_particles(const float *d_coords, int64_t n, const float *d_edges, int n_edges,
                         float *d_mass, float *d_metallicity, float *d_age, uint8_t *d_mask,
                         void *stream);

/* filter_particles + spaxel_assignment in one pass over coords (rubix/core/telescope.py:155-174 followed
 * by rubix/telescope/utils.py:138-151).  With d_mass / d_metallicity / d_age given they are zeroed in
 * place outside the aperture like rbx_filter_particles; with all three NULL nothing is modified and
 * particles outside the aperture get pixel = -1 instead, which rbx_build_cube / rbx_segment_sum drop --
 * the same cube, because a zero-mass particle contributes exactly 0. */
int rbx_filter_and_assign(const float *d_coords, int64_t n, const float *d_edges, int n_edges, float *d_mass,
                          float *d_metallicity, float *d_age, int32_t *d_pixel, uint8_t *d_mask, void *stream);

/* Stable device radix sort of particles by spaxel id (SURVEY 8b minimum set; the grouping step that stands in for
 * the scatter of jax.ops.segment_sum, rubix/spectra/ifu.py:286-287) -- the sort the cube build runs internally
 * (csrc/sort.cu: one kernel per 8-bit pass, decoupled look-back), for a plain id array.
 *   d_pixel   (n,) int32 from rbx_spaxel_assign / rbx_filter_and_assign.  Ids outside [0, num_segments) -- the ones
 *             segment_sum drops, -1 of rbx_filter_and_assign included -- count as num_segments and sort to the end.
 *   d_order   (n,) int32: the permutation, bit for bit numpy.argsort(clamped ids, kind="stable").
 *   d_sorted  (n,) int32 or NULL: the clamped ids in sorted order (d_pixel[d_order] where that is in range).
 *   d_offsets (num_segments + 1,) int32 or NULL: d_offsets[s] = first sorted position with id >= s, so segment s is
 *             [d_offsets[s], d_offsets[s+1]) and the dropped particles are [d_offsets[num_segments], n).
 * Integer work, bit-exact and bit-reproducible.  n <= 2^30. */
size_t rbx_sort_by_spaxel_workspace_bytes(int64_t n, int num_segments);
int rbx_sort_by_spaxel(const int32_t *d_pixel, int64_t n, int num_segments, int32_t *d_order, int32_t *d_sorted,
                       int32_t *d_offsets, void *d_workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Stage calls: one per reference stage, materialising the same intermediates as the reference.
 * Used by the stepwise (notebook) path and for stage-level parity.
 * ------------------------------------------------------------------------------------------- */
/* a1 calculate_spectra (rubix/core/ifu.py:95-118): d_spectra (n, L) = interp2d(Z, age). */
int rbx_ssp_lookup(const rbx_plan *plan, const float *d_metallicity, const float *d_age, int64_t n,
                   float *d_spectra, void *stream);
/* a2 scale_spectrum_by_mass (rubix/core/ifu.py:152-154): d_out[p,l] = d_spectra[p,l] * d_mass[p]
 * (d_out may alias d_spectra). */
int rbx_scale_by_mass(const float *d_spectra, const float *d_mass, int64_t n, int L, float *d_out,
                      void *stream);
/* a3+a4 doppler_shift_and_resampling (rubix/core/ifu.py:269-293, rubix/spectra/ifu.py:190,241-260):
 * d_velocity (n,3); d_spectra (n, L) -> d_out (n, W). */
int rbx_doppler_resample(const rbx_plan *plan, const float *d_spectra, const float *d_velocity,
                         int64_t n, float *d_out, void *stream);
/* a5 calculate_cube (rubix/spectra/ifu.py:286-287): d_cube (num_segments, W) += segment_sum of
 * d_spectra (n, W) by d_pixel; indices outside [0, num_segments) are dropped.  The caller zeroes
 * d_cube (or passes zero_first = 1). */
int rbx_segment_sum(const float *d_spectra, const int32_t *d_pixel, int64_t n, int W,
                    int num_segments, float *d_cube, int zero_first, void *stream);

/* a5, bit-reproducible: d_order / d_offsets from rbx_sort_by_spaxel; d_cube (num_segments, W) is OVERWRITTEN with
 * the float32 sums formed one particle after the other in particle order inside every segment -- the sum
 * jax.ops.segment_sum forms on the CPU backend, so the result is bit-identical to a sequential
 * float32 scatter-add of the same spectra.  No atomics; empty segments give 0 (d_spectra / d_order may be NULL when
 * every segment is empty). */
int rbx_segment_sum_sorted(const float *d_spectra, const int32_t *d_order, const int32_t *d_offsets, int W,
                           int num_segments, float *d_cube, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Fused a1..a5: particles -> cube without materialising (n, L) or (n, W).
 *   d_velocity (n,3), d_mass/d_metallicity/d_age (n,), d_pixel (n,) from rbx_spaxel_assign.
 *   d_cube (num_spaxels^2, W) is overwritten.  Workspace: rbx_build_cube_workspace_bytes().
 * Particles with mass == 0, (Z, age) outside the SSP grid or pixel outside [0, S^2) contribute
 * exactly 0, as in the reference (extrap=0 -> zero spectrum -> nan_to_num(0/0) = 0; segment_sum
 * drops out-of-range ids).
 * ------------------------------------------------------------------------------------------- */
size_t rbx_build_cube_workspace_bytes(const rbx_plan *plan, int64_t n, int num_spaxels);
int rbx_build_cube(const rbx_plan *plan, const float *d_velocity, const float *d_mass,
                   const float *d_metallicity, const float *d_age, const int32_t *d_pixel, int64_t n,
                   int num_spaxels, float *d_cube, void *d_workspace, size_t workspace_bytes,
                   void *stream);

/* a0 + a1..a5 in one call: spaxel assignment (and, with apply_filter, the aperture filter: particles outside
 * get pixel -1, the same cube as zeroing their mass) happens inside the first kernel of the cube build, so the
 * particle arrays are read once.  d_pixel_out (n,) may be NULL.  Same workspace as rbx_build_cube. */
int rbx_assign_build_cube(const rbx_plan *plan, const float *d_coords, const float *d_edges, int n_edges,
                          int apply_filter, const float *d_velocity, const float *d_mass,
                          const float *d_metallicity, const float *d_age, int64_t n, int num_spaxels,
                          int32_t *d_pixel_out, float *d_cube, void *d_workspace, size_t workspace_bytes,
                          void *stream);

/* The same with structure-of-arrays particles: x, y and the line-of-sight velocity (the component
 * ifu.doppler.velocity_direction selects) as arrays of their own -- the 24 bytes per particle this path reads
 * instead of the 40 of the (n, 3) arrays. */
int rbx_assign_build_cube_packed(const rbx_plan *plan, const float *d_x, const float *d_y, const float *d_edges,
                                 int n_edges, int apply_filter, const float *d_vlos, const float *d_mass,
                                 const float *d_metallicity, const float *d_age, int64_t n, int num_spaxels,
                                 int32_t *d_pixel_out, float *d_cube, void *d_workspace, size_t workspace_bytes,
                                 void *stream);

/* What the last build on this workspace did (synchronises `stream`).  *h_error: 0 ok; 1 work-item tables too small,
 * 2 SSP knot window wider than the kernels hold, 3 Doppler range too wide for the chunk geometry -- for != 0 the
 * cube was filled with NaN rather than left silently wrong.  *h_impl: 0 = fused_cube_warp_kernel ran, 1 = the
 * general fused_cube_kernel (chosen on the device from the knot window and Doppler range actually present). */
int rbx_build_cube_status(const rbx_plan *plan, int64_t n, int num_spaxels, const void *d_workspace, int *h_error,
                          int *h_impl, void *stream);

/* Which shared-memory cell layout the warp kernel of the last build on this workspace used (synchronises `stream`):
 * *h_transposed = 1 for the transposed block layout (a cell's bank is its block: DESIGN.md section 5), 0 for cells in
 * channel order or when the general kernel ran.  Results agree to rounding; the choice is made on the device from the
 * Doppler range present (option fused_tr = 0 switches the transposed layout off). */
int rbx_build_cube_cell_layout(const rbx_plan *plan, int64_t n, int num_spaxels, const void *d_workspace,
                               int *h_transposed, void *stream);

/* Slab-major partial cube for the multi-GPU exchange: the cube is stored as nslab wavelength slabs of
 * wslab = ceil(W / nslab) channels, slab r an (S*S, ws) block with ws = wslab + 2 halo: its own channels plus `halo`
 * channels of each neighbour (zeros beyond [0, W): the zero padding of the reference's 'same' convolutions).
 * d_slabs holds nslab * S*S * ws floats.  One rbx_reduce_scatter_cube then leaves rank r with its summed slab
 * including the LSF halo, ready for rbx_psf_lsf_taps_pitched (rubix/core/ifu.py:324-333 -> core/psf.py, core/lsf.py). */
int rbx_slab_geometry(int W, int nslab, int halo, int *wslab, int *ws);
int rbx_assign_build_cube_slabs(const rbx_plan *plan, const float *d_coords, const float *d_edges, int n_edges,
                                int apply_filter, const float *d_velocity, const float *d_mass,
                                const float *d_metallicity, const float *d_age, int64_t n, int num_spaxels, int nslab,
                                int halo, float *d_slabs, void *d_workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * The exchange step: rubix sums the per-device cubes with jnp.sum(ifu_cubes, axis=0) after its pmap
 * (rubix/core/ifu.py:324-333).  Here: one process per GPU, NCCL over NVLink, bound at run time (dlopen of
 * libnccl.so.2 -- the copy the host process already loaded, if any).  rbx_comm_unique_id on one rank, its 128
 * bytes handed to every rank by the host (any out-of-band channel), rbx_comm_init on every rank with its CUDA
 * device current.  Collectives are asynchronous on `stream`; float32 sums.
 *   rbx_reduce_cube          d_recv (root only) = sum over ranks of d_send[count]
 *   rbx_allreduce_cube       every rank gets the sum
 *   rbx_reduce_scatter_cube  d_send holds world * recv_count floats; rank r receives the sum of block r
 *   rbx_allgather_cube       d_recv[world * send_count] = the blocks of all ranks in rank order
 * ------------------------------------------------------------------------------------------- */
#define RBX_COMM_ID_BYTES 128
typedef struct rbx_comm rbx_comm;
int rbx_comm_unique_id(void *id);
int rbx_comm_init(rbx_comm **comm, const void *id, int rank, int world);
int rbx_comm_destroy(rbx_comm *comm);
int rbx_comm_info(const rbx_comm *comm, int *rank, int *world, int *nccl_version);
int rbx_reduce_cube(rbx_comm *comm, const float *d_send, float *d_recv, int64_t count, int root, void *stream);
int rbx_allreduce_cube(rbx_comm *comm, const float *d_send, float *d_recv, int64_t count, void *stream);
int rbx_reduce_scatter_cube(rbx_comm *comm, const float *d_send, float *d_recv, int64_t recv_count, void *stream);
int rbx_allgather_cube(rbx_comm *comm, const float *d_send, float *d_recv, int64_t send_count, void *stream);
int rbx_allreduce_f64(rbx_comm *comm, const double *d_send, double *d_recv, int64_t count, void *stream);

/* ---------------------------------------------------------------------------------------------
 * a6 apply_psf (rubix/telescope/psf/psf.py:56-57): per wavelength slice
 *    convolve2d(slice, kernel, mode="same"); d_kernel is (M, N) row-major on the device.
 * a7 apply_lsf (rubix/telescope/lsf/lsf.py:59-65,96-105): convolve(full) along lambda then the
 *    slice [ext : W+K-1-ext]; d_kernel (K,) on the device; ext = extend_factor (12).
 * rbx_psf_lsf does both in one pass (d_out must not alias d_in).
 * d_in / d_out are (ny, nx, W), lambda fastest.
 * ------------------------------------------------------------------------------------------- */
int rbx_convolve_psf(const float *d_in, float *d_out, int ny, int nx, int W, const float *d_kernel,
                     int M, int N, void *stream);
int rbx_convolve_lsf(const float *d_in, float *d_out, int64_t rows, int W, const float *d_kernel,
                     int K, int ext, void *stream);
int rbx_psf_lsf(const float *d_in, float *d_out, int ny, int nx, int W, const float *d_psf, int M,
                int N, const float *d_lsf, int K, int ext, void *stream);
/* The same pass with HOST taps (the reference builds both kernels on the host from the config:
 * rubix/telescope/psf/kernels.py:5-31, rubix/telescope/lsf/lsf.py:12-26).  h_psf (M, N) or NULL for no PSF,
 * h_lsf (K = 2 ext + 1) or NULL for no LSF.  Knowing the taps at launch lets the library run an
 * outer-product PSF as two 1-D passes and drop LSF taps below 1e-14 of the largest (they cannot change a
 * float32 sum); returns RBX_ERR_UNSUPPORTED when the taps need the general kernels (rbx_psf_lsf). */
int rbx_psf_lsf_taps(const float *d_in, float *d_out, int ny, int nx, int W, const float *h_psf, int M,
                     int N, const float *h_lsf, int K, int ext, void *stream);
/* The same on a wavelength SLAB of a cube (the multi-GPU PSF / LSF stage shards by wavelength):
 * W channels starting at d_in / d_out, consecutive spaxels in_pitch / out_pitch floats apart (>= W).
 * Channels outside the slab count as zero, so a rank passes its slab plus a halo of `ext` channels taken
 * from the summed cube and keeps the interior. */
int rbx_psf_lsf_taps_pitched(const float *d_in, int in_pitch, float *d_out, int out_pitch, int ny, int nx,
                             int W, const float *h_psf, int M, int N, const float *h_lsf, int K, int ext,
                             void *stream);
/* gaussian_kernel_2d (rubix/telescope/psf/kernels.py:26-31) and _get_kernel
 * (rubix/telescope/lsf/lsf.py:12-26) evaluated in float32 on the device. */
int rbx_gaussian_psf_kernel(int m, int n, float sigma, float *d_kernel, void *stream);
int rbx_gaussian_lsf_kernel(float sigma, float wave_res, int factor, float *d_kernel, void *stream);

/* ---------------------------------------------------------------------------------------------
 * rotate_galaxy, the stage right before the path (rubix/core/rotation.py:76-115 ->
 * rubix/galaxy/alignment.py:233-265): moment-of-inertia tensor of the particles within the half-mass
 * radius (alignment.py:67-125, including its index-0 padding), eigenvectors by ascending eigenvalue
 * (:128-146), then (p @ R) @ E for coordinates and velocities (:149-229).  h_euler is the (3, 3) row-major
 * Euler matrix R_z R_y R_x of the configured angles (alignment.py:164-209), float32 on the HOST.
 * d_rotation (9 floats) receives R; eigenvector signs are normalised (largest component positive) because
 * eigh's signs are backend-dependent in the reference.  d_velocity / d_velocity_out may both be NULL.
 * Outputs may not alias inputs.  Workspace: rbx_rotate_galaxy_workspace_bytes().
 * ------------------------------------------------------------------------------------------- */
size_t rbx_rotate_galaxy_workspace_bytes(void);
int rbx_rotate_galaxy(const float *d_coords, const float *d_velocity, const float *d_mass, int64_t n,
                      float halfmass_radius, const float *h_euler, float *d_coords_out,
                      float *d_velocity_out, float *d_rotation, void *d_workspace, size_t workspace_bytes,
                      void *stream);
/* The same in two steps for a galaxy whose particles are sharded over ranks (the reference computes ONE rotation from
 * all particles before its reshape / pmap split, rubix/core/rotation.py:76-115): rbx_rotate_moments writes this
 * shard's inertia sums, 12 doubles (six second moments, n_inside, n, and x, y, z, mass of the galaxy's particle 0
 * from the rank with first_shard != 0 -- the reference's index-0 padding term); the host sums them over the ranks
 * (rbx_allreduce_f64) and rbx_rotate_apply derives the rotation and rotates the shard.  Same workspace size. */
int rbx_rotate_moments(const float *d_coords, const float *d_mass, int64_t n, float halfmass_radius,
                       int first_shard, double *d_moments, void *d_workspace, size_t workspace_bytes, void *stream);
int rbx_rotate_apply(const float *d_coords, const float *d_velocity, int64_t n, const double *d_moments,
                     const float *h_euler, float *d_coords_out, float *d_velocity_out, float *d_rotation,
                     void *stream);

/* ---------------------------------------------------------------------------------------------
 * apply_noise, the stage right after the path (rubix/core/noise.py:15-78 ->
 * rubix/telescope/noise/noise.py:37-115): flux image, median of the non-zero flux (jnp.median propagates the
 * NaN of a flux-less spaxel -> 0, reproduced), S2N map, cube += cube * N * S2N.  N restates
 * jax.random.normal / uniform(PRNGKey(0)): threefry2x32 with the element index as counter (key0 = key1 = 0
 * for PRNGKey(0)); distribution 0 = "normal", 1 = "uniform".  d_out may alias d_in.
 * rbx_noise_samples exposes the raw sample stream (and its 32-bit words) for tests.
 * ------------------------------------------------------------------------------------------- */
size_t rbx_apply_noise_workspace_bytes(int ny, int nx);
int rbx_apply_noise(const float *d_in, float *d_out, int ny, int nx, int W, float signal_to_noise,
                    int distribution, uint32_t key0, uint32_t key1, void *d_workspace,
                    size_t workspace_bytes, void *stream);
int rbx_noise_samples(float *d_out, uint32_t *d_bits, int64_t n, int distribution, uint32_t key0,
                      uint32_t key1, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Dust extinction, the calc_dusty_ifu variant (rubix/core/dust.py:15-65 -> apply_spaxel_extinction,
 * rubix/spectra/dust/dust_extinction.py:169-358).
 *
 * rbx_dust_av: A_V of every star from the gas cells of its spaxel.  Per cell: 12 + log10(O/H) from
 *   d_gas_metals (n_gas, n_metals; hydrogen = column 0, oxygen = column 4), the dust-to-gas ratio
 *   1 / 10^(a + alpha (8.69 - x)) with h_dust_to_gas = {a_high, alpha_high, a_low, alpha_low, x_transition}
 *   (HOST, Remy-Ruyer 2014 table 1; the high pair applies for x > x_transition), times mass * ext_const /
 *   spaxel_area (ext_const = 3 pi Msun_g / kpc_cm^2 / (0.4 ln10 lambda_V[cm] rho_grain), dust_extinction.py:96-166).
 *   Cells are sorted by (pixel, z) (jnp.lexsort, :240-245), accumulated along z per spaxel (:318) and
 *   interpolated at the star's z with jnp.interp(left="extrapolate") on the reference's table, in which the
 *   cells of all other spaxels sit at z * 1e30 with value 0 (:313-336).  d_av (n_star,) in input order; stars
 *   with a pixel outside [0, n_spaxels) get 0.  d_cell_av_out (n_gas,) may be NULL (per-cell A_V, input order).
 * rbx_apply_extinction: d_out[q, w] = d_spectra[q, w] * 10^(-0.4 * d_axav[w] * d_av[q])
 *   (BaseExtRvModel.extinguish, dust_baseclasses.py:126-164; dust_extinction.py:341-356); d_axav (W,) is the
 *   configuration's A(lambda)/A(V) curve.  d_out may alias d_spectra.
 * rbx_build_cube_dusty: Doppler shift + resampling (rubix/spectra/ifu.py:224-266) of already mass-scaled SSP
 *   spectra d_spectra (n, L), times the extinction factor, summed per spaxel into d_cube (num_spaxels^2, W)
 *   (calculate_cube, rubix/spectra/ifu.py:270-288) -- the three stages doppler_shift_and_resampling ->
 *   calculate_extinction -> calculate_datacube without the (n, W) intermediate.  d_av / d_axav NULL = no dust.
 * ------------------------------------------------------------------------------------------- */
size_t rbx_dust_av_workspace_bytes(int64_t n_gas, int n_spaxels);
int rbx_dust_av(const float *d_gas_coords, const int32_t *d_gas_pixel, const float *d_gas_mass,
                const float *d_gas_metals, int n_metals, int64_t n_gas, const float *d_star_coords,
                const int32_t *d_star_pixel, int64_t n_star, int n_spaxels, const float *h_dust_to_gas,
                float ext_const, float spaxel_area, float *d_av, float *d_cell_av_out, void *d_workspace,
                size_t workspace_bytes, void *stream);
int rbx_apply_extinction(const float *d_spectra, const float *d_av, const float *d_axav, int64_t n, int W,
                         float *d_out, void *stream);
/* The dusty cube through rbx_build_cube (the knot-based kernel): stars are binned by A_V (n_bins bins of width step
 * from av0) and the factor is expanded to third order inside a bin.  rbx_dusty_bins writes every star
 * M = rbx_dusty_moments() times: virtual particle m of star i (arrays of M * n entries, moment-major) carries the weight
 * mass * (A_V - A_bin)^m, the star's velocity / metallicity / age, and the virtual spaxel id
 * (pixel * n_bins + bin) * M + m (-1 for dropped ids).  The caller runs rbx_build_cube ONCE on the M * n virtual
 * particles with ceil(sqrt(nseg * n_bins * M))^2 virtual spaxels, and rbx_dusty_combine folds the rows:
 * cube[s, w] = sum_b 10^(-0.4 axav_w A_b) sum_m (-0.4 ln10 axav_w)^m / m! vcube[(s n_bins + b) M + m, w].
 * Truncation < 2.6e-7 relative when 0.4 ln10 max|axav| step <= 0.1. */
int rbx_dusty_moments(void);
int rbx_dusty_bins(const float *d_av, const int32_t *d_pixel, const float *d_mass, const float *d_velocity,
                   const float *d_metallicity, const float *d_age, int64_t n, int nseg, int n_bins, float av0,
                   float step, int32_t *d_vpixel, float *d_wmass, float *d_vvelocity, float *d_vmetallicity,
                   float *d_vage, void *stream);
int rbx_dusty_combine(const float *d_vcube, int nseg, int n_bins, float av0, float step, const float *d_axav, int W,
                      float *d_cube, void *stream);
size_t rbx_build_cube_dusty_workspace_bytes(int64_t n);
int rbx_build_cube_dusty(const rbx_plan *plan, const float *d_spectra, const float *d_velocity,
                         const int32_t *d_pixel, const float *d_av, const float *d_axav, int64_t n,
                         int num_spaxels, float *d_cube, void *d_workspace, size_t workspace_bytes,
                         void *stream);

/* ---------------------------------------------------------------------------------------------
 * Host-buffer convenience call: the whole path (filter -> spaxel -> fused cube -> PSF -> LSF) for
 * callers that hold numpy / host arrays.  Copies inputs H2D, runs the kernels, copies the cube
 * back; h_cube is (num_spaxels, num_spaxels, W).  h_psf (M,N) / h_lsf (K,) may be NULL to skip.
 * ------------------------------------------------------------------------------------------- */
int rbx_pipeline_host(const rbx_plan *plan, const float *h_coords, const float *h_velocity,
                      const float *h_mass, const float *h_metallicity, const float *h_age, int64_t n,
                      const float *h_edges, int n_edges, int num_spaxels, int apply_filter,
                      const float *h_psf, int M, int N, const float *h_lsf, int K, int ext,
                      float *h_cube, void *stream);

/* The same call with structure-of-arrays host buffers (x, y, line-of-sight velocity: 24 bytes per particle over
 * PCIe instead of 40). */
int rbx_pipeline_host_packed(const rbx_plan *plan, const float *h_x, const float *h_y, const float *h_vlos,
                             const float *h_mass, const float *h_metallicity, const float *h_age, int64_t n,
                             const float *h_edges, int n_edges, int num_spaxels, int apply_filter,
                             const float *h_psf, int M, int N, const float *h_lsf, int K, int ext,
                             float *h_cube, void *stream);

/* A rank's host shard -> its PARTIAL cube on the device: the first half of rbx_pipeline_host for the multi-GPU path
 * (rubix/core/ifu.py:299-333: every device bins its particle shard, then jnp.sum over devices -- here
 * rbx_reduce_cube / rbx_reduce_scatter_cube follow on the device).  The shard is copied in contiguous ranges on a
 * second stream while the kernels of the previous range run.  nslab > 1: slab-major cube with `halo` channels as
 * rbx_assign_build_cube_slabs writes it (d_cube then holds nslab * S*S * ws floats), else (S, S, W).  Stream-ordered:
 * returns without synchronising; the host arrays (pinned for the overlap) must stay valid until `stream` has passed
 * this call. */
int rbx_build_cube_host(const rbx_plan *plan, const float *h_coords, const float *h_velocity, const float *h_mass,
                        const float *h_metallicity, const float *h_age, int64_t n, const float *h_edges, int n_edges,
                        int num_spaxels, int apply_filter, int nslab, int halo, float *d_cube, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RUBIX_B200_H */
