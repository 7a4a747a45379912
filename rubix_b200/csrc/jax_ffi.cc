// XLA FFI handlers around the C ABI (include/rubix_b200.h) so that the path can be called with
// jax.ffi.ffi_call from inside rubix's single jax.jit (rubix/pipeline/abstract_pipeline.py:126-130).
//
// NOT BUILT IN THIS IMAGE: neither jax nor the XLA FFI headers (xla/ffi/api/ffi.h, shipped inside
// jaxlib: `python -c "import jax.ffi; print(jax.ffi.include_dir())"`) exist here, so this file is
// compiled only when the header is found (`make -C rubix_b200/csrc jax_ffi XLA_FFI_INCLUDE=...`)
// and has not been exercised.  The C ABI underneath is what the GPU tests cover.
//
// Conventions: every handler takes the CUDA stream from the platform context, device buffers from
// XLA, and writes into XLA-owned result buffers; the workspace of rbx_build_cube is one more result
// buffer (uint8) sized on the Python side with rbx_build_cube_workspace_bytes, so nothing is
// allocated here and nothing synchronises.  The plan handle (rbx_plan*, created once per
// configuration outside jit) travels as an int64 attribute.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define RBX_HAVE_XLA_FFI 1
#endif
#endif

#ifdef RBX_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "rubix_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error status(int rc) {
  if (rc == RBX_OK) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, std::string("rubix_b200: ") + rbx_last_error());
}

// a0: coords (n,3), edges (e,) -> pixel (n,) int32      replaces rubix/telescope/utils.py:138-151
static ffi::Error SpaxelAssignImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> coords, ffi::Buffer<ffi::F32> edges,
                                   ffi::ResultBuffer<ffi::S32> pixel) {
  const int64_t n = coords.dimensions()[0];
  return status(rbx_spaxel_assign(coords.typed_data(), n, edges.typed_data(), (int)edges.element_count(),
                                  pixel->typed_data(), nullptr, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(RbxSpaxelAssign, SpaxelAssignImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

// a1..a5 fused: velocity (n,3), mass, metallicity, age (n,), pixel (n,) -> cube (S,S,W), workspace
// replaces rubix/core/ifu.py:95-118,152-154,269-293,327-339
static ffi::Error BuildCubeImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> velocity, ffi::Buffer<ffi::F32> mass,
                                ffi::Buffer<ffi::F32> metallicity, ffi::Buffer<ffi::F32> age,
                                ffi::Buffer<ffi::S32> pixel, ffi::ResultBuffer<ffi::F32> cube,
                                ffi::ResultBuffer<ffi::U8> workspace, int64_t plan, int32_t num_spaxels) {
  const int64_t n = mass.element_count();
  return status(rbx_build_cube(reinterpret_cast<const rbx_plan *>(plan), velocity.typed_data(), mass.typed_data(),
                               metallicity.typed_data(), age.typed_data(), pixel.typed_data(), n, num_spaxels,
                               cube->typed_data(), workspace->typed_data(), workspace->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(RbxBuildCube, BuildCubeImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<int32_t>("num_spaxels"));

// stage calls for the stepwise path (stars.spectra observable between stages)
static ffi::Error SspLookupImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> metallicity, ffi::Buffer<ffi::F32> age,
                                ffi::ResultBuffer<ffi::F32> spectra, int64_t plan) {
  return status(rbx_ssp_lookup(reinterpret_cast<const rbx_plan *>(plan), metallicity.typed_data(), age.typed_data(),
                               (int64_t)metallicity.element_count(), spectra->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(RbxSspLookup, SspLookupImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("plan"));

static ffi::Error DopplerResampleImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> spectra,
                                      ffi::Buffer<ffi::F32> velocity, ffi::ResultBuffer<ffi::F32> out, int64_t plan) {
  const int64_t n = velocity.element_count() / 3;
  return status(rbx_doppler_resample(reinterpret_cast<const rbx_plan *>(plan), spectra.typed_data(),
                                     velocity.typed_data(), n, out->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(RbxDopplerResample, DopplerResampleImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("plan"));

// a6 + a7: cube (ny,nx,W), psf (M,N), lsf (K,) -> out (ny,nx,W)
// replaces rubix/telescope/psf/psf.py:56-57 and rubix/telescope/lsf/lsf.py:96-105
static ffi::Error PsfLsfImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> cube, ffi::Buffer<ffi::F32> psf,
                             ffi::Buffer<ffi::F32> lsf, ffi::ResultBuffer<ffi::F32> out, int32_t ext) {
  auto d = cube.dimensions();
  auto k = psf.dimensions();
  return status(rbx_psf_lsf(cube.typed_data(), out->typed_data(), (int)d[0], (int)d[1], (int)d[2], psf.typed_data(),
                            (int)k[0], (int)k[1], lsf.typed_data(), (int)lsf.element_count(), ext, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(RbxPsfLsf, PsfLsfImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("ext"));
// a0 + a1..a5 in one call: coords (n,3), edges (e,), velocity (n,3), mass, metallicity, age (n,) -> cube, workspace
// (spaxel assignment and aperture filter inside the first kernel of the build)
static ffi::Error AssignBuildCubeImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> coords, ffi::Buffer<ffi::F32> edges,
                                      ffi::Buffer<ffi::F32> velocity, ffi::Buffer<ffi::F32> mass,
                                      ffi::Buffer<ffi::F32> metallicity, ffi::Buffer<ffi::F32> age,
                                      ffi::ResultBuffer<ffi::F32> cube, ffi::ResultBuffer<ffi::U8> workspace,
                                      int64_t plan, int32_t num_spaxels, int32_t apply_filter) {
  const int64_t n = mass.element_count();
  return status(rbx_assign_build_cube(reinterpret_cast<const rbx_plan *>(plan), coords.typed_data(), edges.typed_data(),
                                      (int)edges.element_count(), apply_filter, velocity.typed_data(), mass.typed_data(),
                                      metallicity.typed_data(), age.typed_data(), n, num_spaxels, nullptr,
                                      cube->typed_data(), workspace->typed_data(), workspace->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(RbxAssignBuildCube, AssignBuildCubeImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<int32_t>("num_spaxels")
                                  .Attr<int32_t>("apply_filter"));

// a6 + a7 with the taps as static attributes (they are config constants inside rubix's closures:
// rubix/core/psf.py:52-60, rubix/core/lsf.py:50-55): the HBM-bound marching kernel
static ffi::Error PsfLsfTapsImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> cube, ffi::ResultBuffer<ffi::F32> out,
                                 ffi::Span<const float> psf, int32_t psf_size, ffi::Span<const float> lsf, int32_t ext) {
  auto d = cube.dimensions();
  int rc = rbx_psf_lsf_taps(cube.typed_data(), out->typed_data(), (int)d[0], (int)d[1], (int)d[2], psf.begin(), psf_size,
                            psf_size, lsf.begin(), (int)lsf.size(), ext, stream);
  if (rc == RBX_ERR_UNSUPPORTED)
    return ffi::Error(ffi::ErrorCode::kUnimplemented, "rubix_b200: taps need the device-tap call rbx_psf_lsf");
  return status(rc);
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(RbxPsfLsfTaps, PsfLsfTapsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<ffi::Span<const float>>("psf")
                                  .Attr<int32_t>("psf_size")
                                  .Attr<ffi::Span<const float>>("lsf")
                                  .Attr<int32_t>("ext"));
// dust variant (calc_dusty_ifu): A_V per star from the gas cells of its spaxel, replaces the lexsort / lax.scan part of
// rubix/spectra/dust/dust_extinction.py:240-337; the dust-to-gas fit, the A_V constant and the spaxel area are static
static ffi::Error DustAvImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> gas_coords, ffi::Buffer<ffi::S32> gas_pixel,
                             ffi::Buffer<ffi::F32> gas_mass, ffi::Buffer<ffi::F32> gas_metals,
                             ffi::Buffer<ffi::F32> star_coords, ffi::Buffer<ffi::S32> star_pixel,
                             ffi::ResultBuffer<ffi::F32> av, ffi::ResultBuffer<ffi::U8> workspace, int32_t n_spaxels,
                             ffi::Span<const float> dust_to_gas, float ext_const, float spaxel_area) {
  if (dust_to_gas.size() != 5)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "rubix_b200: dust_to_gas needs 5 parameters");
  const int64_t n_gas = gas_mass.element_count(), n_star = star_pixel.element_count();
  const int n_metals = n_gas ? (int)(gas_metals.element_count() / n_gas) : 5;
  return status(rbx_dust_av(gas_coords.typed_data(), gas_pixel.typed_data(), gas_mass.typed_data(), gas_metals.typed_data(),
                            n_metals, n_gas, star_coords.typed_data(), star_pixel.typed_data(), n_star, n_spaxels,
                            dust_to_gas.begin(), ext_const, spaxel_area, av->typed_data(), nullptr, workspace->typed_data(),
                            workspace->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(RbxDustAv, DustAvImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("n_spaxels")
                                  .Attr<ffi::Span<const float>>("dust_to_gas")
                                  .Attr<float>("ext_const")
                                  .Attr<float>("spaxel_area"));

// spectra (n, W) * 10^(-0.4 axav (W,) av (n,)): replaces extinguish + the final product, dust_extinction.py:341-356
static ffi::Error ApplyExtinctionImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> spectra, ffi::Buffer<ffi::F32> av,
                                      ffi::Buffer<ffi::F32> axav, ffi::ResultBuffer<ffi::F32> out) {
  const int W = (int)axav.element_count();
  return status(rbx_apply_extinction(spectra.typed_data(), av.typed_data(), axav.typed_data(),
                                     (int64_t)av.element_count(), W, out->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(RbxApplyExtinction, ApplyExtinctionImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());
#endif  // RBX_HAVE_XLA_FFI
