// XLA FFI shim for libgwi.so: lets jax.ffi.ffi_call("gwi_loglike", ...) run the
// fused likelihood + gradient on XLA's own CUDA stream, inside jit / lax.while_loop (NumPyro's NUTS).
//
// COMPILE-GATED: jaxlib's header xla/ffi/api/ffi.h is not present in this image (no jax / jaxlib /
// numpyro, no network), so this translation unit is empty here and is NOT part of the default build
// or of any test.  In an environment with jaxlib:
//
//   g++ -O2 -fPIC -shared -std=c++17 gwi_ffi.cc -I../../include \
//       -I$(python -c "import jax; print(jax.ffi.include_dir())") -L.. -lgwi -o ../libgwi_ffi.so
//
// and register + wrap it as shown in INTEGRATION.md section 4 (custom_vjp: the forward call already
// returns the gradient, the backward rule is `ct * grad`).
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define GWI_HAVE_XLA_FFI 1
#endif
#endif

#ifdef GWI_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include <cstdint>

#include "gwi.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

// `handle`: the gwi_model* obtained from gwi_model_create (ctypes side), passed as an int64 attribute.
// flags: bit 0 marginalize_selection, bit 1 min_neff_cut, bit 2 max_variance_cut (analysis.py:139-163).
static ffi::Error GwiLogLikeImpl(cudaStream_t stream, int64_t handle, int32_t nobs, int32_t flags, ffi::Buffer<ffi::F64> lam,
                                 ffi::ResultBuffer<ffi::F64> out) {
  gwi_model* m = reinterpret_cast<gwi_model*>(static_cast<intptr_t>(handle));
  gwi_like_opts o{nobs, flags & 1, (flags >> 1) & 1, (flags >> 2) & 1};
  const int64_t n_chains = lam.dimensions().size() == 2 ? lam.dimensions()[0] : 1;  // vmap over chains
  const int rc = n_chains == 1 ? gwi_loglike(m, lam.typed_data(), &o, out->typed_data(), stream)
                               : gwi_loglike_batch(m, lam.typed_data(), static_cast<int32_t>(n_chains), &o, out->typed_data(), stream);
  return rc == GWI_OK ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, gwi_last_error());
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(GwiLogLike, GwiLogLikeImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Attr<int32_t>("nobs")
                                  .Attr<int32_t>("flags")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());
#endif  // GWI_HAVE_XLA_FFI
