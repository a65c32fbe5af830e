// XLA-FFI shim (level 2 of the boundary): registers jsso_forward / jsso_backward as XLA
// custom calls.  Zero logic: it unwraps buffers and forwards to the C ABI.
//
// NOT COMPILED IN THIS IMAGE: it needs the headers from jax.ffi.include_dir() (jaxlib), and jax is not
// installed here.  Build on a machine with jax >= 0.5:
//   g++ -shared -fPIC -O2 -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -I../../include \
//       jsso_xla_ffi.cc -L.. -ljsso -o ../libjsso_xla.so
#include <cstdint>

#include "jsso.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

// The handle pointer is passed as an int64 attribute (one handle per frozen model, cached by the Python side).
static ffi::Error ForwardImpl(cudaStream_t stream, int64_t handle, double rtol, ffi::Buffer<ffi::F64> crds,
                              ffi::Buffer<ffi::F64> prop_q, ffi::Buffer<ffi::F64> prop_b,
                              ffi::Buffer<ffi::F64> f, ffi::ResultBuffer<ffi::F64> u) {
  jsso_solve_opts o{rtol, 0, 0, 0, 0};
  jsso_stats st;
  jsso_handle* h = reinterpret_cast<jsso_handle*>(handle);
  int rc = jsso_forward(h, crds.typed_data(), prop_q.typed_data(), prop_b.typed_data(), f.typed_data(),
                        u->typed_data(), &o, &st, stream);
  return rc ? ffi::Error(ffi::ErrorCode::kInternal, jsso_last_error(h)) : ffi::Error::Success();
}

static ffi::Error BackwardImpl(cudaStream_t stream, int64_t handle, double rtol, ffi::Buffer<ffi::F64> crds,
                               ffi::Buffer<ffi::F64> prop_q, ffi::Buffer<ffi::F64> prop_b,
                               ffi::Buffer<ffi::F64> u, ffi::Buffer<ffi::F64> g,
                               ffi::ResultBuffer<ffi::F64> d_crds, ffi::ResultBuffer<ffi::F64> d_prop_q,
                               ffi::ResultBuffer<ffi::F64> d_prop_b) {
  jsso_solve_opts o{rtol, 0, 0, 0, 0};
  jsso_stats st;
  jsso_handle* h = reinterpret_cast<jsso_handle*>(handle);
  // The adjoint solve uses the matrix the handle holds.  In a traced program several fea_solve calls may share one
  // handle before their backward passes run (two designs in one loss, jacrev, XLA reordering), so the matrix of
  // THIS design is re-assembled first (1.2 ms + the numeric multigrid setup at 1M quads, against a solve of > 0.1 s).
  int rc = jsso_assemble(h, crds.typed_data(), prop_q.typed_data(), prop_b.typed_data(), 1, stream);
  if (rc) return ffi::Error(ffi::ErrorCode::kInternal, jsso_last_error(h));
  rc = jsso_backward(h, crds.typed_data(), prop_q.typed_data(), prop_b.typed_data(), u.typed_data(),
                         g.typed_data(), d_crds->typed_data(), d_prop_q->typed_data(), d_prop_b->typed_data(),
                         nullptr, &o, &st, stream);
  return rc ? ffi::Error(ffi::ErrorCode::kInternal, jsso_last_error(h)) : ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(jsso_xla_forward, ForwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Attr<double>("rtol")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(jsso_xla_backward, BackwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Attr<double>("rtol")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());
