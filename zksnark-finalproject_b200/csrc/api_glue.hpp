// Shared glue for the extern "C" translation units.
#pragma once
#include <memory>

#include "common.hpp"

struct b2z_ctx {
  b2z::Ctx impl;
};

namespace b2z {

template <class F>
b2z_status guarded(b2z_ctx* ctx, F&& f) {
  if (ctx == nullptr) return B2Z_EINVAL;
  std::lock_guard<std::mutex> lock(ctx->impl.mu);
  try {
    B2Z_CUDA(cudaSetDevice(ctx->impl.device));
    f(ctx->impl);
    return B2Z_OK;
  } catch (const StatusError& e) {
    ctx->impl.last_error = e.msg;
    // leave the device usable for the next call
    cudaDeviceSynchronize();
    cudaGetLastError();
    return e.code;
  } catch (const std::exception& e) {
    ctx->impl.last_error = e.what();
    return B2Z_EINVAL;
  }
}

}  // namespace b2z
