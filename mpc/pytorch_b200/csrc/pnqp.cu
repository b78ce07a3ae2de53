// pnqp.cu - standalone projected-Newton box QP (reference mpc/pnqp.py:5-82) for small n.
// One thread per QP; same per-problem control flow and arithmetic as inside the step kernel
// (pnqp_lane in lqr_step.cuh).  min 0.5 x'Hx + q'x  s.t. lower <= x <= upper.
#include "../../../include/mpcb200.h"
#include "lqr_step.cuh"

namespace mpcb200 {

struct PnqpArgs {
  int B, n_iter, has_init;
  const void *H, *q, *lo, *hi, *x_init;
  void *x, *Hfree;
  unsigned char* If;
  int* iters;
  int* status;
};

template <typename R, int M>
__global__ void __launch_bounds__(128) pnqp_kernel(const PnqpArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const R* gH = (const R*)a.H + (size_t)b * M * M;
  const R* gq = (const R*)a.q + (size_t)b * M;
  const R* glo = (const R*)a.lo + (size_t)b * M;
  const R* ghi = (const R*)a.hi + (size_t)b * M;
  R H[M][M], q[M], lo[M], hi[M], x[M];
#pragma unroll
  for (int i = 0; i < M; ++i) {
#pragma unroll
    for (int j = 0; j < M; ++j) H[i][j] = gH[i * M + j];
    q[i] = gq[i];
    lo[i] = glo[i];
    hi[i] = ghi[i];
    x[i] = a.has_init ? ((const R*)a.x_init)[(size_t)b * M + i] : R(0);
  }
  Ldl<R, M> fac;
  unsigned fm = 0u;
  int it = 0;
  bool conv = false, badpiv = false;
  pnqp_lane<R, M>(H, q, lo, hi, a.has_init != 0, x, fac, fm, it, conv, badpiv, a.n_iter);
  R* ox = (R*)a.x + (size_t)b * M;
  R* oH = (R*)a.Hfree + (size_t)b * M * M;
#pragma unroll
  for (int i = 0; i < M; ++i) {
    ox[i] = x[i];
    a.If[(size_t)b * M + i] = (fm >> i) & 1u;
#pragma unroll
    for (int j = 0; j < M; ++j) {   // H_ of the returning iteration (reference mpc/pnqp.py:46-48)
      const bool ff = ((fm >> i) & 1u) && ((fm >> j) & 1u);
      oH[i * M + j] = (ff ? H[i][j] : R(0)) + (i == j ? R(1e-11) : R(0));
    }
  }
  a.iters[b] = it;
  if (a.status != nullptr) a.status[b] = (conv ? 0 : 1) | (badpiv ? 4 : 0);
}

template <typename R>
static int pnqp_dispatch(const PnqpArgs& a, int n, cudaStream_t stream) {
  const int grid = (a.B + 127) / 128;
  switch (n) {
#define MPCB_PNQP_CASE(MM) \
  case MM: pnqp_kernel<R, MM><<<grid, 128, 0, stream>>>(a); break;
    MPCB_PNQP_CASE(1) MPCB_PNQP_CASE(2) MPCB_PNQP_CASE(3) MPCB_PNQP_CASE(4)
    MPCB_PNQP_CASE(5) MPCB_PNQP_CASE(6) MPCB_PNQP_CASE(7) MPCB_PNQP_CASE(8)
#undef MPCB_PNQP_CASE
    default: return MPCB200_ERR_UNSUPPORTED_DIMS;
  }
  return cudaGetLastError() == cudaSuccess ? MPCB200_OK : MPCB200_ERR_LAUNCH;
}

template <typename R>
static int pnqp_impl(int32_t B, int32_t n, const R* H, const R* q, const R* lower, const R* upper,
                     const R* x_init, int32_t n_iter, R* x, R* H_free, uint8_t* If, int32_t* iters,
                     int32_t* status, void* stream) {
  if (B <= 0 || n <= 0 || n_iter < 1) return MPCB200_ERR_BAD_DIMS;
  if (!H || !q || !lower || !upper || !x || !H_free || !If || !iters) return MPCB200_ERR_NULL_POINTER;
  PnqpArgs a;
  a.B = B; a.n_iter = n_iter; a.has_init = x_init != nullptr;
  a.H = H; a.q = q; a.lo = lower; a.hi = upper; a.x_init = x_init;
  a.x = x; a.Hfree = H_free; a.If = If; a.iters = iters; a.status = status;
  return pnqp_dispatch<R>(a, n, (cudaStream_t)stream);
}
}  // namespace mpcb200

extern "C" {
int mpcb200_pnqp_f32(int32_t B, int32_t n, const float* H, const float* q, const float* lower,
                     const float* upper, const float* x_init, int32_t n_iter, float* x, float* H_free,
                     uint8_t* If, int32_t* iters, int32_t* status, void* stream) {
  return mpcb200::pnqp_impl<float>(B, n, H, q, lower, upper, x_init, n_iter, x, H_free, If, iters, status, stream);
}
int mpcb200_pnqp_f64(int32_t B, int32_t n, const double* H, const double* q, const double* lower,
                     const double* upper, const double* x_init, int32_t n_iter, double* x, double* H_free,
                     uint8_t* If, int32_t* iters, int32_t* status, void* stream) {
  return mpcb200::pnqp_impl<double>(B, n, H, q, lower, upper, x_init, n_iter, x, H_free, If, iters, status, stream);
}
}
