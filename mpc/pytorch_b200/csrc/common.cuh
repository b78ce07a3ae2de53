// common.cuh - PTX helpers (mbarrier, 1-D bulk TMA) and tiny per-lane linear algebra.
// sm_100a only.  No reference code: the algorithms these serve are cited in lqr_step.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpcb200 {

#define MPCB_DEV __device__ __forceinline__

// ------------------------------------------------------------------ mbarrier / bulk copy
MPCB_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

MPCB_DEV void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
MPCB_DEV void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

MPCB_DEV void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
MPCB_DEV void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
MPCB_DEV bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
MPCB_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk TMA: global -> shared, completion signalled on an mbarrier (UBLKCP in SASS).
// dst/src 16-byte aligned, bytes a multiple of 16.
MPCB_DEV void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
MPCB_DEV void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------ compile-time helpers
__host__ __device__ constexpr int round_up(int v, int a) { return (v + a - 1) / a * a; }
template <typename R>
__host__ __device__ constexpr int vec_elems(int count) {
  // widest vector (in elements) that divides `count` elements and keeps 16/8-byte alignment
  return (count * (int)sizeof(R)) % 16 == 0 ? 16 / (int)sizeof(R)
         : (count * (int)sizeof(R)) % 8 == 0 ? 8 / (int)sizeof(R)
                                             : 1;
}

template <typename R, int V>
struct VecLoad;
template <>
struct VecLoad<float, 4> {
  MPCB_DEV static void ld(const float* p, float* o) {
    float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};
template <>
struct VecLoad<float, 2> {
  MPCB_DEV static void ld(const float* p, float* o) {
    float2 v = *reinterpret_cast<const float2*>(p);
    o[0] = v.x; o[1] = v.y;
  }
};
template <>
struct VecLoad<float, 1> {
  MPCB_DEV static void ld(const float* p, float* o) { o[0] = p[0]; }
};
template <>
struct VecLoad<double, 2> {
  MPCB_DEV static void ld(const double* p, double* o) {
    double2 v = *reinterpret_cast<const double2*>(p);
    o[0] = v.x; o[1] = v.y;
  }
};
template <>
struct VecLoad<double, 1> {
  MPCB_DEV static void ld(const double* p, double* o) { o[0] = p[0]; }
};

// Load CNT contiguous elements whose start is aligned to V elements.
template <typename R, int CNT, int V>
MPCB_DEV void load_vec(const R* p, R (&out)[CNT]) {
  static_assert(CNT % V == 0, "vector width must divide count");
#pragma unroll
  for (int e = 0; e < CNT; e += V) VecLoad<R, V>::ld(p + e, &out[e]);
}

template <typename R>
MPCB_DEV R shfl(R v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// ------------------------------------------------------------------ M x M LDL^T (per lane, registers)
// Factor a symmetric matrix A = L D L^T (unit lower L).  Only the lower triangle of A is read.
// Masked (clamped) indices are presented by the caller as zero rows/cols with a tiny diagonal,
// exactly how the reference builds H_ (mpc/pnqp.py:46-48) and Qt_uu_ (mpc/lqr_step.py:107-116):
// they decouple, and a zero right-hand side gives an exactly zero solution component.
template <typename R, int M>
struct Ldl {
  R L[M][M];
  R d[M];
  R dinv[M];
  bool bad;  // a pivot was <= 0 or not finite

  MPCB_DEV void factor(const R (&A)[M][M]) {
    bad = false;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      R dj = A[j][j];
#pragma unroll
      for (int k = 0; k < j; ++k) dj -= L[j][k] * L[j][k] * d[k];
      bad = bad || !(dj > R(0));
      d[j] = dj;
      dinv[j] = R(1) / dj;
#pragma unroll
      for (int i = j + 1; i < M; ++i) {
        R s = A[i][j];
#pragma unroll
        for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k] * d[k];
        L[i][j] = s * dinv[j];
      }
    }
  }
  // x = A^{-1} b
  MPCB_DEV void solve(const R (&b)[M], R (&x)[M]) const {
    R y[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
      R s = b[i];
#pragma unroll
      for (int k = 0; k < i; ++k) s -= L[i][k] * y[k];
      y[i] = s;
    }
#pragma unroll
    for (int i = M - 1; i >= 0; --i) {
      R s = y[i] * dinv[i];
#pragma unroll
      for (int k = i + 1; k < M; ++k) s -= L[k][i] * x[k];
      x[i] = s;
    }
  }
};

}  // namespace mpcb200
