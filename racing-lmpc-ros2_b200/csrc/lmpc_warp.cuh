// lmpc_warp.cuh -- SIMT building blocks of the solver kernels.
//
// The QP kernel gives one *warp group* (NW warps = NT threads, one CTA) to one MPC instance.  Its code
// is a sequence of *phases*: inside GLANES_BEGIN/GLANES_END every thread of the group runs the body
// with its own `lane` (0..NT-1), then the group synchronises (shared memory written in one phase is
// read in the next).  Values that live in a thread's registers across phases are LaneVar<T, NT>.
// Cross-lane reductions are the group_* collectives, called between phases.  Code outside phases is
// group-uniform (every thread computes the same value).
//
// Under nvcc this is ordinary SIMT code (__syncwarp / bar.sync, shuffles).  With -DLMPC_EMULATE the
// same source compiles with g++ into a lane-loop emulator (tests/emu) so that the kernel's logic can
// be checked against the CPU oracle on a machine without a GPU.  The emulator is a test artefact: it
// is never linked into liblmpc_b200.so and nothing in the product falls back to it.
#pragma once

#include <math.h>
#include <stdint.h>

#define LMPC_UNROLL _Pragma("unroll")   // usable inside macros
#define LMPC_NOUNROLL _Pragma("unroll 1")
#define LMPC_UNROLL2 _Pragma("unroll 2")

#if defined(LMPC_EMULATE)
// ------------------------------------------------------------------ lane-loop emulation (tests)
#define LMPC_DEV static inline
#define LMPC_HD static inline
#define LMPC_HDM inline   // member functions
extern int g_lmpc_emu_reverse;  // run lanes NT-1..0 instead of 0..NT-1 (order-independence check)
#define GLANES_BEGIN(NT)                                      \
  for (int lane_it_ = 0; lane_it_ < (NT); ++lane_it_) {       \
    const int lane = g_lmpc_emu_reverse ? (NT) - 1 - lane_it_ : lane_it_;
#define GLANES_END(NW) }
#define GROUP_SYNC(NW)
#define LANE0_ONLY(stmt) { stmt; }
template <class T, int NT = 32>
struct LaneVar {
  T v[NT];
  inline T& operator()(int l) { return v[l]; }
  inline const T& operator()(int l) const { return v[l]; }
};
#else
// ------------------------------------------------------------------ CUDA
#define LMPC_DEV __device__ __forceinline__
#define LMPC_HD __host__ __device__ __forceinline__
#define LMPC_RED __device__ __forceinline__   // the multi-value reductions (tried as calls, __noinline__, to shrink the 270 KB
                                              // kernel: nvcc inlines them regardless and the kernel grew to 304 KB)
#define LMPC_HDM __host__ __device__ __forceinline__
#define GROUP_SYNC(NW)                       \
  do {                                       \
    if ((NW) == 1) __syncwarp(); else __syncthreads(); \
  } while (0)
#define GLANES_BEGIN(NT) \
  {                      \
    const int lane = (int)threadIdx.x;
#define GLANES_END(NW) \
  }                    \
  GROUP_SYNC(NW);
#define LANE0_ONLY(stmt) { if (threadIdx.x == 0u) { stmt; } }
template <class T, int NT = 32>
struct LaneVar {
  T v;
  __device__ __forceinline__ T& operator()(int) { return v; }
  __device__ __forceinline__ const T& operator()(int) const { return v; }
};
#endif

// running maximum: acc unless v is larger.  A NaN in v is ignored like fmax does; three instructions instead of fmax's
// seven (DSETP.MAX + NaN patch-up)
LMPC_HD double lmpc_max(double acc, double v) { return (v > acc) ? v : acc; }

// ------------------------------------------------------------------ branch-free reciprocal
// 1 / x for normal, finite, non-zero x (what the solver divides by: slacks, multipliers, pivots it has already tested).
// CUDA: the hardware seed (rcp.approx.ftz.f64, MUFU.RCP64H) and two Newton steps -- five instructions and no
// slow-path branch, within 1 ulp of the IEEE quotient; the IEEE division costs about fifteen and a call.  Emulation: 1 / x.
#if defined(LMPC_EMULATE)
LMPC_HD double lmpc_rcp(double x) { return 1.0 / x; }
LMPC_HD double lmpc_rsqrt(double x) { return 1.0 / sqrt(x); }
#else
// 1 / sqrt(x), x > 0 normal: rsqrt.approx.ftz.f64 and two Newton steps r <- r + r (1/2 - x r^2 / 2)
LMPC_DEV double lmpc_rsqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double hx = 0.5 * x;
  r = fma(r, fma(-hx * r, r, 0.5), r);
  r = fma(r, fma(-hx * r, r, 0.5), r);
  return r;
}
LMPC_DEV double lmpc_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}
#endif

// single-warp spellings (the safe-set kernel runs one independent item per warp, several warps per block)
#if defined(LMPC_EMULATE)
#define LANES_BEGIN GLANES_BEGIN(32)
#define LANES_END GLANES_END(1)
#else
#define LANES_BEGIN \
  {                 \
    const int lane = (int)(threadIdx.x & 31u);
#define LANES_END \
  }               \
  __syncwarp();
#endif

// ------------------------------------------------------------------ group collectives over NV values
// Every thread ends with the reduction over all NT = 32*NW lanes.  Order of the additions: butterfly
// inside each warp, then warps 0..NW-1 in sequence -- identical in the CUDA and the emulated build, so
// results are bit-identical.  `scratch` needs NW*NV doubles of shared memory when NW > 1.
enum { LMPC_RED_SUM = 0, LMPC_RED_MAX = 1, LMPC_RED_MIN = 2 };

LMPC_HD double lmpc_red_op(int op, double a, double b) { return op == LMPC_RED_SUM ? a + b : (op == LMPC_RED_MAX ? fmax(a, b) : fmin(a, b)); }

#if defined(LMPC_EMULATE)
template <int NW, int NV>
LMPC_DEV void group_reduce(LaneVar<double, 32 * NW> (&x)[NV], const int (&op)[NV], double* /*scratch*/) {
  for (int q = 0; q < NV; q++) {
    double wres[NW];
    for (int w = 0; w < NW; w++) {
      double t[32];
      for (int l = 0; l < 32; ++l) t[l] = x[q].v[32 * w + l];
      for (int off = 16; off >= 1; off >>= 1) {
        double n[32];
        for (int l = 0; l < 32; ++l) n[l] = lmpc_red_op(op[q], t[l], t[l ^ off]);
        for (int l = 0; l < 32; ++l) t[l] = n[l];
      }
      wres[w] = t[0];
    }
    double r = wres[0];
    for (int w = 1; w < NW; w++) r = lmpc_red_op(op[q], r, wres[w]);
    for (int l = 0; l < 32 * NW; ++l) x[q].v[l] = r;
  }
}
template <int NW, int NV>
LMPC_DEV void group_reduce_sum(LaneVar<double, 32 * NW> (&x)[NV], double* scratch) {
  int op[NV];
  for (int q = 0; q < NV; q++) op[q] = LMPC_RED_SUM;
  group_reduce<NW, NV>(x, op, scratch);
}
// arg-max / arg-min on (value, index), ties to the lowest index
template <int NW>
LMPC_DEV void group_argbest(LaneVar<double, 32 * NW>& val, LaneVar<int, 32 * NW>& idx, bool want_max, double* /*scratch*/) {
  double bv = val.v[0]; int bi = idx.v[0];
  for (int l = 1; l < 32 * NW; ++l) {
    const bool better = want_max ? (val.v[l] > bv) : (val.v[l] < bv);
    if (better || (val.v[l] == bv && idx.v[l] < bi)) { bv = val.v[l]; bi = idx.v[l]; }
  }
  for (int l = 0; l < 32 * NW; ++l) { val.v[l] = bv; idx.v[l] = bi; }
}
template <int NW>
LMPC_DEV void group_or(LaneVar<int, 32 * NW>& x, double* /*scratch*/) {
  int m = 0;
  for (int l = 0; l < 32 * NW; ++l) m |= (x.v[l] != 0);
  for (int l = 0; l < 32 * NW; ++l) x.v[l] = m;
}
#else
LMPC_DEV double shfl_xor_f64(double v, int off) { return __shfl_xor_sync(0xffffffffu, v, off); }
// Sums only, one warp: the butterfly costs 5 NV exchange steps.  Here every exchange step HALVES the slots a lane is
// responsible for (at offset 16 the lanes with bit 4 clear keep slots 0..15 and hand slots 16..31 to their partner, and
// so on): 31 steps per 32 values, lane L ends with the total of slot L, which goes through `scratch` (NV doubles of
// shared memory) to every lane.  The additions are the butterfly's own (same pairs, same order; a + b is commutative),
// so the totals are bit-identical to it.  A remainder of fewer than 8 values takes the butterfly.
template <int NV>
LMPC_RED void warp_reduce_scatter_sum(LaneVar<double, 32> (&x)[NV], double* scratch) {
  const unsigned lane = threadIdx.x & 31u;
  constexpr int NFULL = (NV % 32 >= 8) ? NV : (NV / 32) * 32;   // values that go through the halving scheme
#pragma unroll
  for (int c0 = 0; c0 < NFULL; c0 += 32) {
    double v[32];
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = (c0 + j < NFULL) ? x[c0 + j].v : 0.0;
#pragma unroll
    for (int H = 16; H >= 1; H >>= 1) {
      const bool hi = (lane & (unsigned)H) != 0u;
#pragma unroll
      for (int j = 0; j < H; j++) {
        const double keep = hi ? v[j + H] : v[j];
        const double send = hi ? v[j] : v[j + H];
        v[j] = keep + shfl_xor_f64(send, H);
      }
    }
    if (c0 + (int)lane < NFULL) scratch[c0 + lane] = v[0];
  }
#pragma unroll
  for (int q = NFULL; q < NV; q++) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) x[q].v = x[q].v + shfl_xor_f64(x[q].v, off);
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < NFULL; q++) x[q].v = scratch[q];
  __syncwarp();
}
template <int NW, int NV>
LMPC_RED void group_reduce(LaneVar<double, 32 * NW> (&x)[NV], const int (&op)[NV], double* scratch) {
#pragma unroll
  for (int q = 0; q < NV; q++) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) x[q].v = lmpc_red_op(op[q], x[q].v, shfl_xor_f64(x[q].v, off));
  }
  if (NW > 1) {
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31u) == 0u) {
#pragma unroll
      for (int q = 0; q < NV; q++) scratch[w * NV + q] = x[q].v;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; q++) {
      double r = scratch[q];
#pragma unroll
      for (int ww = 1; ww < NW; ww++) r = lmpc_red_op(op[q], r, scratch[ww * NV + q]);
      x[q].v = r;
    }
    __syncthreads();
  }
}
// all-SUM form: one warp with shared-memory scratch takes the halving scheme, everything else the general path
template <int NW, int NV>
LMPC_DEV void group_reduce_sum(LaneVar<double, 32 * NW> (&x)[NV], double* scratch) {
  if constexpr (NW == 1 && NV >= 8) {
    warp_reduce_scatter_sum<NV>(x, scratch);
  } else {
    int op[NV];
#pragma unroll
    for (int q = 0; q < NV; q++) op[q] = LMPC_RED_SUM;
    group_reduce<NW, NV>(x, op, scratch);
  }
}
template <int NW>
LMPC_DEV void group_argbest(LaneVar<double, 32 * NW>& val, LaneVar<int, 32 * NW>& idx, bool want_max, double* scratch) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const double ov = shfl_xor_f64(val.v, off);
    const int oi = __shfl_xor_sync(0xffffffffu, idx.v, off);
    const bool better = want_max ? (ov > val.v) : (ov < val.v);
    if (better || (ov == val.v && oi < idx.v)) { val.v = ov; idx.v = oi; }
  }
  if (NW > 1) {
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31u) == 0u) { scratch[2 * w] = val.v; scratch[2 * w + 1] = (double)idx.v; }
    __syncthreads();
    double bv = scratch[0]; int bi = (int)scratch[1];
#pragma unroll
    for (int ww = 1; ww < NW; ww++) {
      const double ov = scratch[2 * ww]; const int oi = (int)scratch[2 * ww + 1];
      const bool better = want_max ? (ov > bv) : (ov < bv);
      if (better || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    val.v = bv; idx.v = bi;
    __syncthreads();
  }
}
template <int NW>
LMPC_DEV void group_or(LaneVar<int, 32 * NW>& x, double* /*scratch*/) {
  if (NW == 1) x.v = __any_sync(0xffffffffu, x.v != 0) ? 1 : 0;
  else x.v = __syncthreads_or(x.v != 0) ? 1 : 0;
}
#endif

// ------------------------------------------------------------------ bulk stage-in (TMA)
// group_bulk_load_begin: one elected thread arms an mbarrier with the byte count and issues ONE bulk asynchronous
// copy global -> shared (cp.async.bulk, the TMA engine; SASS UBLKCP) of n doubles; the group goes on with other work
// and calls group_bulk_load_wait before the first read of dst.  Requirements: dst, src 16-byte aligned, 8 n a multiple
// of 16, a group synchronisation (any phase end) between begin and wait so that every thread sees the initialised
// barrier.  bar: 8 bytes of shared memory, used once per kernel (phase parity 0).  Emulation: a plain copy.
#if defined(LMPC_EMULATE)
LMPC_DEV void group_bulk_load_begin(double* dst, const double* src, int n, double* /*bar*/) { for (int i = 0; i < n; i++) dst[i] = src[i]; }
LMPC_DEV void group_bulk_load_wait(double* /*bar*/) {}
#else
LMPC_DEV void group_bulk_load_begin(double* dst, const double* src, int n, double* bar) {
  if (threadIdx.x == 0u) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), d = (uint32_t)__cvta_generic_to_shared(dst);
    const uint32_t bytes = 8u * (uint32_t)n;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src), "r"(bytes), "r"(b)
                 : "memory");
  }
}
LMPC_DEV void group_bulk_load_wait(double* bar) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(b)
        : "memory");
  }
}
#endif

// convenience single-value forms
template <int NW>
LMPC_DEV void group_sum(LaneVar<double, 32 * NW>& x, double* scratch) {
  LaneVar<double, 32 * NW>(&a)[1] = reinterpret_cast<LaneVar<double, 32 * NW>(&)[1]>(x);
  const int op[1] = {LMPC_RED_SUM};
  group_reduce<NW, 1>(a, op, scratch);
}
template <int NW>
LMPC_DEV void group_max(LaneVar<double, 32 * NW>& x, double* scratch) {
  LaneVar<double, 32 * NW>(&a)[1] = reinterpret_cast<LaneVar<double, 32 * NW>(&)[1]>(x);
  const int op[1] = {LMPC_RED_MAX};
  group_reduce<NW, 1>(a, op, scratch);
}

// single-warp spellings used by the safe-set kernel
LMPC_DEV void warp_argmin(LaneVar<double>& val, LaneVar<int>& idx) { group_argbest<1>(val, idx, false, nullptr); }
LMPC_DEV void warp_or(LaneVar<int>& x) { group_or<1>(x, nullptr); }
// arg-min of NON-NEGATIVE values (squared distances; NaN sorts last), ties to the lowest index -- the same answer as
// warp_argmin.  On the GPU: the bit pattern of a non-negative double orders like an unsigned integer, so three integer
// warp reductions (high word, low word among the lanes that matched, index among those) replace five exchange steps of
// (value, index) pairs: 12 instructions instead of 55 in each of the safe-set query's 32 rounds.
LMPC_DEV void warp_argmin_nonneg(LaneVar<double>& val, LaneVar<int>& idx) {
#if defined(LMPC_EMULATE)
  warp_argmin(val, idx);
#else
  const unsigned long long b = (unsigned long long)__double_as_longlong(val.v);
  const unsigned hi = (unsigned)(b >> 32), lo = (unsigned)b;
  const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
  const unsigned mi = __reduce_min_sync(0xffffffffu, (hi == mh && lo == ml) ? (unsigned)idx.v : 0xffffffffu);
  val.v = __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
  idx.v = (int)mi;
#endif
}
