// lmpc_warp.cuh -- warp-synchronous building blocks of the solver kernels.
//
// The QP kernel gives one warp to one MPC instance.  Its code is written as a sequence of
// *phases*: inside LANES_BEGIN/LANES_END every lane runs the body with its own `lane`, then the
// warp synchronises (shared memory written in one phase is read in the next).  Values that
// live in a lane's registers across phases are LaneVar<T>.  Cross-lane reductions are the
// warp_* collectives, called between phases.
//
// Under nvcc this is ordinary SIMT code (__syncwarp, shuffles).  With -DLMPC_EMULATE the same
// source compiles with g++ into a lane-loop emulator (tests/emu) so that the kernel's logic can
// be checked against the CPU oracle on a machine without a GPU.  The emulator is a test
// artefact: it is never linked into liblmpc_b200.so and nothing in the product falls back to it.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(LMPC_EMULATE)
// ------------------------------------------------------------------ lane-loop emulation (tests)
#define LMPC_DEV static inline
#define LMPC_HD static inline
extern int g_lmpc_emu_reverse;  // run lanes 31..0 instead of 0..31 (order-independence check)
#define LANES_BEGIN                                      \
  for (int lane_it_ = 0; lane_it_ < 32; ++lane_it_) {    \
    const int lane = g_lmpc_emu_reverse ? 31 - lane_it_ : lane_it_;
#define LANES_END }
template <class T>
struct LaneVar {
  T v[32];
  inline T& operator()(int l) { return v[l]; }
  inline const T& operator()(int l) const { return v[l]; }
};
#else
// ------------------------------------------------------------------ CUDA
#define LMPC_DEV __device__ __forceinline__
#define LMPC_HD __host__ __device__ __forceinline__
#define LANES_BEGIN \
  {                 \
    const int lane = (int)(threadIdx.x & 31u);
#define LANES_END \
  }               \
  __syncwarp();
template <class T>
struct LaneVar {
  T v;
  __device__ __forceinline__ T& operator()(int) { return v; }
  __device__ __forceinline__ const T& operator()(int) const { return v; }
};
#endif

// ------------------------------------------------------------------ collectives
#if defined(LMPC_EMULATE)
LMPC_DEV void warp_sum(LaneVar<double>& x) {
  // same pairwise (butterfly) order as the shuffle version so that results are bit-identical
  double t[32];
  for (int l = 0; l < 32; ++l) t[l] = x.v[l];
  for (int off = 16; off >= 1; off >>= 1) {
    double n[32];
    for (int l = 0; l < 32; ++l) n[l] = t[l] + t[l ^ off];
    for (int l = 0; l < 32; ++l) t[l] = n[l];
  }
  for (int l = 0; l < 32; ++l) x.v[l] = t[l];
}
LMPC_DEV void warp_min(LaneVar<double>& x) {
  double m = x.v[0];
  for (int l = 1; l < 32; ++l) m = fmin(m, x.v[l]);
  for (int l = 0; l < 32; ++l) x.v[l] = m;
}
LMPC_DEV void warp_max(LaneVar<double>& x) {
  double m = x.v[0];
  for (int l = 1; l < 32; ++l) m = fmax(m, x.v[l]);
  for (int l = 0; l < 32; ++l) x.v[l] = m;
}
LMPC_DEV void warp_or(LaneVar<int>& x) {
  int m = 0;
  for (int l = 0; l < 32; ++l) m |= (x.v[l] != 0);
  for (int l = 0; l < 32; ++l) x.v[l] = m;
}
// arg-max with ties to the lowest index: every lane ends with the winning (value, index)
LMPC_DEV void warp_argmax(LaneVar<double>& val, LaneVar<int>& idx) {
  double bv = val.v[0];
  int bi = idx.v[0];
  for (int l = 1; l < 32; ++l)
    if (val.v[l] > bv || (val.v[l] == bv && idx.v[l] < bi)) { bv = val.v[l]; bi = idx.v[l]; }
  for (int l = 0; l < 32; ++l) { val.v[l] = bv; idx.v[l] = bi; }
}
// arg-min on (value, index) with ties to the lowest index
LMPC_DEV void warp_argmin(LaneVar<double>& val, LaneVar<int>& idx) {
  double bv = val.v[0];
  int bi = idx.v[0];
  for (int l = 1; l < 32; ++l)
    if (val.v[l] < bv || (val.v[l] == bv && idx.v[l] < bi)) { bv = val.v[l]; bi = idx.v[l]; }
  for (int l = 0; l < 32; ++l) { val.v[l] = bv; idx.v[l] = bi; }
}
#else
LMPC_DEV double shfl_xor_f64(double v, int off) { return __shfl_xor_sync(0xffffffffu, v, off); }
LMPC_DEV void warp_sum(LaneVar<double>& x) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) x.v += shfl_xor_f64(x.v, off);
}
LMPC_DEV void warp_min(LaneVar<double>& x) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) x.v = fmin(x.v, shfl_xor_f64(x.v, off));
}
LMPC_DEV void warp_max(LaneVar<double>& x) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) x.v = fmax(x.v, shfl_xor_f64(x.v, off));
}
LMPC_DEV void warp_or(LaneVar<int>& x) { x.v = __any_sync(0xffffffffu, x.v != 0) ? 1 : 0; }
LMPC_DEV void warp_argmax(LaneVar<double>& val, LaneVar<int>& idx) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const double ov = shfl_xor_f64(val.v, off);
    const int oi = __shfl_xor_sync(0xffffffffu, idx.v, off);
    if (ov > val.v || (ov == val.v && oi < idx.v)) { val.v = ov; idx.v = oi; }
  }
}
LMPC_DEV void warp_argmin(LaneVar<double>& val, LaneVar<int>& idx) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const double ov = shfl_xor_f64(val.v, off);
    const int oi = __shfl_xor_sync(0xffffffffu, idx.v, off);
    if (ov < val.v || (ov == val.v && oi < idx.v)) { val.v = ov; idx.v = oi; }
  }
}
#endif
