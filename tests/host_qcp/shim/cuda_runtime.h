// Test-only stand-in for <cuda_runtime.h>: lets g++ compile mdtraj_b200/csrc/qcp.cuh for the host so that the
// device solver's SOURCE can be checked against float64 truth on a machine without a GPU (tests/test_qcp_host.py).
// One "warp" is a single lane here: warp votes degenerate to the lane's own predicate.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fadd_rn(float a, float b) { volatile float s = a + b; return s; }  // volatile: no reassociation, no excess precision
static inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline double __longlong_as_double(long long i) { double d; std::memcpy(&d, &i, 8); return d; }
static inline bool __all_sync(unsigned, bool p) { return p; }
static inline bool __any_sync(unsigned, bool p) { return p; }
