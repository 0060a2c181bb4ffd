// rst_common.cuh -- shared helpers for the B200 (sm_100a) damped-Newton BVP kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define RST_FULL_MASK 0xffffffffu

// status codes shared with include/rst_b200.h
#define RST_OK 0
#define RST_ERR_INVALID (-1)
#define RST_ERR_CUDA (-2)
#define RST_ERR_ZERO_PIVOT (-3)
#define RST_ERR_NAN (-4)
#define RST_ERR_UNSUPPORTED (-5)
#define RST_ERR_NOT_FACTORIZED (-6)

// device-side error flag bits (OR-ed into Solver::d_flags[0])
#define RST_FLAG_ZERO_PIVOT 1u
#define RST_FLAG_NAN_F 2u
#define RST_FLAG_NAN_STEP 4u
#define RST_FLAG_PEER_TIMEOUT 8u

namespace rst {

__host__ __device__ constexpr int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

__device__ __forceinline__ double shfl_d(double v, int src, int width) {
    return __shfl_sync(RST_FULL_MASK, v, src, width);
}

// streaming (evict-first) stores/loads for data touched once per pass
__device__ __forceinline__ void st_cs(double *p, double v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs2(double2 *p, double2 v) { __stcs(p, v); }
__device__ __forceinline__ double ld_cs(const double *p) { return __ldcs(p); }
__device__ __forceinline__ double2 ld_cs2(const double2 *p) { return __ldcs(p); }

// Programmatic dependent launch (the host launches the sweeps with cudaLaunchAttributeProgrammaticStreamSerialization):
// a sweep kernel lets its successor be scheduled at once, pulls the first factors it will need into L2 -- they are static,
// nothing a predecessor writes -- and only then waits for the predecessor's results.  Without the launch attribute both
// instructions are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }


// reciprocal without the slow-path subroutine of __drcp_rn: MUFU.RCP64H seed (2^-23) + two Newton steps (~1 ulp).
// Denormal / zero / non-finite pivots give inf / NaN and are reported through the zero-pivot flag by the caller.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}


}  // namespace rst
