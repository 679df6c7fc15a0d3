// par.cuh — how the per-knot / per-instance device code is written.
//
// Every cooperative routine in csrc/ is a sequence of PARALLEL PHASES over the threads of one CTA:
//
//     SMPC_PAR(tid)  { ... body indexed by tid, strided by SMPC_NT ... }  SMPC_END
//
// On the device a phase is the body executed by all threads followed by __syncthreads(); state that
// crosses a phase boundary lives in shared memory (the `Scratch` structs), never in registers.
// Because of that discipline the same source can be compiled by g++ with the phase turned into a
// serial loop over tid — tests/host_emul/ does exactly this to debug kernel logic on the CPU-only
// development box. That emulation is a TEST harness: it is not compiled into libsmpc_cuda.so, and
// the product has no CPU path.
#pragma once
#define SMPC_SCHED_FENCE asm volatile("" ::: "memory");

#ifdef __CUDACC__
#define SMPC_DEV __device__ __forceinline__
#define SMPC_DEV_NOINLINE __device__ __noinline__
#define SMPC_HD __host__ __device__ __forceinline__ // integer bookkeeping shared with the host-side setup code
#define SMPC_PAR(tid)                                                                                                  \
  {                                                                                                                    \
    const int tid = threadIdx.x;                                                                                       \
    const int SMPC_NT = blockDim.x;                                                                                    \
    (void)SMPC_NT;
#define SMPC_END                                                                                                       \
  }                                                                                                                    \
  __syncthreads();
// the one exception to "no registers across phases": a per-thread array that must survive barriers
// (a matrix row kept in registers during a factorisation). Declared once, fetched inside each phase.
#define SMPC_THREAD_ARRAY(name, n) double name##_reg[n];
#define SMPC_THREAD_IARRAY(name, n) int name##_reg[n];
#define SMPC_THREAD_PTR(name, n, tid) (name##_reg)
#else
#include <cmath>
#include <vector>
#define SMPC_THREAD_ARRAY(name, n) std::vector<double> name##_all((size_t)SMPC_EMUL_NT * (n), 0.0);
#define SMPC_THREAD_IARRAY(name, n) std::vector<int> name##_all((size_t)SMPC_EMUL_NT * (n), 0);
#define SMPC_THREAD_PTR(name, n, tid) (&name##_all[(size_t)(tid) * (n)])
#ifndef SMPC_EMUL_NT
#define SMPC_EMUL_NT 64
#endif
#define SMPC_DEV inline
#define SMPC_DEV_NOINLINE inline
#define SMPC_HD inline
#define SMPC_PAR(tid)                                                                                                  \
  for (int tid = 0; tid < SMPC_EMUL_NT; ++tid)                                                                         \
  {                                                                                                                    \
    const int SMPC_NT = SMPC_EMUL_NT;                                                                                  \
    (void)SMPC_NT;
#define SMPC_END }
#define __restrict__
#endif

// two consecutive doubles from a 16-byte aligned shared address in one LDS.128
SMPC_DEV void ld2(const double * p, double & a, double & b)
{
#ifdef __CUDACC__
  const double2 v = *reinterpret_cast<const double2 *>(p);
  a = v.x;
  b = v.y;
#else
  a = p[0];
  b = p[1];
#endif
}
