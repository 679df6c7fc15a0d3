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

#ifdef __CUDACC__
#define SMPC_DEV __device__ __forceinline__
#define SMPC_DEV_NOINLINE __device__ __noinline__
#define SMPC_PAR(tid)                                                                                                  \
  {                                                                                                                    \
    const int tid = threadIdx.x;                                                                                       \
    const int SMPC_NT = blockDim.x;                                                                                    \
    (void)SMPC_NT;
#define SMPC_END                                                                                                       \
  }                                                                                                                    \
  __syncthreads();
#else
#include <cmath>
#ifndef SMPC_EMUL_NT
#define SMPC_EMUL_NT 64
#endif
#define SMPC_DEV inline
#define SMPC_DEV_NOINLINE inline
#define SMPC_PAR(tid)                                                                                                  \
  for (int tid = 0; tid < SMPC_EMUL_NT; ++tid)                                                                         \
  {                                                                                                                    \
    const int SMPC_NT = SMPC_EMUL_NT;                                                                                  \
    (void)SMPC_NT;
#define SMPC_END }
#define __restrict__
#endif
