// Multi-GPU state of a handle (see mgpu.cu).
#pragma once
#include <vector>
#include "pycs_common.cuh"

#define MG_MAX_WORLD 8

// One block per rank, exported to every peer through CUDA IPC.  All flags are epoch counters, never reset:
//   sflag[d] = fused steps rank d has completed (its MF-PR sum of that step is in psum[steps & 1][d]);
//   dflag[d] = halo exchanges rank d has delivered into this rank's Q arrays;
//   qflag[d] = flushes rank d has completed (it has stopped modifying its current Q array: see k_mg_quiesce).
struct MgSync {
  long long sflag[MG_MAX_WORLD];
  long long dflag[MG_MAX_WORLD];
  long long qflag[MG_MAX_WORLD];
  long long bflag[MG_MAX_WORLD];     // device-side barrier (k_mg_device_barrier): rank d has reached barrier number bflag[d]
  double psum[2][MG_MAX_WORLD];
  int err;                           // != 0: a flag wait of this rank gave up (a peer is gone); the host turns it
                                     // into PYCS_ERR_STATE at the next synchronisation point (k_mg_check)
};
// rectangle [i0,i1) x [j0,j1) of one panel that this rank stores into rank `peer` after every step
struct MgRect { int peer, panel, i0, i1, j0, j1; };
struct MgScatterJob { double* dst; int panel, i0, i1, j0, j1; };

struct StepCtl;
struct MgpuState {
  int rank, world, connected;
  double* alloc[2];                  // this rank's two Q allocations, in export order
  double* peer_q[2][MG_MAX_WORLD];   // peer_q[i][d]: allocation i of rank d, mapped here
  MgSync* sync;
  unsigned* counter;                 // CTAs of the exchange kernel that have finished
  MgSync* peer_sync[MG_MAX_WORLD];
  unsigned long long timeout_ns;     // wall-clock bound of a flag wait (PYCS_MG_TIMEOUT_S, default 30 s)
  std::vector<MgRect>* rects;        // what this rank sends (host copy of the plan)
  MgScatterJob* jobs_dev[2];         // the same with the destination pointers of allocation 0 / 1 resolved
  int njobs;
  int gf_lo, gf_hi;                  // rows whose ghost cells this rank needs: [row_lo - 3, row_hi + 3)
  long long qcount;                  // flushes of this rank (every rank runs the same call sequence)
  long long bcount;                  // device-side barriers issued
};

// Host-side plan (no GPU needed: exercised by the CPU tests).
void pycs_mgpu_rows(int N, int world, int rank, int* row_lo, int* row_hi);
// Rectangles rank `rank` must send after every step so that every peer holds what its next step reads:
// the 3 rows next to its slab and the sources of the Lagrange ghost cells it needs (derived from the
// halo index maps and the stencil table kmin_east (4, P), degree = order - 1).
void pycs_mgpu_plan_rects(int N, int world, int rank, const int* kmin_east, int order, std::vector<MgRect>* out);

int k_mg_init(pycs_handle h, int rank, int world, unsigned char* handles_out);
int k_mg_connect(pycs_handle h, const unsigned char* all_handles);
int k_mg_replan(pycs_handle h);      // after the Lagrange tables changed
void k_mg_release(pycs_handle h);
int k_mg_wait_steps(pycs_handle h, cudaStream_t st);  // until every rank has completed as many steps as this one
// A flush rewrites the whole interior of the current Q array, peers' cells included.  Before a rank may
// store into that array of a peer outside the step protocol (k_fused_share_slab), the peer's flush must be
// over: every flush ends with k_mg_quiesce_raise, and the writer waits with k_mg_quiesce_wait.
int k_mg_quiesce_raise(pycs_handle h, cudaStream_t st);
int k_mg_quiesce_wait(pycs_handle h, cudaStream_t st);
// stream-ordered barrier over the ranks, on the device: what follows on the stream starts within a flag
// latency of the slowest rank reaching this point (pycs_run_timed: the ranks enter the timed steps together)
int k_mg_device_barrier(pycs_handle h, cudaStream_t st);
int k_mg_check(pycs_handle h);       // after a stream synchronisation: did a flag wait time out?
// after the boundary CTAs of a step wrote `qnext`: deliver the rectangles, then raise dflag on every peer
int k_mg_exchange(pycs_handle h, const double* qnext, StepCtl* ctl, cudaStream_t st);
void k_fused_reset_grid(pycs_handle h);
void k_fused_replan_exchange(pycs_handle h);

#ifdef __CUDACC__
// Wait until *f >= epoch.  A peer that is merely late (host work between runs, a slow rank) is waited
// for; a peer that is gone must not hang the GPU and must not kill the context either: after
// timeout_ns of wall clock (%globaltimer) the wait gives up, records it in *err and returns false --
// the kernel carries on with whatever data it has and the host reports PYCS_ERR_STATE.  Once *err is
// set every later wait returns at once, so a dead run drains quickly.
__device__ __forceinline__ unsigned long long mg_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// The flag is read with acquire semantics at system scope: what the peer stored before raising it (its rows,
// its MF-PR sum) is visible to the loads that follow, without a separate fence in every waiting CTA.
__device__ __forceinline__ long long mg_ld_acquire(const long long* f) {
  long long v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
  return v;
}
__device__ __forceinline__ bool mg_wait_flag(const volatile long long* fv, long long epoch, int* err,
                                             unsigned long long timeout_ns) {
  const long long* f = const_cast<const long long*>(fv);
  if (mg_ld_acquire(f) >= epoch) return true;
  if (*(volatile int*)err) return false;
  const unsigned long long t0 = mg_now_ns();
  for (;;) {
    if (mg_ld_acquire(f) >= epoch) return true;
    __nanosleep(32);
    if (mg_now_ns() - t0 > timeout_ns) {
      atomicExch(err, 1);
      return false;
    }
  }
}
#endif
