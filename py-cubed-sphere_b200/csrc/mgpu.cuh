// Multi-GPU state of a handle (see mgpu.cu).
#pragma once
#include "pycs_common.cuh"

#define MG_MAX_WORLD 8
#define MG_MAX_JOBS 40

struct MgSync {
  long long flag[MG_MAX_WORLD];      // flag[d] = last exchange epoch rank d delivered to this rank
  double psum[2][MG_MAX_WORLD];      // per-rank MF-PR sums, slot = epoch & 1
  int err;                           // != 0: a flag wait of this rank gave up (a peer is gone); the host turns it
                                     // into PYCS_ERR_STATE at the next synchronisation point (k_mg_check)
};
struct MgJob { int peer, i0, i1, j0, j1; };   // rectangle of every panel to store into rank `peer`

struct MgpuState {
  int rank, world, connected;
  double* alloc[2];                  // this rank's two Q allocations, in export order
  double* peer_q[2][MG_MAX_WORLD];   // peer_q[i][d]: allocation i of rank d, mapped here
  MgSync* sync;
  unsigned* counter;                 // blocks of the exchange kernel that have finished
  MgSync* peer_sync[MG_MAX_WORLD];
  long long epoch;                   // exchanges issued so far
  unsigned long long timeout_ns;     // wall-clock bound of a flag wait (PYCS_MG_TIMEOUT_S, default 30 s)
  int njobs;
  MgJob jobs[MG_MAX_JOBS];
};

void pycs_mgpu_rows(int N, int world, int rank, int* row_lo, int* row_hi);
int pycs_mgpu_plan_jobs(int N, int world, int rank, MgJob* jobs, int max_jobs);
int k_mg_init(pycs_handle h, int rank, int world, unsigned char* handles_out);
int k_mg_connect(pycs_handle h, const unsigned char* all_handles);
void k_mg_release(pycs_handle h);
int k_mg_wait(pycs_handle h);
int k_mg_check(pycs_handle h);       // after a stream synchronisation: did a flag wait time out?
const double* k_mg_sums(pycs_handle h);
int k_mg_exchange(pycs_handle h, const double* qnext, const double* part, int npart);
void k_fused_reset_grid(pycs_handle h);
struct FusedMg;
int k_mg_fill_args(pycs_handle h, const double* qnext, FusedMg* out);

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
__device__ __forceinline__ bool mg_wait_flag(const volatile long long* f, long long epoch, int* err,
                                             unsigned long long timeout_ns) {
  if (*f >= epoch) return true;
  if (*(volatile int*)err) return false;
  const unsigned long long t0 = mg_now_ns();
  for (;;) {
    if (*f >= epoch) return true;
    __nanosleep(32);
    if (mg_now_ns() - t0 > timeout_ns) {
      atomicExch(err, 1);
      return false;
    }
  }
}
#endif
