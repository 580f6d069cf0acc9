// Multi-GPU state of a handle (see mgpu.cu).
#pragma once
#include "pycs_common.cuh"

#define MG_MAX_WORLD 8
#define MG_MAX_JOBS 40

struct MgSync {
  long long flag[MG_MAX_WORLD];      // flag[d] = last exchange epoch rank d delivered to this rank
  double psum[2][MG_MAX_WORLD];      // per-rank MF-PR sums, slot = epoch & 1
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
  int njobs;
  MgJob jobs[MG_MAX_JOBS];
};

void pycs_mgpu_rows(int N, int world, int rank, int* row_lo, int* row_hi);
int pycs_mgpu_plan_jobs(int N, int world, int rank, MgJob* jobs, int max_jobs);
int k_mg_init(pycs_handle h, int rank, int world, unsigned char* handles_out);
int k_mg_connect(pycs_handle h, const unsigned char* all_handles);
void k_mg_release(pycs_handle h);
int k_mg_wait(pycs_handle h);
const double* k_mg_sums(pycs_handle h);
int k_mg_exchange(pycs_handle h, const double* qnext, const double* part, int npart);
void k_fused_reset_grid(pycs_handle h);
struct FusedMg;
int k_mg_fill_args(pycs_handle h, const double* qnext, FusedMg* out);
