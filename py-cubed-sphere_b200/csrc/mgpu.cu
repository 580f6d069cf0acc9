// Multi-GPU layer of the fused step: row-slab decomposition with peer-mapped halo stores.
//
// The reference is single process ("new design").  Rank k of W owns the rows
// [row_lo, row_hi) of all six panels.  One step needs, besides the own rows,
//   * the 3 rows above and below the slab (the 7x7 dependence box of the split scheme), and
//   * the four 4-wide boundary strips of every panel, which feed the Lagrange ghost fill
//     (src/halo_data.py:15-185 gathers exactly these) -- every rank then fills all ghost
//     cells it needs locally with the unchanged ghost-fill kernels;
//   * one scalar per rank for the MF-PR projection (src/discrete_operators.py:98-101).
// After its step kernel each rank stores those pieces of its new rows directly into the
// peers' Q arrays at their natural positions (CUDA IPC mappings over NVLink, no staging, no
// NCCL call on the data path), publishes its partial sum and raises a per-rank flag with
// system-scope release; the next step starts with a kernel that spins on the W flags.
// Everything is stream ordered on the handle's stream: there is no host synchronisation in
// the step loop.  Flags are epoch counters, so buffers (the two ping-pong Q arrays) are
// reused without resets; a peer can be at most one step ahead.
#include <cstdlib>
#include <cstring>
#include "pycs_common.cuh"
#include "mgpu.cuh"
#include "fused_args.cuh"

namespace {

__global__ void mg_wait_kernel(MgSync* sync, int world, long long epoch, unsigned long long timeout_ns) {
  const int d = threadIdx.x;
  if (d >= world) return;
  mg_wait_flag(&sync->flag[d], epoch, &sync->err, timeout_ns);
  __threadfence_system();
}

struct ScatterJob { double* dst; int i0, i1, j0, j1; };
struct ScatterJobs { int n; ScatterJob job[MG_MAX_JOBS]; };
struct PeerSync { MgSync* s[MG_MAX_WORLD]; };

// One launch per step, one CTA per (job, panel): (1) copy the rectangle [i0,i1) x [j0,j1) of
// one panel of src into the same position of a peer's array (block 0 also delivers this
// rank's MF-PR sum); (2) system-scope fence, ticket; (3) the last CTA raises the flag on every
// rank.  Latency, not bandwidth, is what matters here (<= 3 MB per rank and step).
__global__ void mg_exchange_kernel(Geo g, ScatterJobs jobs, const double* __restrict__ src, PeerSync peers,
                                   int world, int rank, const double* __restrict__ sum, int parity,
                                   long long epoch, unsigned* __restrict__ counter) {
  __shared__ int last;
  {
    const ScatterJob jb = jobs.job[blockIdx.y];
    const int w = jb.j1 - jb.j0, n = (jb.i1 - jb.i0) * w;
    const int p = blockIdx.x;
    if (((w | jb.j0) & 1) == 0) {             // 16-byte aligned rectangle: two cells per store
      const int w2 = w >> 1, n2 = n >> 1;
      for (int t = threadIdx.x; t < n2; t += blockDim.x) {
        const int i = jb.i0 + t / w2, j = jb.j0 + 2 * (t % w2);
        const long long id = gidx(g, p, i, j);
        *reinterpret_cast<double2*>(jb.dst + id) = *reinterpret_cast<const double2*>(src + id);
      }
    } else {
      for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int i = jb.i0 + t / w, j = jb.j0 + t % w;
        const long long id = gidx(g, p, i, j);
        jb.dst[id] = src[id];
      }
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < world)
    peers.s[threadIdx.x]->psum[parity][rank] = *sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    // one system-scope fence per CTA: it is cumulative over the stores the barrier ordered before it
    __threadfence_system();
    last = (atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) *counter = 0;
  if (threadIdx.x < world) {
    __threadfence_system();
    *((volatile long long*)&peers.s[threadIdx.x]->flag[rank]) = epoch;
  }
}

}  // namespace

// ---- host-side plan (no GPU needed: exercised by the CPU tests) ----------------------------
void pycs_mgpu_rows(int N, int world, int rank, int* row_lo, int* row_hi) {
  const int base = N / world, rem = N % world;
  const int off = rank * base + (rank < rem ? rank : rem);
  *row_lo = PYCS_NG + off;
  *row_hi = *row_lo + base + (rank < rem ? 1 : 0);
}

int pycs_mgpu_plan_jobs(int N, int world, int rank, MgJob* jobs, int max_jobs) {
  const int lo = PYCS_NG, hi = PYCS_NG + N;
  int a, b, n = 0;
  pycs_mgpu_rows(N, world, rank, &a, &b);
  auto add = [&](int peer, int i0, int i1, int j0, int j1) {
    if (n < max_jobs) jobs[n] = MgJob{peer, i0, i1, j0, j1};
    ++n;
  };
  for (int d = 0; d < world; ++d) {
    if (d == rank) continue;
    add(d, a, b, lo, lo + PYCS_NG);                           // S strips of my rows
    add(d, a, b, hi - PYCS_NG, hi);                           // N strips
    if (rank == 0) add(d, lo, lo + PYCS_NG, lo, hi);          // W strips (first slab)
    if (rank == world - 1) add(d, hi - PYCS_NG, hi, lo, hi);  // E strips (last slab)
    if (d == rank - 1) add(d, a, a + 3, lo, hi);              // lower neighbour's halo rows
    if (d == rank + 1) add(d, b - 3, b, lo, hi);              // upper neighbour's halo rows
  }
  return n;
}

// ---- device-side state ---------------------------------------------------------------------
int k_mg_init(pycs_handle h, int rank, int world, unsigned char* handles_out) {
  if (world < 2 || world > MG_MAX_WORLD || rank < 0 || rank >= world) {
    pycs_set_error("pycs_mgpu_init: world must be 2..8 and 0 <= rank < world");
    return PYCS_ERR_ARG;
  }
  if (h->g.N / world < PYCS_NG) {
    pycs_set_error("pycs_mgpu_init: fewer than 4 rows per rank");
    return PYCS_ERR_ARG;
  }
  if (h->mg) {
    pycs_set_error("pycs_mgpu_init: already initialised");
    return PYCS_ERR_STATE;
  }
  MgpuState* mg = new MgpuState();
  memset(mg, 0, sizeof *mg);
  mg->rank = rank;
  mg->world = world;
  const char* et = getenv("PYCS_MG_TIMEOUT_S");
  const double tsec = et ? atof(et) : 30.0;
  mg->timeout_ns = (unsigned long long)((tsec > 0.001 ? tsec : 30.0) * 1e9);
  TRY(pycs_field_ptr(h, PYCS_F_Q, &mg->alloc[0]));
  TRY(pycs_field_ptr(h, PYCS_F_Q_NEXT, &mg->alloc[1]));
  CK(cudaMalloc(&mg->sync, sizeof(MgSync)));
  CK(cudaMemset(mg->sync, 0, sizeof(MgSync)));
  CK(cudaMalloc(&mg->counter, sizeof(unsigned)));
  CK(cudaMemset(mg->counter, 0, sizeof(unsigned)));
  CK(cudaStreamSynchronize(h->stream));
  cudaIpcMemHandle_t hd[3];
  CK(cudaIpcGetMemHandle(&hd[0], mg->alloc[0]));
  CK(cudaIpcGetMemHandle(&hd[1], mg->alloc[1]));
  CK(cudaIpcGetMemHandle(&hd[2], mg->sync));
  memcpy(handles_out, hd, sizeof hd);
  pycs_mgpu_rows(h->g.N, world, rank, &h->row_lo, &h->row_hi);
  MgJob jobs[MG_MAX_JOBS];
  mg->njobs = pycs_mgpu_plan_jobs(h->g.N, world, rank, jobs, MG_MAX_JOBS);
  if (mg->njobs > MG_MAX_JOBS) {
    delete mg;
    pycs_set_error("pycs_mgpu_init: too many scatter jobs");
    return PYCS_ERR_ARG;
  }
  memcpy(mg->jobs, jobs, sizeof(MgJob) * mg->njobs);
  h->mg = mg;
  k_fused_reset_grid(h);
  return 0;
}

int k_mg_connect(pycs_handle h, const unsigned char* all_handles) {
  MgpuState* mg = h->mg;
  if (!mg) {
    pycs_set_error("pycs_mgpu_connect before pycs_mgpu_init");
    return PYCS_ERR_STATE;
  }
  for (int d = 0; d < mg->world; ++d) {
    if (d == mg->rank) {
      mg->peer_q[0][d] = mg->alloc[0];
      mg->peer_q[1][d] = mg->alloc[1];
      mg->peer_sync[d] = mg->sync;
      continue;
    }
    cudaIpcMemHandle_t hd[3];
    memcpy(hd, all_handles + (size_t)d * sizeof hd, sizeof hd);
    void* p[3];
    for (int k = 0; k < 3; ++k) CK(cudaIpcOpenMemHandle(&p[k], hd[k], cudaIpcMemLazyEnablePeerAccess));
    mg->peer_q[0][d] = (double*)p[0];
    mg->peer_q[1][d] = (double*)p[1];
    mg->peer_sync[d] = (MgSync*)p[2];
  }
  mg->connected = 1;
  return 0;
}

void k_mg_release(pycs_handle h) {
  MgpuState* mg = h->mg;
  if (!mg) return;
  for (int d = 0; d < mg->world; ++d) {
    if (d == mg->rank || !mg->connected) continue;
    cudaIpcCloseMemHandle(mg->peer_q[0][d]);
    cudaIpcCloseMemHandle(mg->peer_q[1][d]);
    cudaIpcCloseMemHandle(mg->peer_sync[d]);
  }
  if (mg->sync) cudaFree(mg->sync);
  if (mg->counter) cudaFree(mg->counter);
  delete mg;
  h->mg = nullptr;
}

// wait until every rank has delivered the data of exchange `epoch` (stream ordered)
int k_mg_wait(pycs_handle h) {
  MgpuState* mg = h->mg;
  mg_wait_kernel<<<1, 32, 0, h->stream>>>(mg->sync, mg->world, mg->epoch, mg->timeout_ns);
  CKL(h);
  return 0;
}

// Host side of the bounded wait: call after the stream has been synchronised.
int k_mg_check(pycs_handle h) {
  MgpuState* mg = h->mg;
  if (!mg) return 0;
  int e = 0;
  CK(cudaMemcpy(&e, &mg->sync->err, sizeof(int), cudaMemcpyDeviceToHost));
  if (e) {
    pycs_set_error("multi-GPU step: a peer's flag did not arrive within PYCS_MG_TIMEOUT_S (a rank is gone or "
                   "never entered the run); the state of this handle is undefined");
    return PYCS_ERR_STATE;
  }
  return 0;
}

// where the W per-rank MF-PR sums of the last exchange live (this rank's copy)
const double* k_mg_sums(pycs_handle h) { return h->mg->sync->psum[h->mg->epoch & 1]; }

// after the step kernel wrote the own rows of `qnext`: deliver halos, sum (*part, one scalar) and flag
int k_mg_exchange(pycs_handle h, const double* qnext, const double* part, int npart) {
  MgpuState* mg = h->mg;
  if (!mg->connected) {
    pycs_set_error("multi-GPU step before pycs_mgpu_connect");
    return PYCS_ERR_STATE;
  }
  const int idx = (qnext == mg->alloc[0]) ? 0 : 1;
  ScatterJobs js;
  js.n = mg->njobs;
  int nmax = 1;
  for (int k = 0; k < mg->njobs; ++k) {
    const MgJob& j = mg->jobs[k];
    js.job[k] = ScatterJob{mg->peer_q[idx][j.peer], j.i0, j.i1, j.j0, j.j1};
    const int n = (j.i1 - j.i0) * (j.j1 - j.j0);
    if (n > nmax) nmax = n;
  }
  (void)nmax;
  (void)npart;
  PeerSync ps;
  for (int d = 0; d < mg->world; ++d) ps.s[d] = mg->peer_sync[d];
  mg->epoch += 1;
  mg_exchange_kernel<<<dim3(6, mg->njobs), 512, 0, h->stream>>>(h->g, js, qnext, ps, mg->world, mg->rank, part,
                                                                (int)(mg->epoch & 1), mg->epoch, mg->counter);
  CKL(h);
  return 0;
}

// The same exchange done by the step kernel itself (fused2b.cu): fill the multi-GPU part of its
// arguments for a launch that writes `qnext`, and account for the exchange it will perform.
int k_mg_fill_args(pycs_handle h, const double* qnext, FusedMg* out) {
  MgpuState* mg = h->mg;
  if (!mg->connected) {
    pycs_set_error("multi-GPU step before pycs_mgpu_connect");
    return PYCS_ERR_STATE;
  }
  const int idx = (qnext == mg->alloc[0]) ? 0 : 1;
  mg->epoch += 1;
  out->world = mg->world;
  out->rank = mg->rank;
  out->parity = (int)(mg->epoch & 1);
  out->epoch = mg->epoch;
  for (int d = 0; d < MG_MAX_WORLD; ++d) {
    out->peer_qn[d] = d < mg->world ? mg->peer_q[idx][d] : nullptr;
    out->peer_sync[d] = d < mg->world ? mg->peer_sync[d] : nullptr;
  }
  return 0;
}
