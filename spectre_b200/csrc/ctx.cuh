// Host-side types shared by the translation units of libdgrhs.so: the context,
// the per-N launcher table and small helpers.  dgrhs.cu holds the C-ABI and all
// N-independent host logic; per_n.cu is compiled once per number of grid points
// (-DDG_N=2..12, in parallel) and holds every kernel instantiation of that N.
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <vector>

#include "../../include/dgrhs.h"
#include "kernels.cuh"

extern "C" {
// shared error string / launch counter (defined in dgrhs.cu; not in the public header)
void dgrhs_internal_set_error(const char* msg);
void dgrhs_internal_count_launch(void);
}

namespace {

inline int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  dgrhs_internal_set_error(buf);
  return 1;
}

}  // namespace

#define CU(call)                                                            \
  do {                                                                      \
    cudaError_t err__ = (call);                                             \
    if (err__ != cudaSuccess)                                               \
      return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), \
                  __FILE__, __LINE__);                                      \
  } while (0)

#define CHECK_CTX(ctx) \
  if (!(ctx)) return fail("null context")

struct HistoryEntry {
  long long tick;
  int slot;
};

struct SubstepOp {
  enum Kind { kAbStep, kAbEvalOnly, kRestoreU0 } kind;
  int order;
  long long tick, tick_end;
  bool regular = false;  // a step of the evolution proper (not self-start)
};

struct dgrhs_ctx {
  int system = 0, N = 0, nelem = 0, nghost = 0, device = 0;
  int C = 0, S = 0, HC = 0, n = 0, npad = 0, f = 0;
  int n_interior = -1;
  int n_send = 0;  // faces packed for other ranks (0: no exchange needed)
  bool aligned_table_ok = true;
  cudaStream_t stream = nullptr;
  // side stream for the few-CTA, latency-bound Bjorhus kernel: it runs next to the
  // face kernel (disjoint corr slots) and joins before the volume kernel
  // set by launch_faces when the face kernel is the last thing queued: the volume
  // kernel that follows may be launched as its programmatic dependent
  bool pdl_volume = false;
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
  double *u = nullptr, *invjac = nullptr, *coords = nullptr, *stat = nullptr;
  double *corr = nullptr, *D = nullptr, *gH = nullptr, *gdH = nullptr;
  double *halo_send = nullptr, *halo_recv = nullptr;
  int32_t *nbr = nullptr, *halo_map = nullptr, *nbr_face = nullptr;
  std::vector<int32_t> nbr_host;
  std::vector<double*> dt_slots;  // derivative buffers (history ring)
  double* u0 = nullptr;           // saved value (self-start / RK step start)
  double* u_alt = nullptr;        // second state buffer for the fused update
  double* ctxbuf = nullptr;       // [E][26][npad] output of gh_context_kernel
  double* mesh_v = nullptr;       // [E][3][npad] inertial mesh velocity (moving mesh) or null
  void* lts = nullptr;            // local-time-stepping state (lts.cu)
  int lts_mode = 1;               // dgrhs_lts_set_mode
  double* filterF = nullptr;      // [N*N] exponential filter matrix (enabled if set)
  double filterF_host[144] = {};  // the same on the host (kernel parameter of the filter pass)
  int num_sms = 148;
  unsigned long long* violations = nullptr;  // DemandOutgoingCharSpeeds status (device)
  // ConstraintPreservingBjorhus faces (DGRHS_NEIGHBOR_BJORHUS in the neighbour table)
  int n_bjorhus_faces = 0;
  int64_t aux_faces_eval = -1;       // RHS evaluation that already ran the Bjorhus/mortar kernels
  int32_t* bjorhus_faces = nullptr;  // [n][3] element, direction, physical
  // The elements that own a Bjorhus face form the tail [bjorhus_tail_begin, nelem) of the
  // element order (-1: they do not): the volume kernel of the other elements then runs next to
  // the Bjorhus kernel and only the tail waits for it.  (An element list in the kernel
  // arguments instead of a range cost 170 B more spills per thread and 18 % of the fused
  // volume kernel's time at N = 12, profiles/README.md.)
  int bjorhus_tail_begin = -1;
  bool bjorhus_join_pending = false; // set by the face launcher, consumed by rhs_range
  // non-conforming mortars (dgrhs_set_mortars)
  int n_mortar_faces = 0;
  int n_mortar_faces_local = 0;      // groups without a remote side come first
  int32_t* mortar_faces = nullptr;   // [n_mortar_faces][4]
  int32_t* mortar_table = nullptr;   // [n_mortars][4]
  std::vector<int32_t> mortar_faces_host, mortar_table_host;  // the same on the host (lts.cu)
  double* mortar_P = nullptr;        // [3][N*N]
  double* mortar_R = nullptr;        // [3][N*N]
  // faces to a neighbour with a different N (another context): dgrhs_set_p_mortars
  int n_pmortar_faces = 0;
  int32_t* pm_faces = nullptr;       // [n][4] = element, direction, NB, neighbour direction | perm << 3
  double* pm_ghost = nullptr;        // [n][HC][144] the neighbours' faces
  double* pm_P = nullptr;            // [13][144]
  double* pm_R = nullptr;            // [13][144]
  cudaEvent_t pm_event = nullptr;
  int volume_variant = 0;         // 0 default, 1 context + streaming kernels (N <= 10),
                                  // 2 DFMA pair-staged kernel also for N = 12
  bool fuse_update = true;        // fuse UpdateU into the volume kernel
  dg::UpdateArgs pending_upd{};   // filled by begin_substep when fusing
  bool upd_active = false;
  double* dt_last = nullptr;
  int gauge = DGRHS_GAUGE_HARMONIC;
  double gauge_params[8] = {0};
  // stepping
  int stepper = DGRHS_STEPPER_ADAMS_BASHFORTH, order = 1;
  double t0 = 0.0, dt = 0.0;
  // exact slab bookkeeping (dgrhs_set_slab): first slab and steps per slab; 0 = t0 + k dt
  double slab_start = 0.0, slab_end = 0.0;
  int steps_per_slab = 0;
  long long tick_den = 1, step_index = 0;
  std::deque<HistoryEntry> history;
  std::deque<SubstepOp> pending;  // self-start program
  std::vector<int> free_slots;
  int rk_substep = 0;
  int cur_slot = -1;
  SubstepOp cur_op{};
  bool in_substep = false;
  int64_t rhs_evals = 0;
  int range_covered = 0;  // elements already evaluated in the current RHS (range calls)
  // halo exchange inside the library (dgrhs_comm_init): NCCL communicator of the
  // ranks that share the domain, one send/recv pair per peer and RHS on its own stream
  void* nccl_comm = nullptr;  // ncclComm_t
  int comm_rank = 0, comm_world = 1;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_packed = nullptr, ev_halo = nullptr, ev_faces1 = nullptr, ev_faces2 = nullptr;
  std::vector<int> send_counts, recv_counts;  // faces per peer (rank order)
  // optional event timeline of the multi-GPU schedule (dgrhs_set_phase_timing)
  bool phase_timing = false;
  cudaEvent_t phase_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t state_len() const { return (size_t)nelem * C * npad; }
};

namespace {

template <typename T>
int dev_alloc(T** p, size_t count) {
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  // experiment (profiles/README.md, round 2): DGRHS_ALLOC_SKEW=<bytes> staggers the start
  // addresses of the large arrays by k * skew (the arrays are then never freed)
  static const long skew = std::getenv("DGRHS_ALLOC_SKEW") ? std::atol(std::getenv("DGRHS_ALLOC_SKEW")) : 0;
  static int counter = 0;
  if (skew > 0 && bytes > (size_t(64) << 20)) {
    char* base = nullptr;
    CU(cudaMalloc((void**)&base, bytes + 16 * (size_t)skew));
    *p = reinterpret_cast<T*>(base + (size_t)(counter++ % 16) * (size_t)skew);
  } else {
    CU(cudaMalloc((void**)p, bytes));
  }
  CU(cudaMemset(*p, 0, bytes));
  // cudaMemset on device memory is asynchronous and runs on the legacy default stream, which
  // the context's non-blocking streams do not wait for: finish it before anyone can queue
  // work on the new buffer (allocation happens at set-up time, never inside a step)
  CU(cudaStreamSynchronize(cudaStreamLegacy));
  return 0;
}


// Blocking host-to-device copy of a set-up table.  cudaMemcpy from pageable memory may return
// while the DMA is still in flight on the legacy default stream, which the context's
// non-blocking streams do not wait for: finish it before a kernel can be queued.
inline cudaError_t h2d_table(void* dst, const void* src, size_t bytes) {
  cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(cudaStreamLegacy);
}

// DGRHS_NO_PDL=1 in the environment turns programmatic dependent launch off
static const bool g_pdl = [] {
  const char* v = std::getenv("DGRHS_NO_PDL");
  return !(v && v[0] == '1');
}();

// kernel<<<blocks, threads, smem, stream>>>(args), optionally as the programmatic
// dependent of the kernel queued before it
template <typename Kernel, typename Args>
cudaError_t launch_dependent(Kernel k, int blocks, int threads, size_t smem, cudaStream_t stream,
                             bool pdl, const Args& args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k, args);
}

}  // namespace

// Launchers of one N (per_n.cu).  Every entry queues work on c->stream and
// returns 0 or 1 (error text in dgrhs_last_error()).
struct DgNOps {
  int (*faces)(dgrhs_ctx* c, int eb, int ee);
  int (*gauge)(dgrhs_ctx* c, double time);
  int (*volume)(dgrhs_ctx* c, double* dt, int eb, int ee, bool with_corr,
                const dg::UpdateArgs* upd);
  int (*pack)(dgrhs_ctx* c);
  int (*filter)(dgrhs_ctx* c);
  int (*gauge_from_state)(dgrhs_ctx* c, const double* state_dev);
  int (*constraints)(dgrhs_ctx* c, double* sums_dev);
  int (*partial_derivatives)(const dg::DerivArgs* a, int blocks, cudaStream_t stream);
  int (*mesh_velocity_terms)(dgrhs_ctx* c, double* dt, int eb, int ee);
  // local time stepping (lts.cu): volume part + external boundary conditions of a range;
  // face snapshot; boundary deltas of the elements that finish a step
  int (*lts_evaluate)(dgrhs_ctx* c, const int32_t* nbr_external, const uint8_t* mortar_skip,
                      double* dt, int eb, int ee, const dg::UpdateArgs* upd);
  int (*lts_snapshot)(dgrhs_ctx* c, double* fh, const uint8_t* in_history, int depth, int slot,
                      int eb, int ee);
  int (*lts_boundary)(dgrhs_ctx* c, const dg::LtsBoundaryArgs* a);
  int (*lts_mortar)(dgrhs_ctx* c, const dg::LtsMortarArgs* a, int n_groups);
  // exponential filter on ntiles component blocks starting at u (a range of elements)
  int (*filter_range)(dgrhs_ctx* c, double* u, int ntiles);
};
const DgNOps* dgrhs_nops(int N);  // nullptr for an unsupported N

// shared with lts.cu (defined in dgrhs.cu)
// Adams-Bashforth coefficients of history times given in integer ticks of tick_size
// (adams_coefficients::coefficients, AdamsCoefficients.hpp:64-104)
std::vector<double> dgrhs_internal_ab_coefficients_ticks(const std::vector<long long>& ticks,
                                                         long long start, long long end,
                                                         double tick_size);
// u[0, len) = a u + sum_j coef_j v_j[0, len) on the context's stream (len even)
int dgrhs_internal_lincomb_range(dgrhs_ctx* c, double* u, double a,
                                 const std::vector<double>& coef,
                                 const std::vector<const double*>& v, size_t len);
int dgrhs_internal_upload(dgrhs_ctx* c, double* dst, const double* src, int ncomp);
void dgrhs_internal_lts_free(dgrhs_ctx* c);
