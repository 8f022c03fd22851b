// comm.h -- the one inter-rank step of the path, behind the C ABI (included by engine.cu).
//
// Replaces, for one process per GPU:
//   synchronize_bank                      src/eigenvalue.cpp:86-365
//   the global-tally / total-weight sums  src/eigenvalue.cpp:63, src/tallies/tally.cpp:1027-1053
//   reduce_tally_results                  src/tallies/tally.cpp:997-1025
//
// Transport: NCCL (libnccl.so.2, bound with dlopen so that a single-GPU process needs no NCCL
// at all) over the GPUs of one box, or three caller-supplied functions (omcgpu_comm_init_custom)
// -- what a libopenmc built with a CUDA-aware MPI would pass, and what the CPU test tier uses
// to run this very code over gloo.  Everything moves between DEVICE buffers; the host sees only
// the per-rank counts.
#pragma once

#if OMC_CUDA
#include <dlfcn.h>
#include <nccl.h>
#endif

namespace omc {

struct Transport {
  bool active = false;
  int rank = 0, size = 1;
  // custom transport
  omcgpu_allgather_i64_fn allgather = nullptr;
  omcgpu_sendrecv_fn sendrecv = nullptr;
  omcgpu_allreduce_f64_fn allreduce = nullptr;
  void* user = nullptr;
#if OMC_CUDA
  // NCCL, bound at run time
  void* lib = nullptr;
  ncclComm_t comm = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
#endif
  int64_t* d_i64 = nullptr; // device staging for the all-gathers: [size] send slots + [size * 4] receive
  int64_t* h_i64 = nullptr; // pinned
  double* d_f64 = nullptr;  // device staging for the globals
  double* h_f64 = nullptr;
  int64_t bytes_exchanged = 0;
};
static Transport T;

#if OMC_CUDA
static int nccl_fail(ncclResult_t r, const char* what)
{
  std::snprintf(g_err, sizeof(g_err), "NCCL error in %s: %s", what, T.GetErrorString ? T.GetErrorString(r) : "?");
  return OMCGPU_E_NCCL;
}
#define NCCL_OK(call, what) \
  do { \
    ncclResult_t r_ = (call); \
    if (r_ != ncclSuccess) \
      return nccl_fail(r_, what); \
  } while (0)

static int nccl_bind()
{
  if (T.lib)
    return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names)
    if ((T.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL)))
      break;
  if (!T.lib)
    return fail(OMCGPU_E_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define BIND(field, sym) \
  do { \
    *reinterpret_cast<void**>(&T.field) = dlsym(T.lib, sym); \
    if (!T.field) \
      return fail(OMCGPU_E_NCCL, "libnccl lacks %s", sym); \
  } while (0)
  BIND(GetUniqueId, "ncclGetUniqueId");
  BIND(CommInitRank, "ncclCommInitRank");
  BIND(CommDestroy, "ncclCommDestroy");
  BIND(GetErrorString, "ncclGetErrorString");
  BIND(AllGather, "ncclAllGather");
  BIND(AllReduce, "ncclAllReduce");
  BIND(Reduce, "ncclReduce");
  BIND(Send, "ncclSend");
  BIND(Recv, "ncclRecv");
  BIND(GroupStart, "ncclGroupStart");
  BIND(GroupEnd, "ncclGroupEnd");
#undef BIND
  return 0;
}
#endif

static int comm_buffers()
{
  if (T.d_i64)
    return 0;
  int rc = dev_alloc(&T.d_i64, (size_t)T.size * 8 + 8);
  rc = rc ? rc : dev_alloc(&T.d_f64, 64);
  if (rc)
    return rc;
  CUDA_OK(cudaMallocHost((void**)&T.h_i64, ((size_t)T.size * 8 + 8) * sizeof(int64_t)));
  CUDA_OK(cudaMallocHost((void**)&T.h_f64, 64 * sizeof(double)));
  return 0;
}

// every rank contributes n int64 values; recv = [size][n] on the host
static int comm_allgather(const int64_t* send, int n, int64_t* recv)
{
  if (T.allgather)
    return T.allgather(T.user, send, n, recv) ? fail(OMCGPU_E_NCCL, "custom all-gather failed%s") : 0;
#if OMC_CUDA
  if (n > 4)
    return fail(OMCGPU_E_INVALID_ARGUMENT, "comm_allgather: at most 4 values per rank%s");
  for (int k = 0; k < n; ++k)
    T.h_i64[k] = send[k];
  int64_t* d_send = T.d_i64;
  int64_t* d_recv = T.d_i64 + 8;
  CUDA_OK(cudaMemcpyAsync(d_send, T.h_i64, n * sizeof(int64_t), cudaMemcpyHostToDevice, G.stream));
  NCCL_OK(T.AllGather(d_send, d_recv, (size_t)n, ncclInt64, T.comm, G.stream), "ncclAllGather");
  CUDA_OK(cudaMemcpyAsync(T.h_i64 + 8, d_recv, (size_t)n * T.size * sizeof(int64_t), cudaMemcpyDeviceToHost, G.stream));
  CUDA_OK(cudaStreamSynchronize(G.stream));
  for (int k = 0; k < n * T.size; ++k)
    recv[k] = T.h_i64[8 + k];
  return 0;
#else
  return fail(OMCGPU_E_NCCL, "no transport%s");
#endif
}

struct P2P {
  int peer;
  bool send;
  void* buf;
  int64_t bytes;
};
static int comm_sendrecv(const std::vector<P2P>& ops)
{
  if (ops.empty())
    return 0;
  for (const P2P& o : ops)
    if (o.send)
      T.bytes_exchanged += o.bytes;
  if (T.sendrecv) {
    std::vector<int32_t> peer, is_send;
    std::vector<void*> buf;
    std::vector<int64_t> bytes;
    for (const P2P& o : ops) {
      peer.push_back(o.peer);
      is_send.push_back(o.send ? 1 : 0);
      buf.push_back(o.buf);
      bytes.push_back(o.bytes);
    }
    CUDA_OK(cudaStreamSynchronize(G.stream)); // a foreign transport does not know the engine's stream
    return T.sendrecv(T.user, (int)ops.size(), peer.data(), is_send.data(), buf.data(), bytes.data())
             ? fail(OMCGPU_E_NCCL, "custom send/receive failed%s")
             : 0;
  }
#if OMC_CUDA
  NCCL_OK(T.GroupStart(), "ncclGroupStart");
  for (const P2P& o : ops) {
    if (o.send)
      NCCL_OK(T.Send(o.buf, (size_t)o.bytes, ncclInt8, o.peer, T.comm, G.stream), "ncclSend");
    else
      NCCL_OK(T.Recv(o.buf, (size_t)o.bytes, ncclInt8, o.peer, T.comm, G.stream), "ncclRecv");
  }
  NCCL_OK(T.GroupEnd(), "ncclGroupEnd");
  return 0;
#else
  return fail(OMCGPU_E_NCCL, "no transport%s");
#endif
}

// in-place sum of a DEVICE array over the ranks; root < 0: every rank gets the sum, else only `root`
static int comm_allreduce(double* d_buf, int64_t n, int root)
{
  if (n <= 0)
    return 0;
  if (T.allreduce) {
    CUDA_OK(cudaStreamSynchronize(G.stream));
    return T.allreduce(T.user, d_buf, n, root) ? fail(OMCGPU_E_NCCL, "custom all-reduce failed%s") : 0;
  }
#if OMC_CUDA
  if (root < 0)
    NCCL_OK(T.AllReduce(d_buf, d_buf, (size_t)n, ncclDouble, ncclSum, T.comm, G.stream), "ncclAllReduce");
  else
    NCCL_OK(T.Reduce(d_buf, d_buf, (size_t)n, ncclDouble, ncclSum, root, T.comm, G.stream), "ncclReduce");
  return 0;
#else
  return fail(OMCGPU_E_NCCL, "no transport%s");
#endif
}

// bank rows [s0, s0 + n) of `s` -> rows [d0, d0 + n) of `d` (different banks)
struct RangeArgs {
  Bank d, s;
  int64_t d0, s0, n;
};
OMC_D void body_copy_range(int64_t tid, const RangeArgs& a)
{
  if (tid < a.n)
    bank_copy(a.d, a.d0 + tid, a.s, a.s0 + tid);
}
OMC_KERNEL(k_copy_range, RangeArgs, body_copy_range)

static void bank_column_ptrs(const Bank& b, char** cols, int* width)
{
  double* const dcols[9] = {b.x, b.y, b.z, b.u, b.v, b.w, b.E, b.time, b.wgt};
  for (int k = 0; k < 9; ++k) {
    cols[k] = reinterpret_cast<char*>(dcols[k]);
    width[k] = 8;
  }
  cols[9] = reinterpret_cast<char*>(b.delayed_group);
  width[9] = 4;
  cols[10] = reinterpret_cast<char*>(b.parent_id);
  width[10] = 8;
  cols[11] = reinterpret_cast<char*>(b.progeny_id);
  width[11] = 8;
}

} // namespace omc

using namespace omc;

extern "C" {

int omcgpu_comm_unique_id(char* id)
{
#if OMC_CUDA
  int rc = nccl_bind();
  if (rc)
    return rc;
  ncclUniqueId u;
  NCCL_OK(T.GetUniqueId(&u), "ncclGetUniqueId");
  std::memcpy(id, u.internal, OMCGPU_COMM_ID_BYTES);
  return 0;
#else
  (void)id;
  return fail(OMCGPU_E_NCCL, "NCCL is not part of the CPU test build%s");
#endif
}

int omcgpu_comm_finalize(void)
{
#if OMC_CUDA
  if (T.comm && T.CommDestroy)
    T.CommDestroy(T.comm);
#endif
  if (T.h_i64)
    cudaFreeHost(T.h_i64);
  if (T.h_f64)
    cudaFreeHost(T.h_f64);
  // (the device staging buffers belong to the engine's allocation list)
#if OMC_CUDA
  void* lib = T.lib; // keep the library bound
  T = Transport();
  T.lib = lib;
  if (lib)
    nccl_bind();
#else
  T = Transport();
#endif
  return 0;
}

int omcgpu_comm_init(const char* id, int rank, int n_ranks)
{
#if OMC_CUDA
  if (!G.ready)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_comm_init: call omcgpu_init first (the communicator lives on the engine's device)%s");
  if (!id || rank < 0 || rank >= n_ranks)
    return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_comm_init: bad arguments%s");
  if (T.active)
    omcgpu_comm_finalize();
  int rc = nccl_bind();
  if (rc)
    return rc;
  ncclUniqueId u;
  std::memcpy(u.internal, id, OMCGPU_COMM_ID_BYTES);
  NCCL_OK(T.CommInitRank(&T.comm, n_ranks, u, rank), "ncclCommInitRank");
  T.rank = rank;
  T.size = n_ranks;
  rc = comm_buffers();
  if (rc)
    return rc;
  T.active = true;
  // NCCL sets a point-to-point connection up on first use (tens to hundreds of ms per pair), and
  // which neighbours a rank trades sites with changes from generation to generation: connect to
  // every peer now, outside any timed generation
  std::vector<P2P> ops;
  for (int p = 0; p < n_ranks; ++p) {
    if (p == rank)
      continue;
    ops.push_back({p, true, T.d_f64, 8});
    ops.push_back({p, false, T.d_f64 + 8 + p % 32, 8});
  }
  rc = comm_sendrecv(ops);
  if (rc)
    return rc;
  CUDA_OK(cudaStreamSynchronize(G.stream));
  T.bytes_exchanged = 0;
  return 0;
#else
  (void)id;
  (void)rank;
  (void)n_ranks;
  return fail(OMCGPU_E_NCCL, "NCCL is not part of the CPU test build: use omcgpu_comm_init_custom%s");
#endif
}

int omcgpu_comm_init_custom(int rank, int n_ranks, omcgpu_allgather_i64_fn allgather, omcgpu_sendrecv_fn sendrecv,
  omcgpu_allreduce_f64_fn allreduce, void* user)
{
  if (!G.ready)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_comm_init_custom: call omcgpu_init first%s");
  if (rank < 0 || rank >= n_ranks || !allgather || !sendrecv || !allreduce)
    return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_comm_init_custom: bad arguments%s");
  if (T.active)
    omcgpu_comm_finalize();
  T.rank = rank;
  T.size = n_ranks;
  T.allgather = allgather;
  T.sendrecv = sendrecv;
  T.allreduce = allreduce;
  T.user = user;
  int rc = comm_buffers();
  if (rc)
    return rc;
  T.active = true;
  return 0;
}

int omcgpu_comm_allgather_i64(const int64_t* send, int32_t n, int64_t* recv)
{
  if (!T.active)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_comm_init has not been called%s");
  return comm_allgather(send, n, recv);
}

int64_t omcgpu_comm_bytes_exchanged(void)
{
  return T.bytes_exchanged;
}

// synchronize_bank, src/eigenvalue.cpp:86-365, on the device-resident banks of all ranks
int omcgpu_sync_bank(int64_t n_particles, const int64_t* work_index, int64_t seed_id, int64_t master_seed,
  uint64_t stride, double* globals, int32_t n_globals, omcgpu_site* source_bank, int64_t* n_source)
{
  if (!G.ready)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_init has not been called%s");
  if (!T.active)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_comm_init has not been called%s");
  if (n_globals < 0 || n_globals > 32)
    return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_sync_bank: at most 32 globals%s");
  const int R = T.size, me = T.rank;
  // (1) sizes of every rank's fission bank -> start offset and total (MPI_Exscan + MPI_Bcast, :100-112).
  //     A rank that banked nothing would stop alone while its peers wait in the exchange: its flag
  //     travels with the count, so that all ranks fail together.
  std::vector<int64_t> all((size_t)R * 2);
  const int64_t mine[2] = {G.n_fis, G.n_fis == 0 ? 1 : 0};
  int rc = comm_allgather(mine, 2, all.data());
  if (rc)
    return rc;
  int64_t start = 0, total = 0;
  bool empty_somewhere = false;
  for (int r = 0; r < R; ++r) {
    if (r < me)
      start += all[2 * r];
    total += all[2 * r];
    empty_somewhere = empty_somewhere || all[2 * r + 1] != 0;
  }
  if (empty_somewhere)
    return fail(OMCGPU_E_PHYSICS, "No fission sites banked on some rank%s");
  // (2) every rank combs its sorted fission bank with the SAME generation seed (:152-178)
  int64_t n_s = 0;
  int comb_rc = omcgpu_comb_fission_bank(start, total, n_particles, seed_id, master_seed, stride, &n_s);
  // (3) sampled counts of every rank -> bank_position (:186-197); a comb that overflowed on one rank
  //     is reported by all of them
  const int64_t mine2[2] = {n_s, comb_rc ? 1 : 0};
  char keep[sizeof(g_err)];
  std::memcpy(keep, g_err, sizeof(keep));
  rc = comm_allgather(mine2, 2, all.data());
  if (rc)
    return rc;
  std::vector<int64_t> pos((size_t)R + 1, 0);
  bool comb_failed = false;
  for (int r = 0; r < R; ++r) {
    pos[r + 1] = pos[r] + all[2 * r];
    comb_failed = comb_failed || all[2 * r + 1] != 0;
  }
  if (comb_failed) {
    if (comb_rc)
      std::memcpy(g_err, keep, sizeof(keep));
    else
      fail(OMCGPU_E_OUT_OF_BOUNDS, "the comb overflowed the source bank of another rank%s");
    return comb_rc ? comb_rc : OMCGPU_E_OUT_OF_BOUNDS;
  }
  if (pos[R] != n_particles)
    return fail(OMCGPU_E_PHYSICS, "comb did not select n_particles sites over all ranks%s");
  // (4) redistribute: rank r ends up with the global sites [work_index[r], work_index[r+1]) (:224-348).
  //     The new bank is assembled in the (now free) unsorted fission bank: contiguous column ranges
  //     go straight from the sampled bank to their owners, what stays is one range copy; then the
  //     two banks swap roles.
  const int64_t my_lo = pos[me], my_hi = pos[me + 1];
  const int64_t need_lo = work_index[me], need_hi = work_index[me + 1];
  const int64_t n_new = need_hi - need_lo;
  if (n_new > G.fis.cap)
    return fail(OMCGPU_E_OUT_OF_BOUNDS, "this rank's share of the source exceeds its bank%s");
  char *src_cols[12], *dst_cols[12];
  int width[12];
  bank_column_ptrs(G.src, src_cols, width);
  bank_column_ptrs(G.fis, dst_cols, width);
  std::vector<P2P> ops;
  for (int p = 0; p < R; ++p) {
    if (p == me)
      continue;
    int64_t lo = std::max(my_lo, work_index[p]), hi = std::min(my_hi, work_index[p + 1]); // what I hold that p owns
    if (hi > lo)
      for (int k = 0; k < 12; ++k)
        ops.push_back({p, true, src_cols[k] + (lo - my_lo) * width[k], (hi - lo) * width[k]});
    lo = std::max(pos[p], need_lo);
    hi = std::min(pos[p + 1], need_hi); // what p holds that I own
    if (hi > lo)
      for (int k = 0; k < 12; ++k)
        ops.push_back({p, false, dst_cols[k] + (lo - need_lo) * width[k], (hi - lo) * width[k]});
  }
  const int64_t keep_lo = std::max(my_lo, need_lo), keep_hi = std::min(my_hi, need_hi);
  if (keep_hi > keep_lo) {
    RangeArgs ra;
    ra.d = G.fis;
    ra.s = G.src;
    ra.d0 = keep_lo - need_lo;
    ra.s0 = keep_lo - my_lo;
    ra.n = keep_hi - keep_lo;
    OMC_LAUNCH(k_copy_range, ra.n, G.stream, ra);
  }
  rc = comm_sendrecv(ops);
  if (rc)
    return rc;
  // (5) the k / leakage / weight sums of the generation (MPI_Allreduce, src/eigenvalue.cpp:63)
  if (n_globals > 0 && globals) {
    for (int k = 0; k < n_globals; ++k)
      T.h_f64[k] = globals[k];
    CUDA_OK(cudaMemcpyAsync(T.d_f64, T.h_f64, n_globals * sizeof(double), cudaMemcpyHostToDevice, G.stream));
    rc = comm_allreduce(T.d_f64, n_globals, -1);
    if (rc)
      return rc;
    CUDA_OK(cudaMemcpyAsync(T.h_f64, T.d_f64, n_globals * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
  }
  std::swap(G.src, G.fis);
  G.n_src = n_new;
  if (n_source)
    *n_source = n_new;
  if (source_bank) { // mirror for the host's simulation::source_bank
    rc = fetch_bank(G.src, source_bank, n_new);
    if (rc)
      return rc;
  }
  CUDA_OK(cudaStreamSynchronize(G.stream));
  if (n_globals > 0 && globals)
    for (int k = 0; k < n_globals; ++k)
      globals[k] = T.h_f64[k];
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return fail(OMCGPU_E_CUDA, "kernel failure: %s", cudaGetErrorString(e));
  return 0;
}

// reduce_tally_results, src/tallies/tally.cpp:997-1025: the device VALUE arrays of all tallies are
// summed onto `root` (root < 0: onto every rank) and zeroed on the other ranks
int omcgpu_reduce_tallies(int root)
{
  if (!G.ready || !T.active)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_reduce_tallies: engine / communicator not initialised%s");
  if (G.n_tally == 0)
    return 0;
  double* v = nullptr;
  int64_t n = 0;
  int rc = omcgpu_tally_values_device(&v, &n); // folds the replicas
  if (rc)
    return rc;
  rc = comm_allreduce(v, n, root);
  if (rc)
    return rc;
  if (root >= 0 && root != T.rank)
    CUDA_OK(cudaMemsetAsync(v, 0, n * sizeof(double), G.stream));
  CUDA_OK(cudaStreamSynchronize(G.stream));
  return 0;
}

} // extern "C"
