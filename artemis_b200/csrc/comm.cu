// comm.cu -- the multi-rank half of the hot path behind the C ABI: the single-round remote
// ghost exchange (planner + aggregated buffers + NCCL send/recv over NVLink) and the global dt
// all-reduce, so a C++ Parthenon host gets them without Python / torch.
//
// Replaces, for blocks whose neighbour lives on another rank,
//   parthenon::SendBoundBufs / ReceiveBoundBufs / SetBounds   P:bvals/comms/boundary_communication.cpp:48-334
//   CommBuffer::Send / TryReceive (MPI_Isend / Irecv / Iprobe / Test)  P:utils/communication_buffer.hpp:209-420
//   MPI_Allreduce(&dt, 1, MPI_DOUBLE, MPI_MIN)                P:driver/driver.cpp:237
// Where the reference posts one message per (block, neighbour, variable), this sends ONE
// aggregated message per peer rank and stage: faces, rank edges and the rank corner at once
// (<= 26 peers, 7 in a 2x2x2 lattice), one pack launch, one NCCL group, one unpack launch, on a
// library-owned stream ordered against the stage kernels with events -- no host polling.
//
// NCCL is bound at run time (dlopen, preferring a copy the process already loaded -- under
// PyTorch that is torch's bundled libnccl.so.2), so libartemis_b200.so has no link-time
// dependency on it and single-rank hosts never touch it.
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>

#include "ab200_ctx.cuh"

namespace ab200 {

// ---- minimal NCCL surface (types restated so no header is needed at build time) -----------
struct NcclUniqueId { char internal[128]; };
typedef void *NcclComm;
enum { kNcclFloat64 = 8, kNcclMin = 3 };

struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(NcclUniqueId *) = nullptr;
  int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};

static NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.h ? &api : nullptr;
  tried = true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    api.h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // reuse the process's copy
    if (api.h) break;
  }
  for (const char *n : names) {
    if (api.h) break;
    api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
  }
  if (!api.h) return nullptr;
#define AB_SYM(field, name)                                                       \
  *(void **)(&api.field) = dlsym(api.h, name);                                    \
  if (!api.field) { api.h = nullptr; return nullptr; }
  AB_SYM(GetUniqueId, "ncclGetUniqueId")
  AB_SYM(CommInitRank, "ncclCommInitRank")
  AB_SYM(CommDestroy, "ncclCommDestroy")
  AB_SYM(AllReduce, "ncclAllReduce")
  AB_SYM(AllGather, "ncclAllGather")
  AB_SYM(Send, "ncclSend")
  AB_SYM(Recv, "ncclRecv")
  AB_SYM(GroupStart, "ncclGroupStart")
  AB_SYM(GroupEnd, "ncclGroupEnd")
  AB_SYM(GetErrorString, "ncclGetErrorString")
#undef AB_SYM
  return &api;
}

#define AB_NCCL(call)                                                                      \
  do {                                                                                     \
    int e__ = (call);                                                                      \
    if (e__ != 0) {                                                                        \
      set_error(std::string("NCCL: ") + #call + ": " + nccl_api()->GetErrorString(e__));   \
      return AB200_ECUDA;                                                                  \
    }                                                                                      \
  } while (0)

// ---- planner ------------------------------------------------------------------------------
// Single-round exchange (restates artemis_b200/comm.py:plan_direct, which the gloo tests
// cover; tests/test_comm_plan.py requires the two to agree row for row).  For every neighbour
// offset o in {-1,0,1}^3 whose non-zero directions ALL cross onto another rank, the blocks on
// that rank face / edge / corner send their ng innermost layers along the non-zero directions
// and their INTERIOR range along the others (same-level branch of CalcIndices,
// P:bvals/comms/bnd_info.cpp:152-213).  Both ends walk the offsets in one canonical order (the
// receiver in -o), blocks lexicographically, so the aggregated buffers line up without tags.
struct PlanRow {
  int peer, recv, fluid, block, var0, ncomp, si, ei, sj, ej, sk, ek;
  long long offset;  // doubles from the start of the peer's send / receive message
};

static void plan_direct(const int nbd[3], const int nt[3], const int s[3], const int e[3],
                        const int ng[3], int nfl, const int *fl_type, const int *fl_S,
                        const int lay[3], const int rl[3], const int periodic[3],
                        std::vector<PlanRow> &rows, std::map<int, std::pair<long long, long long>> &sizes) {
  rows.clear();
  sizes.clear();
  auto peer_of = [&](const int o[3]) -> int {
    int prc[3] = {rl[0], rl[1], rl[2]};
    for (int d = 0; d < 3; ++d) {
      if (o[d] == 0) continue;
      if (nt[d] == 1 || lay[d] == 1) return -1;
      prc[d] += o[d];
      if (prc[d] < 0 || prc[d] >= lay[d]) {
        if (!periodic[d]) return -1;
        prc[d] = (prc[d] + lay[d]) % lay[d];
      }
    }
    return prc[0] + lay[0] * (prc[1] + lay[1] * prc[2]);
  };
  auto add = [&](const int o[3], bool sending) {
    const int peer = peer_of(o);
    if (peer < 0) return;
    auto &sz = sizes[peer];
    int lo_l[3], hi_l[3];
    for (int d = 0; d < 3; ++d) {
      if (o[d] == 0) { lo_l[d] = 0; hi_l[d] = nbd[d] - 1; }
      else { lo_l[d] = hi_l[d] = o[d] > 0 ? nbd[d] - 1 : 0; }
    }
    for (int lz = lo_l[2]; lz <= hi_l[2]; ++lz)
      for (int ly = lo_l[1]; ly <= hi_l[1]; ++ly)
        for (int lx = lo_l[0]; lx <= hi_l[0]; ++lx) {
          const int b = lx + nbd[0] * (ly + nbd[1] * lz);
          int lo[3], hi[3];
          long long ncell = 1;
          for (int d = 0; d < 3; ++d) {
            if (o[d] == 0) { lo[d] = s[d]; hi[d] = e[d]; }
            else if (sending) {
              if (o[d] > 0) { lo[d] = e[d] - ng[d] + 1; hi[d] = e[d]; }
              else { lo[d] = s[d]; hi[d] = s[d] + ng[d] - 1; }
            } else {
              if (o[d] > 0) { lo[d] = e[d] + 1; hi[d] = e[d] + ng[d]; }
              else { lo[d] = s[d] - ng[d]; hi[d] = s[d] - 1; }
            }
            ncell *= hi[d] - lo[d] + 1;
          }
          for (int q = 0; q < nfl; ++q) {
            // FillGhost pack entries: gas prim rho, v, sie (pressure is not exchanged,
            // src/gas/gas.cpp:243-270); dust prim rho, v
            const int S = fl_S[q];
            const int runs[2][2] = {{0, 4 * S}, {5 * S, S}};
            const int nruns = fl_type[q] == AB200_GAS ? 2 : 1;
            for (int r = 0; r < nruns; ++r) {
              long long &off = sending ? sz.first : sz.second;
              rows.push_back({peer, sending ? 0 : 1, fl_type[q], b, runs[r][0], runs[r][1], lo[0],
                              hi[0], lo[1], hi[1], lo[2], hi[2], off});
              off += (long long)runs[r][1] * ncell;
            }
          }
        }
  };
  for (int pass = 0; pass < 2; ++pass)
    for (int oz = -1; oz <= 1; ++oz)
      for (int oy = -1; oy <= 1; ++oy)
        for (int ox = -1; ox <= 1; ++ox) {
          if (!ox && !oy && !oz) continue;
          const int o[3] = {pass ? -ox : ox, pass ? -oy : oy, pass ? -oz : oz};
          add(o, pass == 0);
        }
}

struct CommState {
  NcclComm comm = nullptr;
  int nranks = 1, rank = 0;
  cudaStream_t stream = nullptr;          // library-owned: pack, NCCL group, unpack
  cudaEvent_t ev_stage = nullptr, ev_done = nullptr;
  bool planned = false, in_flight = false;
  struct Peer { int rank; long long soff, nsend, roff, nrecv; };
  std::vector<Peer> peers;
  double *dsend = nullptr, *drecv = nullptr;
  std::vector<ab200_bnd_desc> send_desc, recv_desc;
  long long bytes_per_exchange = 0;
  // ---- direct transport: the pack kernel stores straight into the PEER's receive slab over
  // NVLink (CUDA IPC mappings exchanged once), a flag per peer replaces the NCCL round
  bool direct = false;
  double *rslab[2] = {nullptr, nullptr};          // local receive slabs, alternating by step
  unsigned long long *flags = nullptr;            // local [nranks]: last step each peer delivered
  int *d_err = nullptr;                           // set by a wait that timed out
  std::vector<double *> peer_rslab[2];            // [peer index] mapped receive slabs of the peers
  std::vector<unsigned long long *> peer_flags;   // [peer index] mapped flag arrays of the peers
  std::vector<void *> ipc_opened;
  std::vector<ab200_bnd_desc> dsend_desc[2], drecv_desc[2];
  int *d_peer_rank = nullptr;                     // device copy of the peer ranks
  unsigned long long **d_peer_flags = nullptr;    // device copy of peer_flags
  unsigned long long step = 0;
};

// signal: my step counter into every peer's flag array (after the pack kernel of this stream
// has completed, i.e. its remote stores are performed); wait: until every peer delivered `step`
__global__ void k_comm_signal(unsigned long long *const *peer_flags, int npeers, int my_rank,
                              unsigned long long step) {
  const int t = threadIdx.x;
  if (t < npeers) {
    __threadfence_system();
    *((volatile unsigned long long *)(peer_flags[t] + my_rank)) = step;
    __threadfence_system();
  }
}
__global__ void k_comm_wait(const unsigned long long *flags, const int *peer_rank, int npeers,
                            unsigned long long step, int *err) {
  const int t = threadIdx.x;
  if (t < npeers) {
    const volatile unsigned long long *f = flags + peer_rank[t];
    long long spins = 0;
    while (*f < step) {
      __nanosleep(200);
      if (++spins > 20000000LL) {  // ~ 4 s: a peer died; do not hang the GPU
        *err = 1;
        break;
      }
    }
    __threadfence_system();
  }
}

static CommState *cs_of(ab200_ctx *c) { return static_cast<CommState *>(c->comm_state); }

}  // namespace ab200

using namespace ab200;

extern "C" {

int ab200_comm_unique_id(char *id128) {
  AB_REQUIRE(id128, AB200_EINVAL, "ab200_comm_unique_id: null output");
  NcclApi *n = nccl_api();
  AB_REQUIRE(n, AB200_ESTATE, "ab200_comm_unique_id: libnccl.so.2 could not be loaded");
  NcclUniqueId id;
  AB_NCCL(n->GetUniqueId(&id));
  memcpy(id128, id.internal, 128);
  return AB200_OK;
}

int ab200_comm_init(ab200_ctx *c, int nranks, int rank, const char *id128) {
  AB_REQUIRE(c && id128, AB200_EINVAL, "ab200_comm_init: null argument");
  AB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, AB200_EINVAL, "ab200_comm_init: bad rank");
  AB_REQUIRE(!c->comm_state, AB200_ESTATE, "ab200_comm_init: communicator already initialised");
  NcclApi *n = nccl_api();
  AB_REQUIRE(n, AB200_ESTATE, "ab200_comm_init: libnccl.so.2 could not be loaded");
  AB_CUDA(cudaSetDevice(c->device));
  CommState *cs = new CommState;
  cs->nranks = nranks;
  cs->rank = rank;
  NcclUniqueId id;
  memcpy(id.internal, id128, 128);
  int e = n->CommInitRank(&cs->comm, nranks, id, rank);
  if (e != 0) {
    set_error(std::string("NCCL: ncclCommInitRank: ") + n->GetErrorString(e));
    delete cs;
    return AB200_ECUDA;
  }
  // default priority; AB200_COMM_HIGH_PRIO=1 raises it (measured: it makes the overlapped cycle
  // slower, see ab200_run_cycles_mr)
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  cudaStreamCreateWithPriority(&cs->stream, cudaStreamNonBlocking,
                               getenv("AB200_COMM_HIGH_PRIO") ? prio_hi : prio_lo);
  cudaEventCreateWithFlags(&cs->ev_stage, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&cs->ev_done, cudaEventDisableTiming);
  c->comm_state = cs;
  return AB200_OK;
}

int ab200_comm_destroy(ab200_ctx *c) {
  AB_REQUIRE(c, AB200_EINVAL, "null context");
  CommState *cs = cs_of(c);
  if (!cs) return AB200_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(cs->stream);
  if (cs->comm) nccl_api()->CommDestroy(cs->comm);
  if (cs->dsend) cudaFree(cs->dsend);
  if (cs->drecv) cudaFree(cs->drecv);
  for (void *p : cs->ipc_opened) cudaIpcCloseMemHandle(p);
  for (int q = 0; q < 2; ++q)
    if (cs->rslab[q]) cudaFree(cs->rslab[q]);
  if (cs->flags) cudaFree(cs->flags);
  if (cs->d_err) cudaFree(cs->d_err);
  if (cs->d_peer_rank) cudaFree(cs->d_peer_rank);
  if (cs->d_peer_flags) cudaFree(cs->d_peer_flags);
  cudaEventDestroy(cs->ev_stage);
  cudaEventDestroy(cs->ev_done);
  cudaStreamDestroy(cs->stream);
  delete cs;
  c->comm_state = nullptr;
  return AB200_OK;
}

// Context-free planner, exported for the CPU tests and for hosts that drive their own
// transport: rows of 13 values (peer, recv?, fluid, block, var0, ncomp, si, ei, sj, ej, sk, ek,
// offset) in plan order; *rows_out is malloc'd, release it with ab200_comm_plan_free.
int ab200_comm_plan_direct(const int *nblk3, const int *nt3, const int *s3, const int *e3,
                           const int *ng3, int nfluids, const int *fluid_type,
                           const int *nspecies, const int *lay3, const int *rl3,
                           const int *periodic3, long long **rows_out, int *nrows_out) {
  AB_REQUIRE(nblk3 && nt3 && s3 && e3 && ng3 && fluid_type && nspecies && lay3 && rl3 &&
                 periodic3 && rows_out && nrows_out,
             AB200_EINVAL, "ab200_comm_plan_direct: null argument");
  std::vector<PlanRow> rows;
  std::map<int, std::pair<long long, long long>> sizes;
  plan_direct(nblk3, nt3, s3, e3, ng3, nfluids, fluid_type, nspecies, lay3, rl3, periodic3, rows,
              sizes);
  long long *out = (long long *)malloc(sizeof(long long) * 13 * std::max<size_t>(rows.size(), 1));
  AB_REQUIRE(out, AB200_ENOMEM, "ab200_comm_plan_direct: out of memory");
  for (size_t q = 0; q < rows.size(); ++q) {
    const PlanRow &r = rows[q];
    const long long v[13] = {r.peer, r.recv, r.fluid, r.block, r.var0, r.ncomp, r.si,
                             r.ei,   r.sj,   r.ej,    r.sk,    r.ek,   r.offset};
    memcpy(out + 13 * q, v, sizeof v);
  }
  *rows_out = out;
  *nrows_out = (int)rows.size();
  return AB200_OK;
}

void ab200_comm_plan_free(long long *rows) { free(rows); }

// Rank lattice of the block-spatial partition: plans the single-round exchange for
// the bound fluids and allocates the two message slabs.  Call after ab200_bind_pack and
// ab200_set_topology (faces towards other ranks carry AB200_BC_NONE).
int ab200_comm_set_layout(ab200_ctx *c, int layx, int layy, int layz, const int *periodic3) {
  AB_REQUIRE(c && c->grid_set && c->topo.set, AB200_ESTATE,
             "ab200_comm_set_layout: bind the grid and the topology first");
  CommState *cs = cs_of(c);
  AB_REQUIRE(cs, AB200_ESTATE, "ab200_comm_set_layout: call ab200_comm_init first");
  AB_REQUIRE(layx * layy * layz == cs->nranks, AB200_EINVAL,
             "ab200_comm_set_layout: rank lattice does not match the communicator size");
  AB_CUDA(cudaSetDevice(c->device));
  const GridDev &g = c->g;
  const int lay[3] = {layx, layy, layz};
  const int rl[3] = {cs->rank % layx, (cs->rank / layx) % layy, cs->rank / (layx * layy)};
  const int per[3] = {periodic3 ? periodic3[0] : 0, periodic3 ? periodic3[1] : 0,
                      periodic3 ? periodic3[2] : 0};
  const int nbd[3] = {c->topo.nbx, c->topo.nby, c->topo.nbz};
  const int nt[3] = {g.ni, g.nj, g.nk};
  const int s[3] = {g.is, g.js, g.ks}, e[3] = {g.ie, g.je, g.ke};
  const int ng[3] = {g.is, g.js, g.ks};  // ghost width per direction (0 in symmetry directions)
  int ftype[2], fS[2], nfl = 0;
  for (int f = 0; f < 2; ++f)
    if (c->fl[f].bound) { ftype[nfl] = f; fS[nfl] = c->fl[f].d.S; ++nfl; }
  std::vector<PlanRow> rows;
  std::map<int, std::pair<long long, long long>> sizes;
  plan_direct(nbd, nt, s, e, ng, nfl, ftype, fS, lay, rl, per, rows, sizes);
  cs->peers.clear();
  long long so = 0, ro = 0;
  std::map<int, size_t> index;
  for (auto &kv : sizes) {  // std::map: sorted by peer rank
    index[kv.first] = cs->peers.size();
    cs->peers.push_back({kv.first, so, kv.second.first, ro, kv.second.second});
    so += kv.second.first;
    ro += kv.second.second;
  }
  if (cs->dsend) cudaFree(cs->dsend);
  if (cs->drecv) cudaFree(cs->drecv);
  cs->dsend = cs->drecv = nullptr;
  AB_CUDA(cudaMalloc((void **)&cs->dsend, sizeof(double) * std::max<long long>(so, 1)));
  AB_CUDA(cudaMalloc((void **)&cs->drecv, sizeof(double) * std::max<long long>(ro, 1)));
  cs->send_desc.clear();
  cs->recv_desc.clear();
  for (const PlanRow &r : rows) {
    const CommState::Peer &p = cs->peers[index[r.peer]];
    ab200_bnd_desc d{r.fluid, r.block, r.var0, r.ncomp, r.si, r.ei, r.sj, r.ej, r.sk, r.ek, nullptr};
    if (r.recv) {
      d.buf = cs->drecv + p.roff + r.offset;
      cs->recv_desc.push_back(d);
    } else {
      d.buf = cs->dsend + p.soff + r.offset;
      cs->send_desc.push_back(d);
    }
  }
  cs->bytes_per_exchange = 8 * so;
  cs->planned = true;
  cs->direct = false;
  if (!cs->peers.empty() && !getenv("AB200_NO_DIRECT")) {
    // ---- direct transport over CUDA IPC peer mappings -----------------------------------------
    // Every rank owns two receive slabs (alternating by step: a sender may write step s+1 while
    // the receiver still unpacks step s; it cannot run two steps ahead because it waits for the
    // receiver's own step-s+1 flag first) and one flag array.  The handles are all-gathered once.
    NcclApi *n = nccl_api();
    const int R = cs->nranks;
    for (int q = 0; q < 2; ++q) {
      if (cs->rslab[q]) cudaFree(cs->rslab[q]);
      AB_CUDA(cudaMalloc((void **)&cs->rslab[q], sizeof(double) * std::max<long long>(ro, 1)));
    }
    if (!cs->flags) AB_CUDA(cudaMalloc((void **)&cs->flags, sizeof(unsigned long long) * R));
    if (!cs->d_err) AB_CUDA(cudaMalloc((void **)&cs->d_err, sizeof(int)));
    AB_CUDA(cudaMemset(cs->flags, 0, sizeof(unsigned long long) * R));
    AB_CUDA(cudaMemset(cs->d_err, 0, sizeof(int)));
    cs->step = 0;
    struct Handles { cudaIpcMemHandle_t slab[2], flags; };
    std::vector<Handles> all(R);
    bool ok = cudaIpcGetMemHandle(&all[cs->rank].slab[0], cs->rslab[0]) == cudaSuccess &&
              cudaIpcGetMemHandle(&all[cs->rank].slab[1], cs->rslab[1]) == cudaSuccess &&
              cudaIpcGetMemHandle(&all[cs->rank].flags, cs->flags) == cudaSuccess;
    if (!ok) cudaGetLastError();
    Handles *d_all = nullptr;
    AB_CUDA(cudaMalloc((void **)&d_all, sizeof(Handles) * R));
    AB_CUDA(cudaMemcpy(d_all + cs->rank, &all[cs->rank], sizeof(Handles), cudaMemcpyHostToDevice));
    AB_NCCL(n->AllGather(d_all + cs->rank, d_all, sizeof(Handles), /*ncclChar*/ 0, cs->comm, c->stream));
    AB_CUDA(cudaStreamSynchronize(c->stream));
    AB_CUDA(cudaMemcpy(all.data(), d_all, sizeof(Handles) * R, cudaMemcpyDeviceToHost));
    AB_CUDA(cudaFree(d_all));
    // every rank must take the same decision: all-reduce the "ok" bit through the dt reducer
    double *d_ok = nullptr;
    AB_CUDA(cudaMalloc((void **)&d_ok, sizeof(double)));
    for (void *p : cs->ipc_opened) cudaIpcCloseMemHandle(p);
    cs->ipc_opened.clear();
    for (int q = 0; q < 2; ++q) cs->peer_rslab[q].assign(cs->peers.size(), nullptr);
    cs->peer_flags.assign(cs->peers.size(), nullptr);
    for (size_t pi = 0; pi < cs->peers.size() && ok; ++pi) {
      const Handles &h = all[cs->peers[pi].rank];
      void *p0 = nullptr, *p1 = nullptr, *pf = nullptr;
      ok = cudaIpcOpenMemHandle(&p0, h.slab[0], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
           cudaIpcOpenMemHandle(&p1, h.slab[1], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
           cudaIpcOpenMemHandle(&pf, h.flags, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
      if (!ok) { cudaGetLastError(); break; }
      cs->ipc_opened.push_back(p0); cs->ipc_opened.push_back(p1); cs->ipc_opened.push_back(pf);
      cs->peer_rslab[0][pi] = (double *)p0;
      cs->peer_rslab[1][pi] = (double *)p1;
      cs->peer_flags[pi] = (unsigned long long *)pf;
    }
    const double okv = ok ? 1.0 : 0.0;
    AB_CUDA(cudaMemcpy(d_ok, &okv, sizeof(double), cudaMemcpyHostToDevice));
    AB_NCCL(n->AllReduce(d_ok, d_ok, 1, kNcclFloat64, kNcclMin, cs->comm, c->stream));
    double all_ok = 0.0;
    AB_CUDA(cudaMemcpyAsync(&all_ok, d_ok, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    AB_CUDA(cudaStreamSynchronize(c->stream));
    AB_CUDA(cudaFree(d_ok));
    if (all_ok == 1.0) {
      // where MY message starts inside each peer's receive slab: the peer's own plan (a pure
      // function of its lattice position; all ranks own tiles of the same shape)
      std::vector<long long> my_off_in_peer(cs->peers.size(), 0);
      for (size_t pi = 0; pi < cs->peers.size(); ++pi) {
        const int pr = cs->peers[pi].rank;
        const int prl[3] = {pr % layx, (pr / layx) % layy, pr / (layx * layy)};
        std::vector<PlanRow> prow;
        std::map<int, std::pair<long long, long long>> psz;
        plan_direct(nbd, nt, s, e, ng, nfl, ftype, fS, lay, prl, per, prow, psz);
        long long off = 0;
        for (auto &kv : psz) {
          if (kv.first == cs->rank) break;
          off += kv.second.second;
        }
        my_off_in_peer[pi] = off;
      }
      for (int q = 0; q < 2; ++q) {
        cs->dsend_desc[q].clear();
        cs->drecv_desc[q].clear();
      }
      for (const PlanRow &r : rows) {
        const size_t pi = index[r.peer];
        const CommState::Peer &p = cs->peers[pi];
        for (int q = 0; q < 2; ++q) {
          ab200_bnd_desc d{r.fluid, r.block, r.var0, r.ncomp, r.si, r.ei, r.sj, r.ej, r.sk, r.ek, nullptr};
          if (r.recv) {
            d.buf = cs->rslab[q] + p.roff + r.offset;
            cs->drecv_desc[q].push_back(d);
          } else {
            d.buf = cs->peer_rslab[q][pi] + my_off_in_peer[pi] + r.offset;
            cs->dsend_desc[q].push_back(d);
          }
        }
      }
      std::vector<int> pr(cs->peers.size());
      for (size_t pi = 0; pi < cs->peers.size(); ++pi) pr[pi] = cs->peers[pi].rank;
      if (cs->d_peer_rank) cudaFree(cs->d_peer_rank);
      if (cs->d_peer_flags) cudaFree(cs->d_peer_flags);
      AB_CUDA(cudaMalloc((void **)&cs->d_peer_rank, sizeof(int) * pr.size()));
      AB_CUDA(cudaMalloc((void **)&cs->d_peer_flags, sizeof(void *) * pr.size()));
      AB_CUDA(cudaMemcpy(cs->d_peer_rank, pr.data(), sizeof(int) * pr.size(), cudaMemcpyHostToDevice));
      AB_CUDA(cudaMemcpy(cs->d_peer_flags, cs->peer_flags.data(), sizeof(void *) * pr.size(),
                         cudaMemcpyHostToDevice));
      cs->direct = true;
    }
  }
  return AB200_OK;
}

int ab200_comm_is_direct(ab200_ctx *c) {
  CommState *cs = c ? cs_of(c) : nullptr;
  return cs && cs->direct ? 1 : 0;
}

long long ab200_comm_bytes_per_exchange(ab200_ctx *c) {
  CommState *cs = c ? cs_of(c) : nullptr;
  return cs ? cs->bytes_per_exchange : 0;
}

// Start the remote round of the stage whose kernels are queued on the context's stream: the
// comm stream waits for them, packs every descriptor in ONE launch, posts ONE NCCL group of
// sends / receives (one message per peer) and unpacks in ONE launch.  Runs concurrently with
// whatever the caller queues next on the context's stream (ab200_fill_ghosts_local reads
// interior zones only and writes none of the cells the unpack writes).
int ab200_comm_exchange_begin(ab200_ctx *c) {
  AB_REQUIRE(c, AB200_EINVAL, "null context");
  CommState *cs = cs_of(c);
  AB_REQUIRE(cs && cs->planned, AB200_ESTATE, "ab200_comm_exchange_begin: no exchange planned");
  AB_REQUIRE(!cs->in_flight, AB200_ESTATE, "ab200_comm_exchange_begin: a round is in flight");
  if (cs->peers.empty()) return AB200_OK;
  AB_CUDA(cudaSetDevice(c->device));
  NcclApi *n = nccl_api();
  AB_CUDA(cudaEventRecord(cs->ev_stage, c->stream));
  AB_CUDA(cudaStreamWaitEvent(cs->stream, cs->ev_stage, 0));
  const cudaStream_t saved = c->halo_stream;
  const bool saved_set = c->halo_stream_set;
  c->halo_stream = cs->stream;
  c->halo_stream_set = true;
  if (cs->direct) {
    // pack = remote stores into the peers' receive slabs; signal; wait for every peer's signal;
    // unpack from the own slab.  No NCCL on the data path.
    const int q = (int)(cs->step & 1);
    const unsigned long long stepno = cs->step + 1;
    const int np = (int)cs->peers.size();
    int rc = launch_halo(c, cs->dsend_desc[q].data(), (int)cs->dsend_desc[q].size(), 0);
    if (rc == AB200_OK) {
      k_comm_signal<<<1, 32 * ((np + 31) / 32), 0, cs->stream>>>(cs->d_peer_flags, np, cs->rank, stepno);
      k_comm_wait<<<1, 32 * ((np + 31) / 32), 0, cs->stream>>>(cs->flags, cs->d_peer_rank, np, stepno, cs->d_err);
      c->launches += 2;
      rc = launch_halo(c, cs->drecv_desc[q].data(), (int)cs->drecv_desc[q].size(), 1);
    }
    c->halo_stream = saved;
    c->halo_stream_set = saved_set;
    AB_TRY(rc);
    AB_CUDA(cudaGetLastError());
    AB_CUDA(cudaEventRecord(cs->ev_done, cs->stream));
    cs->step++;
    cs->in_flight = true;
    return AB200_OK;
  }
  int rc = launch_halo(c, cs->send_desc.data(), (int)cs->send_desc.size(), 0);
  if (rc == AB200_OK) {
    int e = n->GroupStart();
    for (const auto &p : cs->peers) {
      if (e == 0 && p.nsend)
        e = n->Send(cs->dsend + p.soff, (size_t)p.nsend, kNcclFloat64, p.rank, cs->comm, cs->stream);
      if (e == 0 && p.nrecv)
        e = n->Recv(cs->drecv + p.roff, (size_t)p.nrecv, kNcclFloat64, p.rank, cs->comm, cs->stream);
    }
    const int e2 = n->GroupEnd();
    if (e == 0) e = e2;
    if (e != 0) {
      set_error(std::string("NCCL: grouped send/recv: ") + n->GetErrorString(e));
      rc = AB200_ECUDA;
    }
  }
  if (rc == AB200_OK) rc = launch_halo(c, cs->recv_desc.data(), (int)cs->recv_desc.size(), 1);
  c->halo_stream = saved;
  c->halo_stream_set = saved_set;
  AB_TRY(rc);
  AB_CUDA(cudaEventRecord(cs->ev_done, cs->stream));
  cs->in_flight = true;
  return AB200_OK;
}

int ab200_comm_exchange_end(ab200_ctx *c) {
  AB_REQUIRE(c, AB200_EINVAL, "null context");
  CommState *cs = cs_of(c);
  AB_REQUIRE(cs, AB200_ESTATE, "ab200_comm_exchange_end: no communicator");
  if (!cs->in_flight) return AB200_OK;
  AB_CUDA(cudaSetDevice(c->device));
  AB_CUDA(cudaStreamWaitEvent(c->stream, cs->ev_done, 0));
  cs->in_flight = false;
  return AB200_OK;
}

// MPI_Allreduce(&dt, 1, MPI_DOUBLE, MPI_MIN) of EvolutionDriver::SetGlobalTimeStep
// (P:driver/driver.cpp:237) on a DEVICE scalar, in place, on the context's stream.  Without a
// communicator (single rank) it is the identity.
int ab200_allreduce_min(ab200_ctx *c, double *dev_scalar) {
  AB_REQUIRE(c && dev_scalar, AB200_EINVAL, "ab200_allreduce_min: null argument");
  CommState *cs = cs_of(c);
  if (!cs || cs->nranks == 1) return AB200_OK;
  AB_CUDA(cudaSetDevice(c->device));
  AB_NCCL(nccl_api()->AllReduce(dev_scalar, dev_scalar, 1, kNcclFloat64, kNcclMin, cs->comm, c->stream));
  return AB200_OK;
}

// The device-resident cycle of ab200_run_cycles for a rank of a multi-rank job
// (ArtemisDriver::Step, src/artemis_driver.cpp:101-121, + SetGlobalTimeStep): per stage the
// fused stage, then the remote round on the comm stream concurrently with the same-GPU ghost
// fill, then the finish pass; per cycle the all-reduce(MIN) of the device dt scalar.  No host
// round trip unless tlim is finite.
int ab200_run_cycles_mr(ab200_ctx *c, int integrator, int ncycles, double tlim) {
  AB_REQUIRE(c && c->grid_set && c->topo.set, AB200_ESTATE, "ab200_run_cycles_mr: nothing bound");
  AB_REQUIRE(integrator >= 0 && integrator <= 3, AB200_EINVAL, "unknown integrator");
  CommState *cs = cs_of(c);
  AB_REQUIRE(topology_is_local(c) || (cs && cs->planned), AB200_ESTATE,
             "ab200_run_cycles_mr: remote faces but no planned exchange (ab200_comm_set_layout)");
  struct Stage { double g0, g1, b; };
  static const Stage tabs[4][3] = {
      {{0.0, 1.0, 1.0}, {}, {}},
      {{0.0, 1.0, 1.0}, {0.5, 0.5, 0.5}, {}},
      {{0.0, 1.0, 0.5}, {0.0, 1.0, 1.0}, {}},
      {{0.0, 1.0, 1.0}, {0.25, 0.75, 0.25}, {2.0 / 3.0, 1.0 / 3.0, 2.0 / 3.0}}};
  const int nst = integrator == 0 ? 1 : integrator == 3 ? 3 : 2;
  const Stage *st = tabs[integrator];
  const bool finite_tlim = tlim < 1.0e300;
  const bool lazy_before = c->ghost_cons_lazy;
  c->ghost_cons_lazy = true;
  struct Restore {
    ab200_ctx *c; bool v;
    ~Restore() { c->ghost_cons_lazy = v; }
  } restore{c, lazy_before};
  const bool remote = cs && cs->planned && !cs->peers.empty();
  // Overlap (opt-in, AB200_OVERLAP=1): stage the blocks that touch another rank first, start the
  // remote round as soon as they are done and run the interior blocks' share of the last pass
  // underneath it.  Measured on B200 (r02, A/B inside one box, 256^3 per GPU): N = 4: 4.467 ms
  // per cycle without overlap, 4.570 with it, 4.679 with a high-priority comm stream -- the stage
  // kernels are latency-bound at low occupancy, so sharing SMs with the transport kernels and
  // paying one extra launch costs more than the 0.13 ms round it hides; N = 2: 4.556 vs 4.597
  // (marginal gain).  Hence off by default.
  bool split = remote && getenv("AB200_OVERLAP") && !c->has_sources && fused_supports_subsets(c) &&
               c->n_blist[0] > 0 && c->n_blist[1] > 0;
  for (int f = 0; f < 2 && split; ++f)
    if (c->fl[f].bound && sweep_eligible(c, f)) split = false;
  for (int cyc = 0; cyc < ncycles; ++cyc) {
    for (int s = 0; s < nst; ++s) {
      const int pcm = (s == 0 && integrator == 2);
      if (split) {
        const int flags = AB200_STAGE_DEVICE_DT | AB200_STAGE_PINGPONG |
                          (s == nst - 1 ? AB200_STAGE_REDUCE_DT : 0);
        AB_TRY(ab200_fused_stage(c, st[s].g0, st[s].g1, st[s].b, 0.0, pcm, s == 0,
                                 flags | AB200_STAGE_SURFACE));
        AB_TRY(ab200_comm_exchange_begin(c));
        AB_TRY(ab200_fused_stage(c, st[s].g0, st[s].g1, st[s].b, 0.0, pcm, s == 0,
                                 flags | AB200_STAGE_INTERIOR));
      } else {
        AB_TRY(run_stage(c, st[s].g0, st[s].g1, st[s].b, pcm, s == 0, s == nst - 1));
        if (remote) AB_TRY(ab200_comm_exchange_begin(c));
      }
      AB_TRY(ab200_fill_ghosts_local(c));
      if (remote) {
        AB_TRY(ab200_comm_exchange_end(c));
        AB_TRY(ab200_finish_remote_ghosts(c));
      }
    }
    AB_TRY(ab200_allreduce_min(c, c->d_time + 1));
    AB_TRY(ab200_set_global_timestep_device(c, tlim, 1));
    if (finite_tlim) {
      double ts[4];
      AB_TRY(ab200_read_time_state(c, ts));
      if (ts[2] >= tlim) break;
    }
  }
  AB_TRY(ab200_sync_prim(c));
  if (!lazy_before) AB_TRY(ab200_sync_ghost_cons(c));
  return AB200_OK;
}

}  // extern "C"
