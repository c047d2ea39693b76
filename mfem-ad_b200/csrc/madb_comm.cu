// madb_comm.cu -- shared-dof exchange between element partitions behind the C ABI (include/mfemad_b200.h).
//
// What the reference gets from MFEM's ParMesh / ParFiniteElementSpace (ex4.cpp:85,99-101,136 [MFEM-upstream]):
//   P    owner -> sharers   (ParGridFunction::Distribute / the halo of x before the element loop)
//   P^T  sharers -> owner   (ParNonlinearForm::Mult: contributions to a shared dof are summed on its owner)
// Messages are neighbour point-to-point: one ncclGroup of ncclSend / ncclRecv pairs per exchange on a communication
// stream, packed / unpacked by small kernels on the context stream (events order the two streams; nothing is ordered
// by the host).  A dof that receives from several peers is ONE destination whose sources are added in ascending peer
// rank: the result does not depend on arrival order.
// NCCL is loaded with dlopen at the first use (libnccl.so.2: the copy already in the process -- e.g. torch's -- wins), so
// libmadb.so itself has no link-time dependency on it and loads on machines without NCCL.
#include "../../include/mfemad_b200.h"
#include "madb_host.hpp"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace madb
{

struct NcclApi
{
   void *h = nullptr;
   ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
   ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
   ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
   ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*GroupStart)() = nullptr;
   ncclResult_t (*GroupEnd)() = nullptr;
   ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
   const char *(*GetErrorString)(ncclResult_t) = nullptr;
   ncclResult_t (*GetVersion)(int *) = nullptr;
};

static NcclApi *nccl_api()
{
   static NcclApi api;
   static std::once_flag once;
   static bool ok = false;
   std::call_once(once, []()
   {
      const char *names[] = {getenv("MADB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
      for (const char *n : names)
      {
         if (!n) { continue; }
         api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
         if (api.h) { break; }
      }
      if (!api.h) { return; }
#define MADB_SYM(field, name) *(void **)(&api.field) = dlsym(api.h, name)
      MADB_SYM(GetUniqueId, "ncclGetUniqueId");
      MADB_SYM(CommInitRank, "ncclCommInitRank");
      MADB_SYM(CommDestroy, "ncclCommDestroy");
      MADB_SYM(Send, "ncclSend");
      MADB_SYM(Recv, "ncclRecv");
      MADB_SYM(GroupStart, "ncclGroupStart");
      MADB_SYM(GroupEnd, "ncclGroupEnd");
      MADB_SYM(AllReduce, "ncclAllReduce");
      MADB_SYM(GetErrorString, "ncclGetErrorString");
      MADB_SYM(GetVersion, "ncclGetVersion");
#undef MADB_SYM
      ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.GroupStart && api.GroupEnd &&
           api.AllReduce && api.GetErrorString;
   });
   return ok ? &api : nullptr;
}

#define NCCL_OK(call)                                                                                        \
   do {                                                                                                      \
      ncclResult_t r_ = (call);                                                                              \
      if (r_ != ncclSuccess)                                                                                 \
      {                                                                                                      \
         set_error(std::string(#call) + ": " + nccl_api()->GetErrorString(r_));                              \
         return 2;                                                                                           \
      }                                                                                                      \
   } while (0)
#define CUDA_OKC(call)                                                                                       \
   do {                                                                                                      \
      cudaError_t e_ = (call);                                                                               \
      if (e_ != cudaSuccess)                                                                                 \
      {                                                                                                      \
         set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                                      \
         return 2;                                                                                           \
      }                                                                                                      \
   } while (0)

struct Comm
{
   Ctx *ctx = nullptr;
   ncclComm_t comm = nullptr;
   int rank = 0, world = 1;
   cudaStream_t stream = nullptr; // communication stream
   double *d_scratch = nullptr;   // allreduce staging
};

struct Direction // one direction of an exchange: what I send, what I receive
{
   std::vector<int> send_peer, send_count, recv_peer, recv_count;
   int nsend = 0, nrecv = 0, ndst = 0;
   int *d_send_idx = nullptr; // [nsend] local indices packed into the send buffer, peer by peer (ascending rank)
   int *d_dst = nullptr;      // [ndst] destinations of the receive buffer
   int4 *d_src4 = nullptr;    // [ndst] up to 4 receive-buffer positions per destination (-1: none), ascending peer rank
   double *d_sbuf = nullptr, *d_rbuf = nullptr;
};

struct Exchange
{
   Comm *comm = nullptr;
   Direction dir[2]; // 0: forward (owner -> sharers, P), 1: reverse (sharers -> owner, P^T)
   cudaEvent_t ev_packed = nullptr, ev_done = nullptr;
   int pending = -1;
   ~Exchange()
   {
      for (Direction &d : dir)
      {
         cudaFree(d.d_send_idx); cudaFree(d.d_dst); cudaFree(d.d_src4); cudaFree(d.d_sbuf); cudaFree(d.d_rbuf);
      }
      if (ev_packed) { cudaEventDestroy(ev_packed); cudaEventDestroy(ev_done); }
   }
};

__global__ void k_xpack(int n, const int *__restrict__ idx, const double *__restrict__ src, double *__restrict__ dst)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { dst[i] = src[idx[i]]; }
}
__global__ void k_xunpack(int n, const int4 *__restrict__ src4, const int *__restrict__ idx, const double *__restrict__ src, double *dst,
                          int add)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   const int4 s = src4[i];
   const int d = idx[i];
   const double a0 = src[s.x], a1 = (s.y >= 0) ? src[s.y] : 0.0, a2 = (s.z >= 0) ? src[s.z] : 0.0, a3 = (s.w >= 0) ? src[s.w] : 0.0;
   double v = add ? dst[d] + a0 : a0;
   if (s.y >= 0) { v += a1; }
   if (s.z >= 0) { v += a2; }
   if (s.w >= 0) { v += a3; }
   dst[d] = v;
}

template <class T> static int up(const std::vector<T> &h, T **d)
{
   if (cudaMalloc((void **)d, std::max<size_t>(h.size(), 1) * sizeof(T)) != cudaSuccess) { return 2; }
   return cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess ? 0 : 2;
}

static int build_direction(Direction &D, int nsp, const int *sp, const int *sc, const int32_t *sidx, int nrp, const int *rp,
                           const int *rc, const int32_t *ridx)
{
   D.send_peer.assign(sp, sp + nsp); D.send_count.assign(sc, sc + nsp);
   D.recv_peer.assign(rp, rp + nrp); D.recv_count.assign(rc, rc + nrp);
   for (int k = 1; k < nsp; k++) { if (sp[k] <= sp[k - 1]) { set_error("madb_exchange_create: peers must be in ascending rank order"); return 1; } }
   for (int k = 1; k < nrp; k++) { if (rp[k] <= rp[k - 1]) { set_error("madb_exchange_create: peers must be in ascending rank order"); return 1; } }
   D.nsend = 0;
   for (int c : D.send_count) { D.nsend += c; }
   D.nrecv = 0;
   for (int c : D.recv_count) { D.nrecv += c; }
   std::vector<int> sidxv(sidx, sidx + D.nsend);
   // destinations: unique local indices of the receive list, sources in ascending receive-buffer position (= peer rank)
   std::vector<std::pair<int, int>> t(D.nrecv);
   for (int k = 0; k < D.nrecv; k++) { t[k] = {ridx[k], k}; }
   std::stable_sort(t.begin(), t.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first < b.first; });
   std::vector<int> dst;
   std::vector<int4> src4;
   for (size_t k = 0; k < t.size();)
   {
      size_t e = k;
      while (e < t.size() && t[e].first == t[k].first) { e++; }
      if (e - k > 4) { set_error("madb_exchange_create: a dof receives from more than 4 peers"); return 1; }
      int4 s = make_int4(-1, -1, -1, -1);
      int *sv = &s.x;
      for (size_t q = k; q < e; q++) { sv[q - k] = t[q].second; }
      dst.push_back(t[k].first);
      src4.push_back(s);
      k = e;
   }
   D.ndst = (int)dst.size();
   if (up(sidxv, &D.d_send_idx) || up(dst, &D.d_dst) || up(src4, &D.d_src4)) { set_error("madb_exchange_create: device allocation failed"); return 2; }
   if (cudaMalloc((void **)&D.d_sbuf, std::max(D.nsend, 1) * sizeof(double)) != cudaSuccess ||
       cudaMalloc((void **)&D.d_rbuf, std::max(D.nrecv, 1) * sizeof(double)) != cudaSuccess)
   {
      set_error("madb_exchange_create: device allocation failed");
      return 2;
   }
   return 0;
}

} // namespace madb

using namespace madb;

struct madb_ctx : Ctx {};
struct madb_comm : Comm {};
struct madb_exchange : Exchange {};

extern "C"
{
   int madb_comm_unique_id(unsigned char *id128)
   {
      NcclApi *N = nccl_api();
      if (!N) { set_error("NCCL is not available (libnccl.so.2 could not be loaded; set MADB_NCCL_LIB)"); return 2; }
      static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
      ncclUniqueId id;
      NCCL_OK(N->GetUniqueId(&id));
      std::memcpy(id128, &id, 128);
      return 0;
   }

   int madb_comm_create(madb_ctx *ctx, const unsigned char *id128, int rank, int world, madb_comm **out)
   {
      NcclApi *N = nccl_api();
      if (!N) { set_error("NCCL is not available (libnccl.so.2 could not be loaded; set MADB_NCCL_LIB)"); return 2; }
      if (!ctx || !id128 || rank < 0 || rank >= world) { set_error("madb_comm_create: bad arguments"); return 1; }
      CUDA_OKC(cudaSetDevice(ctx->device));
      madb_comm *c = new madb_comm;
      c->ctx = ctx; c->rank = rank; c->world = world;
      ncclUniqueId id;
      std::memcpy(&id, id128, 128);
      ncclResult_t r = N->CommInitRank(&c->comm, world, id, rank);
      if (r != ncclSuccess) { set_error(std::string("ncclCommInitRank: ") + N->GetErrorString(r)); delete c; return 2; }
      CUDA_OKC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
      CUDA_OKC(cudaMalloc((void **)&c->d_scratch, 64 * sizeof(double)));
      *out = c;
      return 0;
   }
   int madb_comm_destroy(madb_comm *c)
   {
      if (!c) { return 0; }
      if (c->comm) { nccl_api()->CommDestroy(c->comm); }
      if (c->stream) { cudaStreamDestroy(c->stream); }
      cudaFree(c->d_scratch);
      delete c;
      return 0;
   }

   int madb_comm_allreduce_sum(madb_comm *c, int n, double *values)
   {
      // global sums of a few scalars (Newton residual norms, the L1 change of lambda at ex4.cpp:205): one ncclAllReduce;
      // NCCL's reduction order is fixed for a given communicator, so every rank gets the same bits
      NcclApi *N = nccl_api();
      if (n < 1 || n > 64) { set_error("madb_comm_allreduce_sum: 1 <= n <= 64"); return 1; }
      CUDA_OKC(cudaSetDevice(c->ctx->device));
      cudaPointerAttributes at;
      const bool dev = cudaPointerGetAttributes(&at, values) == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
      cudaGetLastError();
      double *d = dev ? values : c->d_scratch;
      if (!dev) { CUDA_OKC(cudaMemcpyAsync(d, values, n * sizeof(double), cudaMemcpyHostToDevice, c->ctx->stream)); }
      NCCL_OK(N->AllReduce(d, d, (size_t)n, ncclFloat64, ncclSum, c->comm, c->ctx->stream));
      if (!dev)
      {
         CUDA_OKC(cudaMemcpyAsync(values, d, n * sizeof(double), cudaMemcpyDeviceToHost, c->ctx->stream));
         CUDA_OKC(cudaStreamSynchronize(c->ctx->stream));
      }
      return 0;
   }

   int madb_exchange_create(madb_comm *c, int nown_peers, const int *own_peer, const int *own_count, const int32_t *own_idx,
                            int nghost_peers, const int *ghost_peer, const int *ghost_count, const int32_t *ghost_idx,
                            madb_exchange **out)
   {
      if (!c || nown_peers < 0 || nghost_peers < 0) { set_error("madb_exchange_create: bad arguments"); return 1; }
      CUDA_OKC(cudaSetDevice(c->ctx->device));
      madb_exchange *x = new madb_exchange;
      x->comm = c;
      // forward (P): I send the dofs I own to the peers that hold copies, and receive my copies from their owners;
      // reverse (P^T): the same lists with the roles swapped
      int rc = build_direction(x->dir[0], nown_peers, own_peer, own_count, own_idx, nghost_peers, ghost_peer, ghost_count, ghost_idx);
      if (!rc) { rc = build_direction(x->dir[1], nghost_peers, ghost_peer, ghost_count, ghost_idx, nown_peers, own_peer, own_count, own_idx); }
      if (rc) { delete x; return rc; }
      CUDA_OKC(cudaEventCreateWithFlags(&x->ev_packed, cudaEventDisableTiming));
      CUDA_OKC(cudaEventCreateWithFlags(&x->ev_done, cudaEventDisableTiming));
      *out = x;
      return 0;
   }
   int madb_exchange_destroy(madb_exchange *x) { delete x; return 0; }

   int madb_exchange_begin(madb_exchange *x, const double *vec, int reverse)
   {
      NcclApi *N = nccl_api();
      if (x->pending >= 0) { set_error("madb_exchange_begin: the previous exchange has not been ended"); return 1; }
      Comm &C = *x->comm;
      Direction &D = x->dir[reverse ? 1 : 0];
      CUDA_OKC(cudaSetDevice(C.ctx->device));
      cudaStream_t cs = C.ctx->stream;
      if (D.nsend > 0) { k_xpack<<<(D.nsend + 255) / 256, 256, 0, cs>>>(D.nsend, D.d_send_idx, vec, D.d_sbuf); }
      CUDA_OKC(cudaEventRecord(x->ev_packed, cs));
      CUDA_OKC(cudaStreamWaitEvent(C.stream, x->ev_packed, 0));
      if (!D.send_peer.empty() || !D.recv_peer.empty())
      {
         NCCL_OK(N->GroupStart());
         size_t so = 0, ro = 0;
         for (size_t k = 0; k < D.send_peer.size(); k++)
         {
            if (D.send_count[k] > 0) { NCCL_OK(N->Send(D.d_sbuf + so, (size_t)D.send_count[k], ncclFloat64, D.send_peer[k], C.comm, C.stream)); }
            so += D.send_count[k];
         }
         for (size_t k = 0; k < D.recv_peer.size(); k++)
         {
            if (D.recv_count[k] > 0) { NCCL_OK(N->Recv(D.d_rbuf + ro, (size_t)D.recv_count[k], ncclFloat64, D.recv_peer[k], C.comm, C.stream)); }
            ro += D.recv_count[k];
         }
         NCCL_OK(N->GroupEnd());
      }
      CUDA_OKC(cudaEventRecord(x->ev_done, C.stream));
      x->pending = reverse ? 1 : 0;
      return 0;
   }

   int madb_exchange_end(madb_exchange *x, double *vec, int add)
   {
      if (x->pending < 0) { set_error("madb_exchange_end: no exchange in flight"); return 1; }
      Comm &C = *x->comm;
      Direction &D = x->dir[x->pending];
      x->pending = -1;
      CUDA_OKC(cudaSetDevice(C.ctx->device));
      cudaStream_t cs = C.ctx->stream;
      CUDA_OKC(cudaStreamWaitEvent(cs, x->ev_done, 0));
      if (D.ndst > 0) { k_xunpack<<<(D.ndst + 255) / 256, 256, 0, cs>>>(D.ndst, D.d_src4, D.d_dst, D.d_rbuf, vec, add); }
      CUDA_OKC(cudaGetLastError());
      return 0;
   }
} // extern "C"
