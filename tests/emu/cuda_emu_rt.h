// TEST INFRASTRUCTURE: CUDA runtime API and NCCL stand-ins for running the product's HOST driver
// (csrc/jsso_api.cu, kernel launches rewritten by tests/emu/build_emu.py) on top of the SIMT emulator.
// "Device memory" is the heap; streams are no-ops (every emulated launch is synchronous within its thread);
// the device reports 2 SMs and occupancy 1, so persistent grids stay tiny.  The fake NCCL connects RANK THREADS of
// one process (each thread drives its own handle): buffered sends, blocking receives, rank-ordered all-reduce.
// Not part of the product; nothing here is reachable from libjsso.so.
#pragma once
#include "cuda_emu.h"

#include <nccl.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>

namespace emurt {

inline cudaError_t ok() { return cudaSuccess; }

// EMU_POISON=1: fresh "device" memory is filled with 0xFF (NaN as a double, -1 as an index) instead of zeros, so that a
// kernel reading something nobody computed (e.g. a ghost row missing from a distributed setup plan) shows up
inline cudaError_t Malloc(void** p, size_t n) {
  static const bool poison = std::getenv("EMU_POISON") && std::getenv("EMU_POISON")[0] == '1';
  *p = std::calloc(n ? n : 1, 1);
  if (*p && poison) std::memset(*p, 0xFF, n ? n : 1);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T> inline cudaError_t Malloc(T** p, size_t n) { return Malloc((void**)p, n); }
inline cudaError_t Free(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t MallocHost(void** p, size_t n) { return Malloc(p, n); }
template <class T> inline cudaError_t MallocHost(T** p, size_t n) { return Malloc((void**)p, n); }
inline cudaError_t HostAlloc(void** p, size_t n, unsigned) { return Malloc(p, n); }
inline cudaError_t FreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t Memcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t MemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
  if (emu::capturing) { std::fprintf(stderr, "emu: memcpy inside a stream capture is not emulated\n"); std::abort(); }
  if (n) std::memmove(d, s, n);
  return cudaSuccess;
}
inline cudaError_t Memcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr) {
  for (size_t r = 0; r < h; ++r) std::memmove((char*)d + r * dp, (const char*)s + r * sp, w);
  return cudaSuccess;
}
inline cudaError_t Memset(void* d, int v, size_t n) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t MemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t SetDevice(int) { return cudaSuccess; }
inline cudaError_t GetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t DeviceSynchronize() { return cudaSuccess; }
inline cudaError_t StreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t StreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t StreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t StreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline cudaError_t EventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)std::calloc(1, 8); return cudaSuccess; }
inline cudaError_t EventCreateWithFlags(cudaEvent_t* e, unsigned) { return EventCreate(e); }
inline cudaError_t EventRecord(cudaEvent_t e, cudaStream_t = nullptr) { *(double*)e = now_ms(); return cudaSuccess; }   // launches are synchronous
inline cudaError_t EventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t EventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(*(double*)b - *(double*)a); return cudaSuccess; }
inline cudaError_t EventDestroy(cudaEvent_t e) { std::free((void*)e); return cudaSuccess; }
inline cudaError_t GetLastError() { return cudaSuccess; }
inline const char* GetErrorString(cudaError_t) { return "emulated CUDA runtime"; }
inline cudaError_t GetDeviceProperties(cudaDeviceProp* p, int) { std::memset(p, 0, sizeof *p); p->multiProcessorCount = 2; return cudaSuccess; }
inline cudaError_t DeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 0; return cudaSuccess; }   // no cooperative launch
template <class F> inline cudaError_t OccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 1; return cudaSuccess; }
template <class F> inline cudaError_t FuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
// every "device" and host pointer is heap memory; EMU_PAGEABLE=1 reports host pointers as pageable so that the driver's
// staging path (threaded memcpy through its pinned buffers) runs instead of the direct DMA
inline cudaError_t PointerGetAttributes(cudaPointerAttributes* a, const void*) {
  static const bool pageable = std::getenv("EMU_PAGEABLE") != nullptr;
  std::memset(a, 0, sizeof *a);
  a->type = pageable ? cudaMemoryTypeUnregistered : cudaMemoryTypeHost;
  return cudaSuccess;
}
// "IPC" between rank threads of one process: the handle carries the pointer
inline cudaError_t IpcGetMemHandle(cudaIpcMemHandle_t* hd, void* p) { std::memset(hd, 0, sizeof *hd); std::memcpy(hd, &p, sizeof p); return cudaSuccess; }
inline cudaError_t IpcOpenMemHandle(void** p, cudaIpcMemHandle_t hd, unsigned) { std::memcpy(p, &hd, sizeof *p); return cudaSuccess; }
inline cudaError_t IpcCloseMemHandle(void*) { return cudaSuccess; }
inline cudaError_t LaunchCooperativeKernel(const void*, dim3, dim3, void**, size_t, cudaStream_t) { return cudaErrorNotSupported; }

// CUDA graphs: capture = record the launches of this thread, launch = replay them
inline cudaError_t StreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) {
  if (emu::capturing) return cudaErrorIllegalState;
  emu::capturing = new emu::Graph;
  return cudaSuccess;
}
inline cudaError_t StreamEndCapture(cudaStream_t, cudaGraph_t* g) {
  if (!emu::capturing) return cudaErrorIllegalState;
  *g = (cudaGraph_t)emu::capturing;
  emu::capturing = nullptr;
  return cudaSuccess;
}
inline cudaError_t GraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long = 0) {
  *e = (cudaGraphExec_t) new emu::Graph(*(emu::Graph*)g);
  return cudaSuccess;
}
inline cudaError_t GraphLaunch(cudaGraphExec_t e, cudaStream_t) {
  for (const emu::GraphNode& nd : ((emu::Graph*)e)->nodes) emu::launch(nd.grid, nd.block, nd.smem, nd.body);
  return cudaSuccess;
}
inline cudaError_t GraphDestroy(cudaGraph_t g) { delete (emu::Graph*)g; return cudaSuccess; }
inline cudaError_t GraphExecDestroy(cudaGraphExec_t e) { delete (emu::Graph*)e; return cudaSuccess; }

// ---------------------------------------------------------------- fake NCCL between rank threads
struct Shared {
  int n = 0;
  std::mutex mu;
  std::condition_variable cv;
  std::map<std::pair<int, int>, std::deque<std::vector<char>>> box;   // (src, dst) -> FIFO of messages
  std::vector<std::vector<double>> red;                               // all-reduce contributions by rank
  int red_arrived = 0, red_left = 0;
  long long red_gen = 0;
};
struct Comm { Shared* sh; int rank; };
struct PendingRecv { void* buf; size_t bytes; int peer; Comm* c; };
struct Registry {
  std::mutex mu;
  std::map<std::string, Shared*> by_id;
  long long next_id = 1;
};
inline Registry& reg() { static Registry r; return r; }
inline thread_local int group_depth = 0;
inline thread_local std::vector<PendingRecv> pending;

inline ncclResult_t NGetUniqueId(ncclUniqueId* id) {
  std::memset(id, 0, sizeof *id);
  std::lock_guard<std::mutex> g(reg().mu);
  const long long v = reg().next_id++;
  std::memcpy(id->internal, &v, sizeof v);
  return ncclSuccess;
}
inline ncclResult_t NCommInitRank(ncclComm_t* comm, int n, ncclUniqueId id, int rank) {
  std::lock_guard<std::mutex> g(reg().mu);
  Shared*& s = reg().by_id[std::string(id.internal, sizeof id.internal)];
  if (!s) { s = new Shared; s->n = n; s->red.resize(n); }
  *comm = (ncclComm_t) new Comm{s, rank};
  return ncclSuccess;
}
inline ncclResult_t NCommDestroy(ncclComm_t c) { delete (Comm*)c; return ncclSuccess; }
inline void do_recv(const PendingRecv& r) {
  Shared& s = *r.c->sh;
  std::unique_lock<std::mutex> lk(s.mu);
  auto key = std::make_pair(r.peer, r.c->rank);
  s.cv.wait(lk, [&] { return !s.box[key].empty(); });
  std::vector<char> m = std::move(s.box[key].front());
  s.box[key].pop_front();
  lk.unlock();
  if (m.size() != r.bytes) { std::fprintf(stderr, "emu nccl: size mismatch %zu vs %zu\n", m.size(), r.bytes); std::abort(); }
  std::memcpy(r.buf, m.data(), r.bytes);
}
inline ncclResult_t NGroupStart() { ++group_depth; return ncclSuccess; }
inline ncclResult_t NGroupEnd() {
  if (--group_depth == 0) {
    for (const PendingRecv& r : pending) do_recv(r);
    pending.clear();
  }
  return ncclSuccess;
}
inline size_t tsize(ncclDataType_t t) { return t == ncclDouble ? 8 : (t == ncclFloat ? 4 : (t == ncclInt32 ? 4 : 1)); }
inline ncclResult_t NSend(const void* buf, size_t count, ncclDataType_t t, int peer, ncclComm_t comm, cudaStream_t) {
  Comm* c = (Comm*)comm;
  std::vector<char> m((const char*)buf, (const char*)buf + count * tsize(t));   // buffered: a send never blocks
  {
    std::lock_guard<std::mutex> g(c->sh->mu);
    c->sh->box[std::make_pair(c->rank, peer)].push_back(std::move(m));
  }
  c->sh->cv.notify_all();
  return ncclSuccess;
}
inline ncclResult_t NRecv(void* buf, size_t count, ncclDataType_t t, int peer, ncclComm_t comm, cudaStream_t) {
  PendingRecv r{buf, count * tsize(t), peer, (Comm*)comm};
  if (group_depth > 0) pending.push_back(r); else do_recv(r);
  return ncclSuccess;
}
inline ncclResult_t NAllReduce(const void* send, void* recv, size_t count, ncclDataType_t t, ncclRedOp_t op, ncclComm_t comm, cudaStream_t) {
  if (t != ncclDouble || op != ncclSum) return ncclInvalidArgument;
  Comm* c = (Comm*)comm;
  Shared& s = *c->sh;
  std::unique_lock<std::mutex> lk(s.mu);
  s.cv.wait(lk, [&] { return s.red_left == 0; });                 // the previous reduction has been read by everybody
  s.red[c->rank].assign((const double*)send, (const double*)send + count);
  const long long gen = s.red_gen;
  if (++s.red_arrived == s.n) { s.red_left = s.n; s.red_arrived = 0; ++s.red_gen; s.cv.notify_all(); }
  else s.cv.wait(lk, [&] { return s.red_gen != gen; });
  for (size_t k = 0; k < count; ++k) {
    double acc = 0.0;
    for (int r = 0; r < s.n; ++r) acc += s.red[r][k];             // rank order: identical on every rank
    ((double*)recv)[k] = acc;
  }
  if (--s.red_left == 0) s.cv.notify_all();
  return ncclSuccess;
}
inline const char* NGetErrorString(ncclResult_t) { return "emulated NCCL"; }

}  // namespace emurt

#define cudaMalloc emurt::Malloc
#define cudaFree emurt::Free
#define cudaMallocHost emurt::MallocHost
#define cudaHostAlloc emurt::HostAlloc
#define cudaFreeHost emurt::FreeHost
#define cudaMemcpy emurt::Memcpy
#define cudaMemcpyAsync emurt::MemcpyAsync
#define cudaMemcpy2DAsync emurt::Memcpy2DAsync
#define cudaMemset emurt::Memset
#define cudaMemsetAsync emurt::MemsetAsync
#define cudaSetDevice emurt::SetDevice
#define cudaGetDeviceCount emurt::GetDeviceCount
#define cudaDeviceSynchronize emurt::DeviceSynchronize
#define cudaStreamSynchronize emurt::StreamSynchronize
#define cudaStreamCreateWithFlags emurt::StreamCreateWithFlags
#define cudaStreamDestroy emurt::StreamDestroy
#define cudaStreamWaitEvent emurt::StreamWaitEvent
#define cudaEventCreate emurt::EventCreate
#define cudaEventCreateWithFlags emurt::EventCreateWithFlags
#define cudaEventRecord emurt::EventRecord
#define cudaEventSynchronize emurt::EventSynchronize
#define cudaEventElapsedTime emurt::EventElapsedTime
#define cudaEventDestroy emurt::EventDestroy
#define cudaGetLastError emurt::GetLastError
#define cudaGetErrorString emurt::GetErrorString
#define cudaGetDeviceProperties emurt::GetDeviceProperties
#define cudaDeviceGetAttribute emurt::DeviceGetAttribute
#define cudaOccupancyMaxActiveBlocksPerMultiprocessor emurt::OccupancyMaxActiveBlocksPerMultiprocessor
#define cudaFuncSetAttribute emurt::FuncSetAttribute
#define cudaPointerGetAttributes emurt::PointerGetAttributes
#define cudaIpcGetMemHandle emurt::IpcGetMemHandle
#define cudaIpcOpenMemHandle emurt::IpcOpenMemHandle
#define cudaIpcCloseMemHandle emurt::IpcCloseMemHandle
#define cudaLaunchCooperativeKernel emurt::LaunchCooperativeKernel
#define cudaStreamBeginCapture emurt::StreamBeginCapture
#define cudaStreamEndCapture emurt::StreamEndCapture
#define cudaGraphInstantiate emurt::GraphInstantiate
#define cudaGraphLaunch emurt::GraphLaunch
#define cudaGraphDestroy emurt::GraphDestroy
#define cudaGraphExecDestroy emurt::GraphExecDestroy
