// fa_ring.cu — the data plane of the K/V ring (BASELINE.json configs[4]) on the copy engines.
//
// New functionality (the reference is single-GPU). Round 1 rotated the shards with NCCL send/recv from Python; at 8
// GPUs the per-step trace (profiles/r2_ring.md) showed every step waiting ~1.3 ms for a 134 MB transfer behind 1.05 ms
// of attention: the NCCL kernels compete for SMs with attention CTAs that fill every SM (213 KB of shared memory
// each). Here one rank's shard goes to the next rank as plain `cudaMemcpyAsync` peer copies on a dedicated stream —
// copy engines over NVLink 5 / NVSwitch, no SM involved — into receive slots the neighbour exported with CUDA IPC:
//
//   fa_ring_create   one cudaMalloc per rank: n_slots receive slots + flag words; an IPC handle blob to hand to the
//                    neighbours (the caller moves the blobs with whatever it has: torch.distributed, MPI, a file)
//   fa_ring_connect  opens the next rank's blob (data + `ready` flags are written there) and the previous rank's (its
//                    `free` flags are written there)
//   fa_ring_send     copy stream: wait for the caller's stream (event), for the source slot's data when a received
//                    shard is being forwarded, and for the destination slot to be free; copy; raise `ready` at the receiver
//   fa_ring_recv_wait     the given stream waits (cuStreamWaitValue32, no kernel, no host sync) until the slot is filled
//   fa_ring_recv_release  once the stream's work so far and the forwarding copy are done, raise `free` at the sender
//
// Flags are 32-bit fill / drain counters per slot; a flag is raised by a 4-byte peer copy from a local table of
// integers (so signalling also stays on the copy engine and is ordered behind the payload in the same stream).
// Every rank makes the same sequence of calls, so the expected counter values are known on both sides without a
// handshake. Nothing here launches a kernel.
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include <vector>

#include "../../include/fa_b200.h"

namespace {

constexpr int kMaxSlots = 8;
constexpr int kTable = 1 << 16;   // fill counts wrap far beyond any run; the table is replenished modulo its size

struct Blob {             // what fa_ring_create hands out (FA_RING_HANDLE_BYTES)
  cudaIpcMemHandle_t mem;
  uint64_t slot_bytes;
  int32_t n_slots, rank;
};
static_assert(sizeof(Blob) <= FA_RING_HANDLE_BYTES, "handle blob larger than the public constant");

typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

}  // namespace

struct fa_ring {
  int rank = 0, world = 1, n_slots = 0;
  size_t slot_bytes = 0, flags_off = 0, total = 0;
  char* local = nullptr;        // [slots][ready flags][free flags][table]
  char* next = nullptr;         // peer mapping of the next rank's allocation (data + ready flags)
  char* prev = nullptr;         // peer mapping of the previous rank's allocation (free flags)
  cudaStream_t copy = nullptr;
  std::vector<cudaEvent_t> events;
  size_t ev_next = 0;
  uint32_t sent[kMaxSlots] = {}, expected[kMaxSlots] = {}, released[kMaxSlots] = {};
  WaitValue32Fn wait32 = nullptr;
  // layout helpers
  uint32_t* ready(char* base, int s) const { return reinterpret_cast<uint32_t*>(base + flags_off) + s; }
  uint32_t* freed(char* base, int s) const { return reinterpret_cast<uint32_t*>(base + flags_off) + kMaxSlots + s; }
  const uint32_t* table(uint32_t v) const {
    return reinterpret_cast<const uint32_t*>(local + flags_off + 2 * kMaxSlots * 4) + (v % kTable);
  }
  cudaEvent_t event() {
    if (events.size() < 64) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
      events.push_back(e);
      return e;
    }
    return events[ev_next++ % events.size()];
  }
};

namespace {
int cu_fail() { return FA_ECUDA; }
#define RING_CU(x)                         \
  do {                                     \
    if ((x) != cudaSuccess) return cu_fail(); \
  } while (0)
}  // namespace

extern "C" {

int fa_ring_create(int32_t rank, int32_t world, size_t slot_bytes, int32_t n_slots, fa_ring_t** out, void* handle_blob) {
  if (!out || !handle_blob) return FA_EINVAL_NULL;
  if (world < 1 || rank < 0 || rank >= world || n_slots < 1 || n_slots > kMaxSlots || slot_bytes == 0)
    return FA_EINVAL_SHAPE;
  fa_ring* r = new fa_ring();
  r->rank = rank;
  r->world = world;
  r->n_slots = n_slots;
  r->slot_bytes = (slot_bytes + 255) & ~size_t(255);
  r->flags_off = r->slot_bytes * n_slots;
  r->total = r->flags_off + 2 * kMaxSlots * 4 + size_t(kTable) * 4;
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
    delete r;
    return FA_ENODEVICE;
  }
  r->wait32 = reinterpret_cast<WaitValue32Fn>(f);
  if (cudaMalloc(&r->local, r->total) != cudaSuccess) {
    delete r;
    return cu_fail();
  }
  std::vector<uint32_t> init((2 * kMaxSlots) + kTable, 0u);
  for (int i = 0; i < kTable; ++i) init[2 * kMaxSlots + i] = uint32_t(i);
  Blob b{};
  if (cudaMemcpy(r->local + r->flags_off, init.data(), init.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaStreamCreateWithFlags(&r->copy, cudaStreamNonBlocking) != cudaSuccess ||
      cudaIpcGetMemHandle(&b.mem, r->local) != cudaSuccess) {
    fa_ring_destroy(r);
    return cu_fail();
  }
  b.slot_bytes = r->slot_bytes;
  b.n_slots = n_slots;
  b.rank = rank;
  memset(handle_blob, 0, FA_RING_HANDLE_BYTES);
  memcpy(handle_blob, &b, sizeof(b));
  *out = r;
  return FA_OK;
}

int fa_ring_connect(fa_ring_t* r, const void* next_blob, const void* prev_blob) {
  if (!r) return FA_EINVAL_NULL;
  if (r->world == 1) return FA_OK;
  if (!next_blob || !prev_blob) return FA_EINVAL_NULL;
  Blob nb, pb;
  memcpy(&nb, next_blob, sizeof(nb));
  memcpy(&pb, prev_blob, sizeof(pb));
  if (nb.slot_bytes != r->slot_bytes || nb.n_slots != r->n_slots || pb.slot_bytes != r->slot_bytes ||
      pb.n_slots != r->n_slots || nb.rank != (r->rank + 1) % r->world || pb.rank != (r->rank + r->world - 1) % r->world)
    return FA_EINVAL_SHAPE;
  RING_CU(cudaIpcOpenMemHandle(reinterpret_cast<void**>(&r->next), nb.mem, cudaIpcMemLazyEnablePeerAccess));
  if (r->world == 2) {
    r->prev = r->next;   // the same neighbour on both sides: one mapping
  } else {
    RING_CU(cudaIpcOpenMemHandle(reinterpret_cast<void**>(&r->prev), pb.mem, cudaIpcMemLazyEnablePeerAccess));
  }
  return FA_OK;
}

void* fa_ring_slot(fa_ring_t* r, int32_t slot) {
  if (!r || slot < 0 || slot >= r->n_slots) return nullptr;
  return r->local + size_t(slot) * r->slot_bytes;
}

int fa_ring_send(fa_ring_t* r, int32_t dst_slot, int32_t n_parts, const void* const* src, const size_t* bytes,
                 int32_t forwarded_slot, void* after_stream) {
  if (!r || !src || !bytes) return FA_EINVAL_NULL;
  if (r->world == 1) return FA_EINVAL_SHAPE;
  if (dst_slot < 0 || dst_slot >= r->n_slots || forwarded_slot >= r->n_slots || n_parts < 1) return FA_EINVAL_SHAPE;
  size_t total = 0;
  for (int i = 0; i < n_parts; ++i) total += (bytes[i] + 255) & ~size_t(255);
  if (total > r->slot_bytes) return FA_EINVAL_WORKSPACE;
  // the payload is final once the caller's stream reaches this point ...
  cudaEvent_t e = r->event();
  if (!e) return cu_fail();
  RING_CU(cudaEventRecord(e, static_cast<cudaStream_t>(after_stream)));
  RING_CU(cudaStreamWaitEvent(r->copy, e, 0));
  // ... and, when it sits in one of our receive slots, once that slot has been filled
  if (forwarded_slot >= 0 &&
      r->wait32(r->copy, reinterpret_cast<CUdeviceptr>(r->ready(r->local, forwarded_slot)), r->expected[forwarded_slot],
                CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
    return cu_fail();
  // the receiver has drained what we put into this slot before
  if (r->sent[dst_slot] > 0 &&
      r->wait32(r->copy, reinterpret_cast<CUdeviceptr>(r->freed(r->local, dst_slot)), r->sent[dst_slot],
                CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
    return cu_fail();
  char* dst = r->next + size_t(dst_slot) * r->slot_bytes;
  for (int i = 0; i < n_parts; ++i) {
    RING_CU(cudaMemcpyAsync(dst, src[i], bytes[i], cudaMemcpyDeviceToDevice, r->copy));
    dst += (bytes[i] + 255) & ~size_t(255);
  }
  r->sent[dst_slot] += 1;
  RING_CU(cudaMemcpyAsync(r->ready(r->next, dst_slot), r->table(r->sent[dst_slot]), 4, cudaMemcpyDeviceToDevice, r->copy));
  return FA_OK;
}

int fa_ring_recv_wait(fa_ring_t* r, int32_t slot, void* stream) {
  if (!r) return FA_EINVAL_NULL;
  if (slot < 0 || slot >= r->n_slots || r->world == 1) return FA_EINVAL_SHAPE;
  r->expected[slot] += 1;
  if (r->wait32(static_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(r->ready(r->local, slot)), r->expected[slot],
                CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
    return cu_fail();
  return FA_OK;
}

int fa_ring_recv_release(fa_ring_t* r, int32_t slot, void* stream) {
  if (!r) return FA_EINVAL_NULL;
  if (slot < 0 || slot >= r->n_slots || r->world == 1) return FA_EINVAL_SHAPE;
  cudaEvent_t e = r->event();
  if (!e) return cu_fail();
  RING_CU(cudaEventRecord(e, static_cast<cudaStream_t>(stream)));
  RING_CU(cudaStreamWaitEvent(r->copy, e, 0));      // the copy stream also holds the forwarding copy of this slot
  r->released[slot] += 1;
  RING_CU(cudaMemcpyAsync(r->freed(r->prev, slot), r->table(r->released[slot]), 4, cudaMemcpyDeviceToDevice, r->copy));
  return FA_OK;
}

int fa_ring_destroy(fa_ring_t* r) {
  if (!r) return FA_OK;
  if (r->copy) cudaStreamSynchronize(r->copy);
  if (r->next) cudaIpcCloseMemHandle(r->next);
  if (r->prev && r->prev != r->next) cudaIpcCloseMemHandle(r->prev);
  for (auto e : r->events) cudaEventDestroy(e);
  if (r->copy) cudaStreamDestroy(r->copy);
  if (r->local) cudaFree(r->local);
  delete r;
  return FA_OK;
}

}  // extern "C"
