// Peer-memory ring context: the NVSwitch-native replacement for the reference's lax.ppermute K/V rotation
// (un-vendored `ringattention` package; call site lwm/llama.py:539-569, ring exchange described in SURVEY.md §2.1 /
// Appendix A).
//
// On an NVSwitch box every GPU maps every peer's HBM at full NVLink bandwidth, so the "ring" needs no two-sided
// send/recv: each rank owns one heap (cudaMalloc + cudaIpc handle), maps its peers' heaps, and the executor
//   * PULLS K/V (and Q / dO) blocks out of the owners' heaps with copy-engine cudaMemcpyAsync — no SM is taken from
//     the attention kernels, nothing has to be matched by the peer;
//   * PUTS results (dK/dV partials, O / dQ chunks) into landing slots of the owner's heap the same way;
//   * orders everything with 32-bit flags in the heaps: a remote flag write enqueued after the payload copy on the
//     same stream, and cuStreamWaitValue32(GEQ) on the local flag — stream-ordered, no host synchronisation.
// Ownership: the context owns the heap, the peer mappings and a small table of constants; tensors stay caller-owned.
// One host thread per rank; calls are asynchronous on the given stream.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include "capi_internal.h"
#include "../../include/lwm_b200.h"

namespace {

typedef CUresult (*wait32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned);
typedef CUresult (*write32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned);
typedef CUresult (*memset32_fn)(CUdeviceptr, unsigned, size_t, CUstream);

void* driver_entry(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return fn;
}

constexpr int kMaxWorld = LWM_RING_MAX_WORLD;
constexpr int kValTable = 4096;   // constants for the 4-byte-copy signal (value modulo table size)

}  // namespace

struct lwm_ring_ctx {
  int rank, world, device;
  long long heap_bytes;       // including the flag page
  unsigned char* heap[kMaxWorld];   // heap[rank] = own allocation, others = IPC mappings (null until opened)
  bool opened;
  int signal_mode;            // 0 cuStreamWriteValue32 on the peer pointer, 1 cuMemsetD32Async, 2 4-byte copy
  unsigned* val_table;        // device: val_table[i] = i
  wait32_fn f_wait;
  write32_fn f_write;
  memset32_fn f_memset;
};

static bool ctx_ok(lwm_ring_ctx* c, const char* who) {
  if (!c) { lwm_fail(LWM_ERR_ARG, "ring ctx: null context"); return false; }
  (void)who;
  return true;
}

extern "C" int lwm_ring_ctx_create(int rank, int world, long long heap_bytes, int signal_mode, lwm_ring_ctx** out) {
  if (!out) return lwm_fail(LWM_ERR_ARG, "ring_ctx_create: null out");
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world)
    return lwm_fail(LWM_ERR_ARG, "ring_ctx_create: bad rank/world (world <= 16)");
  if (heap_bytes < 0 || signal_mode < 0 || signal_mode > 2) return lwm_fail(LWM_ERR_ARG, "ring_ctx_create: bad heap size / signal mode");
  if (!lwm_check_device()) return LWM_ERR_DEVICE;
  *out = nullptr;
  lwm_ring_ctx* c = new lwm_ring_ctx();
  memset(c, 0, sizeof(*c));
  c->rank = rank; c->world = world; c->signal_mode = signal_mode;
  cudaGetDevice(&c->device);
  c->heap_bytes = LWM_RING_FLAG_BYTES + ((heap_bytes + 255) & ~255LL);
  c->f_wait = (wait32_fn)driver_entry("cuStreamWaitValue32");
  c->f_write = (write32_fn)driver_entry("cuStreamWriteValue32");
  c->f_memset = (memset32_fn)driver_entry("cuMemsetD32Async");
  if (!c->f_wait || (signal_mode == 0 && !c->f_write) || (signal_mode == 1 && !c->f_memset)) {
    delete c;
    return lwm_fail(LWM_ERR_CUDA, "ring_ctx_create: stream memory operations are not available in this driver");
  }
  void* p = nullptr;
  if (cudaMalloc(&p, size_t(c->heap_bytes)) != cudaSuccess) {
    cudaGetLastError();
    delete c;
    return lwm_fail(LWM_ERR_CUDA, "ring_ctx_create: cudaMalloc of the heap failed");
  }
  c->heap[rank] = reinterpret_cast<unsigned char*>(p);
  unsigned h[kValTable];
  for (int i = 0; i < kValTable; ++i) h[i] = unsigned(i);
  if (cudaMemset(p, 0, LWM_RING_FLAG_BYTES) != cudaSuccess || cudaMalloc(&c->val_table, sizeof(h)) != cudaSuccess ||
      cudaMemcpy(c->val_table, h, sizeof(h), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(p);
    delete c;
    return lwm_fail(LWM_ERR_CUDA, "ring_ctx_create: flag page / constant table setup failed");
  }
  *out = c;
  return LWM_OK;
}

extern "C" int lwm_ring_ctx_get_handle(lwm_ring_ctx* c, void* handle64) {
  if (!ctx_ok(c, "get_handle") || !handle64) return lwm_fail(LWM_ERR_ARG, "ring_ctx_get_handle: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == LWM_RING_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, c->heap[c->rank]) != cudaSuccess) {
    cudaGetLastError();
    return lwm_fail(LWM_ERR_CUDA, "ring_ctx_get_handle: cudaIpcGetMemHandle failed");
  }
  memcpy(handle64, &h, sizeof(h));
  return LWM_OK;
}

extern "C" int lwm_ring_ctx_open_peers(lwm_ring_ctx* c, const void* handles) {
  if (!ctx_ok(c, "open_peers") || !handles) return lwm_fail(LWM_ERR_ARG, "ring_ctx_open_peers: null argument");
  if (c->opened) return lwm_fail(LWM_ERR_ARG, "ring_ctx_open_peers: already opened");
  const unsigned char* hs = reinterpret_cast<const unsigned char*>(handles);
  for (int p = 0; p < c->world; ++p) {
    if (p == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hs + size_t(p) * LWM_RING_HANDLE_BYTES, sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      char buf[256];
      snprintf(buf, sizeof(buf), "ring_ctx_open_peers: cudaIpcOpenMemHandle(rank %d) failed: %s", p, cudaGetErrorString(e));
      return lwm_fail(LWM_ERR_CUDA, buf);
    }
    c->heap[p] = reinterpret_cast<unsigned char*>(ptr);
  }
  c->opened = true;
  return LWM_OK;
}

// Base address (in THIS process) of rank `peer`'s heap payload area (after the flag page); null on error.
extern "C" void* lwm_ring_ctx_heap(lwm_ring_ctx* c, int peer) {
  if (!c || peer < 0 || peer >= c->world || !c->heap[peer]) {
    lwm_fail(LWM_ERR_ARG, "ring_ctx_heap: bad peer or peers not opened");
    return nullptr;
  }
  return c->heap[peer] + LWM_RING_FLAG_BYTES;
}

extern "C" long long lwm_ring_ctx_heap_bytes(lwm_ring_ctx* c) { return c ? c->heap_bytes - LWM_RING_FLAG_BYTES : 0; }

// Copy-engine transfer between any two mapped addresses (local <-> peer heap): pull or put.
extern "C" int lwm_ring_copy(void* dst, const void* src, long long bytes, void* stream) {
  if (!dst || !src || bytes < 0) return lwm_fail(LWM_ERR_ARG, "ring_copy: bad arguments");
  if (bytes == 0) return LWM_OK;
  if (cudaMemcpyAsync(dst, src, size_t(bytes), cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)) !=
      cudaSuccess) {
    cudaError_t e = cudaGetLastError();
    char buf[200];
    snprintf(buf, sizeof(buf), "ring_copy: cudaMemcpyAsync failed: %s", cudaGetErrorString(e));
    return lwm_fail(LWM_ERR_CUDA, buf);
  }
  return LWM_OK;
}

// flags[flag] of rank `peer` := value, after everything enqueued on `stream` so far (payload copies included).
extern "C" int lwm_ring_signal(lwm_ring_ctx* c, int peer, int flag, unsigned value, void* stream) {
  if (!ctx_ok(c, "signal")) return LWM_ERR_ARG;
  if (peer < 0 || peer >= c->world || flag < 0 || flag >= LWM_RING_NUM_FLAGS || !c->heap[peer])
    return lwm_fail(LWM_ERR_ARG, "ring_signal: bad peer / flag index, or peers not opened");
  unsigned* dst = reinterpret_cast<unsigned*>(c->heap[peer]) + flag;
  CUresult r = CUDA_SUCCESS;
  if (c->signal_mode == 0) {
    r = c->f_write(reinterpret_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(dst), value, 0);
  } else if (c->signal_mode == 1) {
    r = c->f_memset(reinterpret_cast<CUdeviceptr>(dst), value, 1, reinterpret_cast<CUstream>(stream));
  } else {
    if (value >= unsigned(kValTable)) return lwm_fail(LWM_ERR_ARG, "ring_signal: value exceeds the constant table (mode 2)");
    if (cudaMemcpyAsync(dst, c->val_table + value, 4, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)) !=
        cudaSuccess) {
      cudaGetLastError();
      r = CUDA_ERROR_UNKNOWN;
    }
  }
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof(buf), "ring_signal: remote flag write failed (mode %d, CUresult %d)", c->signal_mode, int(r));
    return lwm_fail(LWM_ERR_CUDA, buf);
  }
  return LWM_OK;
}

// `stream` does not proceed until my flags[flag] >= value (unsigned 32-bit compare).
extern "C" int lwm_ring_wait(lwm_ring_ctx* c, int flag, unsigned value, void* stream) {
  if (!ctx_ok(c, "wait")) return LWM_ERR_ARG;
  if (flag < 0 || flag >= LWM_RING_NUM_FLAGS) return lwm_fail(LWM_ERR_ARG, "ring_wait: bad flag index");
  unsigned* src = reinterpret_cast<unsigned*>(c->heap[c->rank]) + flag;
  CUresult r = c->f_wait(reinterpret_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(src), value,
                         CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS) {
    char buf[120];
    snprintf(buf, sizeof(buf), "ring_wait: cuStreamWaitValue32 failed (CUresult %d)", int(r));
    return lwm_fail(LWM_ERR_CUDA, buf);
  }
  return LWM_OK;
}

extern "C" int lwm_ring_ctx_destroy(lwm_ring_ctx* c) {
  if (!c) return LWM_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int p = 0; p < c->world; ++p)
    if (p != c->rank && c->heap[p]) cudaIpcCloseMemHandle(c->heap[p]);
  if (c->heap[c->rank]) cudaFree(c->heap[c->rank]);
  if (c->val_table) cudaFree(c->val_table);
  cudaGetLastError();
  delete c;
  return LWM_OK;
}
