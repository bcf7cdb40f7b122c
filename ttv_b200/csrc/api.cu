// api.cu -- the C-ABI of include/ttv_b200.h: validation, pointer classification, host staging, workspace, launch.
//
// This is the drop-in boundary for the reference's low-level interface tlib::ttv::ttv (include/tlib/ttv.h:54-92).
// There is no CPU fallback: if no CUDA device can be used every compute entry returns TTV_B200_ERR_CUDA.
#include "../../include/ttv_b200.h"
#include "launch.h"
#include "plan.h"
#include "copy_pool.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>
#include <atomic>
#include <condition_variable>
#include <memory>
#include <thread>
#include <algorithm>
#include <new>

using namespace ttvb;

namespace {

thread_local std::string g_last_error;

int fail(int status, const char* fmt = nullptr, ...)
{
  g_last_error = status_message(status);
  if (fmt) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error += " [";
    g_last_error += buf;
    g_last_error += "]";
  }
  return status;
}

int fail_cuda(cudaError_t e, const char* what)
{
  cudaGetLastError();   // clear the sticky-less error state
  return fail(TTV_B200_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define CUDA_TRY(expr, what) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return fail_cuda(e__, what); } while (0)

// ---- per-device state ------------------------------------------------------------------------------------------
struct Buffer {
  void*  ptr = nullptr;
  size_t bytes = 0;
};

struct DeviceState {
  int sm_count = 0;
  // Host-pointer calls share the staging buffers, the chunk ring and the copy stream of their device: one such call at a
  // time per device (they are synchronous and PCIe-bound, so nothing is lost).  Always taken BEFORE g_mutex.
  std::mutex host_mutex;
  Buffer stage_a, stage_b, stage_c;           // staging for host-pointer calls
  // chunked host path: a copy stream and a ring of chunk buffers with their events
  cudaStream_t copy_stream = nullptr;
  Buffer ring[3];
  cudaEvent_t ready[3] = {nullptr, nullptr, nullptr}, freed[3] = {nullptr, nullptr, nullptr};
  // pageable host memory: pinned bounce buffers, one per ring slot, filled by the copy threads
  void*  bounce[3] = {nullptr, nullptr, nullptr};
  size_t bounce_bytes = 0;
  // ... and the way back: chunks of C on their way into pageable host memory
  void*  bounce_out[3] = {nullptr, nullptr, nullptr};
  size_t bounce_out_bytes = 0;
  cudaEvent_t out_done[3] = {nullptr, nullptr, nullptr};
  // stream-ordered pool for the intermediates of a chain of products (ttv_b200_ttvs)
  cudaMemPool_t pool = nullptr;
};

std::mutex g_mutex;                                        // guards g_devices and the buffer tables inside a DeviceState
std::map<int, std::unique_ptr<DeviceState>> g_devices;     // entries are never erased: a DeviceState* stays valid

// caller holds g_mutex
int device_state(int device, DeviceState** out)
{
  auto it = g_devices.find(device);
  if (it == g_devices.end()) {
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device), "cudaDeviceGetAttribute");
    std::unique_ptr<DeviceState> st(new DeviceState);
    st->sm_count = sms;
    it = g_devices.emplace(device, std::move(st)).first;
  }
  *out = it->second.get();
  return TTV_B200_OK;
}

int device_state_locked(int device, DeviceState** out)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  return device_state(device, out);
}

int ensure(Buffer& buf, size_t bytes)
{
  if (buf.bytes >= bytes && buf.ptr) return TTV_B200_OK;
  if (buf.ptr) { cudaFree(buf.ptr); buf.ptr = nullptr; buf.bytes = 0; }
  if (bytes == 0) bytes = 256;
  CUDA_TRY(cudaMalloc(&buf.ptr, bytes), "cudaMalloc");
  buf.bytes = bytes;
  return TTV_B200_OK;
}

enum class Where { Host, Device };

int classify(const void* p, Where* where, int* device)
{
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) return fail_cuda(e, "cudaPointerGetAttributes (is a CUDA device visible?)");
  if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) { *where = Where::Device; *device = attr.device; }
  else { *where = Where::Host; *device = -1; }
  return TTV_B200_OK;
}

bool is_pinned_host(const void* p)
{
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return attr.type == cudaMemoryTypeHost;
}

CopyPool& copy_pool()
{
  static std::unique_ptr<CopyPool> pool;
  static std::once_flag once;
  std::call_once(once, [] {
    int n = 0;
    if (const char* e = std::getenv("TTV_B200_COPY_THREADS")) n = std::atoi(e);
    if (n <= 0) n = (int)std::min<unsigned>(12, std::max<unsigned>(2, std::thread::hardware_concurrency() * 3 / 4));
    pool.reset(new CopyPool(n - 1));                          // the caller is the n-th copier
  });
  return *pool;
}

uint64_t alignment_of(const void* p)
{
  const uintptr_t x = reinterpret_cast<uintptr_t>(p);
  return x ? (uint64_t)(x & (~x + 1)) : 256;
}

struct DeviceGuard {
  int prev = -1;
  bool active = false;
  cudaError_t set(int device) {
    cudaError_t e = cudaGetDevice(&prev);
    if (e != cudaSuccess) return e;
    if (prev != device) { e = cudaSetDevice(device); active = (e == cudaSuccess); }
    return e;
  }
  ~DeviceGuard() { if (active) cudaSetDevice(prev); }
};

// the stream-ordered pool of a device: split-n_q partials and the intermediates of a chain (ttv_b200_ttvs)
int chain_pool(int device, cudaMemPool_t* out)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  DeviceState* st = nullptr;
  if (int rc = device_state(device, &st)) return rc;
  if (!st->pool) {
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    CUDA_TRY(cudaMemPoolCreate(&st->pool, &props), "cudaMemPoolCreate");
    // blocks of up to 1 GiB stay with the pool across synchronisations (a chain on the same shapes allocates nothing new);
    // anything above goes back to the driver so that other allocators of the process can have it
    uint64_t keep = 1ull << 30;
    CUDA_TRY(cudaMemPoolSetAttribute(st->pool, cudaMemPoolAttrReleaseThreshold, &keep), "cudaMemPoolSetAttribute");
  }
  *out = st->pool;
  return TTV_B200_OK;
}

// runs the canonical view with device pointers on `device`
int run_view_device(int dtype, const View& v, const void* a, const void* b, void* c, const ttv_b200_opts* opts,
                    int device, bool sync)
{
  cudaStream_t stream = opts ? static_cast<cudaStream_t>(opts->stream) : nullptr;
  const bool accumulate = opts && (opts->flags & TTV_B200_FLAG_ACCUMULATE);

  DeviceState* st = nullptr;
  if (int rc = device_state_locked(device, &st)) return rc;

  if (v.strided) {
    CUDA_TRY(launch_strided(dtype, v, a, b, c, accumulate, st->sm_count, stream), "kernel launch");
    if (sync) CUDA_TRY(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
    return TTV_B200_OK;
  }
  Launch l;
  int rc = choose_launch(dtype, v, opts, alignment_of(a), alignment_of(b), alignment_of(c), st->sm_count, &l);
  if (rc) return fail(rc);

  // The partials of a split n_q live in a stream-ordered allocation of this call only: the tile kernel writes them, the
  // reduce kernel reads them, cudaFreeAsync hands the block back behind the reduce.  Calls on the same stream from several
  // host threads (torch's default stream is shared by all of them) therefore never see each other's partials, whichever way
  // their launches interleave, and nothing is freed under a kernel that still reads it.
  void* ws = nullptr;
  if (l.workspace_bytes) {
    cudaMemPool_t pool = nullptr;
    if (int r2 = chain_pool(device, &pool)) return r2;
    CUDA_TRY(cudaMallocFromPoolAsync(&ws, (size_t)l.workspace_bytes, pool, stream), "cudaMallocFromPoolAsync (split-n_q partials)");
  }
  cudaError_t le = launch_view(dtype, v, l, a, b, c, ws, accumulate, st->sm_count, stream);
  if (ws && cudaFreeAsync(ws, stream) != cudaSuccess) cudaGetLastError();
  CUDA_TRY(le, "kernel launch");
  if (sync) CUDA_TRY(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  return TTV_B200_OK;
}

int env_mb(const char* name, int fallback)
{
  const char* e = std::getenv(name);
  return (e && *e) ? std::atoi(e) : fallback;
}

// Host pointers, large A: stream A across PCIe in chunks of whole slabs of its slowest mode and run the kernels of chunk
// i while chunk i+1 is still in flight.  A product whose mode q is NOT the slowest mode gets the
// matching rows of C from every chunk (free split, C chunk goes back right away: PCIe is full duplex); a product that
// contracts the slowest mode accumulates its chunk of n_q into the whole C, which goes back at the end.  Only three chunk
// buffers of A live on the device, so the tensor may be larger than what is free in HBM.
// Returns -1 when the call is not eligible (small, strided, one indivisible slab): the caller takes the plain path.
int run_host_pipelined(int dtype, uint64_t count, const View* views, const void* a, const void* const* b, void* const* c,
                       const ttv_b200_opts* opts, int device, bool bc_dev = false, char* resident = nullptr)
{
  // resident: the chunks of A land side by side in this device buffer (|A| bytes) instead of the three-slot ring, and stay
  // there for the products that follow (ttv_b200_run_resident); nothing is ever overwritten, so no slot has to be freed
  // bc_dev: the vectors and the results are DEVICE buffers (the first product of a chain): nothing of them is staged
  const size_t s = (size_t)dtype_size(dtype);
  const View& v0 = views[0];
  const uint64_t total = v0.outer * v0.nq * v0.inner;
  const uint64_t slow = v0.slow_extent ? v0.slow_extent : (v0.outer > 1 ? v0.outer : v0.nq);
  // pageable A: smaller chunks, bounced through pinned buffers by the copy threads (memcpy of chunk i+1 overlaps the DMA
  // of chunk i); pinned A: DMA straight from the caller's buffer
  const bool bounce = !is_pinned_host(a) && env_mb("TTV_B200_BOUNCE", 1) != 0;
  size_t chunk_target = bounce ? std::min<size_t>((size_t)env_mb("TTV_B200_H2D_CHUNK_MB", 128), (size_t)env_mb("TTV_B200_BOUNCE_CHUNK_MB", 32)) << 20
                               : (size_t)env_mb("TTV_B200_H2D_CHUNK_MB", 128) << 20;
  // Pageable tensors of a few hundred MB: with 32 MiB chunks the memcpy of the first chunk and the DMA of the last one are
  // a large share of the whole transfer.  About 16 chunks per tensor (not below 4 MiB each) keep the pipeline full.
  if (bounce && env_mb("TTV_B200_BOUNCE_ADAPT", 1) != 0)
    chunk_target = std::min(chunk_target, std::max<size_t>((size_t)4 << 20, (size_t)(total * s) / 16));
  const bool debug = env_mb("TTV_B200_DEBUG", 0) != 0;
  if (debug) fprintf(stderr, "[ttv_b200] host path: total=%llu slow=%llu chunk_target=%zu count=%llu\n", (unsigned long long)total,
                     (unsigned long long)slow, chunk_target, (unsigned long long)count);
  if (chunk_target == 0 || slow < 2 || total * s < 2 * chunk_target) return -1;
  const uint64_t slab = total / slow;                                       // elements per index of the slowest mode
  if (slab * s > 4 * chunk_target) return -1;
  for (uint64_t i = 0; i < count; ++i) {
    const View& v = views[i];
    if (v.strided || v.outer * v.nq * v.inner != total) return -1;
    const bool nq_split = v.outer == 1 || (v.outer % slow) != 0;             // mode q is the slowest mode
    if (nq_split && !(v.outer == 1 && v.nq == slow)) return -1;
  }
  const uint64_t per = std::max<uint64_t>(1, chunk_target / (slab * s));     // slabs per chunk
  const uint64_t chunks = (slow + per - 1) / per;
  if (debug) fprintf(stderr, "[ttv_b200] host path: pipelined, %llu chunks of %llu slabs\n", (unsigned long long)chunks, (unsigned long long)per);
  const size_t chunk_bytes = (size_t)(per * slab) * s;

  cudaStream_t stream = opts ? static_cast<cudaStream_t>(opts->stream) : nullptr;
  const bool accumulate = opts && (opts->flags & TTV_B200_FLAG_ACCUMULATE);
  ttv_b200_opts local = opts ? *opts : ttv_b200_opts{-1, 0, 0, 0, 0, 0, 0, 0, nullptr};

  DeviceState* st = nullptr;
  char *db_ = nullptr, *dc_ = nullptr;
  size_t max_b = 0, sum_c = 0;
  for (uint64_t i = 0; i < count; ++i) {
    max_b = std::max(max_b, ((size_t)views[i].nq * s + 255) / 256 * 256);
    sum_c += ((size_t)(views[i].outer * views[i].inner) * s + 255) / 256 * 256;
  }
  // Results of free-split products that go to PAGEABLE memory leave chunk by chunk through pinned buffers too: one D2H of
  // the whole C into pageable memory at the end runs at ~11 GB/s on these hosts, and with a short contraction (n_q = 2..16)
  // C is a large fraction of A.  The copy threads move chunk ch-3 of C to the caller's array before slot r is reused.
  std::vector<char> c_pinned(count), c_bounce(count, 0);
  std::vector<size_t> out_off(count, 0);
  size_t out_chunk_bytes = 0;
  for (uint64_t i = 0; i < count; ++i) {
    c_pinned[i] = (!bc_dev && is_pinned_host(c[i])) ? 1 : 0;
    if (bc_dev || c_pinned[i] || views[i].outer == 1 || env_mb("TTV_B200_BOUNCE", 1) == 0 || env_mb("TTV_B200_BOUNCE_OUT", 1) == 0) continue;
    c_bounce[i] = 1;
    out_off[i] = out_chunk_bytes;
    out_chunk_bytes += ((size_t)(per * (views[i].outer / slow) * views[i].inner) * s + 255) / 256 * 256;
  }
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (int rc = device_state(device, &st)) return rc;
    if (!st->copy_stream) {
      CUDA_TRY(cudaStreamCreateWithFlags(&st->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate");
      for (int r = 0; r < 3; ++r) {
        CUDA_TRY(cudaEventCreateWithFlags(&st->ready[r], cudaEventDisableTiming), "cudaEventCreate");
        CUDA_TRY(cudaEventCreateWithFlags(&st->freed[r], cudaEventDisableTiming), "cudaEventCreate");
      }
    }
    if (!resident) for (int r = 0; r < 3; ++r) if (int rc = ensure(st->ring[r], chunk_bytes)) return rc;
    if (bounce && st->bounce_bytes < chunk_bytes) {
      for (int r = 0; r < 3; ++r) {
        if (st->bounce[r]) { cudaFreeHost(st->bounce[r]); st->bounce[r] = nullptr; }
        CUDA_TRY(cudaHostAlloc(&st->bounce[r], chunk_bytes, cudaHostAllocDefault), "cudaHostAlloc");
      }
      st->bounce_bytes = chunk_bytes;
    }
    if (out_chunk_bytes) {
      for (int r = 0; r < 3; ++r)
        if (!st->out_done[r]) CUDA_TRY(cudaEventCreateWithFlags(&st->out_done[r], cudaEventDisableTiming), "cudaEventCreate");
      if (st->bounce_out_bytes < out_chunk_bytes) {
        for (int r = 0; r < 3; ++r) {
          if (st->bounce_out[r]) { cudaFreeHost(st->bounce_out[r]); st->bounce_out[r] = nullptr; }
          st->bounce_out_bytes = 0;
          CUDA_TRY(cudaHostAlloc(&st->bounce_out[r], out_chunk_bytes, cudaHostAllocDefault), "cudaHostAlloc");
        }
        st->bounce_out_bytes = out_chunk_bytes;
      }
    }
    if (!bc_dev) {
      if (int rc = ensure(st->stage_b, max_b * count)) return rc;
      if (int rc = ensure(st->stage_c, sum_c)) return rc;
      db_ = static_cast<char*>(st->stage_b.ptr); dc_ = static_cast<char*>(st->stage_c.ptr);
    }
  }
  struct Piece { char* dst; const char* src; size_t bytes; };
  std::vector<Piece> pending[3];                        // chunks of C sitting in bounce_out[r], D2H queued behind out_done[r]
  auto drain = [&](int r) -> int {
    if (pending[r].empty()) return TTV_B200_OK;
    CUDA_TRY(cudaEventSynchronize(st->out_done[r]), "cudaEventSynchronize");
    for (const Piece& pc : pending[r]) copy_pool().copy(pc.dst, pc.src, pc.bytes);
    pending[r].clear();
    return TTV_B200_OK;
  };
  // the vectors (and C when it is accumulated into) go first, on the compute stream
  std::vector<char*> dci(count);
  std::vector<const char*> dbi(count);
  size_t coff = 0;
  for (uint64_t i = 0; i < count; ++i) {
    if (bc_dev) { dci[i] = static_cast<char*>(c[i]); dbi[i] = static_cast<const char*>(b[i]); continue; }
    const size_t bytes_c = (size_t)(views[i].outer * views[i].inner) * s;
    dci[i] = dc_ + coff;
    dbi[i] = db_ + i * max_b;
    coff += (bytes_c + 255) / 256 * 256;
    CUDA_TRY(cudaMemcpyAsync(db_ + i * max_b, b[i], (size_t)views[i].nq * s, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync H2D b");
    if (accumulate) CUDA_TRY(cudaMemcpyAsync(dci[i], c[i], bytes_c, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync H2D C");
  }
  // the copy stream must not run ahead of work already queued on the caller's stream that still reads the ring
  CUDA_TRY(cudaEventRecord(st->freed[0], stream), "cudaEventRecord");
  CUDA_TRY(cudaStreamWaitEvent(st->copy_stream, st->freed[0], 0), "cudaStreamWaitEvent");

  const char* ah = static_cast<const char*>(a);
  for (uint64_t ch = 0; ch < chunks; ++ch) {
    const int r = (int)(ch % 3);
    const uint64_t s0 = ch * per, s1 = std::min(slow, s0 + per), ns = s1 - s0;
    if (ch >= 3 && !resident) CUDA_TRY(cudaStreamWaitEvent(st->copy_stream, st->freed[r], 0), "cudaStreamWaitEvent");
    if (int rc = drain(r)) return rc;                   // chunk ch-3 of C leaves bounce_out[r] before it is refilled
    void* const slot = resident ? static_cast<void*>(resident + (size_t)(s0 * slab) * s) : st->ring[r].ptr;
    const char* src = ah + (size_t)(s0 * slab) * s;
    if (bounce) {
      // bounce[r] was last read by the DMA of chunk ch-3, whose completion is ready[r]
      if (ch >= 3) CUDA_TRY(cudaEventSynchronize(st->ready[r]), "cudaEventSynchronize");
      copy_pool().copy(st->bounce[r], src, (size_t)(ns * slab) * s);
      src = static_cast<const char*>(st->bounce[r]);
    }
    CUDA_TRY(cudaMemcpyAsync(slot, src, (size_t)(ns * slab) * s, cudaMemcpyHostToDevice, st->copy_stream),
             "cudaMemcpyAsync H2D A chunk");
    CUDA_TRY(cudaEventRecord(st->ready[r], st->copy_stream), "cudaEventRecord");
    CUDA_TRY(cudaStreamWaitEvent(stream, st->ready[r], 0), "cudaStreamWaitEvent");
    for (uint64_t i = 0; i < count; ++i) {
      View v = views[i];
      const bool nq_split = v.outer == 1;
      char* cdst = dci[i];
      const char* bsrc = dbi[i];
      if (nq_split) {
        v.nq = ns;
        bsrc += (size_t)s0 * s;
        local.flags = (opts ? opts->flags : 0u) | ((ch > 0 || accumulate) ? (uint32_t)TTV_B200_FLAG_ACCUMULATE : 0u);
      } else {
        const uint64_t rows = v.outer / slow;                               // rows of `outer` per slab
        v.outer = ns * rows;
        cdst += (size_t)(s0 * rows * v.inner) * s;
        local.flags = opts ? opts->flags : 0u;
      }
      local.flags &= ~(uint32_t)TTV_B200_FLAG_ASYNC;
      local.stream = stream;
      if (int rc = run_view_device(dtype, v, slot, bsrc, cdst, &local, device, false)) return rc;
      if (!nq_split && c_pinned[i]) {      // (a copy into pageable memory would block the host and with it the next chunk)
        const size_t bytes = (size_t)(v.outer * v.inner) * s;
        CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(c[i]) + (cdst - dci[i]), cdst, bytes, cudaMemcpyDeviceToHost, stream),
                 "cudaMemcpyAsync D2H C chunk");
      } else if (!nq_split && c_bounce[i]) {
        const size_t bytes = (size_t)(v.outer * v.inner) * s;
        char* hb = static_cast<char*>(st->bounce_out[r]) + out_off[i];
        CUDA_TRY(cudaMemcpyAsync(hb, cdst, bytes, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync D2H C chunk (bounce)");
        pending[r].push_back(Piece{static_cast<char*>(c[i]) + (cdst - dci[i]), hb, bytes});
      }
    }
    if (!pending[r].empty()) CUDA_TRY(cudaEventRecord(st->out_done[r], stream), "cudaEventRecord");
    CUDA_TRY(cudaEventRecord(st->freed[r], stream), "cudaEventRecord");
  }
  for (uint64_t i = 0; i < count && !bc_dev; ++i)
    if (views[i].outer == 1 || (!c_pinned[i] && !c_bounce[i]))
      CUDA_TRY(cudaMemcpyAsync(c[i], dci[i], (size_t)(views[i].outer * views[i].inner) * s, cudaMemcpyDeviceToHost, stream),
               "cudaMemcpyAsync D2H C");
  for (uint64_t ch = chunks > 3 ? chunks - 3 : 0; ch < chunks; ++ch)      // the last chunks of C, oldest first
    if (int rc = drain((int)(ch % 3))) return rc;
  CUDA_TRY(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  return TTV_B200_OK;
}

// host pointers: H2D(A, b) -> kernel -> D2H(C), all on one stream, then wait
int run_view_host(int dtype, const View& v, const void* a, const void* b, void* c, const ttv_b200_opts* opts, bool bc_dev = false)
{
  int device = opts ? opts->device : -1;
  if (device < 0) CUDA_TRY(cudaGetDevice(&device), "cudaGetDevice (is a CUDA device visible?)");
  DeviceGuard guard;
  CUDA_TRY(guard.set(device), "cudaSetDevice");
  DeviceState* hst = nullptr;
  if (int rc = device_state_locked(device, &hst)) return rc;
  std::lock_guard<std::mutex> host_lock(hst->host_mutex);      // staging buffers are per device: one host call at a time

  {
    const void* bs[1] = {b};
    void* cs[1] = {c};
    const int rc = run_host_pipelined(dtype, 1, &v, a, bs, cs, opts, device, bc_dev);
    if (rc >= 0) return rc;
  }
  const size_t s = (size_t)dtype_size(dtype);
  // non-packed strides: the whole span goes across, padding included (C's padding must come back unchanged)
  const size_t bytes_a = (size_t)(v.strided ? v.span_a : v.outer * v.nq * v.inner) * s;
  const size_t bytes_b = (size_t)v.nq * s;
  const size_t bytes_c = (size_t)(v.strided ? v.span_c : v.outer * v.inner) * s;
  cudaStream_t stream = opts ? static_cast<cudaStream_t>(opts->stream) : nullptr;
  const bool c_goes_up = (opts && (opts->flags & TTV_B200_FLAG_ACCUMULATE)) || v.strided;   // C is read (accumulate) or has padding to keep

  void *da = nullptr, *db = nullptr, *dc = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceState* st = nullptr;
    if (int rc = device_state(device, &st)) return rc;
    if (int rc = ensure(st->stage_a, bytes_a)) return rc;
    da = st->stage_a.ptr;
    if (bc_dev) { db = const_cast<void*>(b); dc = c; }
    else {
      if (int rc = ensure(st->stage_b, bytes_b)) return rc;
      if (int rc = ensure(st->stage_c, bytes_c)) return rc;
      db = st->stage_b.ptr; dc = st->stage_c.ptr;
    }
  }
  CUDA_TRY(cudaMemcpyAsync(da, a, bytes_a, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync H2D A");
  if (!bc_dev) {
    CUDA_TRY(cudaMemcpyAsync(db, b, bytes_b, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync H2D b");
    if (c_goes_up) CUDA_TRY(cudaMemcpyAsync(dc, c, bytes_c, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync H2D C");
  }
  if (int rc = run_view_device(dtype, v, da, db, dc, opts, device, false)) return rc;
  if (!bc_dev) CUDA_TRY(cudaMemcpyAsync(c, dc, bytes_c, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync D2H C");
  CUDA_TRY(cudaStreamSynchronize(stream), "cudaStreamSynchronize");   // the staging of A is free again when this returns
  return TTV_B200_OK;
}

int run_any(int dtype, const View& v, const void* a, const void* b, void* c, const ttv_b200_opts* opts)
{
  Where wa = Where::Host, wb = Where::Host, wc = Where::Host;
  int da = -1, db = -1, dc = -1;
  if (int rc = classify(a, &wa, &da)) return rc;
  if (int rc = classify(b, &wb, &db)) return rc;
  if (int rc = classify(c, &wc, &dc)) return rc;
  if (wa != wb || wa != wc) return fail(TTV_B200_ERR_MIXED_POINTERS);
  if (wa == Where::Host) return run_view_host(dtype, v, a, b, c, opts);
  if (da != db || da != dc) return fail(TTV_B200_ERR_MIXED_POINTERS, "a, b, c live on devices %d, %d, %d", da, db, dc);
  DeviceGuard guard;
  CUDA_TRY(guard.set(da), "cudaSetDevice");
  const bool async = opts && (opts->flags & TTV_B200_FLAG_ASYNC);
  return run_view_device(dtype, v, a, b, c, opts, da, !async);
}

// ---- plain copies between host and device, pageable memory pipelined through the pinned bounce buffers --------------------
// Caller holds the device's host_mutex and has the device current.  Both return after the copy has completed.
int ensure_bounce_in(DeviceState* st, size_t chunk)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  if (!st->copy_stream) {
    CUDA_TRY(cudaStreamCreateWithFlags(&st->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate");
    for (int r = 0; r < 3; ++r) {
      CUDA_TRY(cudaEventCreateWithFlags(&st->ready[r], cudaEventDisableTiming), "cudaEventCreate");
      CUDA_TRY(cudaEventCreateWithFlags(&st->freed[r], cudaEventDisableTiming), "cudaEventCreate");
    }
  }
  if (st->bounce_bytes < chunk) {
    for (int r = 0; r < 3; ++r) {
      if (st->bounce[r]) { cudaFreeHost(st->bounce[r]); st->bounce[r] = nullptr; }
      st->bounce_bytes = 0;
      CUDA_TRY(cudaHostAlloc(&st->bounce[r], chunk, cudaHostAllocDefault), "cudaHostAlloc");
    }
    st->bounce_bytes = chunk;
  }
  return TTV_B200_OK;
}

int ensure_bounce_out(DeviceState* st, size_t chunk)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  for (int r = 0; r < 3; ++r)
    if (!st->out_done[r]) CUDA_TRY(cudaEventCreateWithFlags(&st->out_done[r], cudaEventDisableTiming), "cudaEventCreate");
  if (st->bounce_out_bytes < chunk) {
    for (int r = 0; r < 3; ++r) {
      if (st->bounce_out[r]) { cudaFreeHost(st->bounce_out[r]); st->bounce_out[r] = nullptr; }
      st->bounce_out_bytes = 0;
      CUDA_TRY(cudaHostAlloc(&st->bounce_out[r], chunk, cudaHostAllocDefault), "cudaHostAlloc");
    }
    st->bounce_out_bytes = chunk;
  }
  return TTV_B200_OK;
}

size_t bounce_chunk(size_t bytes)
{
  // ~16 chunks per transfer, 4 .. 32 MiB each: the memcpy of the first chunk and the DMA of the last one are the only parts
  // of the pipeline that do not overlap
  return std::min<size_t>((size_t)32 << 20, std::max<size_t>((size_t)4 << 20, (bytes / 16 + 4095) / 4096 * 4096));
}

int copy_h2d(DeviceState* st, void* dst, const void* src, size_t bytes, cudaStream_t stream)
{
  if (bytes == 0) return TTV_B200_OK;
  if (bytes < ((size_t)8 << 20) || is_pinned_host(src) || env_mb("TTV_B200_BOUNCE", 1) == 0) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync H2D");
    CUDA_TRY(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
    return TTV_B200_OK;
  }
  const size_t chunk = bounce_chunk(bytes);
  if (int rc = ensure_bounce_in(st, chunk)) return rc;
  size_t off = 0;
  for (uint64_t ch = 0; off < bytes; ++ch, off += chunk) {
    const int r = (int)(ch % 3);
    const size_t n = std::min(chunk, bytes - off);
    if (ch >= 3) CUDA_TRY(cudaEventSynchronize(st->ready[r]), "cudaEventSynchronize");      // the DMA of chunk ch-3 has read bounce[r]
    copy_pool().copy(st->bounce[r], static_cast<const char*>(src) + off, n);
    CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(dst) + off, st->bounce[r], n, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync H2D chunk");
    CUDA_TRY(cudaEventRecord(st->ready[r], stream), "cudaEventRecord");
  }
  CUDA_TRY(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  return TTV_B200_OK;
}

int copy_d2h(DeviceState* st, void* dst, const void* src, size_t bytes, cudaStream_t stream)
{
  if (bytes == 0) return TTV_B200_OK;
  if (bytes < ((size_t)8 << 20) || is_pinned_host(dst) || env_mb("TTV_B200_BOUNCE", 1) == 0 || env_mb("TTV_B200_BOUNCE_OUT", 1) == 0) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync D2H");
    CUDA_TRY(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
    return TTV_B200_OK;
  }
  const size_t chunk = bounce_chunk(bytes);
  if (int rc = ensure_bounce_out(st, chunk)) return rc;
  const uint64_t chunks = (bytes + chunk - 1) / chunk;
  auto land = [&](uint64_t ch) -> int {                   // chunk ch sits in bounce_out[ch % 3] once out_done[ch % 3] has fired
    const int r = (int)(ch % 3);
    CUDA_TRY(cudaEventSynchronize(st->out_done[r]), "cudaEventSynchronize");
    copy_pool().copy(static_cast<char*>(dst) + ch * chunk, st->bounce_out[r], std::min(chunk, bytes - (size_t)ch * chunk));
    return TTV_B200_OK;
  };
  for (uint64_t ch = 0; ch < chunks; ++ch) {
    const int r = (int)(ch % 3);
    if (ch >= 3) if (int rc = land(ch - 3)) return rc;
    const size_t n = std::min(chunk, bytes - (size_t)ch * chunk);
    CUDA_TRY(cudaMemcpyAsync(st->bounce_out[r], static_cast<const char*>(src) + ch * chunk, n, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync D2H chunk");
    CUDA_TRY(cudaEventRecord(st->out_done[r], stream), "cudaEventRecord");
  }
  for (uint64_t ch = chunks > 3 ? chunks - 3 : 0; ch < chunks; ++ch) if (int rc = land(ch)) return rc;
  return TTV_B200_OK;
}

} // namespace

// a host tensor's twin in HBM (include/ttv_b200.h: ttv_b200_resident)
struct ttv_b200_resident {
  std::mutex  mutex;
  int         device = -1;
  void*       dev = nullptr;       // |A| bytes on `device`
  size_t      capacity = 0;
  const void* host = nullptr;      // the host tensor the copy was made from ...
  size_t      bytes = 0;           // ... and how much of it
  bool        valid = false;
};

namespace {

// ---- the chain of p-1 products (reference ttvpy/src/wrapped_ttv.cpp:83-198) ----------------------------------------
// Which mode each step contracts and with which vector: vector j belongs to mode r = j+1 (j+1 < q) or j+2.
//   backward: r = p, p-1, ... (skipping q)      wrapped_ttv.cpp:135-146
//   forward : r = 1, 2, ...   (skipping q)      wrapped_ttv.cpp:147-156
//   optimal : longest vector first, so that the tensor shrinks as fast as possible (:157-192).  The reference sorts its
//             (vector, mode) pairs ASCENDING by length (std::sort, which is an insertion sort -- stable -- below 16
//             elements) and walks them from the back, so among equal lengths the LARGER mode goes first.
// modes[i] is numbered in the tensor that is left when step i runs (contracted modes drop out, later ones move down).
int chain_plan(uint64_t q, uint64_t p, const uint64_t* na, int order, uint64_t* modes, uint64_t* vectors)
{
  std::vector<uint64_t> seq;                                     // original modes in the order they are contracted
  for (uint64_t r = 1; r <= p; ++r) if (r != q) seq.push_back(r);
  if (order == TTV_B200_CHAIN_BACKWARD) std::reverse(seq.begin(), seq.end());
  else if (order == TTV_B200_CHAIN_OPTIMAL) {
    std::stable_sort(seq.begin(), seq.end(), [na](uint64_t x, uint64_t y) { return na[x - 1] < na[y - 1]; });
    std::reverse(seq.begin(), seq.end());
  }
  std::vector<uint64_t> alive;
  for (uint64_t r = 1; r <= p; ++r) alive.push_back(r);
  for (size_t i = 0; i < seq.size(); ++i) {
    const uint64_t r = seq[i];
    const auto it = std::find(alive.begin(), alive.end(), r);
    modes[i] = (uint64_t)(it - alive.begin()) + 1;
    vectors[i] = r < q ? r - 1 : r - 2;
    alive.erase(it);
  }
  return TTV_B200_OK;
}

// stream-ordered allocations of one chain; whatever is still held goes back to the pool when the chain leaves
struct PoolAllocs {
  cudaMemPool_t pool = nullptr;
  cudaStream_t stream = nullptr;
  std::vector<void*> held;
  int get(void** out, size_t bytes)
  {
    CUDA_TRY(cudaMallocFromPoolAsync(out, std::max<size_t>(bytes, 256), pool, stream), "cudaMallocFromPoolAsync");
    held.push_back(*out);
    return TTV_B200_OK;
  }
  void put(void* ptr)
  {
    auto it = std::find(held.begin(), held.end(), ptr);
    if (it == held.end()) return;
    held.erase(it);
    if (cudaFreeAsync(ptr, stream) != cudaSuccess) cudaGetLastError();
  }
  ~PoolAllocs() { while (!held.empty()) put(held.back()); }
};

} // namespace

// ================================================================================================================
extern "C" {

int ttv_b200_run(int dtype, uint64_t q, uint64_t p,
                 const void* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                 const void* b, const uint64_t* nb,
                 void* c, const uint64_t* nc, const uint64_t* wc, const uint64_t* pic,
                 const ttv_b200_opts* opts)
{
  if (dtype_size(dtype) == 0) return fail(TTV_B200_ERR_DTYPE);
  View v;
  if (int rc = validate_and_fold(q, p, a, na, wa, pia, b, nb, c, nc, wc, pic, opts ? opts->flags : 0u, &v)) return fail(rc);
  return run_any(dtype, v, a, b, c, opts);
}

int ttv_b200_multi(int dtype, uint64_t p,
                   const void* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                   uint64_t count, const uint64_t* q, const void* const* b, void* const* c,
                   const ttv_b200_opts* opts)
{
  if (dtype_size(dtype) == 0) return fail(TTV_B200_ERR_DTYPE);
  if (count == 0) return TTV_B200_OK;
  if (!q || !b || !c) return fail(TTV_B200_ERR_OPTS, "ttv_b200_multi: q, b and c arrays must not be null");
  if (p == 0) return fail(TTV_B200_ERR_ORDER_ZERO);
  if (p > (uint64_t)kMaxOrder) return fail(TTV_B200_ERR_OPTS);
  if (!na) return fail(TTV_B200_ERR_NA_NULL);
  if (!pia) return fail(TTV_B200_ERR_PIA_NULL);

  // fold every product first: nothing is copied or launched unless all of them are valid
  std::vector<View> views(count);
  for (uint64_t i = 0; i < count; ++i) {
    if (q[i] == 0 || q[i] > p) return fail(TTV_B200_ERR_MODE);
    if (!is_valid_shape(na, p)) return fail(TTV_B200_ERR_SHAPE_A);
    if (!is_valid_layout(pia, p)) return fail(TTV_B200_ERR_LAYOUT_A);
    if (p < 2) return fail(TTV_B200_ERR_SHAPE_C);
    uint64_t nc[kMaxOrder], pic[kMaxOrder], wc[kMaxOrder];
    output_shape(na, p, q[i], nc);
    output_layout(pia, p, q[i], pic);
    uint64_t stride = 1;                                   // packed strides of (nc, pic)
    for (uint64_t r = 0; r + 1 < p; ++r) { wc[pic[r] - 1] = stride; stride *= nc[pic[r] - 1]; }
    const uint64_t nb = na[q[i] - 1];
    if (int rc = validate_and_fold(q[i], p, a, na, wa, pia, b[i], &nb, c[i], nc, wc, pic, opts ? opts->flags : 0u, &views[i])) return fail(rc);
  }

  Where wa_ = Where::Host, wx = Where::Host;
  int da = -1, dx = -1;
  if (int rc = classify(a, &wa_, &da)) return rc;
  for (uint64_t i = 0; i < count; ++i) {
    if (int rc = classify(b[i], &wx, &dx)) return rc;
    if (wx != wa_ || (wa_ == Where::Device && dx != da)) return fail(TTV_B200_ERR_MIXED_POINTERS);
    if (int rc = classify(c[i], &wx, &dx)) return rc;
    if (wx != wa_ || (wa_ == Where::Device && dx != da)) return fail(TTV_B200_ERR_MIXED_POINTERS);
  }
  cudaStream_t stream = opts ? static_cast<cudaStream_t>(opts->stream) : nullptr;
  const bool accumulate = opts && (opts->flags & TTV_B200_FLAG_ACCUMULATE);

  if (wa_ == Where::Device) {
    DeviceGuard guard;
    CUDA_TRY(guard.set(da), "cudaSetDevice");
    for (uint64_t i = 0; i < count; ++i)
      if (int rc = run_view_device(dtype, views[i], a, b[i], c[i], opts, da, false)) return rc;
    if (!(opts && (opts->flags & TTV_B200_FLAG_ASYNC))) CUDA_TRY(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
    return TTV_B200_OK;
  }

  // host pointers: A goes to the device once; b_i / C_i are staged per product
  int device = opts ? opts->device : -1;
  if (device < 0) CUDA_TRY(cudaGetDevice(&device), "cudaGetDevice (is a CUDA device visible?)");
  DeviceGuard guard;
  CUDA_TRY(guard.set(device), "cudaSetDevice");
  DeviceState* hst = nullptr;
  if (int rc = device_state_locked(device, &hst)) return rc;
  std::lock_guard<std::mutex> host_lock(hst->host_mutex);      // staging buffers are per device: one host call at a time
  {
    const int rc = run_host_pipelined(dtype, count, views.data(), a, b, c, opts, device);
    if (rc >= 0) return rc;
  }
  const size_t s = (size_t)dtype_size(dtype);
  size_t bytes_a = (size_t)(views[0].outer * views[0].nq * views[0].inner) * s;
  for (uint64_t i = 0; i < count; ++i)
    if (views[i].strided) bytes_a = std::max(bytes_a, (size_t)views[i].span_a * s);     // padded A: the whole span
  size_t max_b = 0, sum_c = 0;
  for (uint64_t i = 0; i < count; ++i) {
    max_b = std::max(max_b, (size_t)views[i].nq * s);
    sum_c += ((size_t)(views[i].outer * views[i].inner) * s + 255) / 256 * 256;
  }
  char *da_ = nullptr, *db_ = nullptr, *dc_ = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceState* st = nullptr;
    if (int rc = device_state(device, &st)) return rc;
    if (int rc = ensure(st->stage_a, bytes_a)) return rc;
    if (int rc = ensure(st->stage_b, max_b * count)) return rc;
    if (int rc = ensure(st->stage_c, sum_c)) return rc;
    da_ = static_cast<char*>(st->stage_a.ptr); db_ = static_cast<char*>(st->stage_b.ptr); dc_ = static_cast<char*>(st->stage_c.ptr);
  }
  CUDA_TRY(cudaMemcpyAsync(da_, a, bytes_a, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync H2D A");
  size_t coff = 0;
  for (uint64_t i = 0; i < count; ++i) {
    const size_t bytes_b = (size_t)views[i].nq * s, bytes_c = (size_t)(views[i].outer * views[i].inner) * s;
    char* dbi = db_ + i * max_b;
    char* dci = dc_ + coff;
    CUDA_TRY(cudaMemcpyAsync(dbi, b[i], bytes_b, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync H2D b");
    if (accumulate) CUDA_TRY(cudaMemcpyAsync(dci, c[i], bytes_c, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync H2D C");
    if (int rc = run_view_device(dtype, views[i], da_, dbi, dci, opts, device, false)) return rc;
    CUDA_TRY(cudaMemcpyAsync(c[i], dci, bytes_c, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync D2H C");
    coff += (bytes_c + 255) / 256 * 256;
  }
  CUDA_TRY(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  return TTV_B200_OK;
}

int ttv_b200_chain_plan(uint64_t q, uint64_t p, const uint64_t* na, int order, uint64_t* modes, uint64_t* vectors)
{
  if (p == 0) return fail(TTV_B200_ERR_ORDER_ZERO);
  if (q == 0 || q > p) return fail(TTV_B200_ERR_MODE);
  if (!na) return fail(TTV_B200_ERR_NA_NULL);
  if (!modes || !vectors) return fail(TTV_B200_ERR_OPTS, "ttv_b200_chain_plan: modes and vectors must not be null");
  if (order < TTV_B200_CHAIN_OPTIMAL || order > TTV_B200_CHAIN_FORWARD) return fail(TTV_B200_ERR_OPTS, "ttv_b200_chain_plan: unknown order %d", order);
  return chain_plan(q, p, na, order, modes, vectors);
}

int ttv_b200_ttvs(int dtype, uint64_t q, uint64_t p, const void* a, const uint64_t* na, const uint64_t* pia,
                  const void* const* b, int order, void* c, const ttv_b200_opts* opts)
{
  const size_t s = (size_t)dtype_size(dtype);
  if (s == 0) return fail(TTV_B200_ERR_DTYPE);
  if (p == 0) return fail(TTV_B200_ERR_ORDER_ZERO);
  if (q == 0 || q > p) return fail(TTV_B200_ERR_MODE);
  if (!a) return fail(TTV_B200_ERR_A_NULL);
  if (!b) return fail(TTV_B200_ERR_B_NULL);
  if (!c) return fail(TTV_B200_ERR_C_NULL);
  if (!na) return fail(TTV_B200_ERR_NA_NULL);
  if (!pia) return fail(TTV_B200_ERR_PIA_NULL);
  if (p > (uint64_t)kMaxOrder) return fail(TTV_B200_ERR_OPTS, "order above %d", kMaxOrder);
  if (!is_valid_shape(na, p)) return fail(TTV_B200_ERR_SHAPE_A);
  if (!is_valid_layout(pia, p)) return fail(TTV_B200_ERR_LAYOUT_A);
  if (p < 2) return fail(TTV_B200_ERR_SHAPE_C);                      // nothing to contract
  if (order < TTV_B200_CHAIN_OPTIMAL || order > TTV_B200_CHAIN_FORWARD) return fail(TTV_B200_ERR_OPTS, "ttv_b200_ttvs: unknown order %d", order);
  const uint64_t steps = p - 1;
  for (uint64_t j = 0; j < steps; ++j) if (!b[j]) return fail(TTV_B200_ERR_B_NULL);

  Where where = Where::Host, wx = Where::Host;
  int dev_a = -1, dx = -1;
  if (int rc = classify(a, &where, &dev_a)) return rc;
  if (int rc = classify(c, &wx, &dx)) return rc;
  if (wx != where || (where == Where::Device && dx != dev_a)) return fail(TTV_B200_ERR_MIXED_POINTERS);
  for (uint64_t j = 0; j < steps; ++j) {
    if (int rc = classify(b[j], &wx, &dx)) return rc;
    if (wx != where || (where == Where::Device && dx != dev_a)) return fail(TTV_B200_ERR_MIXED_POINTERS);
  }
  const bool host = where == Where::Host;
  int device = host ? (opts ? opts->device : -1) : dev_a;
  if (device < 0) CUDA_TRY(cudaGetDevice(&device), "cudaGetDevice (is a CUDA device visible?)");
  DeviceGuard guard;
  CUDA_TRY(guard.set(device), "cudaSetDevice");

  ttv_b200_opts local = opts ? *opts : ttv_b200_opts{-1, TTV_B200_EXEC_PAR_LOOP, TTV_B200_SUBTENSOR, TTV_B200_FUSE_ALL, 0, 0, 0, 0, nullptr};
  const bool async = !host && (local.flags & TTV_B200_FLAG_ASYNC);
  local.device = device;
  local.kernel = 0; local.ksplit = 0;                                  // every step picks its own kernel
  local.flags &= (uint32_t)TTV_B200_FLAG_NO_VEC;                       // no accumulate, no strides, steps never wait
  cudaStream_t stream = static_cast<cudaStream_t>(local.stream);

  uint64_t modes[kMaxOrder], vecs[kMaxOrder];
  chain_plan(q, p, na, order, modes, vecs);

  PoolAllocs mem;
  mem.stream = stream;
  if (int rc = chain_pool(device, &mem.pool)) return rc;

  // host vectors go up once, packed into one block
  std::vector<const void*> dvec(steps);
  if (host) {
    std::vector<size_t> off(steps);
    size_t total = 0;
    for (uint64_t j = 0; j < steps; ++j) {
      const uint64_t r = j + 1 < q ? j + 1 : j + 2;
      off[j] = total;
      total += ((size_t)na[r - 1] * s + 255) / 256 * 256;
    }
    void* block = nullptr;
    if (int rc = mem.get(&block, total)) return rc;
    for (uint64_t j = 0; j < steps; ++j) {
      const uint64_t r = j + 1 < q ? j + 1 : j + 2;
      dvec[j] = static_cast<char*>(block) + off[j];
      CUDA_TRY(cudaMemcpyAsync(const_cast<void*>(dvec[j]), b[j], (size_t)na[r - 1] * s, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync H2D b");
    }
  } else {
    for (uint64_t j = 0; j < steps; ++j) dvec[j] = b[j];
  }

  uint64_t cn[kMaxOrder], cpi[kMaxOrder], cw[kMaxOrder], nn[kMaxOrder], npi[kMaxOrder], nw[kMaxOrder];
  std::copy(na, na + p, cn);
  std::copy(pia, pia + p, cpi);
  uint64_t cp = p;
  const void* cur = a;                 // host pointer at step 0 of a host chain, device pointer otherwise
  void* cur_owned = nullptr;           // the pool block behind `cur`, if any
  for (uint64_t i = 0; i < steps; ++i) {
    const uint64_t m = modes[i];
    output_shape(cn, cp, m, nn);
    output_layout(cpi, cp, m, npi);
    uint64_t stride = 1, elems = 1;                                     // packed strides of both tensors
    for (uint64_t r = 0; r < cp; ++r) { cw[cpi[r] - 1] = stride; stride *= cn[cpi[r] - 1]; }
    for (uint64_t r = 0; r + 1 < cp; ++r) { nw[npi[r] - 1] = elems; elems *= nn[npi[r] - 1]; }
    const bool last = i + 1 == steps;
    void* dst = nullptr;
    void* dst_owned = nullptr;
    if (last && !host) dst = c;
    else { if (int rc = mem.get(&dst, (size_t)elems * s)) return rc; dst_owned = dst; }
    const uint64_t nb = cn[m - 1];
    View v;
    if (int rc = validate_and_fold(m, cp, cur, cn, cw, cpi, dvec[vecs[i]], &nb, dst, nn, nw, npi, 0u, &v)) return fail(rc);
    if (i == 0 && host) {
      // the only step that reads all of A: streamed across PCIe in chunks under its own kernels; the result stays in HBM
      if (int rc = run_view_host(dtype, v, cur, dvec[vecs[i]], dst, &local, /*bc_dev=*/true)) return rc;
    } else {
      if (int rc = run_view_device(dtype, v, cur, dvec[vecs[i]], dst, &local, device, false)) return rc;
    }
    if (cur_owned) mem.put(cur_owned);                                  // stream-ordered: free after the step that read it
    cur = dst; cur_owned = dst_owned;
    std::copy(nn, nn + cp - 1, cn);
    std::copy(npi, npi + cp - 1, cpi);
    --cp;
  }
  if (host) {
    CUDA_TRY(cudaMemcpyAsync(c, cur, (size_t)na[q - 1] * s, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync D2H c");
    CUDA_TRY(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  } else if (!async) {
    CUDA_TRY(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  }
  return TTV_B200_OK;
}

#define TTV_B200_TYPED(NAME, CODE, CT)                                                                              \
  int NAME(uint64_t q, uint64_t p, const CT* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,        \
           const CT* b, const uint64_t* nb, CT* c, const uint64_t* nc, const uint64_t* wc, const uint64_t* pic,     \
           const ttv_b200_opts* opts)                                                                               \
  { return ttv_b200_run(CODE, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic, opts); }

TTV_B200_TYPED(ttv_b200_f32,  TTV_B200_F32,  float)
TTV_B200_TYPED(ttv_b200_f64,  TTV_B200_F64,  double)
TTV_B200_TYPED(ttv_b200_c64,  TTV_B200_C64,  void)
TTV_B200_TYPED(ttv_b200_c128, TTV_B200_C128, void)
TTV_B200_TYPED(ttv_b200_i32,  TTV_B200_I32,  int32_t)
TTV_B200_TYPED(ttv_b200_i64,  TTV_B200_I64,  int64_t)
#undef TTV_B200_TYPED

int ttv_b200_plan(int dtype, uint64_t q, uint64_t p,
                  const void* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                  const void* b, const uint64_t* nb,
                  const void* c, const uint64_t* nc, const uint64_t* wc, const uint64_t* pic,
                  const ttv_b200_opts* opts, ttv_b200_plan_t* plan)
{
  if (dtype_size(dtype) == 0) return fail(TTV_B200_ERR_DTYPE);
  View v;
  if (int rc = validate_and_fold(q, p, a, na, wa, pia, b, nb, c, nc, wc, pic, opts ? opts->flags : 0u, &v)) return fail(rc);
  Launch l;
  if (!v.strided)
    if (int rc = choose_launch(dtype, v, opts, 256, 256, 256, 148, &l)) return fail(rc);
  if (plan) fill_plan(dtype, v, l, plan);
  return TTV_B200_OK;
}

int ttv_b200_plan_view(int dtype, uint64_t outer, uint64_t nq, uint64_t inner, const ttv_b200_opts* opts,
                       ttv_b200_plan_t* plan)
{
  if (dtype_size(dtype) == 0) return fail(TTV_B200_ERR_DTYPE);
  if (outer == 0 || nq == 0 || inner == 0) return fail(TTV_B200_ERR_SHAPE_A);
  View v;
  v.outer = outer; v.nq = nq; v.inner = inner;
  v.k = 0; v.ref_case = inner == 1 ? 6 : (outer == 1 ? 7 : 8);
  Launch l;
  if (int rc = choose_launch(dtype, v, opts, 256, 256, 256, 148, &l)) return fail(rc);
  if (plan) fill_plan(dtype, v, l, plan);
  return TTV_B200_OK;
}

int ttv_b200_view(int dtype, uint64_t outer, uint64_t nq, uint64_t inner, const void* a, const void* b, void* c,
                  const ttv_b200_opts* opts)
{
  if (dtype_size(dtype) == 0) return fail(TTV_B200_ERR_DTYPE);
  if (outer == 0 || nq == 0 || inner == 0) return fail(TTV_B200_ERR_SHAPE_A);
  if (!a) return fail(TTV_B200_ERR_A_NULL);
  if (!b) return fail(TTV_B200_ERR_B_NULL);
  if (!c) return fail(TTV_B200_ERR_C_NULL);
  View v;
  v.outer = outer; v.nq = nq; v.inner = inner;
  v.k = 0; v.ref_case = inner == 1 ? 6 : (outer == 1 ? 7 : 8);
  return run_any(dtype, v, a, b, c, opts);
}

int ttv_b200_view_scatter(int dtype, uint64_t outer, uint64_t nq, uint64_t inner, const void* a, const void* b,
                          void* const* peer_ws, uint32_t world, uint32_t rank, uint64_t blk, const ttv_b200_opts* opts)
{
  const uint64_t s = (uint64_t)dtype_size(dtype);
  if (s == 0) return fail(TTV_B200_ERR_DTYPE);
  if (outer == 0 || nq == 0 || inner == 0) return fail(TTV_B200_ERR_SHAPE_A);
  if (!a) return fail(TTV_B200_ERR_A_NULL);
  if (!b) return fail(TTV_B200_ERR_B_NULL);
  if (!peer_ws) return fail(TTV_B200_ERR_C_NULL);
  if (world == 0 || world > 16 || rank >= world || blk == 0 || blk * world < outer * inner)
    return fail(TTV_B200_ERR_OPTS, "ttv_b200_view_scatter: need 1 <= world <= 16, rank < world, world*blk >= outer*inner");
  for (uint32_t j = 0; j < world; ++j) if (!peer_ws[j]) return fail(TTV_B200_ERR_C_NULL);
  Where w; int dev = -1;
  if (int rc = classify(a, &w, &dev)) return rc;
  if (w != Where::Device) return fail(TTV_B200_ERR_MIXED_POINTERS, "ttv_b200_view_scatter needs device pointers");
  // widest vector that divides inner and blk and fits every alignment
  uint64_t align = std::min(alignment_of(a), (uint64_t)256);
  for (uint32_t j = 0; j < world; ++j) align = std::min(align, alignment_of(peer_ws[j]));
  uint64_t vec = s >= 16 ? 1 : 16 / s;
  while (vec > 1 && !((inner % vec) == 0 && (blk % vec) == 0 && (align % (vec * s)) == 0)) vec /= 2;
  DeviceGuard guard;
  CUDA_TRY(guard.set(dev), "cudaSetDevice");
  cudaStream_t stream = opts ? static_cast<cudaStream_t>(opts->stream) : nullptr;
  int sm = 148;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceState* st = nullptr;
    if (int rc = device_state(dev, &st)) return rc;
    sm = st->sm_count;
  }
  View v;
  v.outer = outer; v.nq = nq; v.inner = inner;
  CUDA_TRY(launch_scatter(dtype, v, a, b, peer_ws, world, rank, blk, (int)vec, sm, stream), "scatter launch");
  if (!(opts && (opts->flags & TTV_B200_FLAG_ASYNC))) CUDA_TRY(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  return TTV_B200_OK;
}

int ttv_b200_view_exchange(int dtype, uint64_t outer, uint64_t nq, uint64_t inner, const void* a, const void* b,
                           void* const* peer_ws, void* const* peer_flags, uint32_t world, uint32_t rank, uint64_t blk,
                           void* c_block, uint64_t n_block, uint32_t token, void* scratch, uint32_t max_ctas,
                           const ttv_b200_opts* opts)
{
  const uint64_t s = (uint64_t)dtype_size(dtype);
  if (s == 0) return fail(TTV_B200_ERR_DTYPE);
  if (outer == 0 || nq == 0 || inner == 0) return fail(TTV_B200_ERR_SHAPE_A);
  if (!a) return fail(TTV_B200_ERR_A_NULL);
  if (!b) return fail(TTV_B200_ERR_B_NULL);
  if (!peer_ws || !peer_flags || !scratch || (!c_block && n_block)) return fail(TTV_B200_ERR_C_NULL);
  if (world == 0 || world > 16 || rank >= world || blk == 0 || blk * world < outer * inner || n_block > blk || token == 0)
    return fail(TTV_B200_ERR_OPTS, "ttv_b200_view_exchange: need 1 <= world <= 16, rank < world, world*blk >= outer*inner, n_block <= blk, token > 0");
  for (uint32_t j = 0; j < world; ++j) if (!peer_ws[j] || !peer_flags[j]) return fail(TTV_B200_ERR_C_NULL);
  Where w; int dev = -1;
  if (int rc = classify(a, &w, &dev)) return rc;
  if (w != Where::Device) return fail(TTV_B200_ERR_MIXED_POINTERS, "ttv_b200_view_exchange needs device pointers");
  uint64_t align = std::min(alignment_of(a), (uint64_t)256);
  for (uint32_t j = 0; j < world; ++j) align = std::min(align, alignment_of(peer_ws[j]));
  if (c_block) align = std::min(align, alignment_of(c_block));
  uint64_t vec = s >= 16 ? 1 : 16 / s;
  while (vec > 1 && !((inner % vec) == 0 && (blk % vec) == 0 && (align % (vec * s)) == 0)) vec /= 2;
  DeviceGuard guard;
  CUDA_TRY(guard.set(dev), "cudaSetDevice");
  cudaStream_t stream = opts ? static_cast<cudaStream_t>(opts->stream) : nullptr;
  const bool accumulate = opts && (opts->flags & TTV_B200_FLAG_ACCUMULATE);
  DeviceState* st = nullptr;
  if (int rc = device_state_locked(dev, &st)) return rc;
  View v;
  v.outer = outer; v.nq = nq; v.inner = inner;
  const uint64_t timeout_ns = (uint64_t)std::max(1, env_mb("TTV_B200_EXCHANGE_TIMEOUT_MS", 10000)) * 1000000ull;
  char* sc = static_cast<char*>(scratch);
  CUDA_TRY(launch_exchange(dtype, v, a, b, peer_ws, peer_flags, world, rank, blk, (int)vec, c_block, n_block, token, sc, sc + 8,
                           accumulate, timeout_ns, max_ctas, st->sm_count, stream), "exchange launch");
  if (!(opts && (opts->flags & TTV_B200_FLAG_ASYNC))) {
    CUDA_TRY(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
    uint32_t err = 0;
    CUDA_TRY(cudaMemcpy(&err, sc + 8, sizeof err, cudaMemcpyDeviceToHost), "cudaMemcpy");
    if (err) return fail(TTV_B200_ERR_CUDA, "ttv_b200_view_exchange: timed out waiting for the other GPUs (did every rank launch round %u?)", token);
  }
  return TTV_B200_OK;
}

int ttv_b200_reduce_slots(int dtype, const void* ws, void* c, uint64_t n, uint64_t blk, uint32_t slots, const ttv_b200_opts* opts)
{
  if (dtype_size(dtype) == 0) return fail(TTV_B200_ERR_DTYPE);
  if (!ws) return fail(TTV_B200_ERR_A_NULL);
  if (!c) return fail(TTV_B200_ERR_C_NULL);
  if (slots == 0 || n > blk) return fail(TTV_B200_ERR_OPTS, "ttv_b200_reduce_slots: need slots >= 1 and n <= blk");
  Where w; int dev = -1;
  if (int rc = classify(c, &w, &dev)) return rc;
  if (w != Where::Device) return fail(TTV_B200_ERR_MIXED_POINTERS, "ttv_b200_reduce_slots needs device pointers");
  DeviceGuard guard;
  CUDA_TRY(guard.set(dev), "cudaSetDevice");
  cudaStream_t stream = opts ? static_cast<cudaStream_t>(opts->stream) : nullptr;
  const bool accumulate = opts && (opts->flags & TTV_B200_FLAG_ACCUMULATE);
  int sm = 148;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceState* st = nullptr;
    if (int rc = device_state(dev, &st)) return rc;
    sm = st->sm_count;
  }
  CUDA_TRY(launch_reduce_slots(dtype, ws, c, n, blk, slots, accumulate, sm, stream), "reduce launch");
  if (!(opts && (opts->flags & TTV_B200_FLAG_ASYNC))) CUDA_TRY(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  return TTV_B200_OK;
}

int ttv_b200_fill(int dtype, void* x, uint64_t first, uint64_t count, uint64_t seed, const ttv_b200_opts* opts)
{
  if (dtype_size(dtype) == 0) return fail(TTV_B200_ERR_DTYPE);
  if (!x) return fail(TTV_B200_ERR_A_NULL);
  Where w; int dev = -1;
  if (int rc = classify(x, &w, &dev)) return rc;
  if (w != Where::Device) return fail(TTV_B200_ERR_MIXED_POINTERS, "ttv_b200_fill needs a device pointer");
  DeviceGuard guard;
  CUDA_TRY(guard.set(dev), "cudaSetDevice");
  cudaStream_t stream = opts ? static_cast<cudaStream_t>(opts->stream) : nullptr;
  int sm = 148;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceState* st = nullptr;
    if (int rc = device_state(dev, &st)) return rc;
    sm = st->sm_count;
  }
  CUDA_TRY(launch_fill(dtype, x, first, count, seed, sm, stream), "fill launch");
  if (!(opts && (opts->flags & TTV_B200_FLAG_ASYNC))) CUDA_TRY(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  return TTV_B200_OK;
}

// ---- device / pinned memory helpers ---------------------------------------------------------------------------------
int ttv_b200_device_alloc(void** ptr, uint64_t bytes, int device, int zero)
{
  if (!ptr) return fail(TTV_B200_ERR_OPTS, "ttv_b200_device_alloc: ptr must not be null");
  *ptr = nullptr;
  if (device < 0) CUDA_TRY(cudaGetDevice(&device), "cudaGetDevice (is a CUDA device visible?)");
  DeviceGuard guard;
  CUDA_TRY(guard.set(device), "cudaSetDevice");
  CUDA_TRY(cudaMalloc(ptr, std::max<size_t>((size_t)bytes, 256)), "cudaMalloc");
  if (zero) {
    const cudaError_t e = cudaMemset(*ptr, 0, (size_t)bytes);
    if (e != cudaSuccess) { cudaFree(*ptr); *ptr = nullptr; return fail_cuda(e, "cudaMemset"); }
  }
  return TTV_B200_OK;
}

int ttv_b200_device_free(void* ptr)
{
  if (!ptr) return TTV_B200_OK;
  CUDA_TRY(cudaFree(ptr), "cudaFree");
  return TTV_B200_OK;
}

int ttv_b200_host_alloc(void** ptr, uint64_t bytes)
{
  if (!ptr) return fail(TTV_B200_ERR_OPTS, "ttv_b200_host_alloc: ptr must not be null");
  *ptr = nullptr;
  CUDA_TRY(cudaHostAlloc(ptr, std::max<size_t>((size_t)bytes, 64), cudaHostAllocPortable), "cudaHostAlloc");
  return TTV_B200_OK;
}

int ttv_b200_host_free(void* ptr)
{
  if (!ptr) return TTV_B200_OK;
  CUDA_TRY(cudaFreeHost(ptr), "cudaFreeHost");
  return TTV_B200_OK;
}

int ttv_b200_copy(void* dst, const void* src, uint64_t bytes, const ttv_b200_opts* opts)
{
  if (bytes == 0) return TTV_B200_OK;
  if (!dst || !src) return fail(TTV_B200_ERR_A_NULL);
  Where wd = Where::Host, ws = Where::Host;
  int dd = -1, ds = -1;
  if (int rc = classify(dst, &wd, &dd)) return rc;
  if (int rc = classify(src, &ws, &ds)) return rc;
  cudaStream_t stream = opts ? static_cast<cudaStream_t>(opts->stream) : nullptr;
  if (wd == Where::Host && ws == Where::Host) { copy_pool().copy(dst, src, (size_t)bytes); return TTV_B200_OK; }
  const int device = wd == Where::Device ? dd : ds;
  DeviceGuard guard;
  CUDA_TRY(guard.set(device), "cudaSetDevice");
  if (wd == Where::Device && ws == Where::Device) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, stream), "cudaMemcpyAsync D2D");
    CUDA_TRY(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
    return TTV_B200_OK;
  }
  DeviceState* st = nullptr;
  if (int rc = device_state_locked(device, &st)) return rc;
  std::lock_guard<std::mutex> host_lock(st->host_mutex);        // the bounce buffers are per device
  return wd == Where::Device ? copy_h2d(st, dst, src, (size_t)bytes, stream) : copy_d2h(st, dst, src, (size_t)bytes, stream);
}

// ---- resident tensors ---------------------------------------------------------------------------------------------
int ttv_b200_resident_create(ttv_b200_resident** r, int device)
{
  if (!r) return fail(TTV_B200_ERR_OPTS, "ttv_b200_resident_create: r must not be null");
  *r = nullptr;
  if (device < 0) CUDA_TRY(cudaGetDevice(&device), "cudaGetDevice (is a CUDA device visible?)");
  ttv_b200_resident* obj = new (std::nothrow) ttv_b200_resident;
  if (!obj) return fail(TTV_B200_ERR_CUDA, "out of host memory");
  obj->device = device;
  *r = obj;
  return TTV_B200_OK;
}

void ttv_b200_resident_destroy(ttv_b200_resident* r)
{
  if (!r) return;
  {
    std::lock_guard<std::mutex> lock(r->mutex);
    if (r->dev) {
      DeviceGuard guard;
      if (guard.set(r->device) == cudaSuccess) cudaFree(r->dev);      // cudaFree waits for work that still reads the buffer
      cudaGetLastError();
      r->dev = nullptr;
    }
  }
  delete r;
}

void ttv_b200_resident_invalidate(ttv_b200_resident* r)
{
  if (!r) return;
  std::lock_guard<std::mutex> lock(r->mutex);
  r->valid = false;
}

int ttv_b200_resident_valid(const ttv_b200_resident* r)
{
  if (!r) return 0;
  std::lock_guard<std::mutex> lock(const_cast<ttv_b200_resident*>(r)->mutex);
  return r->valid ? 1 : 0;
}

int ttv_b200_run_resident(ttv_b200_resident* r, int dtype, uint64_t q, uint64_t p,
                          const void* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                          const void* b, const uint64_t* nb,
                          void* c, const uint64_t* nc, const uint64_t* wc, const uint64_t* pic,
                          const ttv_b200_opts* opts)
{
  if (!r) return fail(TTV_B200_ERR_OPTS, "ttv_b200_run_resident: r must not be null");
  const size_t s = (size_t)dtype_size(dtype);
  if (s == 0) return fail(TTV_B200_ERR_DTYPE);
  View v;
  if (int rc = validate_and_fold(q, p, a, na, wa, pia, b, nb, c, nc, wc, pic, opts ? opts->flags : 0u, &v)) return fail(rc);
  Where w = Where::Host; int dx = -1;
  for (const void* ptr : {a, b, static_cast<const void*>(c)}) {
    if (int rc = classify(ptr, &w, &dx)) return rc;
    if (w != Where::Host) return fail(TTV_B200_ERR_MIXED_POINTERS, "ttv_b200_run_resident takes host pointers (device tensors need no twin)");
  }
  std::lock_guard<std::mutex> rlock(r->mutex);
  const int device = r->device;
  DeviceGuard guard;
  CUDA_TRY(guard.set(device), "cudaSetDevice");
  DeviceState* st = nullptr;
  if (int rc = device_state_locked(device, &st)) return rc;
  std::lock_guard<std::mutex> host_lock(st->host_mutex);

  const size_t bytes_a = (size_t)(v.strided ? v.span_a : v.outer * v.nq * v.inner) * s;
  const size_t bytes_b = (size_t)v.nq * s;
  const size_t bytes_c = (size_t)(v.strided ? v.span_c : v.outer * v.inner) * s;
  cudaStream_t stream = opts ? static_cast<cudaStream_t>(opts->stream) : nullptr;
  const bool c_goes_up = (opts && (opts->flags & TTV_B200_FLAG_ACCUMULATE)) || v.strided;
  ttv_b200_opts local = opts ? *opts : ttv_b200_opts{-1, 0, 0, 0, 0, 0, 0, 0, nullptr};
  local.device = device;
  local.flags &= ~(uint32_t)TTV_B200_FLAG_ASYNC;

  if (r->valid && (r->host != a || r->bytes != bytes_a)) r->valid = false;      // another tensor (or another extent of it)
  if (!r->valid) {
    if (r->capacity < bytes_a) {
      if (r->dev) { CUDA_TRY(cudaFree(r->dev), "cudaFree"); r->dev = nullptr; r->capacity = 0; }
      CUDA_TRY(cudaMalloc(&r->dev, std::max<size_t>(bytes_a, 256)), "cudaMalloc (resident copy of A)");
      r->capacity = bytes_a;
    }
    // first product: A crosses PCIe chunk by chunk under this product's own kernels, into the buffer that stays
    const void* bs[1] = {b};
    void* cs[1] = {c};
    const int rc = run_host_pipelined(dtype, 1, &v, a, bs, cs, &local, device, false, static_cast<char*>(r->dev));
    if (rc > 0) return rc;
    if (rc == 0) { r->host = a; r->bytes = bytes_a; r->valid = true; return TTV_B200_OK; }
    // not eligible for the chunked path (small, strided, one indivisible slab): one copy, then the device path below
    if (int rc2 = copy_h2d(st, r->dev, a, bytes_a, stream)) return rc2;
    r->host = a; r->bytes = bytes_a; r->valid = true;
  }
  // A is in HBM: only b goes up and C comes back
  void *db = nullptr, *dc = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (int rc = ensure(st->stage_b, bytes_b)) return rc;
    if (int rc = ensure(st->stage_c, bytes_c)) return rc;
    db = st->stage_b.ptr; dc = st->stage_c.ptr;
  }
  CUDA_TRY(cudaMemcpyAsync(db, b, bytes_b, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync H2D b");
  if (c_goes_up) if (int rc = copy_h2d(st, dc, c, bytes_c, stream)) return rc;
  if (int rc = run_view_device(dtype, v, r->dev, db, dc, &local, device, false)) return rc;
  return copy_d2h(st, c, dc, bytes_c, stream);
}

// ---- one host tensor over several GPUs ---------------------------------------------------------------------------------
int ttv_b200_run_devices(int dtype, uint64_t q, uint64_t p,
                         const void* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                         const void* b, const uint64_t* nb,
                         void* c, const uint64_t* nc, const uint64_t* wc, const uint64_t* pic,
                         const ttv_b200_opts* opts, const int32_t* devices, uint32_t n_devices)
{
  const size_t s = (size_t)dtype_size(dtype);
  if (s == 0) return fail(TTV_B200_ERR_DTYPE);
  if (!devices || n_devices == 0 || n_devices > 64) return fail(TTV_B200_ERR_OPTS, "ttv_b200_run_devices: need 1..64 devices");
  const int visible = ttv_b200_device_count();
  for (uint32_t d = 0; d < n_devices; ++d)
    if (devices[d] < 0 || devices[d] >= visible) return fail(TTV_B200_ERR_OPTS, "ttv_b200_run_devices: device %d is not visible (%d devices)", (int)devices[d], visible);
  View v;
  if (int rc = validate_and_fold(q, p, a, na, wa, pia, b, nb, c, nc, wc, pic, opts ? opts->flags : 0u, &v)) return fail(rc);
  Where w = Where::Host; int dx = -1;
  for (const void* ptr : {a, b, static_cast<const void*>(c)}) {
    if (int rc = classify(ptr, &w, &dx)) return rc;
    if (w != Where::Host) return fail(TTV_B200_ERR_MIXED_POINTERS, "ttv_b200_run_devices takes host pointers");
  }
  ttv_b200_opts base = opts ? *opts : ttv_b200_opts{-1, 0, 0, 0, 0, 0, 0, 0, nullptr};
  base.flags &= ~(uint32_t)TTV_B200_FLAG_ASYNC;
  base.stream = nullptr;                                          // a stream belongs to one device
  const bool accumulate = (base.flags & TTV_B200_FLAG_ACCUMULATE) != 0;

  const uint64_t total = v.outer * v.nq * v.inner;
  const uint64_t slow = v.slow_extent ? v.slow_extent : (v.outer > 1 ? v.outer : v.nq);
  const bool nq_split = v.outer == 1 || (v.outer % slow) != 0;
  const bool divisible = !v.strided && slow >= 2 && (!nq_split || (v.outer == 1 && v.nq == slow));
  const uint64_t min_bytes = (uint64_t)env_mb("TTV_B200_MULTI_MIN_MB", 64) << 20;
  // Pageable memory reaches the GPUs through the copy threads' memcpy into pinned buffers, which ONE link already matches
  // (measured on 2 GPUs, 4 GiB: free split 48 -> 58 GB/s, n_q split 49 -> 30 GB/s): a pageable n_q split stays on one device.
  const bool pageable_nq = nq_split && !is_pinned_host(a) && env_mb("TTV_B200_MULTI_PAGEABLE_NQ", 0) == 0;
  if (n_devices == 1 || !divisible || total * s < min_bytes || pageable_nq) {
    base.device = devices[0];
    return run_view_host(dtype, v, a, b, c, &base);
  }
  const uint32_t G = (uint32_t)std::min<uint64_t>(n_devices, slow);
  const uint64_t slab = total / slow;                             // elements of A per index of the slowest mode
  std::vector<uint64_t> begin(G + 1);
  for (uint32_t d = 0; d <= G; ++d) begin[d] = slow / G * d + std::min<uint64_t>(d, slow % G);

  std::vector<int> status(G, TTV_B200_OK);
  std::vector<std::string> message(G);
  std::vector<void*> partial(G, nullptr);                          // n_q split: device d's partial C, in ITS memory
  const uint64_t n_out = v.outer * v.inner;

  auto worker = [&](uint32_t d) {
    ttv_b200_opts o = base;
    o.device = devices[d];
    View sub = v;
    const char* ap = static_cast<const char*>(a) + (size_t)(begin[d] * slab) * s;
    const uint64_t count = begin[d + 1] - begin[d];
    int rc = TTV_B200_OK;
    if (!nq_split) {
      const uint64_t rows = v.outer / slow;                        // rows of `outer` per index of the slowest mode
      sub.outer = count * rows;
      sub.slow_extent = count;
      char* cp = static_cast<char*>(c) + (size_t)(begin[d] * rows * v.inner) * s;
      rc = run_view_host(dtype, sub, ap, b, cp, &o);
    } else {
      // rows [begin, end) of the contraction with the matching slice of b; the partial stays on the device
      sub.nq = count;
      sub.slow_extent = count;
      o.flags &= ~(uint32_t)TTV_B200_FLAG_ACCUMULATE;
      DeviceGuard guard;
      cudaError_t e = guard.set(devices[d]);
      void *pb = nullptr, *pc = nullptr;
      if (e == cudaSuccess) e = cudaMalloc(&pb, std::max<size_t>((size_t)count * s, 256));
      if (e == cudaSuccess) e = cudaMalloc(&pc, std::max<size_t>((size_t)n_out * s, 256));
      if (e == cudaSuccess) e = cudaMemcpy(pb, static_cast<const char*>(b) + (size_t)begin[d] * s, (size_t)count * s, cudaMemcpyHostToDevice);
      if (e != cudaSuccess) { rc = fail_cuda(e, "ttv_b200_run_devices: staging of one GPU's share"); if (pc) cudaFree(pc); pc = nullptr; }
      else rc = run_view_host(dtype, sub, ap, pb, pc, &o, /*bc_dev=*/true);
      if (pb) cudaFree(pb);
      if (rc != TTV_B200_OK && pc) { cudaFree(pc); pc = nullptr; }
      partial[d] = pc;
    }
    status[d] = rc;
    if (rc != TTV_B200_OK) message[d] = g_last_error;              // thread-local: carry it to the caller's thread
  };
  {
    std::vector<std::thread> threads;
    for (uint32_t d = 1; d < G; ++d) threads.emplace_back(worker, d);
    worker(0);
    for (auto& t : threads) t.join();
  }
  int rc = TTV_B200_OK;
  for (uint32_t d = 0; d < G; ++d)
    if (status[d] != TTV_B200_OK && rc == TTV_B200_OK) { rc = status[d]; g_last_error = message[d]; }

  if (nq_split && rc == TTV_B200_OK) {
    // the partials meet on devices[0]: slots [G][n_out], summed in device order (deterministic), then C goes back
    DeviceGuard guard;
    cudaError_t e = guard.set(devices[0]);
    void *slots = nullptr, *dc = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&slots, std::max<size_t>((size_t)G * n_out * s, 256));
    if (e == cudaSuccess) e = cudaMalloc(&dc, std::max<size_t>((size_t)n_out * s, 256));
    for (uint32_t d = 0; d < G && e == cudaSuccess; ++d)
      e = cudaMemcpyPeer(static_cast<char*>(slots) + (size_t)d * n_out * s, devices[0], partial[d], devices[d], (size_t)n_out * s);
    if (e == cudaSuccess && accumulate) e = cudaMemcpy(dc, c, (size_t)n_out * s, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
      DeviceState* st = nullptr;
      rc = device_state_locked(devices[0], &st);
      if (rc == TTV_B200_OK) {
        e = launch_reduce_slots(dtype, slots, dc, n_out, n_out, G, accumulate, st->sm_count, nullptr);
        if (e == cudaSuccess) {
          std::lock_guard<std::mutex> host_lock(st->host_mutex);
          rc = copy_d2h(st, c, dc, (size_t)n_out * s, nullptr);
        }
      }
    }
    if (e != cudaSuccess) rc = fail_cuda(e, "ttv_b200_run_devices: gathering the partial sums");
    if (slots) cudaFree(slots);
    if (dc) cudaFree(dc);
  }
  for (uint32_t d = 0; d < G; ++d)
    if (partial[d]) { DeviceGuard guard; if (guard.set(devices[d]) == cudaSuccess) cudaFree(partial[d]); cudaGetLastError(); }
  return rc;
}

// ---- L0 helpers ------------------------------------------------------------------------------------------------
int ttv_b200_is_valid_shape(const uint64_t* n, uint64_t p)   { return n && is_valid_shape(n, p) ? 1 : 0; }
int ttv_b200_is_valid_layout(const uint64_t* pi, uint64_t p) { return pi && is_valid_layout(pi, p) ? 1 : 0; }
int ttv_b200_is_valid_strides(const uint64_t* pi, uint64_t p, const uint64_t* w)
{
  if (!pi || !w || !is_valid_layout(pi, p)) return -1;   // the reference throws here (strides.h:79-80)
  return is_valid_strides(pi, p, w) ? 1 : 0;
}
int ttv_b200_compute_strides(const uint64_t* n, const uint64_t* pi, uint64_t p, uint64_t* w) { return compute_strides(n, pi, p, w); }
int ttv_b200_output_shape(const uint64_t* na, uint64_t p, uint64_t q, uint64_t* nc)          { return output_shape(na, p, q, nc); }
int ttv_b200_output_layout(const uint64_t* pia, uint64_t p, uint64_t q, uint64_t* pic)       { return output_layout(pia, p, q, pic); }
int ttv_b200_k_order_layout(uint64_t p, uint64_t k, uint64_t* pi)                            { return k_order_layout(p, k, pi); }

// ---- diagnostics ---------------------------------------------------------------------------------------------
const char* ttv_b200_strerror(int status) { return status_message(status); }
const char* ttv_b200_last_error(void)     { return g_last_error.c_str(); }
int         ttv_b200_version(void)        { return TTV_B200_VERSION; }
int         ttv_b200_dtype_size(int dtype) { return dtype_size(dtype); }
uint64_t    ttv_b200_launch_count(void)   { return launch_count(); }

int ttv_b200_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void ttv_b200_release(void)
{
  // collect the states first: host_mutex is always taken before g_mutex
  std::vector<std::pair<int, DeviceState*>> states;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    for (auto& kv : g_devices) states.emplace_back(kv.first, kv.second.get());
  }
  for (auto& ds : states) {
    DeviceState& st = *ds.second;
    std::lock_guard<std::mutex> host_lock(st.host_mutex);       // no host-pointer call is using the buffers
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceGuard guard;
    if (guard.set(ds.first) != cudaSuccess) { cudaGetLastError(); continue; }
    cudaDeviceSynchronize();                                    // queued kernels may still read a workspace
    for (Buffer* b : {&st.stage_a, &st.stage_b, &st.stage_c, &st.ring[0], &st.ring[1], &st.ring[2]})
      if (b->ptr) { cudaFree(b->ptr); b->ptr = nullptr; b->bytes = 0; }
    for (int r = 0; r < 3; ++r) {
      if (st.bounce[r]) { cudaFreeHost(st.bounce[r]); st.bounce[r] = nullptr; }
      if (st.bounce_out[r]) { cudaFreeHost(st.bounce_out[r]); st.bounce_out[r] = nullptr; }
      if (st.out_done[r]) { cudaEventDestroy(st.out_done[r]); st.out_done[r] = nullptr; }
      if (st.ready[r]) { cudaEventDestroy(st.ready[r]); st.ready[r] = nullptr; }
      if (st.freed[r]) { cudaEventDestroy(st.freed[r]); st.freed[r] = nullptr; }
    }
    st.bounce_bytes = 0;
    st.bounce_out_bytes = 0;
    if (st.copy_stream) { cudaStreamDestroy(st.copy_stream); st.copy_stream = nullptr; }
    if (st.pool) cudaMemPoolTrimTo(st.pool, 0);
    cudaGetLastError();
  }
}

} // extern "C"
