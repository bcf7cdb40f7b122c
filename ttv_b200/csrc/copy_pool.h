// copy_pool.h -- host threads that copy a chunk of memory in parallel (pure C++, no CUDA: tests/copy_pool_harness.cpp
// runs it under ThreadSanitizer).  Used by the host-pointer path of api.cu to move chunks of pageable tensors into pinned
// bounce buffers and chunks of results back out.
#pragma once

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace ttvb {

void host_copy(void* dst, const void* src, size_t bytes);   // hostcopy.cpp: memcpy with streaming stores

// Host threads that copy a chunk of pageable memory into a pinned bounce buffer in parallel.  cudaMemcpy from pageable
// memory runs at ~11 GB/s on these hosts (the driver stages it through one thread); a handful of threads saturate PCIe.
class CopyPool {
 public:
  explicit CopyPool(int n) { for (int i = 0; i < n; ++i) workers_.emplace_back([this] { run(); }); }
  ~CopyPool()
  {
    { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void copy(void* dst, const void* src, size_t bytes)
  {
    if (bytes <= ((size_t)256 << 10)) { std::memcpy(dst, src, bytes); return; }   // not worth waking anybody
    auto job = std::make_shared<Job>();
    job->dst = static_cast<char*>(dst); job->src = static_cast<const char*>(src); job->bytes = bytes;
    // about four parts per copier, between 64 KiB and 2 MiB each: a chunk of a few MiB still gets every thread
    const size_t copiers = workers_.size() + 1;
    job->part = std::min(kPart, std::max<size_t>((size_t)64 << 10, (bytes / (4 * copiers) + 4095) / 4096 * 4096));
    job->parts = (bytes + job->part - 1) / job->part;
    job->remaining.store(job->parts);
    if (job->parts == 0) return;
    { std::lock_guard<std::mutex> lk(m_); cur_ = job; ++generation_; }
    cv_.notify_all();
    work(*job);                                              // the calling thread helps
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [&] { return job->remaining.load() == 0; });
  }

 private:
  static constexpr size_t kPart = 2u << 20;
  struct Job {
    char* dst = nullptr; const char* src = nullptr; size_t bytes = 0, parts = 0, part = kPart;
    std::atomic<size_t> next{0}, remaining{0};
  };
  void work(Job& j)
  {
    for (;;) {
      const size_t i = j.next.fetch_add(1);
      if (i >= j.parts) return;
      const size_t off = i * j.part;
      host_copy(j.dst + off, j.src + off, std::min(j.part, j.bytes - off));
      if (j.remaining.fetch_sub(1) == 1) { std::lock_guard<std::mutex> lk(m_); done_.notify_all(); }
    }
  }
  void run()
  {
    uint64_t seen = 0;
    for (;;) {
      std::shared_ptr<Job> job;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
        if (stop_) return;
        seen = generation_;
        job = cur_;
      }
      work(*job);
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  std::shared_ptr<Job> cur_;
  uint64_t generation_ = 0;
  bool stop_ = false;
};

} // namespace ttvb
