// staged_copy.cuh -- host <-> device transfers of the flattened Java containers.
//
// The arrays a JNI caller hands over (GetPrimitiveArrayCritical / GetDoubleArrayRegion) are
// PAGEABLE: a plain cudaMemcpy stages them through the driver's single bounce buffer on the calling thread
// (~11 GB/s measured on the B200 box, profiles/r1).  StagedCopier splits the transfer into chunks that a few
// host threads copy through their own pinned double buffers and DMA on their own streams, so the host-side
// memcpy runs on several cores and overlaps the PCIe transfers.  Memory that is already pinned
// (cudaMallocHost / cudaHostRegister, e.g. a direct ByteBuffer registered by the caller) is DMA'd directly.
// All calls block until the data has arrived; the caller orders them against its own stream.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace cars {

struct CopySeg {
  void* dev;
  void* host;
  size_t bytes;
};

// Pinned staging chunks are kept for the life of the process: cudaMallocHost / cudaFreeHost cost 5-20 ms a piece and
// synchronise the device (measured: cars_destroy 8-245 ms for the 12 chunks of one handle).  A JVM training K folds
// creates K handles; the chunks of a destroyed handle serve the next one.  At most kMaxCached chunks are retained.
class PinnedChunkCache {
 public:
  static constexpr size_t kMaxCached = 32;
  static cudaError_t get(size_t bytes, char** out) {
    {
      std::lock_guard<std::mutex> g(mu());
      if (!free_list().empty()) {
        *out = free_list().back();
        free_list().pop_back();
        return cudaSuccess;
      }
    }
    return cudaMallocHost((void**)out, bytes);
  }
  static void put(char* p) {
    {
      std::lock_guard<std::mutex> g(mu());
      if (free_list().size() < kMaxCached) {
        free_list().push_back(p);
        return;
      }
    }
    cudaFreeHost(p);
  }

 private:
  static std::mutex& mu() { static std::mutex m; return m; }
  static std::vector<char*>& free_list() { static std::vector<char*> v; return v; }
};

class StagedCopier {
 public:
  static constexpr int kMaxWorkers = 8;
  static constexpr size_t kChunk = (size_t)4 << 20;     // bytes per staging buffer
  static constexpr size_t kDirectBelow = (size_t)1 << 20;  // small pageable copies: one cudaMemcpy

  ~StagedCopier() { destroy(); }

  // threads = 0: chosen from the host's core count
  cudaError_t init(int device, int threads = 0) {
    device_ = device;
    unsigned hc = std::thread::hardware_concurrency();
    workers_ = hc >= 16 ? 6 : hc >= 8 ? 4 : hc >= 4 ? 2 : 1;
    if (threads >= 1 && threads <= kMaxWorkers) workers_ = threads;
    return cudaSuccess;
  }

  void destroy() {
    for (int w = 0; w < kMaxWorkers; w++) {
      for (int b = 0; b < 2; b++) {
        if (buf_[w][b]) PinnedChunkCache::put(buf_[w][b]);
        if (ev_[w][b]) cudaEventDestroy(ev_[w][b]);
        buf_[w][b] = nullptr;
        ev_[w][b] = nullptr;
      }
      if (stream_[w]) cudaStreamDestroy(stream_[w]);
      stream_[w] = nullptr;
    }
    ready_ = false;
  }

  // Copies every segment (to_device: host -> dev, else dev -> host); returns when all bytes have landed.
  cudaError_t run(const CopySeg* segs, int nseg, bool to_device) {
    cudaError_t e = cudaSetDevice(device_);
    if (e != cudaSuccess) return e;
    std::vector<Piece> pieces;
    bool any_staged = false;
    cudaStream_t s0 = nullptr;
    for (int i = 0; i < nseg; i++) {
      const CopySeg& sg = segs[i];
      if (sg.bytes == 0) continue;
      if (is_pinned(sg.host) || sg.bytes < kDirectBelow) {
        if (!s0) {
          e = ensure_stream(0);
          if (e != cudaSuccess) return e;
          s0 = stream_[0];
        }
        e = to_device ? cudaMemcpyAsync(sg.dev, sg.host, sg.bytes, cudaMemcpyHostToDevice, s0)
                      : cudaMemcpyAsync(sg.host, sg.dev, sg.bytes, cudaMemcpyDeviceToHost, s0);
        if (e != cudaSuccess) return e;
        continue;
      }
      any_staged = true;
      for (size_t off = 0; off < sg.bytes; off += kChunk) {
        const size_t len = sg.bytes - off < kChunk ? sg.bytes - off : kChunk;
        pieces.push_back(Piece{(char*)sg.dev + off, (char*)sg.host + off, len});
      }
    }
    if (any_staged) {
      e = ensure_buffers();
      if (e != cudaSuccess) return e;
      std::atomic<size_t> next{0};
      cudaError_t errs[kMaxWorkers];
      std::vector<std::thread> th;
      const int nw = (int)(pieces.size() < (size_t)workers_ ? pieces.size() : (size_t)workers_);
      for (int w = 0; w < nw; w++) {
        errs[w] = cudaSuccess;
        th.emplace_back([&, w]() { errs[w] = worker(w, pieces, next, to_device); });
      }
      for (auto& t : th) t.join();
      for (int w = 0; w < nw; w++)
        if (errs[w] != cudaSuccess) return errs[w];
    }
    if (s0) {
      e = cudaStreamSynchronize(s0);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  }

  cudaError_t h2d(void* dev, const void* host, size_t bytes) {
    CopySeg s{dev, const_cast<void*>(host), bytes};
    return run(&s, 1, true);
  }
  cudaError_t d2h(void* host, const void* dev, size_t bytes) {
    CopySeg s{const_cast<void*>(dev), host, bytes};
    return run(&s, 1, false);
  }

  static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return a.type == cudaMemoryTypeHost;
  }

 private:
  struct Piece {
    char* dev;
    char* host;
    size_t bytes;
  };

  cudaError_t ensure_stream(int w) {
    if (stream_[w]) return cudaSuccess;
    return cudaStreamCreateWithFlags(&stream_[w], cudaStreamNonBlocking);
  }

  cudaError_t ensure_buffers() {
    if (ready_) return cudaSuccess;
    for (int w = 0; w < workers_; w++) {
      cudaError_t e = ensure_stream(w);
      if (e != cudaSuccess) return e;
      for (int b = 0; b < 2; b++) {
        if (!buf_[w][b] && (e = PinnedChunkCache::get(kChunk, &buf_[w][b])) != cudaSuccess) return e;
        if (!ev_[w][b] && (e = cudaEventCreateWithFlags(&ev_[w][b], cudaEventDisableTiming)) != cudaSuccess) return e;
      }
    }
    ready_ = true;
    return cudaSuccess;
  }

  cudaError_t worker(int w, const std::vector<Piece>& pieces, std::atomic<size_t>& next, bool to_device) {
    cudaError_t e = cudaSetDevice(device_);
    if (e != cudaSuccess) return e;
    cudaStream_t st = stream_[w];
    bool used[2] = {false, false};
    const Piece* pending = nullptr;  // D2H: piece whose DMA into buf[pb] is in flight
    int pb = 0;
    int k = 0;
    for (;;) {
      const size_t c = next.fetch_add(1);
      if (c >= pieces.size()) break;
      const Piece& p = pieces[c];
      const int b = k & 1;
      if (to_device) {
        if (used[b] && (e = cudaEventSynchronize(ev_[w][b])) != cudaSuccess) return e;
        memcpy(buf_[w][b], p.host, p.bytes);
        if ((e = cudaMemcpyAsync(p.dev, buf_[w][b], p.bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(ev_[w][b], st)) != cudaSuccess) return e;
        used[b] = true;
      } else {
        // buffer b was drained when its piece (two turns ago) was finished below
        if ((e = cudaMemcpyAsync(buf_[w][b], p.dev, p.bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(ev_[w][b], st)) != cudaSuccess) return e;
        if (pending) {
          if ((e = cudaEventSynchronize(ev_[w][pb])) != cudaSuccess) return e;
          memcpy(pending->host, buf_[w][pb], pending->bytes);
        }
        pending = &p;
        pb = b;
      }
      k++;
    }
    if (!to_device && pending) {
      if ((e = cudaEventSynchronize(ev_[w][pb])) != cudaSuccess) return e;
      memcpy(pending->host, buf_[w][pb], pending->bytes);
    }
    return cudaStreamSynchronize(st);
  }

  int device_ = 0;
  int workers_ = 4;
  bool ready_ = false;
  char* buf_[kMaxWorkers][2] = {};
  cudaEvent_t ev_[kMaxWorkers][2] = {};
  cudaStream_t stream_[kMaxWorkers] = {};
};

}  // namespace cars
