// msm_lib.cu -- translation unit holding the G1 MSM kernels.
#include "internal.h"
#include "msm_host.cuh"
#ifndef ALEO_EMU
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#endif

namespace aleo {
cudaError_t msm_upload_constants() { return aleo_upload_field_constants(); }
bool msm_size_supported(size_t n) { return msm::size_supported(n); }
int msm_window_bits(size_t n) { return (int)msm::make_params(n).c; }
int msm_ba_levels(size_t n) {
  if (n == 0) return 0;
  msm::Session ss;
  ss.begin(n, n, 1, nullptr, nullptr, true);  // dry: plans only
  return (int)ss.ba_L;
}

// ---- host-pointer calls: the MSM arrives in point ranges so that the copy of range k + 1 overlaps the
// accumulation of range k (PCIe moves 136 bytes per point about 2.5x faster than the integer pipe adds it) ----
namespace {
thread_local ThreadStream t_copy_stream;  // released when the calling thread ends (internal.h)

cudaError_t copy_stream(cudaStream_t* out) { return t_copy_stream.get(out); }

// Point ranges of a host-pointer MSM.  The first range is small (its copy is the only one not hidden) and the
// later ones grow: a range's accumulation must outlast the copy of the next one.
int chunk_schedule(size_t n, size_t* sizes, bool pinned = false) {
  const char* env = getenv("ALEO_B200_MSM_CHUNKS");  // read per call: tests and sweeps switch it
  const long k_env = env ? atol(env) : 0L;
  // from 2^22 points: 4 ranges, 1 : 2 : 3 : 5 for pageable caller memory (the default: a Rust Vec), 1 : 2 : 4 : 8 when the
  // caller's buffers are pinned (copies at the full 55 GB/s).  Re-swept after the batch-affine levels made the
  // accumulation faster (2^24, profiles/r03k_*: pageable 3:4:5:7 98.4 ms, 2:3:5:9 97.0, 1:2:3:5 95.7, 1:2:4:8 97.2;
  // pinned 1:2:4 94.1 ms, 1:2:4:8 93.1, 1:3:6 93.8): the first range's copy is the only one that is exposed.
  int k = n < ((size_t)1 << 19) ? 1 : (n < ((size_t)1 << 22) ? 2 : 4);
  if (k_env >= 1 && k_env <= 4 && n >= 16) k = (int)k_env;
  // explicit weights, e.g. ALEO_B200_MSM_SPLIT=1,2,4 (sweeps): range i gets w_i / sum(w) of the points
  if (const char* sp = getenv("ALEO_B200_MSM_SPLIT")) {
    unsigned w[4] = {0, 0, 0, 0};
    const int got = sscanf(sp, "%u,%u,%u,%u", &w[0], &w[1], &w[2], &w[3]);
    size_t tot = 0;
    for (int i = 0; i < got; i++) tot += w[i];
    if (got >= 1 && tot > 0 && n >= 16 * tot) {
      size_t used = 0;
      for (int i = 0; i < got - 1; i++) {
        sizes[i] = n / tot * w[i];
        if (sizes[i] == 0) sizes[i] = 1;
        used += sizes[i];
      }
      sizes[got - 1] = n - used;
      return got;
    }
  }
  if (k == 1) {
    sizes[0] = n;
  } else if (k == 4) {
    // A range may be about 1.5x its predecessor before the GPU waits for its copy when the caller's memory is PAGEABLE
    // (staged copies run at 33-46 GB/s on the B200 boxes, tools/feed_probe.cu: 243-340 Mpts/s at 136 B per point against
    // 210 Mpts/s of accumulation), about twice when it is pinned.
    const size_t w[4] = {1, 2, pinned ? (size_t)4 : 3, pinned ? (size_t)8 : 5};
    const size_t tot = w[0] + w[1] + w[2] + w[3];
    sizes[0] = (n * w[0]) / tot;
    sizes[1] = (n * w[1]) / tot;
    sizes[2] = (n * w[2]) / tot;
    sizes[3] = n - sizes[0] - sizes[1] - sizes[2];
  } else if (k == 2) {
    sizes[0] = n / 4;
    sizes[1] = n - sizes[0];
  } else {
    // 1 : 2 : 4 -- a range may be about twice its predecessor (2.57 ns per point on PCIe against 5.3 ns on the
    // integer pipe) before the GPU waits for its copy; measured 2^24: 97.4 ms against 100.9 ms for 1 : 3 : 4
    sizes[0] = n / 7;
    sizes[1] = (n * 2) / 7;
    sizes[2] = n - sizes[0] - sizes[1];
  }
  return k;
}

struct EventSet {
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // ev[4]: allocation ordered; 0..3, 5: ranges
  cudaError_t init() {
    for (auto& e : ev) {
      cudaError_t r = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      if (r != cudaSuccess) return r;
    }
    return cudaSuccess;
  }
  ~EventSet() {
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
  }
};
// ---- host <-> device copies of caller memory that may be PAGEABLE (a Rust Vec is) --------------------------------------
// Pinned / registered memory goes out as one cudaMemcpyAsync.  A large pageable block would be staged by the driver
// through one bounce buffer at ~10 GB/s (tools/feed_probe.cu on the B200 box: pinned 55.6 GB/s, driver bounce 10.4 GB/s,
// cudaHostRegister + copy + unregister 7.3 GB/s) -- slower than the MSM itself -- so it is staged here instead: a
// Feeder owns feed_threads() PERSISTENT helper threads, each with two pinned 4 MB buffers, a stream and two events;
// helper t memcpys slices t, t + T, ... into its buffers and enqueues the DMA from there (the CPU copy of the next
// slice overlaps the DMA of the previous one; measured 42-46 GB/s with 6-12 helpers).  One Feeder per calling thread
// (thread_local, released by its destructor when the thread ends); it is rebuilt when the thread's device changes.
#ifndef ALEO_EMU
constexpr int FEED_THREADS = 16;  // upper bound; feed_threads() of them run
// below this the driver's own pageable copy is used (ALEO_B200_FEED_MIN_KB overrides)
static size_t feed_min_bytes() {
  static const size_t v = [] {
    const char* env = getenv("ALEO_B200_FEED_MIN_KB");
    const long kb = env ? atol(env) : 0L;
    return (kb >= 64 && kb <= (1 << 20)) ? (size_t)kb << 10 : (size_t)8 << 20;
  }();
  return v;
}
#define FEED_MIN_BYTES feed_min_bytes()

// co-located ranks (one process per GPU under torchrun)
long local_world_size() {
  const char* lw = getenv("LOCAL_WORLD_SIZE");
  const long v = lw ? atol(lw) : 1L;
  return v > 0 ? v : 1L;
}

// Bytes per staging slice (two per helper).  One process on the host: 4 MB (fewest synchronisations: 42.5 GB/s against
// 35 GB/s with 1 MB, tools/feed_probe.cu).  Several ranks share the host's memory bandwidth, and every staged byte is
// read, written and read again by the DMA: slices small enough to stay in the last-level cache between the CPU's copy
// and the DMA's read take the second and third pass off the DRAM (ALEO_B200_FEED_SLICE_KB overrides).
size_t feed_slice_bytes() {
  static const size_t v = [] {
    const char* env = getenv("ALEO_B200_FEED_SLICE_KB");
    long kb = env ? atol(env) : 0L;
    if (kb < 64 || kb > 65536) kb = local_world_size() > 1 ? 1024 : 4096;
    return (size_t)kb << 10;
  }();
  return v;
}

// helper threads per staged copy: one core copies ~8 GB/s, PCIe takes ~55 GB/s
int feed_threads() {
  static const int n = [] {
    const char* env = getenv("ALEO_B200_FEED_THREADS");
    long v = env ? atol(env) : 0L;
    if (v < 1) {
      // one process per GPU (torchrun sets LOCAL_WORLD_SIZE): the ranks of a box share its cores -- 8 ranks x 6 helpers
      // on 32 cores ran the staged copies slower than 3 helpers each
      unsigned hc = std::thread::hardware_concurrency();
      const long ranks = local_world_size();
      if (ranks > 1 && hc > 0) hc = hc / (unsigned)ranks > 0 ? hc / (unsigned)ranks : 1u;
      v = hc >= 12 ? 6 : (hc >= 4 ? 4 : (hc >= 3 ? 3 : 2));
    }
    return (int)(v > FEED_THREADS ? FEED_THREADS : v);
  }();
  return n;
}

class Feeder {
 public:
  explicit Feeder(int device) : dev_(device), nt_(feed_threads()), slice_(feed_slice_bytes()) {}
  ~Feeder() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& w : workers_)
      if (w.joinable()) w.join();
    // at process exit the runtime may already be gone: the calls then fail harmlessly
    for (int t = 0; t < nt_; t++) {
      for (int b = 0; b < 2; b++) {
        if (ev_[t][b]) cudaEventDestroy(ev_[t][b]);
        if (buf_[t][b]) cudaFreeHost(buf_[t][b]);
      }
      if (st_[t]) cudaStreamDestroy(st_[t]);
    }
  }
  int device() const { return dev_; }

  cudaError_t init() {
    for (int t = 0; t < nt_; t++) {
      cudaError_t e = cudaStreamCreateWithFlags(&st_[t], cudaStreamNonBlocking);
      for (int b = 0; b < 2 && e == cudaSuccess; b++) {
        e = cudaMallocHost((void**)&buf_[t][b], slice_);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_[t][b], cudaEventDisableTiming);
      }
      if (e != cudaSuccess) return e;
    }
    for (int t = 0; t < nt_; t++) workers_.emplace_back([this, t]() { loop(t); });
    return cudaSuccess;
  }

  // dev <- host: returns when every slice is enqueued; `cs` then waits for the helpers' streams
  cudaError_t h2d(unsigned char* dst_dev, const unsigned char* src_host, size_t bytes, cudaStream_t cs, cudaEvent_t ready) {
    cudaError_t e = run(Job{true, dst_dev, const_cast<unsigned char*>(src_host), bytes, ready});
    for (int t = 0; t < nt_ && e == cudaSuccess; t++)
      for (int b = 0; b < 2 && e == cudaSuccess; b++) e = cudaStreamWaitEvent(cs, ev_[t][b], 0);
    return e;
  }
  // host <- dev: `gate` orders the helpers' streams after the producer; returns when the data is in dst_host
  cudaError_t d2h(unsigned char* dst_host, const unsigned char* src_dev, size_t bytes, cudaEvent_t gate) {
    return run(Job{false, const_cast<unsigned char*>(src_dev), dst_host, bytes, gate});
  }

 private:
  struct Job {
    bool to_device;
    unsigned char* dev;
    unsigned char* host;
    size_t bytes;
    cudaEvent_t gate;
  };

  cudaError_t run(const Job& j) {
    std::unique_lock<std::mutex> lk(mu_);
    job_ = j;
    pending_ = nt_;
    err_ = cudaSuccess;
    generation_++;
    cv_.notify_all();
    done_cv_.wait(lk, [this] { return pending_ == 0; });
    return err_;
  }

  void loop(int t) {
    cudaSetDevice(dev_);
    unsigned long long seen = 0;
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
        if (stop_) return;
        seen = generation_;
        j = job_;
      }
      const cudaError_t e = j.to_device ? push(t, j) : pull(t, j);
      std::lock_guard<std::mutex> lk(mu_);
      if (e != cudaSuccess) err_ = e;
      if (--pending_ == 0) done_cv_.notify_all();
    }
  }

  cudaError_t push(int t, const Job& j) {
    const size_t nslices = (j.bytes + slice_ - 1) / slice_;
    cudaError_t e = cudaStreamWaitEvent(st_[t], j.gate, 0);
    for (size_t i = (size_t)t, round = 0; i < nslices && e == cudaSuccess; i += (size_t)nt_, round++) {
      const int b = (int)(round & 1);
      const size_t off = i * slice_, len = (j.bytes - off < slice_) ? j.bytes - off : slice_;
      e = cudaEventSynchronize(ev_[t][b]);  // the DMA that last read this buffer is done
      if (e != cudaSuccess) break;
      std::memcpy(buf_[t][b], j.host + off, len);
      e = cudaMemcpyAsync(j.dev + off, buf_[t][b], len, cudaMemcpyHostToDevice, st_[t]);
      if (e == cudaSuccess) e = cudaEventRecord(ev_[t][b], st_[t]);
    }
    return e;
  }

  cudaError_t pull(int t, const Job& j) {
    const size_t nslices = (j.bytes + slice_ - 1) / slice_;
    cudaError_t e = cudaStreamWaitEvent(st_[t], j.gate, 0);
    for (int b = 0; b < 2 && e == cudaSuccess; b++) e = cudaEventSynchronize(ev_[t][b]);  // buffers idle
    size_t prev_off = 0, prev_len = 0;
    int prev_b = -1;
    for (size_t i = (size_t)t, round = 0; e == cudaSuccess; i += (size_t)nt_, round++) {
      const int b = (int)(round & 1);
      const bool more = i < nslices;
      size_t off = 0, len = 0;
      if (more) {
        off = i * slice_;
        len = (j.bytes - off < slice_) ? j.bytes - off : slice_;
        e = cudaMemcpyAsync(buf_[t][b], j.dev + off, len, cudaMemcpyDeviceToHost, st_[t]);
        if (e == cudaSuccess) e = cudaEventRecord(ev_[t][b], st_[t]);
      }
      if (prev_b >= 0 && e == cudaSuccess) {  // copy the previous slice out while this one is in flight
        e = cudaEventSynchronize(ev_[t][prev_b]);
        if (e == cudaSuccess) std::memcpy(j.host + prev_off, buf_[t][prev_b], prev_len);
      }
      if (!more) break;
      prev_off = off;
      prev_len = len;
      prev_b = b;
    }
    return e;
  }

  const int dev_, nt_;
  const size_t slice_;
  unsigned char* buf_[FEED_THREADS][2] = {};
  cudaEvent_t ev_[FEED_THREADS][2] = {};
  cudaStream_t st_[FEED_THREADS] = {};
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  Job job_{};
  int pending_ = 0;
  unsigned long long generation_ = 0;
  cudaError_t err_ = cudaSuccess;
  bool stop_ = false;
};
thread_local std::unique_ptr<Feeder> t_feeder;

cudaError_t feeder_for_current_device(Feeder** out) {
  int dev = 0;
  MSM_CK(cudaGetDevice(&dev));
  if (!t_feeder || t_feeder->device() != dev) {
    t_feeder.reset();  // a thread that moved to another device drops the old staging context first
    std::unique_ptr<Feeder> f(new Feeder(dev));
    MSM_CK(f->init());
    t_feeder = std::move(f);
  }
  *out = t_feeder.get();
  return cudaSuccess;
}

bool host_pointer_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();  // older drivers report unregistered memory as an error: clear it
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
#endif

}  // namespace

// dst_dev <- src_host on `cs` (after `ready`, the event that orders dst's allocation); may block the caller while it
// stages pageable memory, never waits for the GPU beyond its own double buffers
cudaError_t feed_h2d(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t cs, cudaEvent_t ready) {
  if (bytes == 0) return cudaSuccess;
#ifndef ALEO_EMU
  static const bool force_plain = getenv("ALEO_B200_NO_STAGING") != nullptr;
  if (bytes >= FEED_MIN_BYTES && !force_plain && !host_pointer_is_pinned(src_host)) {
    Feeder* f = nullptr;
    MSM_CK(feeder_for_current_device(&f));
    return f->h2d((unsigned char*)dst_dev, (const unsigned char*)src_host, bytes, cs, ready);
  }
#endif
  (void)ready;
  return cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, cs);
}

// dst_host <- src_dev, everything enqueued on `s` before the call is complete when the data is read.  Pageable
// destinations are staged like feed_h2d (DMA into pinned double buffers, helper threads copy out).  Blocks until done.
cudaError_t feed_d2h_sync(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return cudaStreamSynchronize(s);
#ifndef ALEO_EMU
  static const bool force_plain = getenv("ALEO_B200_NO_STAGING") != nullptr;
  if (bytes >= FEED_MIN_BYTES && !force_plain && !host_pointer_is_pinned(dst_host)) {
    Feeder* f = nullptr;
    MSM_CK(feeder_for_current_device(&f));
    cudaEvent_t done;
    MSM_CK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(done, s);
    if (e == cudaSuccess) e = f->d2h((unsigned char*)dst_host, (const unsigned char*)src_dev, bytes, done);
    cudaEventDestroy(done);
    cudaError_t e2 = cudaStreamSynchronize(s);
    return e != cudaSuccess ? e : e2;
  }
#endif
  cudaError_t e = cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, s);
  cudaError_t e2 = cudaStreamSynchronize(s);
  return e != cudaSuccess ? e : e2;
}

int msm_host_chunks(size_t n) {
  size_t sizes[4];
  return chunk_schedule(n, sizes);
}

int msm_host_window_bits(size_t n) { return (int)msm::make_params(n, nullptr, (u32)msm_host_chunks(n)).c; }

// out_host: 144 bytes.  Synchronises `s` before returning.
cudaError_t msm_run_host(const void* bases_host, u32 stride, const void* scalars_host, size_t n, void* out_host, cudaStream_t s) {
  unsigned char* d = nullptr;
  const size_t bb = n * stride, sb = n * 32;
  const size_t o_s = (bb + 255) & ~(size_t)255, o_out = o_s + ((sb + 255) & ~(size_t)255);
  MSM_CK(aleo::pool_malloc_async((void**)&d, o_out + 256, s));
  cudaError_t e = cudaSuccess;
  if (n == 0) {
    e = msm::run(nullptr, stride, nullptr, 0, d + o_out, s, false, nullptr);
  } else {
    size_t sizes[4];
#ifndef ALEO_EMU
    const bool pinned = host_pointer_is_pinned(bases_host);
#else
    const bool pinned = false;
#endif
    const int k = chunk_schedule(n, sizes, pinned);
    cudaStream_t cs = nullptr;
    EventSet evs;
    e = copy_stream(&cs);
    if (e == cudaSuccess) e = evs.init();
    msm::Session ss;
    if (e == cudaSuccess) e = cudaEventRecord(evs.ev[4], s);          // the allocation is ordered on s
    if (e == cudaSuccess) e = cudaStreamWaitEvent(cs, evs.ev[4], 0);
    size_t max_chunk = 0;
    for (int i = 0; i < k; i++) max_chunk = sizes[i] > max_chunk ? sizes[i] : max_chunk;
    if (e == cudaSuccess) e = ss.begin(n, max_chunk, (u32)k, nullptr, s, false);
    // range by range: copy (asynchronous for pinned memory, staged for pageable memory -- then this thread is busy
    // feeding range i + 1 while the GPU accumulates range i), then enqueue the range's sort + accumulation
    size_t first = 0;
    for (int i = 0; i < k && e == cudaSuccess; i++) {
      const size_t m = sizes[i];
      e = feed_h2d(d + o_s + first * 32, (const unsigned char*)scalars_host + first * 32, m * 32, cs, evs.ev[4]);
      if (e == cudaSuccess)
        e = feed_h2d(d + first * stride, (const unsigned char*)bases_host + first * stride, m * stride, cs, evs.ev[4]);
      if (e == cudaSuccess) e = cudaEventRecord(evs.ev[i < 4 ? i : 5], cs);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(s, evs.ev[i < 4 ? i : 5], 0);
      if (e == cudaSuccess) e = ss.add_chunk(d + first * stride, stride, (const u32*)(d + o_s + first * 32), m, first, s);
      first += m;
    }
    if (e == cudaSuccess) e = ss.finish(d + o_out, s);
    else ss.release(s);
    if (e != cudaSuccess) cudaStreamSynchronize(cs);  // nothing may still be writing into d when it is freed
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d + o_out, 144, cudaMemcpyDeviceToHost, s);
  cudaFreeAsync(d, s);
  cudaError_t e2 = cudaStreamSynchronize(s);
  return e != cudaSuccess ? e : e2;
}

cudaError_t msm_run(const void* bases_dev, u32 stride, const void* scalars_dev, size_t n, void* out144_dev, cudaStream_t s,
                    bool dry, int* launches_out, float* phase_ms) {
  if (!phase_ms || dry || n == 0)
    return msm::run((const unsigned char*)bases_dev, stride, (const u32*)scalars_dev, n, (unsigned char*)out144_dev, s, dry,
                    launches_out);
  cudaEvent_t ev[4];
  for (int i = 0; i < 4; i++) cudaEventCreate(&ev[i]);
  cudaError_t e = msm::run((const unsigned char*)bases_dev, stride, (const u32*)scalars_dev, n, (unsigned char*)out144_dev,
                           s, false, launches_out, ev);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  for (int i = 0; i < 3; i++) {
    phase_ms[i] = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&phase_ms[i], ev[i], ev[i + 1]);
  }
  for (int i = 0; i < 4; i++) cudaEventDestroy(ev[i]);
  return e;
}

// ---- resident SRS ---------------------------------------------------------------------------------
struct Srs {
  int device;
  unsigned char* pre;   // [W][n] packed 96-byte affine, pre[w][i] = 2^(c w) * P_i
  size_t n;
  u32 c, W;
};

cudaError_t srs_create(const void* bases_dev, u32 stride, size_t n, cudaStream_t s, void** handle_out) {
  if (n == 0 || n >= ((size_t)1 << 31)) return cudaErrorInvalidValue;
  Srs* h = new Srs();
  cudaGetDevice(&h->device);
  h->n = n;
  h->c = msm::choose_window_srs(n);
  h->W = msm::windows_for_srs(h->c);
  if (h->W > msm::SRS_MAX_WINDOWS || (unsigned long long)n * h->W >= (1ull << 31)) {
    delete h;
    return cudaErrorInvalidValue;
  }
  cudaError_t e = cudaMalloc((void**)&h->pre, n * h->W * 96);
  if (e != cudaSuccess) {
    delete h;
    return e;
  }
  LAUNCH_NOSYNC(msm::srs_expand_kernel, dim3((u32)((n + 127) / 128)), dim3(128), 0, s, (const unsigned char*)bases_dev, stride,
                (u32)n, h->c, h->W, h->pre);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) {
    cudaFree(h->pre);
    delete h;
    return e;
  }
  *handle_out = h;
  return cudaSuccess;
}

void srs_destroy(void* handle) {
  Srs* h = (Srs*)handle;
  if (!h) return;
  cudaFree(h->pre);
  delete h;
}

int srs_device(const void* handle) { return ((const Srs*)handle)->device; }

bool srs_size_supported(size_t n) {
  if (n == 0 || n >= ((size_t)1 << 31)) return false;
  const u32 c = msm::choose_window_srs(n), W = msm::windows_for_srs(c);
  return W <= msm::SRS_MAX_WINDOWS && (unsigned long long)n * W < (1ull << 31);
}

void srs_info(const void* handle, size_t* n, int* c, int* W, size_t* bytes) {
  const Srs* h = (const Srs*)handle;
  if (n) *n = h->n;
  if (c) *c = (int)h->c;
  if (W) *W = (int)h->W;
  if (bytes) *bytes = h->n * h->W * 96;
}

cudaError_t srs_msm(const void* handle, const void* scalars_dev, size_t n_used, void* out144_dev, cudaStream_t s, bool dry,
                    int* launches_out, float* phase_ms) {
  const Srs* h = (const Srs*)handle;
  if (n_used > h->n) return cudaErrorInvalidValue;
  msm::SrsView v{h->pre, (u32)h->n, h->c, h->W};
  if (!phase_ms || dry || n_used == 0)
    return msm::run(nullptr, 96, (const u32*)scalars_dev, n_used, (unsigned char*)out144_dev, s, dry, launches_out, nullptr, &v);
  cudaEvent_t ev[4];
  for (int i = 0; i < 4; i++) cudaEventCreate(&ev[i]);
  cudaError_t e = msm::run(nullptr, 96, (const u32*)scalars_dev, n_used, (unsigned char*)out144_dev, s, false, launches_out, ev, &v);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  for (int i = 0; i < 3; i++) {
    phase_ms[i] = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&phase_ms[i], ev[i], ev[i + 1]);
  }
  for (int i = 0; i < 4; i++) cudaEventDestroy(ev[i]);
  return e;
}

// Host scalars against a resident SRS, copied in point ranges like msm_run_host.  montgomery_in: the scalars
// are Fr in Montgomery form (KZG10::commit's polynomial) and are converted range by range on the device.
// out_bytes = 144 (normalised Jacobian) or 48 (compressed commitment).  Synchronises `s`.
cudaError_t srs_msm_host(const void* handle, const void* in_host, size_t n, bool montgomery_in, void* out_host, size_t out_bytes,
                         cudaStream_t s) {
  const Srs* h = (const Srs*)handle;
  if (n > h->n) return cudaErrorInvalidValue;
  msm::SrsView v{h->pre, (u32)h->n, h->c, h->W};
  unsigned char* d = nullptr;
  const size_t sb = (n * 32 + 255) & ~(size_t)255;
  MSM_CK(aleo::pool_malloc_async((void**)&d, sb + 512, s));
  cudaError_t e = cudaSuccess;
  if (n == 0) {
    e = msm::run(nullptr, 96, nullptr, 0, d + sb, s, false, nullptr);
  } else {
    size_t sizes[4];
    const int k = chunk_schedule(n, sizes);
    cudaStream_t cs = nullptr;
    EventSet evs;
    e = copy_stream(&cs);
    if (e == cudaSuccess) e = evs.init();
    msm::Session ss;
    if (e == cudaSuccess) e = cudaEventRecord(evs.ev[4], s);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(cs, evs.ev[4], 0);
    size_t first = 0, max_chunk = 0;
    for (int i = 0; i < k; i++) max_chunk = sizes[i] > max_chunk ? sizes[i] : max_chunk;
    if (e == cudaSuccess) e = ss.begin(n, max_chunk, (u32)k, &v, s, false);
    for (int i = 0; i < k && e == cudaSuccess; i++) {
      e = feed_h2d(d + first * 32, (const unsigned char*)in_host + first * 32, sizes[i] * 32, cs, evs.ev[4]);
      if (e == cudaSuccess) e = cudaEventRecord(evs.ev[i < 4 ? i : 5], cs);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(s, evs.ev[i < 4 ? i : 5], 0);
      if (e == cudaSuccess && montgomery_in) e = fr_to_bigint(d + first * 32, d + first * 32, sizes[i], s);
      if (e == cudaSuccess) e = ss.add_chunk(nullptr, 96, (const u32*)(d + first * 32), sizes[i], first, s);
      first += sizes[i];
    }
    if (e == cudaSuccess) e = ss.finish(d + sb, s);
    else ss.release(s);
    if (e != cudaSuccess) cudaStreamSynchronize(cs);
  }
  if (e == cudaSuccess && out_bytes == 48) {
    e = g1_compress(d + sb, d + sb + 256, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d + sb + 256, 48, cudaMemcpyDeviceToHost, s);
  } else if (e == cudaSuccess) {
    e = cudaMemcpyAsync(out_host, d + sb, 144, cudaMemcpyDeviceToHost, s);
  }
  cudaFreeAsync(d, s);
  cudaError_t e2 = cudaStreamSynchronize(s);
  return e != cudaSuccess ? e : e2;
}

// `count` independent commitments against one resident SRS in ONE launch sequence (KZG10 commits the polynomials of a
// proof against the same powers): the coefficient vectors (device, Montgomery when montgomery_in)
// are converted into one back-to-back scalar buffer, sorted together, accumulated by one kernel (MSM m owns bucket set
// m), reduced with the members as "windows", and normalised / compressed by one CTA per member.
// out_dev: count x 48 bytes (compressed) or count x 144 bytes.  count <= 64.
cudaError_t srs_msm_batch(const void* handle, const void* const* scalars_dev_ptrs, const size_t* n_each, size_t count,
                          bool montgomery_in, void* out_dev, bool compressed, cudaStream_t s) {
  const Srs* h = (const Srs*)handle;
  if (count == 0) return cudaSuccess;
  if (count > 64) return cudaErrorInvalidValue;
  std::vector<u32> off(count + 1, 0);
  size_t total = 0;
  for (size_t m = 0; m < count; m++) {
    if (n_each[m] > h->n) return cudaErrorInvalidValue;
    off[m] = (u32)total;
    total += n_each[m];
  }
  off[count] = (u32)total;
  if (total * h->W >= ((size_t)1 << 32) || total >= ((size_t)1 << 31)) return cudaErrorInvalidValue;
  msm::SrsView v{h->pre, (u32)h->n, h->c, h->W};
  unsigned char* d = nullptr;
  const size_t sb = ((total ? total : 1) * 32 + 255) & ~(size_t)255, ob = ((count + 1) * 4 + 255) & ~(size_t)255;
  MSM_CK(aleo::pool_malloc_async((void**)&d, sb + ob, s));
  u32* off_dev = (u32*)(d + sb);
  cudaError_t e = cudaMemcpyAsync(off_dev, off.data(), (count + 1) * 4, cudaMemcpyHostToDevice, s);
  for (size_t m = 0; m < count && e == cudaSuccess; m++) {
    if (n_each[m] == 0) continue;
    if (montgomery_in)
      e = fr_to_bigint(scalars_dev_ptrs[m], d + (size_t)off[m] * 32, n_each[m], s);
    else
      e = cudaMemcpyAsync(d + (size_t)off[m] * 32, scalars_dev_ptrs[m], n_each[m] * 32, cudaMemcpyDeviceToDevice, s);
  }
  // (`off` is pageable host memory: cudaMemcpyAsync has staged it by the time it returns, the vector may go away)
  msm::Session ss;
  if (e == cudaSuccess) e = ss.begin(total ? total : 1, total ? total : 1, 1, &v, s, false, (u32)count, off_dev);
  if (e == cudaSuccess) e = ss.add_chunk(nullptr, 96, (const u32*)d, total, 0, s);
  if (e == cudaSuccess) e = ss.finish((unsigned char*)out_dev, s, nullptr, compressed);
  else ss.release(s);
  cudaFreeAsync(d, s);
  return e;
}

cudaError_t fr_to_bigint(const void* in_dev, void* out_dev, size_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  LAUNCH_NOSYNC(msm::fr_to_bigint_kernel, dim3((u32)((n + 255) / 256)), dim3(256), 0, s, (const Fr*)in_dev, (Fr*)out_dev, (u32)n);
  return cudaGetLastError();
}

cudaError_t g1_compress(const void* jac144_dev, void* out48_dev, cudaStream_t s) {
  LAUNCH_NOSYNC(msm::g1_compress_kernel, dim3(1), dim3(1), 0, s, (const unsigned char*)jac144_dev, (unsigned char*)out48_dev);
  return cudaGetLastError();
}

cudaError_t g1_sum(const void* points144_dev, u32 count, void* out144_dev, cudaStream_t s) {
  LAUNCH_NOSYNC(msm::g1_sum_kernel, dim3(1), dim3(1), 0, s, (const unsigned char*)points144_dev, count,
                (unsigned char*)out144_dev);
  return cudaGetLastError();
}
}  // namespace aleo
