// msm_lib.cu -- translation unit holding the G1 MSM kernels.
#include "internal.h"
#include "msm_host.cuh"
#ifndef ALEO_EMU
#include <thread>
#endif

namespace aleo {
cudaError_t msm_upload_constants() { return aleo_upload_field_constants(); }
bool msm_size_supported(size_t n) { return msm::size_supported(n); }
int msm_window_bits(size_t n) { return (int)msm::make_params(n).c; }

// ---- host-pointer calls: the MSM arrives in point ranges so that the copy of range k + 1 overlaps the
// accumulation of range k (PCIe moves 136 bytes per point about 2.5x faster than the integer pipe adds it) ----
namespace {
thread_local cudaStream_t t_copy_stream = nullptr;
thread_local int t_copy_dev = -1;

cudaError_t copy_stream(cudaStream_t* out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (t_copy_stream == nullptr || t_copy_dev != dev) {
    e = cudaStreamCreateWithFlags(&t_copy_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return e;
    t_copy_dev = dev;
  }
  *out = t_copy_stream;
  return cudaSuccess;
}

// Point ranges of a host-pointer MSM.  The first range is small (its copy is the only one not hidden) and the
// later ones grow: a range's accumulation must outlast the copy of the next one.
int chunk_schedule(size_t n, size_t* sizes) {
  const char* env = getenv("ALEO_B200_MSM_CHUNKS");  // read per call: tests and sweeps switch it
  const long k_env = env ? atol(env) : 0L;
  int k = n < ((size_t)1 << 19) ? 1 : (n < ((size_t)1 << 22) ? 2 : 3);
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
    sizes[0] = n / 16;
    sizes[1] = (n * 3) / 16;
    sizes[2] = n / 4;
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
// ---- host -> device copies of caller memory that may be PAGEABLE (a Rust Vec is) ---------------------------------------
// Pinned / registered memory goes out as one cudaMemcpyAsync.  A large pageable block would be staged by the driver
// through one bounce buffer at a few GB/s -- slower than the MSM itself -- so it is staged here instead: FEED_THREADS
// helper threads memcpy 4 MB slices into their own pinned double buffers and enqueue the DMA from there (the CPU copy
// of slice i + T overlaps the DMA of slice i).  The call returns when everything is enqueued; `cs` then waits on the
// helpers' streams.  Staging buffers and streams live per calling thread (8 MB pinned per helper; feed_threads()
// helpers, ALEO_B200_FEED_THREADS overrides).
#ifndef ALEO_EMU
constexpr int FEED_THREADS = 8;  // upper bound; feed_threads() of them run
constexpr size_t FEED_SLICE = (size_t)4 << 20;
constexpr size_t FEED_MIN_BYTES = (size_t)8 << 20;  // below this the plain pageable copy is fine

// helper threads per staged copy: one core copies ~10 GB/s, PCIe takes ~53 GB/s
int feed_threads() {
  static const int n = [] {
    const char* env = getenv("ALEO_B200_FEED_THREADS");
    long v = env ? atol(env) : 0L;
    if (v < 1) {
      const unsigned hc = std::thread::hardware_concurrency();
      v = hc >= 12 ? 6 : 4;
    }
    return (int)(v > FEED_THREADS ? FEED_THREADS : v);
  }();
  return n;
}

struct FeedState {
  int dev = -1;
  unsigned char* buf[FEED_THREADS][2] = {};
  cudaEvent_t ev[FEED_THREADS][2] = {};
  cudaStream_t st[FEED_THREADS] = {};
  cudaError_t init(int device) {
    if (dev == device) return cudaSuccess;
    for (int t = 0; t < feed_threads(); t++) {
      cudaError_t e = cudaStreamCreateWithFlags(&st[t], cudaStreamNonBlocking);
      for (int b = 0; b < 2 && e == cudaSuccess; b++) {
        e = cudaMallocHost((void**)&buf[t][b], FEED_SLICE);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[t][b], cudaEventDisableTiming);
      }
      if (e != cudaSuccess) return e;
    }
    dev = device;
    return cudaSuccess;
  }
};
thread_local FeedState t_feed;


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
    int dev = 0;
    MSM_CK(cudaGetDevice(&dev));
    MSM_CK(t_feed.init(dev));
    FeedState* fs = &t_feed;
    const size_t nslices = (bytes + FEED_SLICE - 1) / FEED_SLICE;
    const int NT = feed_threads();
    cudaError_t errs[FEED_THREADS];
    std::thread workers[FEED_THREADS];
    for (int t = 0; t < NT; t++) {
      errs[t] = cudaSuccess;
      workers[t] = std::thread([=, &errs]() {
        cudaError_t e = cudaSetDevice(dev);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(fs->st[t], ready, 0);
        for (size_t i = (size_t)t, round = 0; i < nslices && e == cudaSuccess; i += (size_t)NT, round++) {
          const int b = (int)(round & 1);
          const size_t off = i * FEED_SLICE, len = (bytes - off < FEED_SLICE) ? bytes - off : FEED_SLICE;
          e = cudaEventSynchronize(fs->ev[t][b]);  // the DMA that last read this buffer is done
          if (e != cudaSuccess) break;
          std::memcpy(fs->buf[t][b], (const unsigned char*)src_host + off, len);
          e = cudaMemcpyAsync((unsigned char*)dst_dev + off, fs->buf[t][b], len, cudaMemcpyHostToDevice, fs->st[t]);
          if (e == cudaSuccess) e = cudaEventRecord(fs->ev[t][b], fs->st[t]);
        }
        errs[t] = e;
      });
    }
    cudaError_t e = cudaSuccess;
    for (int t = 0; t < NT; t++) {
      workers[t].join();
      if (errs[t] != cudaSuccess) e = errs[t];
    }
    // cs continues after every helper stream's last copy (events of both buffers cover the tail of each stream)
    for (int t = 0; t < NT && e == cudaSuccess; t++)
      for (int b = 0; b < 2 && e == cudaSuccess; b++) e = cudaStreamWaitEvent(cs, fs->ev[t][b], 0);
    return e;
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
    int dev = 0;
    MSM_CK(cudaGetDevice(&dev));
    MSM_CK(t_feed.init(dev));
    FeedState* fs = &t_feed;
    cudaEvent_t done;
    MSM_CK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    MSM_CK(cudaEventRecord(done, s));
    const size_t nslices = (bytes + FEED_SLICE - 1) / FEED_SLICE;
    const int NT = feed_threads();
    cudaError_t errs[FEED_THREADS];
    std::thread workers[FEED_THREADS];
    for (int t = 0; t < NT; t++) {
      errs[t] = cudaSuccess;
      workers[t] = std::thread([=, &errs]() {
        cudaError_t e = cudaSetDevice(dev);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(fs->st[t], done, 0);
        for (int b = 0; b < 2 && e == cudaSuccess; b++) e = cudaEventSynchronize(fs->ev[t][b]);  // buffers idle
        size_t prev_off = 0, prev_len = 0;
        int prev_b = -1;
        for (size_t i = (size_t)t, round = 0; e == cudaSuccess; i += (size_t)NT, round++) {
          const int b = (int)(round & 1);
          const bool more = i < nslices;
          size_t off = 0, len = 0;
          if (more) {
            off = i * FEED_SLICE;
            len = (bytes - off < FEED_SLICE) ? bytes - off : FEED_SLICE;
            e = cudaMemcpyAsync(fs->buf[t][b], (const unsigned char*)src_dev + off, len, cudaMemcpyDeviceToHost, fs->st[t]);
            if (e == cudaSuccess) e = cudaEventRecord(fs->ev[t][b], fs->st[t]);
          }
          if (prev_b >= 0 && e == cudaSuccess) {  // copy the previous slice out while this one is in flight
            e = cudaEventSynchronize(fs->ev[t][prev_b]);
            if (e == cudaSuccess) std::memcpy((unsigned char*)dst_host + prev_off, fs->buf[t][prev_b], prev_len);
          }
          if (!more) break;
          prev_off = off;
          prev_len = len;
          prev_b = b;
        }
        errs[t] = e;
      });
    }
    cudaError_t e = cudaSuccess;
    for (int t = 0; t < NT; t++) {
      workers[t].join();
      if (errs[t] != cudaSuccess) e = errs[t];
    }
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
  MSM_CK(cudaMallocAsync((void**)&d, o_out + 256, s));
  cudaError_t e = cudaSuccess;
  if (n == 0) {
    e = msm::run(nullptr, stride, nullptr, 0, d + o_out, s, false, nullptr);
  } else {
    size_t sizes[4];
    const int k = chunk_schedule(n, sizes);
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
  MSM_CK(cudaMallocAsync((void**)&d, sb + 512, s));
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
  MSM_CK(cudaMallocAsync((void**)&d, sb + ob, s));
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
