// host_extract.cpp -- channel 0 of the reference's 3-channel interleaved host images, taken on the HOST.
//
// The reference keeps every pyramid level as 3-channel u8 on the host and its kernels read channel 0
// (g_conv_3ch_1ch_constant_..., OptFlowGpu.cu:1040-1090: `3 * (...)`).  The batched host entry point used to
// upload all three channels (6 B per pixel pair) and drop two of them on the device.  Measured on the B200 boxes
// the end-to-end path is bound by the bytes that cross PCIe in BOTH directions together (about 75 GB/s for one
// GPU: 48 GB/s out while 27 GB/s go in), so the two unused channels cost a quarter of the throughput.  With
// ofb_ctx_set_host_threads(ctx, n > 0) a small pool of host threads extracts channel 0 into pinned planar staging
// buffers while the previous sub-batches are in flight, and only 2 B per pixel pair are uploaded.
#include "host_extract.hpp"

#include <cstring>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define OFB_X86 1
#endif

namespace ofb {

#ifdef OFB_X86
// 48 interleaved bytes -> 16 planar bytes with three byte shuffles
__attribute__((target("ssse3"))) static void extract_ssse3(const uint8_t *src, uint8_t *dst, size_t npix)
{
    const __m128i m0 = _mm_setr_epi8(0, 3, 6, 9, 12, 15, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m128i m1 = _mm_setr_epi8(-1, -1, -1, -1, -1, -1, 2, 5, 8, 11, 14, -1, -1, -1, -1, -1);
    const __m128i m2 = _mm_setr_epi8(-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 1, 4, 7, 10, 13);
    size_t i = 0;
    for (; i + 16 <= npix; i += 16) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 3 * i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 3 * i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 3 * i + 32));
        const __m128i r = _mm_or_si128(_mm_or_si128(_mm_shuffle_epi8(a, m0), _mm_shuffle_epi8(b, m1)), _mm_shuffle_epi8(c, m2));
        _mm_storeu_si128(reinterpret_cast<__m128i *>(dst + i), r);
    }
    for (; i < npix; i++) dst[i] = src[3 * i];
}
#endif

void c3_extract_channel0(const uint8_t *src, uint8_t *dst, size_t npix)
{
#ifdef OFB_X86
    static const bool ssse3 = __builtin_cpu_supports("ssse3");
    if (ssse3) {
        extract_ssse3(src, dst, npix);
        return;
    }
#endif
    for (size_t i = 0; i < npix; i++) dst[i] = src[3 * i];
}

// ---- a minimal pool: the caller's thread takes part, workers sleep between jobs ----------------------------------
HostPool::HostPool(int threads)
{
    const int extra = threads > 1 ? threads - 1 : 0;
    for (int t = 0; t < extra; t++) workers_.emplace_back([this] { worker(); });
}

HostPool::~HostPool()
{
    {
        std::lock_guard<std::mutex> lk(m_);
        stop_ = true;
    }
    cv_.notify_all();
    for (auto &w : workers_) w.join();
}

void HostPool::drain(const Job &j)
{
    for (;;) {
        const size_t k = next_.fetch_add(1, std::memory_order_relaxed);
        if (k >= j.chunks) break;
        j.fn(j.arg, k);
        if (done_.fetch_add(1, std::memory_order_acq_rel) + 1 == j.chunks) {
            std::lock_guard<std::mutex> lk(m_);
            cv_done_.notify_all();
        }
    }
}

void HostPool::worker()
{
    unsigned long long seen = 0;
    for (;;) {
        Job j;
        {
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [&] { return stop_ || epoch_ != seen; });
            if (stop_) return;
            seen = epoch_;
            j = job_;
            active_++;
        }
        drain(j);
        {
            std::lock_guard<std::mutex> lk(m_);
            active_--;
            cv_done_.notify_all();
        }
    }
}

void HostPool::run(size_t chunks, void (*fn)(void *, size_t), void *arg)
{
    if (chunks == 0) return;
    Job j{fn, arg, chunks};
    {
        std::lock_guard<std::mutex> lk(m_);
        job_ = j;
        next_.store(0, std::memory_order_relaxed);
        done_.store(0, std::memory_order_relaxed);
        epoch_++;
    }
    cv_.notify_all();
    drain(j);
    // every chunk done AND every worker out of this job (a late worker must not take chunk indices of the next one)
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [&] { return done_.load(std::memory_order_acquire) >= chunks && active_ == 0; });
}

namespace {
struct ExtractJob {
    const uint8_t *src[2];
    uint8_t *dst[2];
    size_t npix_per_frame_set; // pixels of one frame set (cnt images of w*h)
    size_t chunk_px;
    size_t chunks_per_set;
};
void extract_chunk(void *arg, size_t k)
{
    const ExtractJob *e = static_cast<const ExtractJob *>(arg);
    const size_t set = k / e->chunks_per_set, c = k - set * e->chunks_per_set;
    const size_t p0 = c * e->chunk_px;
    const size_t n = p0 + e->chunk_px <= e->npix_per_frame_set ? e->chunk_px : e->npix_per_frame_set - p0;
    c3_extract_channel0(e->src[set] + 3 * p0, e->dst[set] + p0, n);
}
} // namespace

void c3_extract_pair_sets(HostPool *pool, const uint8_t *prev_c3, const uint8_t *next_c3, uint8_t *prev_planar,
                          uint8_t *next_planar, size_t npix)
{
    ExtractJob e;
    e.src[0] = prev_c3;
    e.src[1] = next_c3;
    e.dst[0] = prev_planar;
    e.dst[1] = next_planar;
    e.npix_per_frame_set = npix;
    e.chunk_px = 1u << 18; // 256 K pixels: 768 KB read, 256 KB written per chunk
    e.chunks_per_set = (npix + e.chunk_px - 1) / e.chunk_px;
    if (pool) {
        pool->run(2 * e.chunks_per_set, extract_chunk, &e);
    } else {
        for (size_t k = 0; k < 2 * e.chunks_per_set; k++) extract_chunk(&e, k);
    }
}

} // namespace ofb
