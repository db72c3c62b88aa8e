// host_extract.hpp -- host-side channel extraction for the batched host entry point (host_extract.cpp).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

namespace ofb {

// dst[i] = src[3 * i] for i < npix (channel 0 of interleaved 3-channel bytes)
void c3_extract_channel0(const uint8_t *src, uint8_t *dst, size_t npix);

// `threads` host threads including the caller's; run() returns when every chunk has been processed.
class HostPool {
  public:
    explicit HostPool(int threads);
    ~HostPool();
    HostPool(const HostPool &) = delete;
    HostPool &operator=(const HostPool &) = delete;
    void run(size_t chunks, void (*fn)(void *, size_t), void *arg);
    int threads() const { return (int)workers_.size() + 1; }

  private:
    struct Job {
        void (*fn)(void *, size_t) = nullptr;
        void *arg = nullptr;
        size_t chunks = 0;
    };
    void worker();
    void drain(const Job &j);
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, cv_done_;
    Job job_;
    std::atomic<size_t> next_{0}, done_{0};
    unsigned long long epoch_ = 0;
    int active_ = 0;
    bool stop_ = false;
};

// channel 0 of two frame sets of npix pixels each (prev and next images of a sub-batch), split over the pool
// (pool == nullptr: on the caller's thread)
void c3_extract_pair_sets(HostPool *pool, const uint8_t *prev_c3, const uint8_t *next_c3, uint8_t *prev_planar,
                          uint8_t *next_planar, size_t npix);

} // namespace ofb
