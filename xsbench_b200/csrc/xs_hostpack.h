// xs_hostpack.h -- host side of xs_gpu_lookup_samples: narrow the caller's int materials to bytes before
// they cross PCIe.
//
// The host-sample call is bound by the host->device copy (17 M samples: 136 MB of energies + 68 MB of int
// materials = 3.9 ms at 52 GB/s, against 2.9 ms of GPU work).  A material is one of 12 values
// (cuda/Materials.cu; pick_mat, cuda/Simulation.cu:311-362): as bytes the materials are 17 MB, the call moves
// 153 MB instead of 204.  A few host threads do the narrowing, chunk by chunk, into a pinned staging buffer while
// the DMA engine is busy with the chunk's energies -- the host work hides behind the copy it shortens.
// A value outside [0, 255] becomes 255: the device-side validation (sanitize_sample) then rejects it like
// any other bad material, so the call still fails with XS_ERR_ARG.
#pragma once

#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

namespace xs {

inline void narrow_materials(const int *src, uint8_t *dst, long n)
{
    for (long i = 0; i < n; i++) {
        const unsigned v = (unsigned)src[i];
        dst[i] = (uint8_t)(v > 255u ? 255u : v);
    }
}

// A small persistent pool: run() splits [0, n) over the workers and the calling thread and returns when all
// of it is done.  One run at a time (the C ABI is not re-entrant per context).
class PackPool {
public:
    explicit PackPool(int workers)
    {
        for (int i = 0; i < workers; i++) threads_.emplace_back([this, i] { loop(i); });
    }
    ~PackPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        go_.notify_all();
        for (auto &t : threads_) t.join();
    }
    PackPool(const PackPool &) = delete;
    PackPool &operator=(const PackPool &) = delete;

    void run(const int *src, uint8_t *dst, long n)
    {
        const int parts = (int)threads_.size() + 1;
        if (parts == 1 || n < 65536) { narrow_materials(src, dst, n); return; }
        {
            std::lock_guard<std::mutex> lk(mu_);
            src_ = src; dst_ = dst; n_ = n;
            pending_ = parts - 1;
            generation_++;
        }
        go_.notify_all();
        share(src, dst, n, parts - 1, parts);                // the caller takes the last share
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }

private:
    static void share(const int *src, uint8_t *dst, long n, int i, int parts)
    {
        // shares start on 64-byte lines of the destination
        const long a = (n * i / parts) & ~63L, b = i + 1 == parts ? n : (n * (i + 1) / parts) & ~63L;
        if (b > a) narrow_materials(src + a, dst + a, b - a);
    }
    void loop(int i)
    {
        unsigned long seen = 0;
        for (;;) {
            const int *src; uint8_t *dst; long n;
            {
                std::unique_lock<std::mutex> lk(mu_);
                go_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                src = src_; dst = dst_; n = n_;
            }
            share(src, dst, n, i, (int)threads_.size() + 1);
            std::lock_guard<std::mutex> lk(mu_);
            if (--pending_ == 0) done_.notify_one();
        }
    }

    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable go_, done_;
    const int *src_ = nullptr;
    uint8_t *dst_ = nullptr;
    long n_ = 0;
    int pending_ = 0;
    unsigned long generation_ = 0;
    bool stop_ = false;
};

}  // namespace xs
