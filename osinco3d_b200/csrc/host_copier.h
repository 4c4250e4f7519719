// host_copier.h -- worker threads that produce HOST-side outputs of the pipelined host-pointer
// procedures without a device -> host copy (pipeline.cu).
//
// predict_velocity leaves three history levels per component behind (src/integration.f90:176-188):
//     level 3 = level 2 as it came in,  level 2 = level 1 = the new right-hand side.
// Two of the three are copies of arrays the HOST already holds (the old level 2) or is about to
// receive (the new level 1): downloading them again costs 6 of the 13 fields the call sends back
// over PCIe.  The copier makes them with memcpy on a few worker threads instead, overlapped with
// the transfers, chunk by chunk:
//     job A (chunk c, component k):  level3[c] <- level2[c]   once the H2D copies of chunk c are done
//                                                             (the DMA engine still reads level3[c])
//     job B (chunk c, component k):  level2[c] <- level1[c]   once the D2H copy of level1[c] has
//                                                             landed AND job A of the same chunk has
//                                                             read the old level2[c]
// and nu_t = 0.d0 of a DNS call (src/integration.f90:112) with memset.
//
// Plain C++ (no CUDA types): a job waits on up to two opaque "events" through a caller-supplied
// function, so tests/cpu/host_copier_test.cpp drives the same code with fake events under
// -fsanitize=thread.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace o3d {

struct HostJob {
    void* ev[2] = {nullptr, nullptr};         // device work to wait for (nullptr: none)
    const std::atomic<int>* after = nullptr;  // host job that must have finished first
    std::atomic<int>* done = nullptr;         // set to 1 when this job has finished
    void* dst = nullptr;
    const void* src = nullptr;                // nullptr: zero fill
    size_t bytes = 0;
};

class HostCopier {
public:
    // wait(ev) blocks until the device work behind `ev` is complete; returns 0 on success
    typedef int (*WaitFn)(void* ev);
    // enter() runs once on every worker before its first job (cudaSetDevice)
    typedef void (*EnterFn)(int arg);

    HostCopier(int nthreads, WaitFn wait, EnterFn enter = nullptr, int enter_arg = 0)
        : wait_(wait), enter_(enter), enter_arg_(enter_arg) {
        if (nthreads < 1) nthreads = 1;
        th_.reserve(nthreads);
        try {
            for (int i = 0; i < nthreads; ++i) th_.emplace_back([this] { run(); });
        } catch (...) {
            // thread limit reached: work with the threads that did start; none at all -> throw
            // (no joinable thread is left behind then)
            if (th_.empty()) throw;
        }
    }
    ~HostCopier() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (auto& t : th_) t.join();
    }
    HostCopier(const HostCopier&) = delete;
    HostCopier& operator=(const HostCopier&) = delete;

    int threads() const { return (int)th_.size(); }

    // Jobs are taken in the order they were pushed.  A job named by `after` must have been pushed
    // EARLIER: it is then already running (or done) when the dependent job is taken, and it never
    // waits on a host job itself unless that one was pushed earlier still -- no deadlock.
    void push(const HostJob& j) {
        {
            std::lock_guard<std::mutex> lk(m_);
            q_.push_back(j);
        }
        cv_work_.notify_one();
    }
    // all jobs pushed so far have finished; returns the first non-zero status a wait reported
    int drain() {
        std::unique_lock<std::mutex> lk(m_);
        cv_idle_.wait(lk, [this] { return q_.empty() && busy_ == 0; });
        const int e = err_;
        err_ = 0;
        return e;
    }

private:
    void run() {
        if (enter_) enter_(enter_arg_);
        for (;;) {
            HostJob j;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_work_.wait(lk, [this] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;  // stop_
                j = q_.front();
                q_.pop_front();
                ++busy_;
            }
            int e = 0;
            for (int i = 0; i < 2; ++i)
                if (j.ev[i]) {
                    const int r = wait_(j.ev[i]);
                    if (r && !e) e = r;
                }
            if (j.after)
                while (!j.after->load(std::memory_order_acquire)) std::this_thread::yield();
            // (after a failed wait the call reports an error and its outputs are undefined anyway;
            // the copy still runs so that dependent jobs are released)
            if (j.bytes) {
                if (j.src)
                    std::memcpy(j.dst, j.src, j.bytes);
                else
                    std::memset(j.dst, 0, j.bytes);
            }
            if (j.done) j.done->store(1, std::memory_order_release);
            {
                std::lock_guard<std::mutex> lk(m_);
                if (e && !err_) err_ = e;
                --busy_;
                if (q_.empty() && busy_ == 0) cv_idle_.notify_all();
            }
        }
    }

    WaitFn wait_;
    EnterFn enter_;
    int enter_arg_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_work_, cv_idle_;
    std::deque<HostJob> q_;
    int busy_ = 0;
    int err_ = 0;
    bool stop_ = false;
};

// The jobs of one chunk [off, off + bytes/8) of one history array f(nx,ny,nz,3) (N = nx*ny*nz
// doubles per level), src/integration.f90:176-188:
// job A: f(:,:,:,3) = f(:,:,:,2) -- waits for the H2D copies of the chunk (both upload lanes)
inline HostJob shift_job_a(double* f, long long N, long long off, size_t bytes, void* ev_up0,
                           void* ev_up1, std::atomic<int>* done) {
    HostJob j;
    j.ev[0] = ev_up0, j.ev[1] = ev_up1;
    j.dst = f + 2 * N + off, j.src = f + N + off, j.bytes = bytes;
    j.done = done;
    return j;
}
// job B: f(:,:,:,2) = f(:,:,:,1) -- waits for the D2H copy of level 1 and (itscheme = 3) for job A
// of the same chunk, which reads the old level 2
inline HostJob shift_job_b(double* f, long long N, long long off, size_t bytes, void* ev_l1,
                           const std::atomic<int>* after) {
    HostJob j;
    j.ev[0] = ev_l1;
    j.after = after;
    j.dst = f + N + off, j.src = f + off, j.bytes = bytes;
    return j;
}
inline HostJob zero_job(double* a, long long off, size_t bytes) {
    HostJob j;
    j.dst = a + off, j.bytes = bytes;
    return j;
}

}  // namespace o3d
