// TEST INFRASTRUCTURE ONLY -- the extra layer the reference's AVX thread-pool path needs on top of
// ref_shim.h (probe P2): the two MSVC intrinsics of its per-8-pixel spin lock
// (projekt.cpp:1381, 1405, 2211, 2235) and the platform work queue it submits rows to
// (Platform.AddEntry, projekt.cpp:3609, 3809), whose implementation is not in the snapshot.
//
// The reference text itself is never written into this repository: oracle/Makefile streams
// /root/reference/projekt.cpp through five sed rules that turn MSVC's vector member access
// (X.m256i_u32[n] and friends, 494 occurrences) into the equivalent pointer cast, into g++.
#ifndef B200R_REF_SHIM_AVX_H
#define B200R_REF_SHIM_AVX_H

#include "ref_shim.h"

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

// char _InterlockedCompareExchange8(char volatile *Destination, char Exchange, char Comparand): returns the
// initial value of *Destination (projekt.cpp:1381: CAS 0 -> 1 on the ZMask byte of an 8-pixel group)
static inline char _InterlockedCompareExchange8(char volatile *Destination, char Exchange, char Comparand)
{
    char Expected = Comparand;
    __atomic_compare_exchange_n(Destination, &Expected, Exchange, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE);
    return Expected;
}
// projekt.cpp:1405: the stores to colour and depth must not sink below the unlocking store
#define _WriteBarrier() __atomic_thread_fence(__ATOMIC_RELEASE)

typedef void platform_work_queue_callback(platform_work_queue *Queue, void *Data);

// A minimal work queue: Threads workers pop (callback, data) entries in FIFO order.
struct platform_work_queue
{
    struct entry { platform_work_queue_callback *Callback; void *Data; };
    std::mutex Mutex;
    std::condition_variable HaveWork, AllDone;
    std::deque<entry> Entries;
    std::vector<std::thread> Workers;
    unsigned InFlight = 0;
    bool Quit = false;

    void Start(unsigned Threads)
    {
        for(unsigned T = 0; T < Threads; ++T)
            Workers.emplace_back([this]() {
                for(;;)
                {
                    entry E;
                    {
                        std::unique_lock<std::mutex> Lock(Mutex);
                        HaveWork.wait(Lock, [this]() { return Quit || !Entries.empty(); });
                        if(Entries.empty()) return;
                        E = Entries.front(); Entries.pop_front();
                    }
                    E.Callback(this, E.Data);
                    {
                        std::unique_lock<std::mutex> Lock(Mutex);
                        if(--InFlight == 0) AllDone.notify_all();
                    }
                }
            });
    }
    void Stop()
    {
        { std::unique_lock<std::mutex> Lock(Mutex); Quit = true; }
        HaveWork.notify_all();
        for(auto &W : Workers) W.join();
        Workers.clear(); Quit = false;
    }
    void CompleteAllWork()
    {
        std::unique_lock<std::mutex> Lock(Mutex);
        AllDone.wait(Lock, [this]() { return InFlight == 0; });
    }
};

static void RefAddEntry(platform_work_queue *Queue, platform_work_queue_callback *Callback, void *Data)
{
    { std::unique_lock<std::mutex> Lock(Queue->Mutex); Queue->Entries.push_back({Callback, Data}); ++Queue->InFlight; }
    Queue->HaveWork.notify_one();
}
struct platform_api { void (*AddEntry)(platform_work_queue *, platform_work_queue_callback *, void *); };
static platform_api Platform = { RefAddEntry };

#endif
