// oracle/shim/tbbshim/tbb/tbb.h -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the three Intel TBB facilities the reference's first-party sources use (src/cvo.cpp:116,170,246,
// src/adaptive_cvo.cpp): tbb::parallel_for(first, last, body), tbb::spin_mutex, tbb::concurrent_vector and
// tbb::task_scheduler_init::default_num_threads().  parallel_for runs the index range on OpenMP threads when the
// translation unit is compiled with -fopenmp and CVO_SHIM_THREADS (environment) is > 1, serially (in index order:
// bit-reproducible results, what the fixtures are generated with) otherwise.
#ifndef CVO_ORACLE_SHIM_TBB_H
#define CVO_ORACLE_SHIM_TBB_H
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace tbb {

inline int shim_threads() {
    static const int n = [] {
        const char* e = std::getenv("CVO_SHIM_THREADS");
        const int v = e ? std::atoi(e) : 1;
        return v > 0 ? v : 1;
    }();
    return n;
}

template <class Index, class Body> void parallel_for(Index first, Index last, const Body& body) {
#ifdef _OPENMP
    const int nt = shim_threads();
    if (nt > 1) {
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
        for (Index i = first; i < last; ++i) body(i);
        return;
    }
#endif
    for (Index i = first; i < last; ++i) body(i);
}

class spin_mutex {
    std::atomic_flag f_ = ATOMIC_FLAG_INIT;

  public:
    void lock() {
        while (f_.test_and_set(std::memory_order_acquire)) {
        }
    }
    void unlock() { f_.clear(std::memory_order_release); }
};

// grow-only vector with a thread-safe push_back (iteration is only done single-threaded, as in the reference)
template <class T> class concurrent_vector {
    std::vector<T> v_;
    spin_mutex m_;

  public:
    typedef typename std::vector<T>::iterator iterator;
    typedef typename std::vector<T>::const_iterator const_iterator;
    void clear() { v_.clear(); }
    void reserve(std::size_t n) { v_.reserve(n); }
    std::size_t size() const { return v_.size(); }
    iterator push_back(const T& t) {
        m_.lock();
        v_.push_back(t);
        iterator it = v_.end() - 1;
        m_.unlock();
        return it;
    }
    iterator begin() { return v_.begin(); }
    iterator end() { return v_.end(); }
    const_iterator begin() const { return v_.begin(); }
    const_iterator end() const { return v_.end(); }
};

class task_scheduler_init {
  public:
    static int default_num_threads() { return shim_threads(); }
};

}  // namespace tbb
#endif
