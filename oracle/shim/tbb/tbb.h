// oracle/shim/tbb/tbb.h — legacy TBB as cvo.cpp uses it, executed sequentially (the reference's results do not
// depend on the schedule except through the order of its locked double additions)
#pragma once
#include "concurrent_vector.h"
namespace tbb {
template <class F> inline void parallel_for(int b, int e, const F &f) { for (int i = b; i < e; i++) f(i); }
struct spin_mutex { void lock() {} void unlock() {} };
struct task_scheduler_init { static int default_num_threads() { return 1; } };
}  // namespace tbb
