// multi_thread_executor.h -- accepted-and-ignored stand-in.
//
// /root/reference/nonlinear_optimizer/multi_thread_executor.h:44-179 is the std::thread pool that
// splits the CPU assembly over correspondence ranges.  On the device the whole scan is one grid,
// so SetMultiThreadExecutor() keeps compiling but the pool is never used.  One behavioural
// difference follows and is deliberate: the reference's threaded path silently drops the
// `N mod num_threads` tail correspondences (mahalanobis_distance_minimizer_analytic.cc:59-73);
// the device path always processes all N, i.e. it matches the reference's single-thread result.
#ifndef NONLINEAR_OPTIMIZER_MULTI_THREAD_EXECUTOR_H_
#define NONLINEAR_OPTIMIZER_MULTI_THREAD_EXECUTOR_H_

class MultiThreadExecutor {
 public:
  explicit MultiThreadExecutor(const int num_threads_in_pool) : num_threads_{num_threads_in_pool} {}
  int GetNumOfTotalThreads() const { return num_threads_; }

 private:
  int num_threads_{0};
};

#endif  // NONLINEAR_OPTIMIZER_MULTI_THREAD_EXECUTOR_H_
