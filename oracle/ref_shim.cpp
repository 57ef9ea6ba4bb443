// ref_shim.cpp -- TEST INFRASTRUCTURE.  Compiles the reference's OWN headers
// (pqt/triangle.cuh, pqt/bitonicSort.cuh, included from /root/reference where
// they lie -- never copied) for the host, so that the oracle's restatement can
// be checked against the real code on a CPU-only box:
//   * triangle.cuh is __device__ __host__ code: it compiles as plain C++.
//   * bitonicSort.cuh's in-block networks are run by real host threads, one per
//     CUDA thread, with __syncthreads() mapped to a pthread barrier and
//     threadIdx / blockDim as (thread-local) globals.  All compare-exchanges
//     of one network stage touch disjoint pairs, so this is race-free.
// Output goes to oracle/_ref/libpqt_ref_host.so (git-ignored).
#include <pthread.h>
#include <sys/types.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct shim_dim3 {
  unsigned x, y, z;
};
static thread_local shim_dim3 threadIdx;
static shim_dim3 blockDim;
static pthread_barrier_t g_bar;

#define __device__
#define __host__
#define __global__
#define __shared__
#define __syncthreads() pthread_barrier_wait(&g_bar)

namespace pqt {
float shm[1];   // "extern __shared__" of the header's sortTestLarge (never called here)
float shmf[1];  // "extern __shared__" of the header's scanTestLarge (never called here)
}  // namespace pqt

#include "pqt/triangle.cuh"
#include "pqt/bitonicSort.cuh"

namespace {
struct sort_job {
  float *val;
  uint *idx;
  uint n;
  int large;
  unsigned tid;
};
void *sort_thread(void *arg) {
  sort_job *j = static_cast<sort_job *>(arg);
  threadIdx.x = j->tid;
  threadIdx.y = threadIdx.z = 0;
  if (j->large)
    pqt::bitonicLarge<float>(j->val, j->idx, j->n);
  else
    pqt::bitonic3<float>(j->val, j->idx, j->n);
  return nullptr;
}
}  // namespace

extern "C" {

unsigned short ref_toUShort(float f) { return pqt::toUShort(f); }
float ref_toFloat(unsigned short s) { return pqt::toFloat(s); }
float ref_dist(float a2, float b2, float c2, float l) { return pqt::dist(a2, b2, c2, l); }
float ref_project(float a2, float b2, float c2) { return pqt::project(a2, b2, c2); }
float ref_project_d(float a2, float b2, float c2, float *d2) {
  volatile float d;
  float l = pqt::project(a2, b2, c2, d);
  *d2 = d;
  return l;
}
int ref_equal(float a, float b) { return pqt::equal(a, b) ? 1 : 0; }

// runs bitonic3 (large == 0, one thread per element like the reference's call
// sites) or bitonicLarge (large != 0, nthreads threads striding over n)
int ref_bitonic(float *val, unsigned *idx, unsigned n, int large, unsigned nthreads) {
  if (!large) nthreads = n;
  if (nthreads == 0 || nthreads > 4096) return -1;
  blockDim.x = nthreads;
  blockDim.y = blockDim.z = 1;
  if (pthread_barrier_init(&g_bar, nullptr, nthreads)) return -2;
  std::vector<pthread_t> th(nthreads);
  std::vector<sort_job> jobs(nthreads);
  pthread_attr_t attr;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, 256 * 1024);
  for (unsigned t = 0; t < nthreads; t++) {
    jobs[t] = sort_job{val, idx, n, large, t};
    if (pthread_create(&th[t], &attr, sort_thread, &jobs[t])) return -3;
  }
  for (unsigned t = 0; t < nthreads; t++) pthread_join(th[t], nullptr);
  pthread_attr_destroy(&attr);
  pthread_barrier_destroy(&g_bar);
  return 0;
}

}  // extern "C"
