// ll_bench.cu -- what an exchange between CTAs costs on this GPU (micro-benchmark behind the design of the
// resident kernel's per-iteration exchange).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ll_bench ll_bench.cu
//   1. one-way latency of a tagged word between two CTAs on different SMs, by access flavour
//   2. period of an "all-gather of records" loop (every CTA stores a 448-byte record of LL words, P of the
//      CTAs gather all of them), by number of records, pollers and loads in flight
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long Timer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

template <int F> __device__ __forceinline__ void St(unsigned long long* p, unsigned long long v) {
  if (F == 0) *reinterpret_cast<volatile unsigned long long*>(p) = v;
  else if (F == 1) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
  else if (F == 2) __stcg(p, v);
  else atomicExch(p, v);
}
template <int F> __device__ __forceinline__ unsigned long long Ld(const unsigned long long* p) {
  unsigned long long v;
  if (F == 0) v = *reinterpret_cast<const volatile unsigned long long*>(p);
  else if (F == 1) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  else if (F == 2) v = __ldcg(p);
  else v = atomicAdd(const_cast<unsigned long long*>(p), 0ULL);
  return v;
}

// CTA 0 <-> CTA `other`: `reps` round trips of one word each way
template <int F>
__global__ void pingpong(unsigned long long* a, unsigned long long* b, int other, int reps, unsigned long long* out) {
  if (threadIdx.x != 0) return;
  if (blockIdx.x == 0) {
    const unsigned long long t0 = Timer();
    for (int i = 1; i <= reps; ++i) {
      St<F>(a, i);
      while (Ld<F>(b) != static_cast<unsigned long long>(i)) {}
    }
    out[0] = Timer() - t0;
  } else if (blockIdx.x == other) {
    for (int i = 1; i <= reps; ++i) {
      while (Ld<F>(a) != static_cast<unsigned long long>(i)) {}
      St<F>(b, i);
    }
  }
}

__device__ __forceinline__ void StLL(unsigned long long* p, unsigned long long lo, unsigned long long hi) {
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ void LdLL(const unsigned long long* p, unsigned long long& lo, unsigned long long& hi) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
}

// every CTA: store record (28 x 16 B) with tag it; pollers (blockIdx % poll_every == 0) gather all records;
// non-pollers only wait for CTA 0's "done" word of the iteration (so that all CTAs stay in lockstep).
// MODE 0: thread (j, l8) walks records l8, l8+8, ... with two loads in flight; MODE 1: one (record, quad) item per
// thread, all loads in flight.
template <int MODE>
__global__ void allgather(unsigned long long* recs /*[2][G][28][2]*/, unsigned long long* done /*[2][G]*/, int iters,
                          int poll_every, unsigned long long* out, double* sink) {
  const int G = gridDim.x, tid = threadIdx.x;
  __shared__ double stage[160][28];
  __shared__ double lanes[8][28];
  double acc = 0.0;
  unsigned long long t0 = 0;
  for (int it = 0; it < iters; ++it) {
    if (it == 8 && tid == 0) t0 = Timer();
    const unsigned long long tag = static_cast<unsigned long long>(it + 1) << 32;
    unsigned long long* base = recs + static_cast<size_t>(it & 1) * G * 56;
    if (tid < 28) StLL(base + blockIdx.x * 56 + 2 * tid, tag | (tid + 1), tag | blockIdx.x);
    if (blockIdx.x % poll_every == 0) {
      if (MODE == 0) {
        const int j = tid >> 3, l8 = tid & 7;
        if (j < 28) {
          double s = 0.0;
          const unsigned long long* src = base + l8 * 56 + 2 * j;
          unsigned long long lo, hi, nlo = 0, nhi = 0;
          if (l8 < G) LdLL(src, lo, hi);
          for (int c = l8; c < G; c += 8) {
            if (c + 8 < G) LdLL(src + 8 * 56, nlo, nhi);
            while ((lo >> 32) != (tag >> 32) || (hi >> 32) != (tag >> 32)) LdLL(src, lo, hi);
            s += static_cast<double>(lo & 0xffffffffULL);
            src += 8 * 56; lo = nlo; hi = nhi;
          }
          lanes[l8][j] = s;
        }
        __syncthreads();
        if (tid < 28) { double s = 0; for (int w = 0; w < 8; ++w) s += lanes[w][tid]; acc += s; }
      } else {
        for (int item = tid; item < G * 7; item += blockDim.x) {
          const int c = item / 7, j0 = 4 * (item - 7 * c);
          const unsigned long long* src = base + c * 56 + 2 * j0;
          unsigned long long lo[4], hi[4];
          for (int k = 0; k < 4; ++k) LdLL(src + 2 * k, lo[k], hi[k]);
          for (int k = 0; k < 4; ++k) {
            while ((lo[k] >> 32) != (tag >> 32) || (hi[k] >> 32) != (tag >> 32)) LdLL(src + 2 * k, lo[k], hi[k]);
            stage[c][j0 + k] = static_cast<double>(lo[k] & 0xffffffffULL);
          }
        }
        __syncthreads();
        const int j = tid >> 3, l8 = tid & 7;
        if (j < 28) { double s = 0; for (int c = l8; c < G; c += 8) s += stage[c][j]; lanes[l8][j] = s; }
        __syncthreads();
        if (tid < 28) { double s = 0; for (int w = 0; w < 8; ++w) s += lanes[w][tid]; acc += s; }
      }
      if (blockIdx.x == 0 && tid == 0) {
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(done + (it & 1)), "l"(static_cast<unsigned long long>(it + 1)) : "memory");
      }
    } else if (tid == 0) {
      unsigned long long v;
      do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(done + (it & 1)) : "memory"); } while (v != static_cast<unsigned long long>(it + 1));
    }
    __syncthreads();
  }
  if (tid == 0 && blockIdx.x == 0) out[0] = Timer() - t0;
  if (tid < 28) sink[blockIdx.x * 28 + tid] = acc;
}

template <int MODE>
void run_allgather(int G, int poll_every, int iters, unsigned long long* recs, unsigned long long* done,
                   unsigned long long* out, double* sink) {
  CK(cudaMemset(recs, 0, 2 * 320 * 56 * 8));
  CK(cudaMemset(done, 0, 2 * 8));
  void* args[] = {&recs, &done, &iters, &poll_every, &out, &sink};
  CK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(&allgather<MODE>), dim3(G), dim3(256), args, 0, nullptr));
  CK(cudaDeviceSynchronize());
  unsigned long long ns;
  CK(cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost));
  printf("all-gather  records %3d  pollers every %d  mode %s : %.2f us / iteration\n", G, poll_every,
         MODE == 0 ? "2-in-flight walk" : "all-in-flight staged", ns / 1e3 / (iters - 8));
}

int main() {
  unsigned long long *a, *b, *out, *recs, *done;
  double* sink;
  CK(cudaMalloc(&a, 256)); CK(cudaMalloc(&b, 256)); CK(cudaMalloc(&out, 64));
  CK(cudaMalloc(&recs, 2 * 320 * 56 * 8)); CK(cudaMalloc(&done, 64)); CK(cudaMalloc(&sink, 320 * 28 * 8));
  const int reps = 2000;
  const char* names[] = {"volatile", "relaxed.gpu", "stcg/ldcg", "atomic"};
  for (int other : {1, 74, 147}) {
    for (int f = 0; f < 4; ++f) {
      CK(cudaMemset(a, 0, 8)); CK(cudaMemset(b, 0, 8));
      void* args[] = {&a, &b, const_cast<int*>(&other), const_cast<int*>(&reps), &out};
      void* fn = f == 0 ? (void*)&pingpong<0> : f == 1 ? (void*)&pingpong<1> : f == 2 ? (void*)&pingpong<2> : (void*)&pingpong<3>;
      CK(cudaLaunchCooperativeKernel(fn, dim3(148), dim3(32), args, 0, nullptr));
      CK(cudaDeviceSynchronize());
      unsigned long long ns;
      CK(cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost));
      printf("ping-pong CTA 0 <-> CTA %3d  %-12s : %.0f ns one way\n", other, names[f], ns / 2.0 / reps);
    }
  }
  for (int G : {16, 33, 66, 132}) {
    for (int pe : {1, 4}) {
      run_allgather<0>(G, pe, 400, recs, done, out, sink);
      run_allgather<1>(G, pe, 400, recs, done, out, sink);
    }
  }
  return 0;
}
