// Developer micro-benchmark (B200): latency of the primitives a sync-free sweep is built from.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/lat_bench tools/lat_bench.cu && gpurun_out/lat_bench
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_cg(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// pointer chase: a[i] = next index
template <int MODE>
__global__ void chase(const unsigned long long *a, int hops, long long *out) {
  unsigned long long i = 0;
  long long          t0 = clock64();
  for (int h = 0; h < hops; ++h) i = MODE == 0 ? ld_relaxed(a + i) : (MODE == 1 ? ld_cg(a + i) : a[i]);
  long long t1 = clock64();
  out[0] = t1 - t0;
  out[1] = (long long)i;
}

// ping-pong between CTA 0 and CTA 1 (different SMs) through global memory
__global__ void pingpong(unsigned long long *flag, int rounds, long long *out) {
  if (threadIdx.x) return;
  const unsigned me = blockIdx.x;
  long long      t0 = clock64();
  for (int r = 1; r <= rounds; ++r) {
    if (me == 0) {
      st_relaxed(flag, 2 * r - 1);
      while (ld_relaxed(flag + 16) != (unsigned long long)(2 * r)) {
      }
    } else {
      while (ld_relaxed(flag) != (unsigned long long)(2 * r - 1)) {
      }
      st_relaxed(flag + 16, 2 * r);
    }
  }
  if (me == 0) out[0] = clock64() - t0;
}

// the same through an atomic counter (one waits for the other's atomicAdd)
__global__ void pingpong_atomic(int *ctr, int rounds, long long *out) {
  if (threadIdx.x) return;
  const unsigned me = blockIdx.x;
  long long      t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    // me == 0 adds when the counter is even, me == 1 when odd
    while ((*(volatile int *)ctr & 1) != (int)me) {
    }
    atomicAdd(ctr, 1);
  }
  if (me == 0) out[0] = clock64() - t0;
}

// ping-pong between the two CTAs of a cluster through distributed shared memory
__global__ void __cluster_dims__(2, 1, 1) pingpong_dsmem(int rounds, long long *out) {
  __shared__ unsigned long long box;
  cg::cluster_group                cl = cg::this_cluster();
  if (threadIdx.x == 0) box = 0;
  cl.sync();
  const unsigned      me   = cl.block_rank();
  unsigned long long *peer = cl.map_shared_rank(&box, me ^ 1u);
  volatile unsigned long long *mine = &box;
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    for (int r = 1; r <= rounds; ++r) {
      if (me == 0) {
        *reinterpret_cast<volatile unsigned long long *>(peer) = 2 * r - 1;
        while (*mine != (unsigned long long)(2 * r)) {
        }
      } else {
        while (*mine != (unsigned long long)(2 * r - 1)) {
        }
        *reinterpret_cast<volatile unsigned long long *>(peer) = 2 * r;
      }
    }
    if (me == 0) out[0] = clock64() - t0;
  }
  cl.sync();
}

// ping-pong between two warps of one CTA through shared memory
__global__ void pingpong_smem(int rounds, long long *out) {
  __shared__ volatile unsigned long long a, b;
  if (threadIdx.x == 0) a = b = 0;
  __syncthreads();
  const unsigned w = threadIdx.x >> 5;
  if (threadIdx.x & 31) return;
  long long t0 = clock64();
  for (int r = 1; r <= rounds; ++r) {
    if (w == 0) {
      a = 2 * r - 1;
      while (b != (unsigned long long)(2 * r)) {
      }
    } else {
      while (a != (unsigned long long)(2 * r - 1)) {
      }
      b = 2 * r;
    }
  }
  if (w == 0) out[0] = clock64() - t0;
}

// gather latency: one warp, 8 independent ld.relaxed per lane from a buffer just written by OTHER SMs
__global__ void writer(unsigned long long *x, size_t n) {
  for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) st_relaxed(x + i, i * 2 + 1);
}
__global__ void gather8(const unsigned long long *x, const unsigned *idx, int rounds, long long *out) {
  unsigned long long acc = 0;
  long long          t0  = clock64();
  for (int r = 0; r < rounds; ++r) {
    unsigned long long g[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) g[u] = ld_relaxed(x + idx[(r * 8 + u) * 32 + threadIdx.x]);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += g[u];
    acc = __shfl_xor_sync(0xffffffffu, acc, 1) + acc;  // dependent: next round's loads wait
    if (acc == 12345) out[3] = 1;
  }
  if (threadIdx.x == 0) out[0] = clock64() - t0;
}

// throughput of random 8-byte gathers from an L2-resident buffer, whole GPU
template <int MODE>
__global__ void gather_rate(const unsigned long long *x, unsigned n, int iters, unsigned long long *sink) {
  unsigned           s   = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  unsigned long long acc = 0;
  for (int it = 0; it < iters; ++it) {
    unsigned long long g[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      s ^= s << 13, s ^= s >> 17, s ^= s << 5;
      const unsigned i = s % n;
      g[u] = MODE == 0 ? ld_relaxed(x + i) : (MODE == 1 ? ld_cg(x + i) : x[i]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += g[u];
  }
  if (acc == 42) sink[0] = acc;
}

int main() {
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("SM clock attr %d kHz\n", clk);
  long long *out;
  cudaMallocManaged(&out, 64);
  const double ns = 1e6 / clk;  // per cycle
  {  // pointer chase over 32 MB (L2 resident) and 1 GB (DRAM)
    for (size_t mb : {32, 1024}) {
      size_t                          n = mb * 1024 * 1024 / 8;
      std::vector<unsigned long long> h(n);
      // random cyclic permutation with large strides
      unsigned long long cur = 0;
      const size_t       stride = 104729 * 16 + 1;
      for (size_t k = 0; k < n; ++k) {
        unsigned long long nxt = (cur + stride) % n;
        h[cur]                 = nxt;
        cur                    = nxt;
      }
      unsigned long long *d;
      cudaMalloc(&d, n * 8);
      cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
      const int hops = 20000;
      for (int rep = 0; rep < 2; ++rep) {
        chase<0><<<1, 1>>>(d, hops, out);
        cudaDeviceSynchronize();
        double a = (double)out[0] / hops;
        chase<1><<<1, 1>>>(d, hops, out);
        cudaDeviceSynchronize();
        double b = (double)out[0] / hops;
        chase<2><<<1, 1>>>(d, hops, out);
        cudaDeviceSynchronize();
        double c = (double)out[0] / hops;
        printf("chase %4zu MB: ld.relaxed.gpu %.0f cyc (%.0f ns)  ld.cg %.0f cyc  ld %.0f cyc\n", mb, a, a * ns, b, c);
      }
      cudaFree(d);
    }
  }
  {
    unsigned long long *flag;
    cudaMalloc(&flag, 4096);
    cudaMemset(flag, 0, 4096);
    const int rounds = 5000;
    pingpong<<<2, 32>>>(flag, rounds, out);
    cudaDeviceSynchronize();
    printf("global ping-pong: round trip %.0f cyc (%.0f ns) -> one way %.0f ns\n", (double)out[0] / rounds,
           out[0] * ns / rounds, out[0] * ns / rounds / 2);
    int *ctr;
    cudaMalloc(&ctr, 4);
    cudaMemset(ctr, 0, 4);
    pingpong_atomic<<<2, 32>>>(ctr, rounds, out);
    cudaDeviceSynchronize();
    printf("atomic ping-pong: per hand-off %.0f cyc (%.0f ns)\n", (double)out[0] / rounds / 2, out[0] * ns / rounds / 2);
    pingpong_dsmem<<<2, 32>>>(rounds, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("dsmem ping-pong (%s): round trip %.0f cyc (%.0f ns) -> one way %.0f ns\n", cudaGetErrorString(e),
           (double)out[0] / rounds, out[0] * ns / rounds, out[0] * ns / rounds / 2);
    pingpong_smem<<<1, 64>>>(rounds, out);
    cudaDeviceSynchronize();
    printf("smem ping-pong: round trip %.0f cyc (%.0f ns) -> one way %.0f ns\n", (double)out[0] / rounds,
           out[0] * ns / rounds, out[0] * ns / rounds / 2);
  }
  {  // gather of freshly written tagged values
    const size_t        n = 4 << 20;  // 32 MB
    unsigned long long *x;
    cudaMalloc(&x, n * 8);
    const int             rounds = 2000;
    std::vector<unsigned> hi(rounds * 8 * 32);
    unsigned long long    s = 88172645463325252ull;
    for (auto &v : hi) {
      s ^= s << 13, s ^= s >> 7, s ^= s << 17;
      v = (unsigned)(s % n);
    }
    unsigned *idx;
    cudaMalloc(&idx, hi.size() * 4);
    cudaMemcpy(idx, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) {
      writer<<<592, 256>>>(x, n);
      gather8<<<1, 32>>>(x, idx, rounds, out);
      cudaDeviceSynchronize();
      printf("gather8 (32 lanes x 8 ld.relaxed.gpu, data written by other SMs): %.0f cyc (%.0f ns) per round\n",
             (double)out[0] / rounds, out[0] * ns / rounds);
    }
  }
  {  // gather throughput
    const unsigned      n = 4u << 20;  // 32 MB
    unsigned long long *x;
    cudaMalloc(&x, (size_t)n * 8);
    cudaMemset(x, 1, (size_t)n * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int threads : {256}) {
      for (int ctas : {148 * 2, 148 * 4, 148 * 8}) {
        const int iters = 200;
        float     ms[3];
        for (int mode = 0; mode < 3; ++mode) {
          for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) gather_rate<0><<<ctas, threads>>>(x, n, iters, (unsigned long long *)out);
            if (mode == 1) gather_rate<1><<<ctas, threads>>>(x, n, iters, (unsigned long long *)out);
            if (mode == 2) gather_rate<2><<<ctas, threads>>>(x, n, iters, (unsigned long long *)out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms[mode], e0, e1);
          }
        }
        const double g = (double)ctas * threads * iters * 8;
        printf("gather rate, %d CTAs x %d thr: ld.relaxed.gpu %.0f G/s  ld.cg %.0f G/s  ld %.0f G/s\n", ctas, threads,
               g / ms[0] / 1e6, g / ms[1] / 1e6, g / ms[2] / 1e6);
      }
    }
  }
  return 0;
}
