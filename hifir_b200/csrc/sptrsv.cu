// sptrsv.cu -- block sync-free sparse triangular solve with TMA-staged factor slabs.
//
// Replaces CCS::solve_as_strict_lower / solve_as_strict_upper of the reference
// (ds/CompressedStorage.hpp:2267-2279, 2356-2369) inside prec_solve_ldu
// (alg/prec_solve.hpp:204-223).
//
// Design (DESIGN.md "triangular sweeps"):
//  * The rows of a factor are cut, in sweep order (forward for L, backward for U), into
//    blocks of <= kRowsMax consecutive rows whose packed "slab" fits the shared-memory
//    budget.  The factors of the reference come out of AMD + Crout with strong index
//    locality: with ~900-row blocks, ~60% of all dependencies and almost the whole
//    critical path are block-local.
//  * At attach time each block's data is packed into one contiguous, 16-byte aligned
//    slab: row pointers, the list of EXTERNAL solution entries the block needs (its
//    halo, global indices), 16-bit block-local column indices and the values.  10 bytes
//    per nonzero instead of CSR's 12.
//  * One CTA per block, block order taken from a ticket counter (everything a CTA can
//    wait for is resident or finished -> no deadlock, no per-level-set launches).  One
//    elected thread pulls the whole slab into shared memory with a single TMA bulk copy
//    (cp.async.bulk + mbarrier); nothing in the dependent loop touches global memory.
//  * Roles: thread t < rows solves row t; the last kPollWarps warps are pollers that
//    fetch the halo entries from global memory (L2) into shared-memory slots as they
//    become ready.  Row threads only ever poll shared memory (~30 cycles), so a chain of
//    dependent rows inside a block advances at shared-memory latency, not L2 latency.
//  * Values carry their own ready bit (common.cuh): one 8-byte load = data + readiness.
//  * Per row the updates are applied in the reference's order (ascending column for L,
//    descending for U); only FMA contraction differs from the CPU code.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "hifgpu.h"

namespace hifgpu {

constexpr unsigned kPollWarps = 4;
constexpr unsigned kPollLanes = 32 * kPollWarps;
constexpr unsigned kRowsMax   = 1024;  // rows per block (the shared-memory budget usually binds first)
constexpr unsigned kSmemBudgetMrhs = 224 * 1024;  // multi-rhs sweeps: one CTA per SM
constexpr unsigned kSmemBudgetMax = 112 * 1024;  // per CTA -> two CTAs per SM (227 KB)
static unsigned smem_budget() {
  const char *e = std::getenv("HIFIR_B200_SMEM_KB");
  const unsigned kb = e ? static_cast<unsigned>(std::atoi(e)) : 112u;
  return std::min(kSmemBudgetMax, std::max(16u, kb) * 1024u);
}
constexpr unsigned kPollChunk  = 8;          // independent polling loads in flight per lane

// block descriptor (32 bytes)
struct alignas(16) SlabInfo {
  unsigned long long off;   // byte offset of the slab in the packed buffer
  unsigned           bytes; // slab bytes (multiple of 16)
  unsigned           rows, nnz, nhalo, s0;
  unsigned           pad;
};

__host__ __device__ inline unsigned pad16(unsigned x) { return (x + 15u) & ~15u; }
// slab layout: [ptr u32[rows+1]] [gidx u32[rows]] [halo u32[nhalo]] [order u16[rows]] [idx u16[nnz]] [val f64[nnz]],
// 16B-aligned segments.  order = the block's rows sorted by dependency depth (see kernel).
// gidx[r] = global (natural) index of block row r: the sweep order is a permutation (below)
__host__ __device__ inline unsigned slab_off_gidx(unsigned rows) { return pad16(4u * (rows + 1u)); }
__host__ __device__ inline unsigned slab_off_halo(unsigned rows) { return slab_off_gidx(rows) + pad16(4u * rows); }
__host__ __device__ inline unsigned slab_off_order(unsigned rows, unsigned nhalo) {
  return slab_off_halo(rows) + pad16(4u * nhalo);
}
__host__ __device__ inline unsigned slab_off_idx(unsigned rows, unsigned nhalo) {
  return slab_off_order(rows, nhalo) + pad16(2u * rows);
}
__host__ __device__ inline unsigned slab_off_val(unsigned rows, unsigned nhalo, unsigned nnz) {
  return slab_off_idx(rows, nhalo) + pad16(2u * nnz);
}
__host__ __device__ inline unsigned slab_bytes(unsigned rows, unsigned nhalo, unsigned nnz) {
  return slab_off_val(rows, nhalo, nnz) + pad16(8u * nnz);
}

// ---- PTX wrappers: mbarrier + TMA bulk copy (SASS: SYNCS.*, UBLKCP) ---------------
__device__ __forceinline__ unsigned smem_addr(const void *p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned phase) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_addr(bar)), "r"(phase)
      : "memory");
  return ok != 0;
}

template <bool UPPER, unsigned T>
__global__ void __launch_bounds__(T + kPollLanes, (T <= 256 ? 4 : 2))
    sptrsv_slab_kernel(const unsigned m, const unsigned char *__restrict__ slabs, const SlabInfo *__restrict__ info,
                       const double *__restrict__ rhs_plain, const unsigned long long *rhs_tagged,
                       const double *__restrict__ diag, unsigned long long *x, const unsigned parity, int *ticket,
                       int *error_flag, unsigned long long *trace, const int backoff, const unsigned poll_sleep,
                       const unsigned spin_burst, const int rhs_by_slot) {
  constexpr unsigned kThreads = T + kPollLanes;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ unsigned                        s_blk, s_done;
  __shared__ unsigned long long              s_pollend;
  __shared__ __align__(8) unsigned long long s_bar;
  const unsigned                             tid = threadIdx.x;
  if (tid == 0) {
    s_blk  = static_cast<unsigned>(atomicAdd(ticket, 1));
    s_done = 0;
    s_pollend = 0;
    mbar_init(&s_bar, 1);
  }
  __syncthreads();
  const SlabInfo bi = info[s_blk];
  if (trace && tid == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    trace[8 * s_blk + 0] = globaltimer_ns();
    trace[8 * s_blk + 3] = smid;
  }
  if (tid == 0) {
    mbar_expect_tx(&s_bar, bi.bytes);
    tma_bulk_g2s(smem, slabs + bi.off, bi.bytes, &s_bar);
  }
  const unsigned rows = bi.rows, nhalo = bi.nhalo, nnz = bi.nnz;
  volatile unsigned long long *xs = reinterpret_cast<volatile unsigned long long *>(smem + bi.bytes);
  const unsigned long long     not_ready = parity ^ 1u;
  for (unsigned i = tid; i < rows + nhalo + 1u; i += kThreads) xs[i] = not_ready;  // +1: dummy slot

  // right-hand side of a row: b_i (L sweep) or (L^{-1}b)_i / d_i with a true division
  // (prec_solve.hpp:219) for the U sweep; fetched one row ahead of its use
  const unsigned *gidx = reinterpret_cast<const unsigned *>(smem + slab_off_gidx(bi.rows));
  // gidx[r] = row code (hifgpu.h): solution slot, zero-rhs flag; slots >= m are the auxiliary
  // unknowns of the merged system (merge.cu), their right-hand side is that of row slot - m
  auto load_rhs = [&](unsigned r, unsigned &gi) -> double {
    const unsigned code = gidx[r];
    gi                  = code & kCodeSlotMask;
    if (rhs_by_slot) return rhs_plain[gi];
    if (code & kCodeZeroRhs) return 0.0;
    const unsigned ri = gi >= m ? gi - m : gi;
    return UPPER ? tag_value(rhs_tagged[ri]) / diag[ri] : rhs_plain[ri];
  };
  __syncthreads();  // xs initialised
  while (!mbar_try_wait(&s_bar, 0)) {
  }
  if (trace && tid == 0) trace[8 * s_blk + 1] = globaltimer_ns();

  const unsigned *      ptr   = reinterpret_cast<const unsigned *>(smem);
  const unsigned *      halo  = reinterpret_cast<const unsigned *>(smem + slab_off_halo(rows));
  const unsigned short *order = reinterpret_cast<const unsigned short *>(smem + slab_off_order(rows, nhalo));
  const unsigned short *idx  = reinterpret_cast<const unsigned short *>(smem + slab_off_idx(rows, nhalo));
  const double *        val  = reinterpret_cast<const double *>(smem + slab_off_val(rows, nhalo, nnz));

  if (tid < T) {
    // ---------------- row thread.  The block's rows are sorted by dependency depth
    // (order[]); thread t solves order[t], order[t+T], ... one after the other, so a
    // thread's next row is never expected to be ready before its current one and only T
    // threads (not one per row) spin at any time.  Rows that follow each other in depth
    // order -- a dependent chain -- sit in neighbouring lanes of ONE warp: the value a lane
    // publishes is seen by its neighbour's very next poll.  To keep that hop short the warp
    // does nothing but poll / consume / publish while any lane can advance; stepping a lane
    // to its next row (dependent shared loads, the right-hand side) is deferred to
    // iterations in which no lane of the warp had anything to consume.
    unsigned q = tid;  // position in order[]
    unsigned r = 0, gi = 0, gi_next = 0;
    double   acc = 0.0, acc_next = 0.0;
    unsigned k = 0, e = 0, polls = 0;
    double   a = 0.0;
    volatile unsigned long long *const dummy = xs + rows + nhalo;  // never becomes ready
    const volatile unsigned long long *pa    = dummy;               // slot this lane waits for
    bool active = q < rows, need = false;
    auto publish = [&]() {
      const unsigned long long bits = tag_set(acc, parity);
      xs[r]                         = bits;
      st_publish(x + gi, bits);
      if (trace) {
        if (atomicAdd(&s_done, 1u) + 1u == rows) {  // the row that finishes last
          trace[8 * s_blk + 2] = globaltimer_ns();
          trace[8 * s_blk + 4] = polls;
          trace[8 * s_blk + 5] = r;
          trace[8 * s_blk + 7] = s_pollend;
        }
      }
    };
    // row at position q: returns true when it has no dependency left (published at once)
    auto setup = [&]() -> bool {
      r   = order[q];
      k = ptr[r], e = ptr[r + 1];
      if (k < e) {
        pa = xs + idx[k], a = val[k];
        return false;
      }
      publish();
      return true;
    };
    if (active) {
      acc = load_rhs(order[q], gi);
      if (q + T < rows) acc_next = load_rhs(order[q + T], gi_next);
      need = setup();
      if (need) pa = dummy;
    }
    unsigned rounds = 0;
    for (;;) {
      const unsigned long long bits = *pa;
      const bool               rdy  = (static_cast<unsigned>(bits) & 1u) == parity;
      if (__any_sync(0xffffffffu, rdy)) {
        if (rdy) {
          acc = fma(-a, tag_value(bits), acc);
          ++k;
          if (k < e) {
            pa = xs + idx[k], a = val[k];
          } else {
            publish();
            need = true;
            pa   = dummy;
          }
        }
        rounds = 0;
        continue;
      }
      if (__any_sync(0xffffffffu, need)) {
        if (need) {
          q += T;
          if (q < rows) {
            acc = acc_next, gi = gi_next;
            if (q + T < rows) acc_next = load_rhs(order[q + T], gi_next);
            need = setup();
          } else {
            active = false;
            need   = false;
          }
        }
        continue;
      }
      if (!__any_sync(0xffffffffu, active)) break;
      // nothing to do: this warp waits for values that are still being computed
      if (++rounds > spin_burst) {
        if (trace) ++polls;
        if (backoff) __nanosleep(static_cast<unsigned>(backoff));
        if (rounds > kSpinLimit) {  // hang guard: flag the error, drain with garbage
          *error_flag = 1;
          if (active && !need) {
            publish();
            need = true;
            pa   = dummy;
          }
          rounds = 0;
        }
      }
    }
  } else {
    // ---------------- poller: halo entries global (L2) -> shared slots
    const unsigned lane = tid - T;
    const unsigned ne   = nhalo > lane ? (nhalo - lane + kPollLanes - 1) / kPollLanes : 0;  // my entries
    // pending entries of this lane as a bit mask (ne <= 64 by construction).  Every pass
    // re-polls ONLY the pending ones, up to kPollChunk loads in flight, so that once the
    // bulk of the halo (long-finished blocks) is in, a pass costs one L2 round trip.
    unsigned long long pend   = ne >= 64 ? ~0ull : ((1ull << ne) - 1ull);
    unsigned           passes = 0;
    while (pend) {
      unsigned long long scan = pend;
      while (scan) {
        unsigned long long v[kPollChunk];
        unsigned           ev[kPollChunk];
#pragma unroll
        for (unsigned u = 0; u < kPollChunk; ++u) {
          ev[u] = 64;
          if (scan) {
            ev[u] = static_cast<unsigned>(__ffsll(static_cast<long long>(scan))) - 1u;
            scan &= scan - 1ull;
            v[u] = ld_poll(x + halo[lane + kPollLanes * ev[u]]);
          }
        }
#pragma unroll
        for (unsigned u = 0; u < kPollChunk; ++u) {
          if (ev[u] < 64 && tag_ready(v[u], parity)) {
            xs[rows + lane + kPollLanes * ev[u]] = v[u];
            pend &= ~(1ull << ev[u]);
          }
        }
      }
      // Pending entries belong to rows that are still being computed.  In a fan-out sweep
      // hundreds of CTAs wait for the same few producer rows; polling them back to back
      // would queue thousands of loads in front of the producer's store at one L2 slice.
      if (pend && poll_sleep) __nanosleep(poll_sleep);
      if (pend && ++passes > (kSpinLimit >> 4)) {  // hang guard
        *error_flag = 1;
        while (pend) {  // release the row threads with garbage
          const unsigned q = static_cast<unsigned>(__ffsll(static_cast<long long>(pend))) - 1u;
          pend &= pend - 1ull;
          xs[rows + lane + kPollLanes * q] = parity;
        }
      }
    }
    if (trace) {
      atomicMax(&s_pollend, globaltimer_ns());
      if (lane == 0) trace[8 * s_blk + 6] = passes;
    }
  }
}

// ============================================================================
// multi-rhs sweep: NR right-hand sides per row, row-interleaved (X[i*NR + c], the
// Array<std::array<T,Nrhs>> layout of hif::HIF::solve_mrhs, builder.hpp:433-445;
// per-column arithmetic = CCS::solve_as_strict_lower/upper_mrhs,
// CompressedStorage.hpp:2286-2301, 2376-2393).  Same design as the single-rhs kernel:
// a solution slot holds NR tagged values (64 B for NR = 8: one coalesced, vectorised
// transaction per dependency), the factor slab is read ONCE for all NR columns.
// ============================================================================
__device__ __forceinline__ void lds_v2(const unsigned long long *p, unsigned long long &a, unsigned long long &b) {
  asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(smem_addr(p)) : "memory");
}
__device__ __forceinline__ void ldg_poll_v2(const unsigned long long *p, unsigned long long &a,
                                            unsigned long long &b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

template <bool UPPER, unsigned T, unsigned NR>
__global__ void __launch_bounds__(T + kPollLanes, 1)
    sptrsv_slab_mrhs_kernel(const unsigned m, const unsigned char *__restrict__ slabs,
                            const SlabInfo *__restrict__ info, const double *__restrict__ rhs_plain,
                            const unsigned long long *rhs_tagged, const double *__restrict__ diag,
                            unsigned long long *x, const unsigned parity, int *ticket, int *error_flag) {
  static_assert(NR % 2 == 0, "NR must be even (128-bit shared/global transactions)");
  constexpr unsigned kThreads = T + kPollLanes;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ unsigned                        s_blk;
  __shared__ __align__(8) unsigned long long s_bar;
  const unsigned                             tid = threadIdx.x;
  if (tid == 0) {
    s_blk = static_cast<unsigned>(atomicAdd(ticket, 1));
    mbar_init(&s_bar, 1);
  }
  __syncthreads();
  const SlabInfo bi = info[s_blk];
  if (tid == 0) {
    mbar_expect_tx(&s_bar, bi.bytes);
    tma_bulk_g2s(smem, slabs + bi.off, bi.bytes, &s_bar);
  }
  const unsigned      rows = bi.rows, nhalo = bi.nhalo, nnz = bi.nnz;
  unsigned long long *xs       = reinterpret_cast<unsigned long long *>(smem + bi.bytes);  // [slot][NR]
  const unsigned long long not_ready = parity ^ 1u;
  for (unsigned i = tid; i < (rows + nhalo + 1u) * NR; i += kThreads) xs[i] = not_ready;
  __syncthreads();
  while (!mbar_try_wait(&s_bar, 0)) {
  }
  const unsigned *      ptr   = reinterpret_cast<const unsigned *>(smem);
  const unsigned *      gidx  = reinterpret_cast<const unsigned *>(smem + slab_off_gidx(rows));
  const unsigned *      halo  = reinterpret_cast<const unsigned *>(smem + slab_off_halo(rows));
  const unsigned short *order = reinterpret_cast<const unsigned short *>(smem + slab_off_order(rows, nhalo));
  const unsigned short *idx   = reinterpret_cast<const unsigned short *>(smem + slab_off_idx(rows, nhalo));
  const double *        val   = reinterpret_cast<const double *>(smem + slab_off_val(rows, nhalo, nnz));

  if (tid < T) {
    // row threads: same structure as the single-rhs kernel -- chain rows in neighbouring
    // lanes of one warp, poll / consume / publish with priority, row switches deferred to
    // iterations in which no lane can advance
    unsigned     q = tid;
    unsigned     r = 0, k = 0, e = 0, rounds = 0;
    std::size_t  gi = 0;
    double       acc[NR], a = 0.0;
    const unsigned long long *const dummy = xs + static_cast<std::size_t>(rows + nhalo) * NR;
    const unsigned long long *      pa    = dummy;
    bool                            active = q < rows, need = false;
    auto publish = [&]() {
#pragma unroll
      for (unsigned c = 0; c < NR; ++c) {
        const unsigned long long bits = tag_set(acc[c], parity);
        xs[static_cast<std::size_t>(r) * NR + c] = bits;
        st_publish(x + gi * NR + c, bits);
      }
    };
    auto setup = [&]() -> bool {  // true: the row had no dependency and was published at once
      r  = order[q];
      gi = static_cast<std::size_t>(gidx[r]);
      if (UPPER) {
        const double d = diag[gi];
#pragma unroll
        for (unsigned c = 0; c < NR; ++c) acc[c] = tag_value(rhs_tagged[gi * NR + c]) / d;  // prec_solve.hpp:256-259
      } else {
#pragma unroll
        for (unsigned c = 0; c < NR; ++c) acc[c] = rhs_plain[gi * NR + c];
      }
      k = ptr[r], e = ptr[r + 1];
      if (k < e) {
        pa = xs + static_cast<std::size_t>(idx[k]) * NR, a = val[k];
        return false;
      }
      publish();
      pa = dummy;
      return true;
    };
    if (active) need = setup();
    for (;;) {
      // cheap poll: only the tag of the LAST value of the slot (the publisher writes it
      // last); a hit is verified on all NR tags after the full 128-bit loads, so no memory
      // ordering between the publisher's stores is assumed
      const unsigned long long hint = *reinterpret_cast<const volatile unsigned long long *>(pa + (NR - 1));
      const bool               maybe = (static_cast<unsigned>(hint) & 1u) == parity;
      if (__any_sync(0xffffffffu, maybe)) {
        unsigned long long v[NR];
        unsigned           ok = maybe ? 1u : 0u;
        if (maybe) {
#pragma unroll
          for (unsigned c = 0; c < NR; c += 2) {
            lds_v2(pa + c, v[c], v[c + 1]);
            ok &= static_cast<unsigned>((static_cast<unsigned>(v[c]) & 1u) == parity) &
                  static_cast<unsigned>((static_cast<unsigned>(v[c + 1]) & 1u) == parity);
          }
        }
        const bool rdy = ok != 0u;
        if (rdy) {
#pragma unroll
          for (unsigned c = 0; c < NR; ++c) acc[c] = fma(-a, tag_value(v[c]), acc[c]);
          ++k;
          if (k < e) {
            pa = xs + static_cast<std::size_t>(idx[k]) * NR, a = val[k];
          } else {
            publish();
            need = true;
            pa   = dummy;
          }
        }
        rounds = 0;
        continue;
      }
      if (__any_sync(0xffffffffu, need)) {
        if (need) {
          q += T;
          if (q < rows) {
            need = setup();
          } else {
            active = false;
            need   = false;
          }
        }
        continue;
      }
      if (!__any_sync(0xffffffffu, active)) break;
      if (++rounds > 4096u) {
        __nanosleep(500u);
        if (rounds > kSpinLimit) {
          *error_flag = 1;
          if (active && !need) {
            publish();
            need = true;
            pa   = dummy;
          }
          rounds = 0;
        }
      }
    }
  } else {
    // pollers: a halo entry is NR consecutive tagged values in global memory
    const unsigned     lane = tid - T;
    const unsigned     ne   = nhalo > lane ? (nhalo - lane + kPollLanes - 1) / kPollLanes : 0;
    unsigned long long pend = ne >= 64 ? ~0ull : ((1ull << ne) - 1ull);
    unsigned           passes = 0;
    while (pend) {
      unsigned long long scan = pend;
      while (scan) {
        const unsigned ev = static_cast<unsigned>(__ffsll(static_cast<long long>(scan))) - 1u;
        scan &= scan - 1ull;
        const unsigned            h   = lane + kPollLanes * ev;
        const unsigned long long *src = x + static_cast<std::size_t>(halo[h]) * NR;
        unsigned long long        v[NR];
        unsigned                  ok = 1u;
#pragma unroll
        for (unsigned c = 0; c < NR; c += 2) ldg_poll_v2(src + c, v[c], v[c + 1]);
#pragma unroll
        for (unsigned c = 0; c < NR; ++c) ok &= static_cast<unsigned>((static_cast<unsigned>(v[c]) & 1u) == parity);
        if (ok) {
#pragma unroll
          for (unsigned c = 0; c < NR; ++c) xs[static_cast<std::size_t>(rows + h) * NR + c] = v[c];
          pend &= ~(1ull << ev);
        }
      }
      if (pend) __nanosleep(100u);
      if (pend && ++passes > (kSpinLimit >> 4)) {
        *error_flag = 1;
        while (pend) {
          const unsigned qq = static_cast<unsigned>(__ffsll(static_cast<long long>(pend))) - 1u;
          pend &= pend - 1ull;
#pragma unroll
          for (unsigned c = 0; c < NR; ++c) xs[static_cast<std::size_t>(rows + lane + kPollLanes * qq) * NR + c] = parity;
        }
      }
    }
  }
}

// ============================================================================
// host: pack a strictly triangular CSR factor into slabs
// ============================================================================
namespace {
struct PackedSweep {
  std::vector<SlabInfo>      infos;
  std::vector<unsigned char> buf;
  unsigned                   max_smem = 0;
  std::size_t                halo_total = 0;
  unsigned                   block_depth = 0;
  std::vector<unsigned>      perm, pos;  // sweep position -> row of the sweep-form matrix, and back
  std::vector<unsigned>      row_of_slot;  // solution slot -> row of the sweep-form matrix
};
}  // namespace

// T is in SWEEP FORM (merge.cu: rows in sweep order, every entry references an earlier row,
// T.gid = row codes).  `upper` only selects the ordering heuristic (fan-out instead of fan-in).
// `rows_mask` (optional): pack only these rows of T; every entry of an included row must
// then reference an included row (the caller splits the matrix accordingly).
static void pack_sweep(const HostCsr &T, bool upper, PackedSweep &out, unsigned slot_bytes = 8u,
                       unsigned budget_override = 0u, const std::vector<char> *rows_mask = nullptr,
                       unsigned rows_max = kRowsMax) {
  const unsigned m = static_cast<unsigned>(T.nrows);
  if (!m) return;
  if (T.gid.size() != m) throw std::logic_error("pack_sweep: factor is not in sweep form");
  auto slot_of = [&](unsigned i) { return T.gid[i] & kCodeSlotMask; };
  std::vector<SlabInfo> &     infos = out.infos;
  std::vector<unsigned char> &buf   = out.buf;
  std::vector<unsigned>       stamp(m, 0u), slot(m, 0u);
  // dependency depth of every row (level set index), in sweep order
  std::vector<unsigned> lev(m, 0u);
  for (unsigned i = 0; i < m; ++i) {
    unsigned l = 0;
    for (unsigned k = T.ptr[i]; k < T.ptr[i + 1]; ++k) l = std::max(l, lev[T.col[k]] + 1u);
    lev[i] = l;
  }
  // ---- sweep order.  The reference's rows come in an elimination-tree post order: the
  // dependency closure of row i is (nearly) the contiguous range [lo(i), i] of the natural
  // sweep.  ~85 % of the rows of the BASELINE factors sit in closures of < 1000 rows, i.e.
  // in small closed subtrees that need NOTHING from outside.  The sweep is therefore
  // re-ordered by (phase, subtree, natural position): phase 0 = rows of closed subtrees that
  // fit one block, phase k = rows whose closure is up to 4^k times larger.  Blocks of phase 0
  // have an empty halo, blocks of phase k only need phases < k (finished long ago) plus
  // rows of their own subtree: almost no CTA ever waits, only the top of the tree is a
  // chain of blocks.  Any such order is a valid topological order of the dependency graph.
  std::vector<unsigned> &perm = out.perm, &pos = out.pos;
  std::vector<unsigned>  grp_rows, grp_nnz;
  unsigned               mt = m;  // rows actually packed
  perm.resize(m);
  pos.resize(m);
  {
    const bool natural_order = std::getenv("HIFIR_B200_SWEEP_ORDER") &&
                               std::string(std::getenv("HIFIR_B200_SWEEP_ORDER")) == "natural";
    auto nat0 = [&](unsigned s) { return s; };  // sweep form: natural sweep position = row
    auto pos0 = [&](unsigned i) { return i; };
    std::vector<unsigned> lo(m), phase(m), parent(m);
    const unsigned        S0 = 640;  // closure span of phase 0
    if (!upper) {
      // fan-in (L): closure of a row = the rows it depends on = [lo, s]
      for (unsigned s = 0; s < m; ++s) {
        const unsigned i = nat0(s);
        unsigned       l = s;
        for (unsigned k = T.ptr[i]; k < T.ptr[i + 1]; ++k) l = std::min(l, lo[pos0(T.col[k])]);
        lo[s] = l;
      }
    } else {
      // fan-out (U): what matters is the set of rows that depend ON a row (its subtree),
      // [s, hi]; rows with a large subtree (the top of the tree) go first, the many small
      // closed subtrees last -- then everything they need (their ancestors) is long finished
      for (unsigned s = 0; s < m; ++s) lo[s] = s;  // lo holds hi here
      for (unsigned s = m; s-- > 0;) {
        const unsigned i = nat0(s);
        for (unsigned k = T.ptr[i]; k < T.ptr[i + 1]; ++k) {
          const unsigned sj = pos0(T.col[k]);
          lo[sj]            = std::max(lo[sj], lo[s]);
        }
      }
    }
    unsigned max_phase = 0;
    for (unsigned s = 0; s < m; ++s) {
      const std::size_t width = upper ? lo[s] - s + 1u : s - lo[s] + 1u;
      unsigned          ph    = 0;
      for (std::size_t span = S0; width > span; span *= 4) ++ph;
      phase[s]  = ph;
      parent[s] = s;
      max_phase = std::max(max_phase, ph);
    }
    if (upper)
      for (unsigned s = 0; s < m; ++s) phase[s] = max_phase - phase[s];  // largest subtrees first
    auto find = [&](unsigned a) {
      while (parent[a] != a) a = parent[a] = parent[parent[a]];
      return a;
    };
    // rows of one phase that depend on each other belong to one subtree group
    for (unsigned s = 0; s < m; ++s) {
      const unsigned i = nat0(s);
      for (unsigned k = T.ptr[i]; k < T.ptr[i + 1]; ++k) {
        const unsigned sj = pos0(T.col[k]);
        if (phase[sj] == phase[s]) {
          const unsigned a = find(sj), b = find(s);
          if (a != b) parent[std::max(a, b)] = std::min(a, b);  // representative = first row of the group
        }
      }
    }
    std::vector<unsigned> ord;
    ord.reserve(m);
    for (unsigned s = 0; s < m; ++s)
      if (!rows_mask || (*rows_mask)[nat0(s)]) ord.push_back(s);
    mt = static_cast<unsigned>(ord.size());
    if (rows_mask)
      for (unsigned s : ord) {
        const unsigned i = nat0(s);
        for (unsigned k = T.ptr[i]; k < T.ptr[i + 1]; ++k)
          if (!(*rows_mask)[T.col[k]]) throw std::logic_error("pack_sweep: entry references an excluded row");
      }
    if (!natural_order) {
      std::vector<unsigned> rep(m);
      for (unsigned s = 0; s < m; ++s) rep[s] = find(s);
      std::stable_sort(ord.begin(), ord.end(), [&](unsigned a, unsigned b) {
        if (phase[a] != phase[b]) return phase[a] < phase[b];
        return rep[a] < rep[b];
      });
    }
    for (unsigned t = 0; t < mt; ++t) {
      perm[t]      = nat0(ord[t]);
      pos[perm[t]] = t;
    }
    // subtree groups in the new order: size of the group that STARTS at position t (0 elsewhere)
    grp_rows.assign(m + 1, 0u);
    grp_nnz.assign(m + 1, 0u);
    unsigned start = 0;
    for (unsigned t = 0; t <= mt; ++t) {
      const bool boundary = t == mt || t == 0 || natural_order || find(ord[t]) != find(ord[t - 1]) ||
                            phase[ord[t]] != phase[ord[t - 1]];
      if (boundary && t > 0) {
        grp_rows[start] = t - start;
        unsigned long long z = 0;
        for (unsigned u = start; u < t; ++u) z += T.ptr[perm[u] + 1] - T.ptr[perm[u]];
        grp_nnz[start] = static_cast<unsigned>(std::min<unsigned long long>(z, 0xffffffffu));
        start          = t;
      }
    }
  }
  std::vector<unsigned short> order;
  std::vector<unsigned>       ent, gidx;
  const bool sort_by_depth = !(std::getenv("HIFIR_B200_ENTRY_ORDER") &&
                               std::string(std::getenv("HIFIR_B200_ENTRY_ORDER")) == "natural");
  std::vector<unsigned>      ptr, halo;
  std::vector<unsigned short> idx;
  std::vector<double>         val;
  unsigned                    max_smem = 0;
  auto nat = [&](unsigned s) { return perm[s]; };

  const unsigned budget = budget_override ? budget_override : smem_budget();
  unsigned s0 = 0, bid = 0;
  while (s0 < mt) {
    ++bid;
    // ---- choose the block's rows: as many as fit the thread and shared-memory budget
    unsigned rows = 0, nnz = 0, nh = 0;
    while (s0 + rows < mt && rows < rows_max) {
      // keep a subtree group in one block when it fits an empty one: a group that is not
      // split has no dependency on a sibling block
      const unsigned gr = grp_rows[s0 + rows], gz = grp_nnz[s0 + rows];
      if (rows && gr > 1) {
        const bool fits_empty = gr <= rows_max && slab_bytes(gr, gr / 4, gz) + slot_bytes * (gr + 2 + gr / 4) <= budget;
        const bool fits_here  = rows + gr <= rows_max &&
                               slab_bytes(rows + gr, nh + gr / 4, nnz + gz) + slot_bytes * (rows + gr + 2 + nh + gr / 4) <= budget;
        if (fits_empty && !fits_here) break;
      }
      const unsigned i = nat(s0 + rows);
      unsigned       add_h = 0;
      for (unsigned k = T.ptr[i]; k < T.ptr[i + 1]; ++k) {
        const unsigned j  = static_cast<unsigned>(T.col[k]);
        const unsigned sj = pos[j];
        if (sj < s0 && stamp[j] != bid) {
          stamp[j] = bid;
          slot[j]  = nh + add_h;
          ++add_h;
        }
      }
      const unsigned rn   = T.ptr[i + 1] - T.ptr[i];
      const unsigned need = slab_bytes(rows + 1, nh + add_h, nnz + rn) + slot_bytes * (rows + 2 + nh + add_h);
      if (need > budget || nh + add_h > 64u * kPollLanes) {
        if (!rows)
          throw std::invalid_argument("triangular factor has a row too long for one shared-memory slab (" +
                                      std::to_string(rn) + " nonzeros)");
        // undo the stamps of the rejected row
        for (unsigned k = T.ptr[i]; k < T.ptr[i + 1]; ++k) {
          const unsigned j = static_cast<unsigned>(T.col[k]);
          if (stamp[j] == bid && slot[j] >= nh) stamp[j] = 0;
        }
        break;
      }
      ++rows;
      nnz += rn;
      nh += add_h;
    }
    if (rows + nh > 65536u) throw std::invalid_argument("block-local index space exceeds 16 bits");
    // ---- pack
    ptr.assign(rows + 1, 0u);
    halo.assign(nh, 0u);
    idx.clear();
    val.clear();
    for (unsigned r = 0; r < rows; ++r) {
      const unsigned i = nat(s0 + r);
      ptr[r]           = static_cast<unsigned>(idx.size());
      const unsigned b = T.ptr[i], e = T.ptr[i + 1];
      // entries in the order the dependencies become available: by dependency depth of
      // the producing row (ties: the reference's sweep order).  A row thread consumes its
      // entries in this order, so the dependency it really waits for comes last and no
      // finished work queues behind it.
      ent.resize(e - b);
      for (unsigned q = 0; q < e - b; ++q) ent[q] = b + q;
      if (sort_by_depth)
        std::stable_sort(ent.begin(), ent.end(), [&](unsigned x, unsigned y) { return lev[T.col[x]] < lev[T.col[y]]; });
      for (unsigned q = 0; q < e - b; ++q) {
        const unsigned k  = ent[q];
        const unsigned j  = static_cast<unsigned>(T.col[k]);
        const unsigned sj = pos[j];
        unsigned       loc;
        if (sj >= s0) {
          loc = sj - s0;  // a row of this block
        } else {
          loc            = rows + slot[j];
          halo[slot[j]] = slot_of(j);  // solution slot of the external entry
        }
        idx.push_back(static_cast<unsigned short>(loc));
        val.push_back(T.val[k]);
      }
    }
    ptr[rows] = static_cast<unsigned>(idx.size());
    order.resize(rows);
    for (unsigned r = 0; r < rows; ++r) order[r] = static_cast<unsigned short>(r);
    std::stable_sort(order.begin(), order.end(),
                     [&](unsigned short x, unsigned short y) { return lev[nat(s0 + x)] < lev[nat(s0 + y)]; });
    SlabInfo bi;
    bi.off   = buf.size();
    bi.bytes = slab_bytes(rows, nh, nnz);
    bi.rows = rows, bi.nnz = nnz, bi.nhalo = nh, bi.s0 = s0, bi.pad = 0;
    buf.resize(buf.size() + bi.bytes, 0);
    unsigned char *base = buf.data() + bi.off;
    std::memcpy(base, ptr.data(), 4u * (rows + 1));
    gidx.resize(rows);
    for (unsigned r = 0; r < rows; ++r) gidx[r] = T.gid[nat(s0 + r)];  // row code
    std::memcpy(base + slab_off_gidx(rows), gidx.data(), 4u * rows);
    if (nh) std::memcpy(base + slab_off_halo(rows), halo.data(), 4u * nh);
    std::memcpy(base + slab_off_order(rows, nh), order.data(), 2u * rows);
    if (nnz) {
      std::memcpy(base + slab_off_idx(rows, nh), idx.data(), 2u * nnz);
      std::memcpy(base + slab_off_val(rows, nh, nnz), val.data(), 8u * nnz);
    }
    infos.push_back(bi);
    max_smem = std::max(max_smem, bi.bytes + slot_bytes * (rows + nh + 1u));
    out.halo_total += nh;
    s0 += rows;
  }
  out.max_smem = max_smem;
  // ---- ticket order: blocks sorted by their depth in the block dependency graph (block b
  // needs the blocks its halo entries live in), ties in sweep order.  CTAs take tickets in
  // this order, so the resident blocks are the least dependent ones that are still
  // unfinished instead of a long chain of blocks that all wait for each other.
  const std::size_t     nb = infos.size();
  std::vector<unsigned> blk_of(m), blev(nb, 0u);
  std::vector<unsigned> &row_of_slot = out.row_of_slot;
  row_of_slot.assign(2 * T.orig_rows, 0u);
  for (unsigned i = 0; i < m; ++i) row_of_slot[slot_of(i)] = i;
  for (std::size_t b = 0; b < nb; ++b)
    for (unsigned r = 0; r < infos[b].rows; ++r) blk_of[infos[b].s0 + r] = static_cast<unsigned>(b);
  for (std::size_t b = 0; b < nb; ++b) {
    const unsigned *hl = reinterpret_cast<const unsigned *>(buf.data() + infos[b].off + slab_off_halo(infos[b].rows));
    unsigned        l  = 0;
    for (unsigned h = 0; h < infos[b].nhalo; ++h) {
      const unsigned j = row_of_slot[hl[h]], sj = pos[j];
      l = std::max(l, blev[blk_of[sj]] + 1u);
    }
    blev[b] = l;
  }
  std::vector<unsigned> tick(nb);
  for (std::size_t b = 0; b < nb; ++b) tick[b] = static_cast<unsigned>(b);
  if (std::getenv("HIFIR_B200_BLOCK_ORDER") && std::string(std::getenv("HIFIR_B200_BLOCK_ORDER")) == "level")
    std::stable_sort(tick.begin(), tick.end(), [&](unsigned x, unsigned y) { return blev[x] < blev[y]; });
  std::vector<SlabInfo> sorted(nb);
  for (std::size_t t = 0; t < nb; ++t) sorted[t] = infos[tick[t]];
  infos.swap(sorted);
  out.block_depth = nb ? *std::max_element(blev.begin(), blev.end()) + 1u : 0u;
}

// CPU emulation of the slab sweep (same packed data, same per-row update order, blocks in
// ticket order) -- lets the host-side packing be tested without a GPU (tests/test_abi.py).
void sweep_host_emulate(const HostCsr &T, bool upper, const double *rhs, const double *diag, double *x,
                        std::size_t stats[4], bool f32) {
  PackedSweep       P;
  const MergeParams mp = MergeParams::from_env();
  MergeStats        ms;
  HostCsr           S = merged_sweep_form(T, upper, mp, &ms);
  const unsigned      m = static_cast<unsigned>(T.nrows);
  std::vector<double> xs, xg(2 * static_cast<std::size_t>(m), 0.0);
  if (stream_sweeps()) {
    stream_host_emulate(S, upper, rhs, diag, xg.data(), stats, f32);
    std::copy(xg.begin(), xg.begin() + m, x);
    return;
  }
  pack_sweep(S, upper, P);
  for (const SlabInfo &bi : P.infos) {
    const unsigned char * base = P.buf.data() + bi.off;
    const unsigned *      ptr  = reinterpret_cast<const unsigned *>(base);
    const unsigned *      halo = reinterpret_cast<const unsigned *>(base + slab_off_halo(bi.rows));
    const unsigned short *idx  = reinterpret_cast<const unsigned short *>(base + slab_off_idx(bi.rows, bi.nhalo));
    const double *        val  = reinterpret_cast<const double *>(base + slab_off_val(bi.rows, bi.nhalo, bi.nnz));
    const unsigned *      gix  = reinterpret_cast<const unsigned *>(base + slab_off_gidx(bi.rows));
    xs.assign(bi.rows + bi.nhalo, 0.0);
    for (unsigned h = 0; h < bi.nhalo; ++h) xs[bi.rows + h] = xg[halo[h]];
    for (unsigned r = 0; r < bi.rows; ++r) {
      const unsigned code = gix[r], gi = code & kCodeSlotMask, ri = gi >= m ? gi - m : gi;
      double         acc  = (code & kCodeZeroRhs) ? 0.0 : (upper ? rhs[ri] / diag[ri] : rhs[ri]);
      for (unsigned k = ptr[r]; k < ptr[r + 1]; ++k) acc -= val[k] * xs[idx[k]];
      xs[r]  = acc;
      xg[gi] = acc;
    }
  }
  std::copy(xg.begin(), xg.begin() + m, x);
  stats[0] = P.infos.size();
  stats[1] = P.halo_total;
  stats[2] = P.buf.size();
  stats[3] = P.max_smem;
}

// Timing model of one sweep on the packed slabs (developer tool, host only): blocks take
// tickets in order and occupy one of `slots` CTA slots; a row starts when its thread is
// free, consumes its dependencies in order (each available c_s after its producer finished
// inside the block, c_g after it finished in another block), t_dep per dependency.
// prm = {slots, T, t_load, c_s, c_g, t_dep, t_pub}; out = {total, sum of block lives, max
// block life, mean halo wait, mean in-block tail, blocks}.  Times in microseconds.
void sweep_simulate(const HostCsr &T, bool upper, const double *prm, double *out) {
  PackedSweep       P;
  HostCsr           S  = to_sweep_form(T, upper);
  const MergeParams mp = MergeParams::from_env();
  MergeStats        ms;
  if (mp.enabled) S = merge_levels(S, mp, &ms, upper);
  pack_sweep(S, upper, P);
  const unsigned      m     = 2u * static_cast<unsigned>(T.nrows);  // solution slots
  const unsigned      slots = static_cast<unsigned>(prm[0]), NT = static_cast<unsigned>(prm[1]);
  const double        t_load = prm[2], c_s = prm[3], c_g = prm[4], t_dep = prm[5], t_pub = prm[6];
  std::vector<double> F(m, 0.0), fin, arrive, thr;
  std::vector<double> slot_free(slots, 0.0);
  double              prev_start = 0.0, total = 0.0, life_sum = 0.0, life_max = 0.0, wait_sum = 0.0, tail_sum = 0.0;
  for (const SlabInfo &bi : P.infos) {
    const unsigned char * base  = P.buf.data() + bi.off;
    const unsigned *      ptr   = reinterpret_cast<const unsigned *>(base);
    const unsigned *      halo  = reinterpret_cast<const unsigned *>(base + slab_off_halo(bi.rows));
    const unsigned short *order = reinterpret_cast<const unsigned short *>(base + slab_off_order(bi.rows, bi.nhalo));
    const unsigned short *idx   = reinterpret_cast<const unsigned short *>(base + slab_off_idx(bi.rows, bi.nhalo));
    const unsigned *      gix   = reinterpret_cast<const unsigned *>(base + slab_off_gidx(bi.rows));
    auto                  it    = std::min_element(slot_free.begin(), slot_free.end());
    const double          start = std::max(prev_start, *it);
    prev_start                  = start;
    const double ready = start + t_load;
    arrive.assign(bi.nhalo, 0.0);
    double last_halo = ready;
    for (unsigned h = 0; h < bi.nhalo; ++h) {
      arrive[h] = std::max(ready, F[halo[h]] + c_g);
      last_halo = std::max(last_halo, arrive[h]);
    }
    fin.assign(bi.rows, 0.0);
    thr.assign(NT, ready);
    double done = ready;
    for (unsigned q = 0; q < bi.rows; ++q) {
      const unsigned r = order[q];
      double         t = thr[q % NT];
      for (unsigned k = ptr[r]; k < ptr[r + 1]; ++k) {
        const unsigned c  = idx[k];
        const double   av = c < bi.rows ? fin[c] + c_s : arrive[c - bi.rows];
        t                 = std::max(t, av) + t_dep;
      }
      fin[r]            = t + t_pub;
      F[gix[r] & kCodeSlotMask] = fin[r];
      thr[q % NT]       = fin[r];
      done              = std::max(done, fin[r]);
    }
    *it = done;
    total = std::max(total, done);
    life_sum += done - start;
    life_max = std::max(life_max, done - start);
    wait_sum += last_halo - ready;
    tail_sum += done - last_halo;
  }
  const double nb = static_cast<double>(P.infos.size());
  out[0] = total, out[1] = life_sum, out[2] = life_max, out[3] = wait_sum / nb, out[4] = tail_sum / nb, out[5] = nb;
  out[6] = static_cast<double>(ms.ext_depth), out[7] = static_cast<double>(P.buf.size());
}

// developer tool: the block dependency graph of a packed sweep, in ticket order
void sweep_block_graph(const HostCsr &T, bool upper, std::vector<unsigned> &info, std::vector<unsigned> &src_ptr,
                       std::vector<unsigned> &src_idx) {
  PackedSweep       P;
  HostCsr           S  = to_sweep_form(T, upper);
  const MergeParams mp = MergeParams::from_env();
  if (mp.enabled) S = merge_levels(S, mp, nullptr, upper);
  pack_sweep(S, upper, P);
  const unsigned        m = static_cast<unsigned>(S.nrows);
  std::vector<unsigned> blk_of(m);
  for (std::size_t b = 0; b < P.infos.size(); ++b)
    for (unsigned r = 0; r < P.infos[b].rows; ++r) blk_of[P.infos[b].s0 + r] = static_cast<unsigned>(b);
  src_ptr.assign(1, 0u);
  std::vector<unsigned> seen(P.infos.size(), 0xffffffffu);
  for (std::size_t b = 0; b < P.infos.size(); ++b) {
    const SlabInfo &bi = P.infos[b];
    info.insert(info.end(), {bi.s0, bi.rows, bi.nhalo, bi.nnz});
    const unsigned *hl = reinterpret_cast<const unsigned *>(P.buf.data() + bi.off + slab_off_halo(bi.rows));
    for (unsigned h = 0; h < bi.nhalo; ++h) {
      const unsigned j = P.row_of_slot[hl[h]], sb = blk_of[P.pos[j]];
      if (seen[sb] != b) {
        seen[sb] = static_cast<unsigned>(b);
        src_idx.push_back(sb);
      }
    }
    src_ptr.push_back(static_cast<unsigned>(src_idx.size()));
  }
}

// Forward (L) sweeps are split in two: "lower" rows = rows of small closed
// subtrees (dependency closure spans <= 640 rows; ~85 % of the rows, blocks without halo),
// "upper" rows = the rest (the top of the elimination tree, ~15 % of the rows but ~50 % of
// the nonzeros, 3/4 of which reference lower rows).  x_l = L_ll^{-1} b_l ; r_u = b_u - L_ul x_l
// (a plain SpMV on finished data) ; x_u = L_uu^{-1} r_u.  The dependent chain of blocks at the
// top of the tree then works on L_uu only: 3-4x less data per row, small halos, 3x fewer
// block-to-block hand-offs.
static bool big_chain_blocks() {
  const char *e = std::getenv("HIFIR_B200_BIG_BLOCKS");
  return !e || std::atoi(e) != 0;
}
constexpr unsigned kRowsBig = 4096;

void split_lower_rows(const HostCsr &T, std::vector<char> &is_lower) {
  const unsigned m = static_cast<unsigned>(T.nrows);
  is_lower.assign(m, 1);
  std::vector<unsigned> lo(m);
  for (unsigned i = 0; i < m; ++i) {
    unsigned l = i;
    for (unsigned k = T.ptr[i]; k < T.ptr[i + 1]; ++k) l = std::min(l, lo[T.col[k]]);
    lo[i]       = l;
    is_lower[i] = (i - l + 1u) <= 640u;
  }
}

static void upload_plan(const PackedSweep &P, unsigned m, bool upper, unsigned nr, SweepPlan &plan,
                        std::size_t *tally) {
  plan.m          = m;
  plan.upper      = upper;
  plan.nr         = nr;
  plan.nblocks    = static_cast<unsigned>(P.infos.size());
  plan.smem_bytes = P.max_smem;
  plan.slab_bytes = P.buf.size();
  plan.halo_total = P.halo_total;
  plan.slabs.upload(P.buf.data(), P.buf.size(), tally);
  plan.info.upload(reinterpret_cast<const unsigned char *>(P.infos.data()), P.infos.size() * sizeof(SlabInfo), tally);
}

bool stream_sweeps() {
  const char *e = std::getenv("HIFIR_B200_SWEEP");
  return !e || std::string(e) != "slab";
}

void build_split_plans(const HostCsr &Tnat, SweepPlan &plan_lo, SweepPlan &plan_up, HostCsr &ul,
                       std::vector<unsigned> &urows, std::size_t *tally) {
  const unsigned    mo = static_cast<unsigned>(Tnat.nrows);
  const MergeParams mp = MergeParams::from_env();
  MergeStats        ms;
  HostCsr           T = merged_sweep_form(Tnat, false, mp, &ms);  // or the stored plan of an arena file
  if (stream_sweeps()) {  // one level-major sweep, nothing to split
    build_stream_plan(T, false, plan_lo, tally);
    plan_lo.merge = ms;
    ul.nrows = 0, ul.ncols = 2 * static_cast<std::size_t>(mo);
    ul.ptr.assign(1, 0u);
    urows.clear();
    return;
  }
  const unsigned    m = static_cast<unsigned>(T.nrows);
  std::vector<char> is_lower, is_upper(m);
  split_lower_rows(T, is_lower);
  for (unsigned i = 0; i < m; ++i) is_upper[i] = !is_lower[i];
  // L_uu (full-size index space, upper rows keep their upper entries) and L_ul (compact rows;
  // columns = solution slots of the finished lower rows, row map = row codes)
  HostCsr uu;
  uu.nrows = uu.ncols = m;
  uu.gid              = T.gid;
  uu.orig_rows        = T.orig_rows;
  uu.ptr.assign(m + 1, 0u);
  ul.nrows = 0;
  ul.ncols = 2 * static_cast<std::size_t>(mo);
  ul.ptr.assign(1, 0u);
  urows.clear();
  for (unsigned i = 0; i < m; ++i) {
    if (is_upper[i]) {
      for (unsigned k = T.ptr[i]; k < T.ptr[i + 1]; ++k) {
        if (is_upper[T.col[k]]) {
          uu.col.push_back(T.col[k]);
          uu.val.push_back(T.val[k]);
        } else {
          ul.col.push_back(static_cast<int>(T.gid[T.col[k]] & kCodeSlotMask));
          ul.val.push_back(T.val[k]);
        }
      }
      urows.push_back(T.gid[i]);
      ul.ptr.push_back(static_cast<unsigned>(ul.col.size()));
    }
    uu.ptr[i + 1] = static_cast<unsigned>(uu.col.size());
  }
  ul.nrows = urows.size();
  PackedSweep Plo, Pup;
  pack_sweep(T, false, Plo, 8u, 0u, &is_lower);
  upload_plan(Plo, mo, false, 1, plan_lo, tally);
  plan_lo.merge = ms;
  if (!urows.empty()) {
    // the top of the tree is a dependent chain of blocks: make them as large as one SM
    // allows (fewer block-to-block hand-offs); occupancy is irrelevant there
    if (big_chain_blocks())
      pack_sweep(uu, false, Pup, 8u, kSmemBudgetMrhs, &is_upper, kRowsBig);
    else
      pack_sweep(uu, false, Pup, 8u, 0u, &is_upper);
    upload_plan(Pup, mo, false, 1, plan_up, tally);
    plan_up.rhs_by_slot = true;  // its right-hand side is r_u, stored by solution slot
  }
}

void build_sweep_plan(const HostCsr &Tnat, bool upper, SweepPlan &plan, std::size_t *tally, unsigned nr) {
  plan.m       = static_cast<unsigned>(Tnat.nrows);
  plan.upper   = upper;
  plan.nblocks = 0;
  plan.nr      = nr;
  if (!plan.m) return;
  HostCsr     T;
  PackedSweep P;
  // multi-rhs plans: one CTA per SM with the whole 227 KB, a solution slot holds nr values
  if (nr > 1) {
    T = to_sweep_form(Tnat, upper);
    pack_sweep(T, upper, P, 8u * nr, kSmemBudgetMrhs);
  } else {
    const MergeParams mp = MergeParams::from_env();
    T = merged_sweep_form(Tnat, upper, mp, &plan.merge);  // or the stored plan of an arena file
    if (stream_sweeps()) {
      const MergeStats ms = plan.merge;
      build_stream_plan(T, upper, plan, tally);
      plan.merge = ms;
      return;
    }
    pack_sweep(T, upper, P);
    // a factor that yields only a few hundred blocks can not fill the GPU anyway and is
    // bound by its dependent chain: prefer few large blocks (one CTA per SM)
    if (big_chain_blocks() && P.infos.size() < 4 * kNumSMs) {
      PackedSweep Q;
      pack_sweep(T, upper, Q, 8u, kSmemBudgetMrhs, nullptr, kRowsBig);
      P = std::move(Q);
    }
  }
  plan.nblocks    = static_cast<unsigned>(P.infos.size());
  plan.smem_bytes = P.max_smem;
  plan.slab_bytes = P.buf.size();
  plan.halo_total = P.halo_total;
  plan.slabs.upload(P.buf.data(), P.buf.size(), tally);
  plan.info.upload(reinterpret_cast<const unsigned char *>(P.infos.data()), P.infos.size() * sizeof(SlabInfo), tally);
}

namespace {
int env_int(const char *name, int dflt) {
  const char *e = std::getenv(name);
  return e ? std::atoi(e) : dflt;
}
template <bool UPPER, unsigned T>
void launch_T(Handle *h, const SweepPlan &plan, const double *rhs_plain, const unsigned long long *rhs_tagged,
              const double *diag, unsigned long long *x, unsigned parity, int *ticket, unsigned long long *trace) {
  static bool configured = false;
  if (!configured) {
    HIF_CUDA(cudaFuncSetAttribute(sptrsv_slab_kernel<UPPER, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(kSmemBudgetMrhs)));
    configured = true;
  }
  sptrsv_slab_kernel<UPPER, T><<<plan.nblocks, T + kPollLanes, plan.smem_bytes, h->stream>>>(
      plan.m, plan.slabs.p, reinterpret_cast<const SlabInfo *>(plan.info.p), rhs_plain, rhs_tagged, diag, x, parity,
      ticket, h->error_flag.p, trace, env_int("HIFIR_B200_BACKOFF", 1000),
      static_cast<unsigned>(env_int("HIFIR_B200_POLL_SLEEP", 100)),
      static_cast<unsigned>(env_int("HIFIR_B200_SPIN_BURST", 4096)), plan.rhs_by_slot ? 1 : 0);
}
template <bool UPPER>
void launch_U(Handle *h, const SweepPlan &plan, const double *rhs_plain, const unsigned long long *rhs_tagged,
              const double *diag, unsigned long long *x, unsigned parity, int *ticket, unsigned long long *trace) {
  static int T = 0;
  if (!T) {
    const char *e = std::getenv("HIFIR_B200_ROW_THREADS");
    T             = e ? std::atoi(e) : 448;
  }
  switch (T) {
    case 256: launch_T<UPPER, 256>(h, plan, rhs_plain, rhs_tagged, diag, x, parity, ticket, trace); break;
    case 640: launch_T<UPPER, 640>(h, plan, rhs_plain, rhs_tagged, diag, x, parity, ticket, trace); break;
    default: launch_T<UPPER, 448>(h, plan, rhs_plain, rhs_tagged, diag, x, parity, ticket, trace); break;
  }
}
}  // namespace

namespace {
template <bool UPPER>
void launch_mrhs(Handle *h, const SweepPlan &plan, const double *rhs_plain, const unsigned long long *rhs_tagged,
                 const double *diag, unsigned long long *x, unsigned parity, int *ticket) {
  constexpr unsigned T = 448;
  static bool        configured = false;
  if (!configured) {
    HIF_CUDA(cudaFuncSetAttribute(sptrsv_slab_mrhs_kernel<UPPER, T, kMrhsWidth>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudgetMrhs)));
    configured = true;
  }
  sptrsv_slab_mrhs_kernel<UPPER, T, kMrhsWidth><<<plan.nblocks, T + kPollLanes, plan.smem_bytes, h->stream>>>(
      plan.m, plan.slabs.p, reinterpret_cast<const SlabInfo *>(plan.info.p), rhs_plain, rhs_tagged, diag, x, parity,
      ticket, h->error_flag.p);
}
}  // namespace

void launch_sweep(Handle *h, const SweepPlan &plan, const double *rhs_plain, const unsigned long long *rhs_tagged,
                  const double *diag, unsigned long long *x, unsigned parity, int *ticket, unsigned long long *trace,
                  unsigned nr) {
  if (!plan.nblocks) return;
  if (plan.stream) {
    launch_stream_sweep(h, plan, rhs_plain, rhs_tagged, diag, x, parity, ticket, trace, nr ? nr : plan.nr);
    return;
  }
  if (plan.nr > 1) {
    if (plan.nr != kMrhsWidth) throw std::logic_error("unsupported multi-rhs plan width");
    if (plan.upper)
      launch_mrhs<true>(h, plan, rhs_plain, rhs_tagged, diag, x, parity, ticket);
    else
      launch_mrhs<false>(h, plan, rhs_plain, rhs_tagged, diag, x, parity, ticket);
    HIF_KERNEL_CHECK();
    ++h->launch_count;
    return;
  }
  if (plan.upper)
    launch_U<true>(h, plan, rhs_plain, rhs_tagged, diag, x, parity, ticket, trace);
  else
    launch_U<false>(h, plan, rhs_plain, rhs_tagged, diag, x, parity, ticket, trace);
  HIF_KERNEL_CHECK();
  ++h->launch_count;
}

}  // namespace hifgpu
