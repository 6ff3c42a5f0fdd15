// cfem_device.cuh -- hand-written device-side building blocks of the fused
// per-sample evaluation kernels (FP64, sm_100a).
//
// A generated model translation unit (colloc_fem_code_b200/codegen.py) defines
// the structure tables in namespace `gen` and CFEM_TILE, then includes this
// header, then emits its kernels in terms of the helpers below:
//
//   Skew           bank-conflict-free shared-memory layout for arrays that are
//                  accessed both row-per-thread and flat;
//   stage_rows     coalesced global -> shared staging of a contiguous slab of
//                  rows of a row-major [rows][CORE] array (the IPOPT-facing
//                  decision vector keeps samples row-major, so the slab of one
//                  tile -- including the one-sample halo that the N-1 row
//                  functions read -- is ONE contiguous range of HBM);
//   warp_put /     per-warp shared-memory transposition of the values one
//   warp_flush     thread (= one sample) produces for a [rows][C] output block:
//                  lanes park their C values in the conflict-free Skew layout,
//                  then the warp streams the 32*C contiguous doubles of the
//                  block to HBM with unit-stride, full-sector stores;
//   block_reduce   deterministic warp-shuffle + shared-memory tree reduction of
//                  the per-thread objective / parameter-gradient partial sums.
//
// Reference being replaced: the NumPy broadcast evaluation of
// /root/reference/symfem.py:50-65 (one temporary array per structural nonzero).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#ifndef CFEM_TILE
#error "CFEM_TILE (threads per CTA == samples per tile) must be defined"
#endif

#include "cfem_args.cuh"

namespace cfem {

constexpr unsigned kF = 1u, kGrad = 2u, kG = 4u, kJac = 8u, kHess = 16u;
constexpr int kWarpsPerCta = CFEM_TILE / 32;
constexpr int kReduceGroup = 256;       // CTAs per first-level reduction group

// ---------------------------------------------------------------------------
// staging
// ---------------------------------------------------------------------------

// Copy `n` contiguous doubles (parameters) into shared memory.
__device__ __forceinline__ void stage_contig(double* __restrict__ dst,
                                             const double* __restrict__ src,
                                             int n, int tid)
{
    for (int e = tid; e < n; e += CFEM_TILE) dst[e] = __ldg(src + e);
}

// Bank-conflict-free shared-memory layout of a [rows][C] array of doubles that
// is written row-wise by one thread per row AND streamed flat (element e of
// the row-major array by lane e mod 32), or the other way round.  Shared
// memory has 16 banks of 8 bytes per half-warp phase.  With g = gcd(C, 16) the
// rows t and t + 16/g collide when stored densely, so every group of
// m = 16/g rows is skewed by one more double:
//     addr(t, j) = t*C + j + t/m           (row access: 16 rows -> 16 banks)
//     addr(e)    = e + e/(C*m)             (flat access: C*m is a multiple of
//                                           16, so the skew is constant inside
//                                           an aligned group of 16 elements)
template <int C>
struct Skew {
    static constexpr int gcd16(int c) { return (c % 16 == 0) ? 16 : (c % 8 == 0) ? 8 : (c % 4 == 0) ? 4 : (c % 2 == 0) ? 2 : 1; }
    static constexpr int g = gcd16(C);
    static constexpr int m = 16 / g;
    static constexpr int L = C * m;
    __host__ __device__ static constexpr int row(int t) { return t * C + t / m; }
    __host__ __device__ static constexpr int flat(int e) { return e + e / L; }
    __host__ __device__ static constexpr int size(int rows) { return rows * C + (rows + m - 1) / m; }
};

// Stage rows [row_begin, row_begin + NROWS) of a row-major [rows_total][CORE]
// array into shared memory in the Skew<CORE> layout.  Rows outside the array
// are skipped.  Global reads are unit-stride over the contiguous slab.
template <int CORE, int NROWS>
__device__ __forceinline__ void stage_rows(double* __restrict__ dst,
                                           const double* __restrict__ src,
                                           long long rows_total,
                                           long long row_begin, int tid)
{
    long long lo = row_begin < 0 ? 0 : row_begin;
    long long hi = row_begin + NROWS;
    if (hi > rows_total) hi = rows_total;
    if (hi <= lo) return;
    const int first = (int)(lo - row_begin) * CORE;
    const int nelem = (int)(hi - lo) * CORE;
    const double* __restrict__ s = src + lo * CORE;
    constexpr int kIter = (NROWS * CORE + CFEM_TILE - 1) / CFEM_TILE;
#pragma unroll
    for (int it = 0; it < kIter; ++it) {
        const int e = tid + it * CFEM_TILE;
        if (e < nelem) dst[Skew<CORE>::flat(first + e)] = __ldg(s + e);
    }
}

// Asynchronous variant (cp.async, 8-byte granules: the IPOPT-facing arrays are
// only 8-byte aligned): the copy goes global -> shared without passing through
// registers, so a tile can be prefetched while the previous one is processed.
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory");
}

template <int CORE, int NROWS>
__device__ __forceinline__ void stage_rows_async(double* __restrict__ dst,
                                                 const double* __restrict__ src,
                                                 long long rows_total,
                                                 long long row_begin, int nrows,
                                                 int tid)
{
    // nrows <= NROWS: a partial tile stages only the rows it evaluates
    long long lo = row_begin < 0 ? 0 : row_begin;
    long long hi = row_begin + nrows;
    if (hi > rows_total) hi = rows_total;
    if (hi <= lo) return;
    const int first = (int)(lo - row_begin) * CORE;
    const int nelem = (int)(hi - lo) * CORE;
    const double* __restrict__ s = src + lo * CORE;
    constexpr int kIter = (NROWS * CORE + CFEM_TILE - 1) / CFEM_TILE;
#pragma unroll
    for (int it = 0; it < kIter; ++it) {
        const int e = tid + it * CFEM_TILE;
#ifndef CFEM_EXPERIMENT_NOLOAD
        if (e < nelem) cp_async8(dst + Skew<CORE>::flat(first + e), s + e);
#endif
    }
}

// Work item -> first sample and samples per warp (cfem_args.cuh).
__device__ __forceinline__ void item_range(const KArgs& a, long long item,
                                           long long& k0, int& spw)
{
    int p = 0;
#pragma unroll
    for (int q = 1; q < kMaxPhases; ++q)
        if (q < a.nphase && item >= a.ph_item0[q]) p = q;
    k0 = a.ph_k0[p] + (item - a.ph_item0[p]) * a.ph_size[p];
    spw = a.ph_size[p] / kWarpsPerCta;
}

// ---------------------------------------------------------------------------
// per-warp output transposition
// ---------------------------------------------------------------------------

// Lane parks its C values of one output block (row access of Skew<C>).
template <int C>
__device__ __forceinline__ void warp_put(double* __restrict__ wb, int lane,
                                         const double (&v)[C])
{
    double* __restrict__ row = wb + Skew<C>::row(lane);
#pragma unroll
    for (int j = 0; j < C; ++j) row[j] = v[j];
}

// The warp writes the first nvalid*C doubles of the block chunk to HBM,
// consecutive lanes -> consecutive addresses (flat access of Skew<C>).  The
// full-warp case (all but the last tile) runs without predicates.
template <int C>
__device__ __forceinline__ void warp_flush(const double* __restrict__ wb,
                                           int lane, double* __restrict__ dst,
                                           int nvalid)
{
#ifndef CFEM_STORE_OP               // streaming (evict-first) stores by default
#define CFEM_STORE_OP(ptr, val) __stcs((ptr), (val))
#endif
#ifdef CFEM_EXPERIMENT_NOSTORE      // tuning experiment: everything but the store
#define CFEM_ST(ptr, val) do { const double v_ = (val); if (v_ == 1.2345e300) CFEM_STORE_OP((ptr), v_); } while (0)
#else
#define CFEM_ST(ptr, val) CFEM_STORE_OP((ptr), (val))
#endif
    const double* __restrict__ src = wb + lane;
    double* __restrict__ out = dst + lane;
    if (nvalid == 32) {
#pragma unroll
        for (int it = 0; it < C; ++it) {
            const int e = lane + it * 32;
            CFEM_ST(out + it * 32, src[it * 32 + e / Skew<C>::L]);
        }
    } else {
        const int total = nvalid * C;
#pragma unroll
        for (int it = 0; it < C; ++it) {
            const int e = lane + it * 32;
            if (e < total) CFEM_ST(out + it * 32, src[it * 32 + e / Skew<C>::L]);
        }
    }
}

// A block whose C entries are the same for every sample is a periodic stream:
// element e of the chunk is pat[e mod C].  No per-sample work and no
// transposition: the warp reads the C-entry pattern table and streams.
template <int C>
__device__ __forceinline__ void warp_store_periodic(const double* __restrict__ pat,
                                                    int lane, double* __restrict__ dst,
                                                    int nvalid)
{
    double* __restrict__ out = dst + lane;
    const int total = nvalid * C;
    if constexpr (32 % C == 0) {
        const double v = pat[lane % C];     // the same entry for every chunk row
#pragma unroll
        for (int it = 0; it < C; ++it)
            if (lane + it * 32 < total) CFEM_ST(out + it * 32, v);
    } else {
        constexpr int kStep = 32 % C;
        int idx = lane % C;
#pragma unroll
        for (int it = 0; it < C; ++it) {
            if (lane + it * 32 < total) CFEM_ST(out + it * 32, pat[idx]);
            idx += kStep;
            if (idx >= C) idx -= C;
        }
    }
}

// C == 1 blocks are already unit stride: no staging needed.
__device__ __forceinline__ void lane_store(double* __restrict__ dst, int lane,
                                           int nvalid, double v)
{
    if (lane < nvalid) CFEM_ST(dst + lane, v);
}

// ---------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        v += __shfl_down_sync(0xffffffffu, v, off);
    return v;
}

// Deterministic CTA reduction of R per-thread values; thread 0 writes out[r].
// `scratch` holds kWarpsPerCta * R doubles.
template <int R>
__device__ __forceinline__ void block_reduce_store(const double (&v)[R],
                                                   double* __restrict__ scratch,
                                                   double* __restrict__ out,
                                                   int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const double s = warp_sum(v[r]);
        if (lane == 0) scratch[warp * R + r] = s;
    }
    __syncthreads();
    if (tid < R) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarpsPerCta; ++w) s += scratch[w * R + tid];
        out[tid] = s;
        __threadfence();        // publish before the retirement counter moves
    }
}

// "Last block done" hand-shake: every CTA of a problem bumps the problem's
// counter once its partial sums are globally visible; the CTA that observes
// count == nblocks - 1 is the last one and returns true (in all its threads).
__device__ __forceinline__ bool last_block_done(unsigned int* counter,
                                                unsigned int nblocks, int tid)
{
    __shared__ unsigned int s_last;
    __syncthreads();            // the writers' fences are behind us
    if (tid == 0) s_last = (atomicAdd(counter, 1u) == nblocks - 1u) ? 1u : 0u;
    __syncthreads();
    return s_last != 0u;
}

// Fixed-order sum over tiles of one reduction slot (one CTA per problem).
template <int THREADS>
__device__ __forceinline__ double reduce_tiles(const double* __restrict__ part,
                                               long long ntiles, int nreduce,
                                               int slot, double* scratch,
                                               int tid)
{
    // 8 independent accumulators keep 8 L2 loads in flight per thread; the
    // association order is fixed, so the sum is reproducible run to run.
    double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    long long t = tid;
    for (; t + 7 * THREADS < ntiles; t += 8 * THREADS) {
#pragma unroll
        for (int u = 0; u < 8; ++u)     // partials were written by other SMs
            acc[u] += __ldcg(part + (t + (long long)u * THREADS) * nreduce + slot);
    }
    for (; t < ntiles; t += THREADS) acc[0] += __ldcg(part + t * nreduce + slot);
    double s = ((acc[0] + acc[1]) + (acc[2] + acc[3])) +
               ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    s = warp_sum(s);
    __syncthreads();
    if ((tid & 31) == 0) scratch[tid >> 5] = s;
    __syncthreads();
    double tot = 0.0;
    if (tid == 0) {
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) tot += scratch[w];
    }
    return tot;     // valid in thread 0
}

// Two-level reduction of the per-thread partial sums `v` of every CTA of
// problem `b` (objective and parameter-gradient terms):
//   level 0  per thread over the CTA's tiles, then CTA tree (warp shuffles +
//            shared memory)                             -> partials[cta]
//   level 1  the last CTA of every group of kReduceGroup CTAs to retire sums
//            the group's partials in CTA order          -> gpartials[group]
//   level 2  the last group to retire returns true: its CTA sums gpartials in
//            group order (cfem_finalize in the generated code).
// No floating-point atomics and a fixed association order: bitwise
// reproducible run to run, independent of CTA scheduling; the serial tail is
// two short rounds of L2 loads however long the trajectory is.
template <int R>
__device__ __forceinline__ bool tree_reduce(const KArgs& a, long long b,
                                            const double (&v)[R],
                                            double* __restrict__ scratch,
                                            int tid)
{
    const long long cta = blockIdx.x;       // one partial per (persistent) CTA
    block_reduce_store<R>(v, scratch, a.partials + (b * a.part_stride + cta) * R, tid);
    const long long g = cta / kReduceGroup;
    const long long first = g * kReduceGroup;
    const long long left = a.nctas - first;
    const unsigned in_group = left < kReduceGroup ? (unsigned)left : (unsigned)kReduceGroup;
    unsigned int* gcount = a.group_count + b * a.group_stride + g;
    if (!last_block_done(gcount, in_group, tid)) return false;
    const double* part = a.partials + (b * a.part_stride + first) * R;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const double s = reduce_tiles<CFEM_TILE>(part, in_group, R, r, scratch, tid);
        if (tid == 0) a.gpartials[(b * a.group_stride + g) * R + r] = s;
    }
    if (tid == 0) {
        *gcount = 0u;
        __threadfence();
    }
    return last_block_done(a.done_count + b, (unsigned)a.ngroups, tid);
}

// ---------------------------------------------------------------------------
// fence-free retirement (generator option retire='flag')
// ---------------------------------------------------------------------------
// The "last block done" tree above costs every tile CTA a __threadfence (which
// waits for the CTA's sample-independent stores, issued just before, to drain),
// an atomic round trip and two CTA-wide barriers: 22 % of the warp-stall
// samples of the round-2 ncu report (profiles/r02_stall_breakdown.md).  Here a
// tile CTA only STORES its partial sums: every 8-byte slot of `partials` is its
// own validity flag -- it holds kSlotEmpty (a NaN that no arithmetic produces;
// computed NaNs are canonicalised before they are published) until the CTA
// writes it, and nothing else has to be ordered with that store, so no fence,
// no atomic and no second barrier.  One extra CTA without a tile (the
// finaliser, the LAST CTA of the grid, so every tile CTA has been dispatched
// before it) polls the slots with relaxed gpu-scope loads while the last tiles
// are still running, adds them in a fixed order (thread t takes slots t,
// t + CFEM_TILE, ... in increasing order, then the usual warp / CTA tree) and
// re-arms every slot it has consumed for the next launch.
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr unsigned long long kSlotEmpty = 0xffffffffffffffffull;    // = cudaMemset(0xff)

__device__ __forceinline__ void st_relaxed_gpu_u64(double* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const double* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Tile CTA: CTA sum of the per-thread values, published to the CTA's slots.
template <int R>
__device__ __forceinline__ void publish_partial(const KArgs& a, long long b,
                                                const double (&v)[R],
                                                double* __restrict__ scratch, int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const double s = warp_sum(v[r]);
        if (lane == 0) scratch[warp * R + r] = s;
    }
    __syncthreads();
    if (tid < R) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarpsPerCta; ++w) s += scratch[w * R + tid];
        unsigned long long bits = (unsigned long long)__double_as_longlong(s);
        if (s != s) bits = 0x7ff8000000000000ull;       // never kSlotEmpty
        st_relaxed_gpu_u64(a.partials + (b * a.part_stride + (long long)blockIdx.x) * R + tid, bits);
    }
}

// Finaliser CTA: fixed-order sum of slot `r` of the `n` tile CTAs of problem
// `b`; valid in thread 0.  Up to kPollBatch slots per thread are polled
// together (independent loads); the spin is bounded: a CTA that never
// publishes yields NaN sums instead of a hung GPU.
constexpr int kPollBatch = 8;
__device__ __forceinline__ double collect_partials(const KArgs& a, long long b, int R, int r,
                                                   double* scratch, int tid)
{
    double* __restrict__ part = a.partials + b * a.part_stride * R + r;
    const long long n = a.nctas;
    double acc = 0.0;
    bool ok = true;
    const unsigned long long t0 = global_ns();
    for (long long base = tid; base < n; base += (long long)kPollBatch * CFEM_TILE) {
        unsigned long long bits[kPollBatch];
        for (;;) {
            bool all = true;
#pragma unroll
            for (int u = 0; u < kPollBatch; ++u) {
                const long long i = base + (long long)u * CFEM_TILE;
                bits[u] = i < n ? ld_relaxed_gpu_u64(part + i * R) : 0ull;
                all = all && bits[u] != kSlotEmpty;
            }
            if (all) break;
            __nanosleep(64);
            if (global_ns() - t0 > 2000000000ull) { ok = false; break; }
        }
#pragma unroll
        for (int u = 0; u < kPollBatch; ++u) {
            const long long i = base + (long long)u * CFEM_TILE;
            if (i < n) {
                acc += __longlong_as_double((long long)bits[u]);
                st_relaxed_gpu_u64(part + i * R, kSlotEmpty);      // re-arm for the next launch
            }
        }
        if (!ok) break;
    }
    if (!ok) acc = __longlong_as_double(0x7ff8000000000000ll);
    acc = warp_sum(acc);
    __syncthreads();
    if ((tid & 31) == 0) scratch[tid >> 5] = acc;
    __syncthreads();
    double tot = 0.0;
    if (tid == 0) {
#pragma unroll
        for (int w = 0; w < kWarpsPerCta; ++w) tot += scratch[w];
    }
    return tot;     // valid in thread 0
}

// The finaliser CTA of a time-sharded launch (KArgs::finaliser): wait until
// all `ngroups` groups of problem `b` have retired.  It is the LAST CTA of the
// grid, so every tile CTA has been dispatched before it and the wait cannot
// starve them; the spin is bounded all the same (a bug must not hang the GPU).
__device__ __forceinline__ void wait_groups_done(const KArgs& a, long long b, int tid)
{
    if (tid == 0) {
        const volatile unsigned int* done = a.done_count + b;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (*done < (unsigned)a.ngroups) {
            __nanosleep(64);
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 2000000000ull) break;
        }
        __threadfence();        // the groups' partial sums are visible
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// fused cross-GPU reduction (time-sharded trajectories)
// ---------------------------------------------------------------------------
// Every rank's kernel ends with the same hand-shake, executed by thread 0 of
// the CTA that finalises the rank's own sums (`tot`, R doubles per problem):
//   1. store `tot` into slot [epoch % kPeerRing][my rank] of EVERY rank's inbox
//      (peer stores through NVLink / NVSwitch-mapped memory);
//   2. one system-scope fence, then the launch epoch into the matching flag
//      of every rank (relaxed system-scope stores);
//   3. spin (relaxed loads, then one fence) until all `world` flags of the own inbox carry the
//      epoch, then sum the inbox rows IN RANK ORDER -- every rank computes the
//      same bits -- and continue with the global sums.
// No NCCL launch, no host round trip: the collective costs two NVLink
// latencies.  Two modes (KArgs::peer_defer):
//   synchronous  steps 1-3 inside the per-sample kernel (peer_allreduce): the
//                kernel returns with the global sums in f / grad.
//   pipelined    the per-sample kernel of epoch E first finishes epoch E-1
//                (step 3 for the PREVIOUS launch -- its posts were made a whole
//                kernel duration ago, so there is normally nothing to wait
//                for -- and writes that launch's f / grad), then does steps 1-2
//                for its own sums.  Ranks whose kernels finish a few
//                microseconds apart no longer wait for each other at every
//                step; a rank can run one launch ahead of the slowest one.
//                The sums of the LATEST launch are finished on demand by
//                cfem_peer_collect_kernel (before a fetch / synchronise).
// Ring hazards: slot e % kPeerRing of rank p's inbox is read by p's collects of
// epoch e (inside p's kernel of epoch e+1 and/or in a collect kernel that is
// stream-ordered before it) and overwritten by the posts of epoch
// e + kPeerRing, which a rank makes only after it has seen every rank's flag
// of epoch e + kPeerRing - 1 (pipelined) or e + kPeerRing - 1 completed
// in-kernel (synchronous): p's kernel of epoch e+1 is long finished then.
// All ranks must issue the same sequence of launches (they do: lock-step
// shards).  Spins are bounded (kPeerSpinNs): a lost peer yields NaN sums
// instead of a hung GPU.
constexpr int kPeerRing = 4;
constexpr unsigned long long kPeerSpinNs = 20ull * 1000ull * 1000ull * 1000ull;

// System-scope relaxed accesses for the flags.  Ordering comes from ONE
// __threadfence_system() per hand-shake side (fence; relaxed store = release,
// relaxed load; fence = acquire) instead of a release/acquire per flag: with W
// peers a st.release.sys per flag is W back-to-back system fences, each
// waiting for the previous remote store to be acknowledged over NVLink.
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// The hand-shake is executed by ONE WARP (all 32 lanes call these functions
// with the same arguments); lane p talks to rank p, so the W remote stores,
// the W flag polls and the W row loads are in flight together: the exchange
// costs about two NVLink latencies whatever the number of ranks.

// Wait until the flags of `epoch` from all ranks are in the own inbox.
__device__ __forceinline__ bool peer_wait(const KArgs& a, unsigned long long epoch, int lane)
{
    const int W = a.peer_world;
    bool ok = true;
    if (lane < W) {
        const unsigned long long* myflag =
            a.peer_flag[a.peer_rank] + (long long)(epoch % kPeerRing) * W + lane;
        const unsigned long long t0 = global_ns();
        while (ld_relaxed_sys(myflag) < epoch) {
            __nanosleep(32);
            if (global_ns() - t0 > kPeerSpinNs) { ok = false; break; }
        }
        __threadfence_system();     // acquire: rank `lane`'s sums are visible now
    }
    return __all_sync(0xffffffffu, ok);
}

template <int R>
__device__ __forceinline__ void peer_post(const KArgs& a, long long b, long long nb,
                                          const double (&tot)[R], int lane)
{
    const int W = a.peer_world, me = a.peer_rank;
    const unsigned long long epoch = a.peer_epoch;
    const long long slot = (long long)(epoch % kPeerRing);
    if (lane < W) {
        double* dst = a.peer_inbox[lane] + (slot * W + me) * nb * R + b * R;
#pragma unroll
        for (int r = 0; r < R; ++r) __stcg(dst + r, tot[r]);
        __threadfence_system();     // release: the sums before the flag
        // one flag per (slot, rank, problem) would be needed for batches; a
        // time-sharded problem has batch == 1, enforced on the host
        st_relaxed_sys(a.peer_flag[lane] + slot * W + me, epoch);
    }
    __syncwarp();
}

// Rank-order sum of the inbox rows of `epoch` (after all flags arrived):
// identical bits on every rank, in every lane.
template <int R>
__device__ __forceinline__ void peer_collect(const KArgs& a, long long b, long long nb,
                                             unsigned long long epoch, double (&tot)[R],
                                             int lane)
{
    const int W = a.peer_world;
    const bool ok = peer_wait(a, epoch, lane);
    const long long slot = (long long)(epoch % kPeerRing);
    double mine[R];
#pragma unroll
    for (int r = 0; r < R; ++r) mine[r] = 0.0;
    if (lane < W) {
        const double* in = a.peer_inbox[a.peer_rank] + (slot * W + lane) * nb * R + b * R;
#pragma unroll
        for (int r = 0; r < R; ++r) mine[r] = __ldcv(in + r);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) tot[r] = ok ? 0.0 : __longlong_as_double(0x7ff8000000000000ll);
    for (int p = 0; p < W; ++p) {
#pragma unroll
        for (int r = 0; r < R; ++r) tot[r] += __shfl_sync(0xffffffffu, mine[r], p);
    }
}

template <int R>
__device__ __forceinline__ void peer_allreduce(const KArgs& a, long long b, double (&tot)[R],
                                               int lane)
{
    peer_post<R>(a, b, (long long)gridDim.y, tot, lane);
    peer_collect<R>(a, b, (long long)gridDim.y, a.peer_epoch, tot, lane);
}

}  // namespace cfem
