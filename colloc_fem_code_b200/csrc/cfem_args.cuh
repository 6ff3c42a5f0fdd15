// cfem_args.cuh -- the kernel argument block shared by every kernel of a
// generated model library.  Included after namespace `gen` has defined the
// structure-table sizes (kNumVars, kNumData, ...).  Kept in its own header so
// that the (slow to compile) parameter-only kernel translation unit depends on
// nothing else that is hand-written.
#pragma once

namespace cfem {

constexpr int kMaxPeers = 16;

template <int N> struct AtLeastOne { static constexpr int value = N > 0 ? N : 1; };

// Kernel argument block: everything is resolved on the host at cfem_create().
struct KArgs {
    long long N;            // samples per problem
    long long ntiles;       // ceil(N / CFEM_TILE)
    long long ndec, ncons, nnz_jac, nnz_hess;   // per-problem strides
    int       nreduce;      // reduction slots per tile
    double    obj_factor;
    const double* dvec;
    const double* lam;
    const double* data[AtLeastOne<gen::kNumData>::value];
    long long     data_rows[AtLeastOne<gen::kNumData>::value];
    double        scalars[AtLeastOne<gen::kNumScalars>::value];
    double* f;
    double* grad;
    double* g;
    double* jac;
    double* hess;
    double* partials;       // [batch][ntiles][nreduce]: one partial sum per TILE
    // two-level deterministic reduction tree (all counters zero between launches)
    long long     nctas;        // CTAs per problem of this launch (<= ntiles)
    // Work items = tiles of CFEM_TILE samples.  Small launches (every tile fits
    // a resident CTA) give tile blockIdx.x to CTA blockIdx.x.  Large launches
    // run ONE resident set of persistent CTAs that draw tiles from a ticket
    // counter (dynamic = 1): SMs progress at different rates under memory
    // contention, a static partition would wait for the slowest one.  The
    // counter is never reset: a launch consumes exactly nitems + nctas tickets
    // (every CTA stops at its first ticket >= nitems), the host advances
    // ticket_base by that much.
    long long           nitems;
    int                 dynamic;
    unsigned long long  ticket_base;
    unsigned long long* ticket;         // [batch]
    long long     ngroups;      // ceil(ntiles / kReduceGroup)
    double*       gpartials;    // [batch][ngroups][nreduce]
    unsigned int* group_count;  // [batch][ngroups] tiles retired per group
    unsigned int* done_count;   // [batch] groups retired per problem
    // fused cross-GPU reduction over peer memory (time-sharded runs)
    int                 peer_rank, peer_world;      // world <= 1: disabled
    int                 peer_defer;                 // 1: pipelined (finish the previous epoch, post this one)
    unsigned            peer_prev_mask;             // result mask of the previous posting launch
    unsigned long long  peer_epoch;                 // launch counter, > 0
    double*             peer_inbox[kMaxPeers];      // rank p's inbox  [kPeerRing][world][batch*nreduce]
    unsigned long long* peer_flag[kMaxPeers];       // rank p's flags  [kPeerRing][world]
    double* reduce;         // [batch][nreduce]
    long long var_off[AtLeastOne<gen::kNumVars>::value];
    long long var_rows[AtLeastOne<gen::kNumVars>::value];
    long long cons_off[AtLeastOne<gen::kNumCons>::value];
    long long fun_rows[AtLeastOne<gen::kNumFuns>::value];
    long long jac_off[AtLeastOne<gen::kNumJacBlocks>::value];
    long long hess_off[AtLeastOne<gen::kNumHessBlocks>::value];
};

}  // namespace cfem
