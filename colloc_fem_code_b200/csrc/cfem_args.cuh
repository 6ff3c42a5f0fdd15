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
    double* partials;       // [batch][ntiles][nreduce]
    // two-level deterministic reduction tree (all counters zero between launches)
    long long     nctas;        // CTAs per problem of this launch (<= ntiles)
    // Balanced persistent schedule in units of warp-rows (32 samples): CTA c
    // runs `rounds` full tiles (it * nctas + c) and then ONE partial tile of
    // tail_q (+1 for c < tail_rem) warp-rows, so that every CTA of a problem
    // gets the same number of samples to within one warp-row and the CTAs
    // retire together (no half-empty last wave, no drain).
    long long     rounds;       // full tiles per CTA
    long long     tail_base;    // first warp-row of the partial tiles
    int           tail_q, tail_rem;
    long long     ngroups;      // ceil(nctas / kReduceGroup)
    double*       gpartials;    // [batch][ngroups][nreduce]
    unsigned int* group_count;  // [batch][ngroups] CTAs retired per group
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
