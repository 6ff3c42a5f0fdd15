// cfem_args.cuh -- the kernel argument block shared by every kernel of a
// generated model library.  Included after namespace `gen` has defined the
// structure-table sizes (kNumVars, kNumData, ...).  Kept in its own header so
// that the (slow to compile) parameter-only kernel translation unit depends on
// nothing else that is hand-written.
#pragma once

namespace cfem {

constexpr int kMaxPeers = 16;
constexpr int kMaxPhases = 4;

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
    double* partials;       // [batch][part_stride][nreduce]
    // two-level deterministic reduction tree (all counters zero between launches)
    long long     nctas;        // tile CTAs per problem of this launch (<= part_stride)
    // Time-sharded runs launch ONE more CTA per problem (blockIdx.x == nctas) that has no tile: it waits
    // for the last group of the reduction tree, finalises and runs the cross-GPU exchange, so that no
    // tile CTA carries the NVLink round trip in front of its stores.
    int           finaliser;
    long long     part_stride;  // partial-sum slots per problem: partials [batch][part_stride][nreduce]
    long long     group_stride; // ceil(part_stride / kReduceGroup): gpartials / group_count rows per problem
    // Work items: item i covers samples [k0, k0 + size), size a multiple of
    // 8 * warps per CTA.  The items come in up to kMaxPhases phases of
    // decreasing size -- full tiles of CFEM_TILE samples first, then one
    // resident set of half tiles, then quarter tiles ... (cfem_host: graded
    // tail) -- so that the CTAs of the LAST resident set, whose slots are not
    // refilled, are short.  (Measured on B200: the extra CTAs cost more than
    // the shorter drain saves, so the host side builds a single phase unless
    // CFEM_TAIL_LEVELS asks for more -- profiles/r02_experiments.)
    // CTA c takes items c, c + nctas, ...
    long long     nitems;
    int           nphase;
    long long     ph_item0[kMaxPhases];     // first item of the phase
    long long     ph_k0[kMaxPhases];        // first sample of the phase
    int           ph_size[kMaxPhases];      // samples per item
    long long     ngroups;      // ceil(nctas / kReduceGroup)
    double*       gpartials;    // [batch][group_stride][nreduce]
    unsigned int* group_count;  // [batch][group_stride] CTAs retired per group
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
