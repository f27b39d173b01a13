// Kernel dispatch record: one per specialised topology (see ptfnn_topologies.h).
#pragma once
#define PTFNN_REGISTRY_VERSION 4      /* layout of PtfnnKernelSet; checked by ptfnn_register_kernels */
struct PtfnnKernelSet {
    const char *name;
    int task, I, H, O, NT;
    const void *chain;
    const void *chain_spec;   // chain kernel with speculative windows (0 for the wide-hidden topologies)
    const void *init, *fwd, *sgd;
    int sgd_threads;   // 32, or NT when the wide-hidden team variant is used
    const void *fwd_tc;   // K5: tcgen05 forward / likelihood of wide-hidden nets (0 = not applicable)
    const void *pack_a;   // data set -> UMMA A tiles for fwd_tc and the chain kernel
    int a_tile_floats, tc_smem_bytes, tc_alias_off;
    int tmem_cols;        // TMEM columns one CTA of the chain kernel allocates (0 = none)
};
