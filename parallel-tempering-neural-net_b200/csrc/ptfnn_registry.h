// Kernel dispatch record: one per specialised topology (see ptfnn_topologies.h).
#pragma once
struct PtfnnKernelSet {
    const char *name;
    int task, I, H, O, NT;
    const void *chain, *init, *fwd, *sgd;
    int sgd_threads;   // 32, or NT when the wide-hidden team variant is used
};
