// Kernel dispatch record: one per specialised topology (see ptfnn_topologies.h).
#pragma once
struct PtfnnKernelSet {
    const char *name;
    int task, I, H, O, NT;
    const void *chain, *init, *fwd, *sgd;
};
