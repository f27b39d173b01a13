// The topologies libptfnn.so is specialised for.  One X(...) row = one translation unit
// (topo_inst.cu compiled with -DPTFNN_T_*), so rows build in parallel; __graft_entry__.build()
// parses this file.  X(name, task, I, H, O, threads_per_block, min_blocks_per_sm)
//   task 0 = regression (R:), 1 = classification (C:)
//   threads_per_block / min_blocks_per_sm size the chain kernel: one warp runs the serial SGD
//   recurrence, the others the row-parallel likelihood; min_blocks bounds registers so that the
//   intended number of temperatures is co-resident per SM.
#pragma once
#define PTFNN_TOPOLOGIES(X)                                                                          \
    X(reg_4_5_1, 0, 4, 5, 1, 128, 4)      /* Data_OneStepAhead suite, paper / drafts (hidden = 5) */  \
    X(reg_4_10_1, 0, 4, 10, 1, 128, 4)    /* checked-in driver, R:915 (hidden = 10) */                \
    X(reg_4_64_1, 0, 4, 64, 1, 128, 7)    /* synthetic time series, BASELINE configs[3]: 1024 temperatures = 7 per SM */ \
    X(cls_4_12_3, 1, 4, 12, 3, 128, 4)    /* Iris, C:920-930 */                                       \
    X(cls_9_12_2, 1, 9, 12, 2, 128, 2)    /* Cancer, C:950-957 */                                     \
    X(cls_34_50_2, 1, 34, 50, 2, 128, 2)  /* Ionosphere, C:942-949 */                                 \
    X(cls_16_30_10, 1, 16, 30, 10, 128, 2) /* PenDigit, C:972-986 */                                  \
    X(cls_16_256_10, 1, 16, 256, 10, 160, 2) /* PenDigit-shaped synthetic, BASELINE configs[4] */
