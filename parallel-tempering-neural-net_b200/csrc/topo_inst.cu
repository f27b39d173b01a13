// One translation unit per topology: instantiates the templated kernels of ptfnn_kernels.cuh for
// -DPTFNN_T_NAME / _TASK / _I / _H / _O / _NT / _MINB and exports their addresses.
#include "ptfnn_kernels.cuh"
#include "ptfnn_registry.h"

#define PTFNN_CAT2(a, b) a##b
#define PTFNN_CAT(a, b) PTFNN_CAT2(a, b)
#define PTFNN_STR2(a) #a
#define PTFNN_STR(a) PTFNN_STR2(a)

using namespace ptfnn;

// the chain kernel with speculative windows is only instantiated where it can be used (not for the team topologies)
template <bool TEAM>
struct SpecChain {
    static const void *get() { return (const void *)chain_kernel<PTFNN_T_I, PTFNN_T_H, PTFNN_T_O, PTFNN_T_TASK, PTFNN_T_NT, PTFNN_T_MINB, true>; }
};
template <>
struct SpecChain<true> {
    static const void *get() { return nullptr; }
};

const PtfnnKernelSet *PTFNN_CAT(ptfnn_kernelset_, PTFNN_T_NAME)() {
    constexpr bool kTc = UseTc<PTFNN_T_I, PTFNN_T_H, PTFNN_T_O, PTFNN_T_NT>::value;
    static const PtfnnKernelSet ks = {
        PTFNN_STR(PTFNN_T_NAME), PTFNN_T_TASK, PTFNN_T_I, PTFNN_T_H, PTFNN_T_O, PTFNN_T_NT,
        (const void *)chain_kernel<PTFNN_T_I, PTFNN_T_H, PTFNN_T_O, PTFNN_T_TASK, PTFNN_T_NT, PTFNN_T_MINB>,
        SpecChain<ptfnn::UseSgdTeam<PTFNN_T_H>::value>::get(),
        (const void *)init_kernel<PTFNN_T_I, PTFNN_T_H, PTFNN_T_O, PTFNN_T_TASK, PTFNN_T_NT>,
        (const void *)op_forward_kernel<PTFNN_T_I, PTFNN_T_H, PTFNN_T_O, PTFNN_T_TASK, PTFNN_T_NT>,
        (const void *)op_sgd_kernel<PTFNN_T_I, PTFNN_T_H, PTFNN_T_O, PTFNN_T_TASK, PTFNN_T_NT>,
        ptfnn::UseSgdTeam<PTFNN_T_H>::value ? ptfnn::kTeamThreads : 32,
        kTc ? (const void *)op_forward_tc_kernel<PTFNN_T_I, PTFNN_T_H, PTFNN_T_O, PTFNN_T_TASK, PTFNN_T_NT> : nullptr,
        kTc ? (const void *)tc::pack_a_kernel<PTFNN_T_I> : nullptr,
        tc::a_tile_floats(PTFNN_T_I), tc::Smem<PTFNN_T_I, PTFNN_T_H, PTFNN_T_O>::total, tc::Smem<PTFNN_T_I, PTFNN_T_H, PTFNN_T_O>::off_a,
        kTc ? tc::kTmemCols : 0,
    };
    return &ks;
}

#ifdef PTFNN_T_STANDALONE
// the same, with a fixed name, for specialisations built on demand into their own shared library
extern "C" const void *ptfnn_topology_kernels(void) { return PTFNN_CAT(ptfnn_kernelset_, PTFNN_T_NAME)(); }
extern "C" int ptfnn_topology_registry_version(void) { return PTFNN_REGISTRY_VERSION; }
#endif
