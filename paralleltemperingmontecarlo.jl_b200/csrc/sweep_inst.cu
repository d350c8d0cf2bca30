// sweep_inst.cu — compiled once per lanes-per-replica value (-DPTMC_G=4|8|16|32); exports the table of
// sweep_kernel instantiations for that G.
#include "sweep_kernel.cuh"

#ifndef PTMC_G
#error "compile with -DPTMC_G=<4|8|16|32>"
#endif
#define PTMC_CAT2(a, b) a##b
#define PTMC_CAT(a, b) PTMC_CAT2(a, b)

namespace ptmc {

// pot: ptmc_pot_kind (ELJ, ELJ_EVEN, EAM); geom: Geom; dbg: replay/trace build
sweep_fn PTMC_CAT(sweep_table_g, PTMC_G)(int pot, int geom, int dbg) {
    constexpr int G = PTMC_G;
#define PTMC_PICK(POT, GEOM)                                                             \
    if (pot == POT && geom == GEOM)                                                      \
        return dbg ? (sweep_fn)sweep_kernel<G, POT, GEOM, true> : (sweep_fn)sweep_kernel<G, POT, GEOM, false>;
    PTMC_PICK(PTMC_POT_ELJ_EVEN, GEOM_SPHERE)
    PTMC_PICK(PTMC_POT_ELJ_EVEN, GEOM_PBC)
    PTMC_PICK(PTMC_POT_ELJ, GEOM_SPHERE)
    PTMC_PICK(PTMC_POT_ELJ, GEOM_PBC)
    PTMC_PICK(PTMC_POT_EAM, GEOM_SPHERE)
#undef PTMC_PICK
    return nullptr;
}

}  // namespace ptmc
