// mcb_walk_tally.cu — the walk kernels of scoring cycles: mcb_walk.cu compiled as flavour 1 (see the note there): one
// block of 512 threads per SM, warps in step (a barrier per track) so that they share the instruction stream of the
// estimator code.
#define MCB_WALK_FLAVOUR 1
#ifndef MCB_TALLY_BLOCK
#define MCB_TALLY_BLOCK 512
#endif
#ifndef MCB_TALLY_MINB
#define MCB_TALLY_MINB 1
#endif
#ifndef MCB_TALLY_SYNC
#define MCB_TALLY_SYNC 1
#endif
#define MCB_BLOCK MCB_TALLY_BLOCK
#define MCB_WALK_MINB MCB_TALLY_MINB
#if MCB_TALLY_SYNC
#define MCB_WALK_SYNC MCB_TALLY_SYNC
#endif
#include "mcb_walk.cu"
