// tree.cu -- the two throughput-critical kernels of the batch-affine tree (k_tree_fwd, k_tree_bwd; accumulate.cuh) in a translation unit of
// their own, one object per curve (-DTREE_CURVE=0..3), compiled WITHOUT nvcc's -split-compile.
//
// Why: -split-compile partitions a module for parallel optimisation, and the partitioning shifts with any edit anywhere in the module.  The
// textually identical k_tree_bwd came out of three builds of b200msm.cu as three different SASS listings, 1.528 / 1.586 / 1.530 ms for tree
// round 0 of a 2^20-point BLS12-381 MSM (profiles/README.md, r2late) -- 4 % on the kernel that is half of the MSM.  Here the kernels' code
// generation depends on their own sources only.  b200msm.cu declares these instantiations `extern template` and launches them.
#include "accumulate.cuh"

namespace b200 {
#if TREE_CURVE == 0
using TC = BLS12_381;
#elif TREE_CURVE == 1
using TC = BN254;
#elif TREE_CURVE == 2
using TC = Fq2<BLS12_381>;
#else
using TC = Fq2<BN254>;
#endif
B200_TREE_INSTANTIATE(template, TC)
}  // namespace b200
