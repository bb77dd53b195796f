/* b200msm_probes.h -- measurement hooks of libb200msm.so (NOT part of the drop-in boundary in b200msm.h).
 *
 * bench.py's roofline needs the integer-multiply peak of the GPU it runs on "measured on the box by a register-resident mad.wide
 * microbenchmark" (SURVEY.md 8d); these three functions are that microbenchmark and ship in the library.  The exploratory probes of
 * round 1 (FP64 pipe, dual-pipe overlap, radix-2^29 multiplier: DESIGN.md sections 4 and 7) are compiled only with -DB200_EXPERIMENTS
 * (B200_EXPERIMENTS=1 python __graft_entry__.py) and are declared at the bottom for that build.
 */
#ifndef B200MSM_PROBES_H
#define B200MSM_PROBES_H
#include "b200msm.h"
#ifdef __cplusplus
extern "C" {
#endif

/* ---- measurement hooks: integer-multiply roofline denominators and the field-multiply rate, measured on this GPU.
 * imad_wide_per_s: 32x32+64 -> 64 multiply-adds per second in the form the field multiplier uses (IMAD.WIDE.U32 carry
 *                  chains, register-resident, all SMs) -- the roofline peak for the accumulate phase;
 * imad32_per_s   : plain 32-bit IMAD per second, for context (twice the wide rate on B200);
 * fqmul_per_s    : dependent Montgomery multiplications per second for the given curve. */
int b200msm_probe_imad(b200msm_ctx* ctx, double* imad_wide_per_s);
int b200msm_probe_imad32(b200msm_ctx* ctx, double* imad32_per_s);
int b200msm_probe_fqmul(b200msm_ctx* ctx, int curve, double* fqmul_per_s);

#ifdef B200_EXPERIMENTS
/* dfma_per_s: FP64 fused multiply-adds per second (8 independent chains per thread) -- the second multiplier pipe of the SM, unused by
 * the integer path; measured to size an FP64-limb multiplier that would run beside the IMAD one. */
int b200msm_probe_dfma(b200msm_ctx* ctx, double* dfma_per_s);
/* Do the FP64 and the integer-multiply pipes overlap?  ms[0]: 256 BLS12-381 Fq multiplications per thread (IMAD.WIDE), ms[1]: 256 blocks of
 * 690 FP64 operations per thread (the instruction mix of a 48-bit-limb FP64 Montgomery multiplication), ms[2]: both in the same thread,
 * ms[3] = 256, ms[4]: odd warps do the integer work and even warps the FP64 work (half of each).  ms[2] ~ max(ms[0], ms[1]) would mean a
 * second multiplier can run beside the first (csrc/probes.cu). */
int b200msm_probe_dualpipe(b200msm_ctx* ctx, double ms[5]);
#endif

#ifdef __cplusplus
}
#endif
#endif /* B200MSM_PROBES_H */
