/*
 * TEST INFRASTRUCTURE ONLY.
 *
 * CPU oracle for the multi-scale deformable attention hot path: a plain-C restatement of
 * the arithmetic of the reference CUDA kernels (file:line cited in msda_oracle_body.inc).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load the library built from this file.  The product (uvhand_b200/) never does.
 *
 * Parity pinning: tests/test_oracle.py checks this file against golden vectors produced by
 * the reference's own ms_deform_attn_core_pytorch (fp64, forward and autograd backward),
 * generated in the build container by oracle/make_golden.py and committed under
 * tests/golden/.
 *
 * Three arithmetic variants are emitted:
 *   _f32    storage float,  carried in float   -- the reference kernel as compiled (fp32)
 *   _f32x   storage float,  carried in double  -- fp32 pixel coordinates/indices (the part
 *                                                 that must be bit-exact), unrounded sums:
 *                                                 the score-card for the fp32/bf16 kernels
 *   _f64    storage double, carried in double  -- comparable with the fp64 grid_sample path
 *
 * Build:  gcc -O2 -fopenmp -fPIC -shared -ffp-contract=off oracle/msda_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define T float
#define ACC float
#define FMA_T fmaf
#define FLOOR_T floorf
#define SFX _f32
#include "msda_oracle_body.inc"
#undef T
#undef ACC
#undef FMA_T
#undef FLOOR_T
#undef SFX

#define T float
#define ACC double
#define FMA_T fmaf
#define FLOOR_T floorf
#define SFX _f32x
#include "msda_oracle_body.inc"
#undef T
#undef ACC
#undef FMA_T
#undef FLOOR_T
#undef SFX

#define T double
#define ACC double
#define FMA_T fma
#define FLOOR_T floor
#define SFX _f64
#include "msda_oracle_body.inc"
#undef T
#undef ACC
#undef FMA_T
#undef FLOOR_T
#undef SFX

int msda_oracle_abi_version(void) { return 1; }
