/*
 * sdf_oracle.c -- CPU oracle for the differentiable SDF depth renderer.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / `--impl reference` legs of bench.py may load liboracle.so;
 * the product path (sdfest_b200/) never does and fails loudly without its
 * CUDA library.
 *
 * Parity status: PINNED against outputs of the reference itself -- the golden
 * vectors under tests/golden/ are produced by importing the reference's
 * simple_renderer.py in the build container (tests/golden/make_golden.py) and
 * the f64 instantiation below reproduces them to ~1e-12; the f32 instantiation
 * mirrors the arithmetic of the reference CUDA kernels (operation order,
 * double/float mixing, finite slab bounds) and is additionally compared with the
 * compiled reference extension (oracle/_ref) on the GPU box.
 *
 * Two instantiations of sdf_oracle_impl.h:
 *   *_f32 : REAL=float,  slab bounds -1e-10/1e10  (sdf_renderer_cuda.cu:164-165)
 *   *_f64 : REAL=double, slab bounds -inf/+inf    (simple_renderer.py:85-86)
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC sdf_oracle.c -o liboracle.so -lm
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL float
#define SUF(name) CAT(name, _f32)
#define ORACLE_FINITE_BOUNDS 1
#include "sdf_oracle_impl.h"
#undef REAL
#undef SUF
#undef ORACLE_FINITE_BOUNDS

#define REAL double
#define SUF(name) CAT(name, _f64)
#define ORACLE_FINITE_BOUNDS 0
#include "sdf_oracle_impl.h"
#undef REAL
#undef SUF
#undef ORACLE_FINITE_BOUNDS

int oracle_abi_version(void) { return 1; }
