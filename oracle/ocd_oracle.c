/*
 * ocd_oracle.c -- CPU oracle for the L4DC-MPC-OCD hot path (TEST INFRASTRUCTURE ONLY;
 * see ocd_oracle.h for scope, parity status and the import rule).
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "ocd_oracle.h"

#define REAL float
#define SUF f32
#define REAL_IS_FLOAT 1
#include "ocd_oracle_impl.inc"
#undef REAL
#undef SUF
#undef REAL_IS_FLOAT

#define REAL double
#define SUF f64
#define REAL_IS_FLOAT 0
#include "ocd_oracle_impl.inc"
#undef REAL
#undef SUF
#undef REAL_IS_FLOAT

int ocdo_num_starts(const ocdo_params *p) { return p->extra_inits ? 6 : 3; }

int ocdo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
