/* lbfgs_ exists only in the reference's Fortran source (ndlfortran.f); no Fortran
 * compiler here. Only "-O quasinew" reaches it, which the oracle never uses. */
#include <stdio.h>
#include <stdlib.h>
void lbfgs_(void) { fprintf(stderr, "lbfgs_ unavailable in oracle/_ref\n"); abort(); }
