/* Minimal stand-in for libf2c's header, needed only to compile the reference's
 * shipped f2c translation (ndlfortran.c) when building oracle/_ref.
 * TEST INFRASTRUCTURE ONLY - never linked into the product library. */
#ifndef GPC_ORACLE_F2C_SHIM_H
#define GPC_ORACLE_F2C_SHIM_H
#include <math.h>
typedef int integer;
typedef double doublereal;
typedef float real;
typedef int logical;
#define TRUE_ 1
#define FALSE_ 0
#ifndef abs
#define abs(x) ((x) >= 0 ? (x) : -(x))
#endif
static inline double d_int(doublereal *x) { return (*x > 0) ? floor(*x) : -floor(-*x); }
#endif
