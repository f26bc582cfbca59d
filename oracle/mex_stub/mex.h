/* Minimal stand-in for MATLAB/Octave's mex.h so that the reference solver core compiles outside MATLAB.
 * TEST INFRASTRUCTURE (oracle/): the reference sources include "mex.h" only for mxIsNaN/mxIsInf/mxGetInf
 * (iLQG_problem.tem:6,11-12; iLQG.c:16-19; back_pass.c:15-18).  Compile with -DHAVE_OCTAVE so matrix.h is
 * not requested. */
#ifndef ORACLE_MEX_STUB_H
#define ORACLE_MEX_STUB_H
#include <math.h>
#include <stdio.h>
#define mxIsNaN(v) isnan(v)
#define mxIsInf(v) isinf(v)
#define mxGetInf() ((double)INFINITY)
#endif
