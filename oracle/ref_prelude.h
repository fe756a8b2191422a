/* ref_prelude.h — force-included (g++ -include) in front of the two reference sources that
 * reach "integratorSolverSelection.h".  Test infrastructure only; contains no reference code.
 * Parsing the include/ copy of implicitNewmarkSparse.h first makes the quoted include resolve to
 * vegafem/include/integratorSolverSelection.h:40 (#define PCG); the copy beside the sources
 * (vegafem/integrator/integratorSolverSelection.h:38) selects PARDISO and needs mkl.h.
 * The integrator/ copy of integratorBaseSparse.h goes first because only it declares
 * setConstrainedDOF, which integratorBaseSparse.cpp defines. */
#include "../integrator/integratorBaseSparse.h"
#include "implicitNewmarkSparse.h"
