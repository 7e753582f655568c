// dcp_resident_lean.cu — instantiates the LEAN (branch-free, fast-math) variants of the persistent
// kernel.  Built WITH FMA contraction (-fmad=true): unlike the EXACT policy, the lean step carries no
// operation-order contract (its results are checked against the oracle at 1e-9, not bit for bit).
#define DCP_RES_LEAN_TU 1
#include "dcp_resident.cu"
