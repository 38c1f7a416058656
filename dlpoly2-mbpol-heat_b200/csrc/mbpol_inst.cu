// mbpol_inst.cu -- one (kernel family, engine) instantiation per object file: compile with -DMBPOL_INST_WHICH=2|3 (dimers /
// trimers) and -DMBPOL_INST_ENGINE=1|3|4.  See mbpol_cluster.cuh.
#ifndef MBPOL_BS2
#define MBPOL_BS2 256
#endif
#ifndef MBPOL_BS3
#define MBPOL_BS3 256
#endif
#if MBPOL_INST_WHICH == 2
#define TAPE_BS MBPOL_BS2
#else
#define TAPE_BS MBPOL_BS3
#endif
#include "mbpol_cluster.cuh"

namespace mbpol_k {
#if MBPOL_INST_WHICH == 2
template __global__ void k_dimers<MBPOL_INST_ENGINE>(const long long*, const int2*, int, const double*, CellInfo, const double*, double*);
template __global__ void k_eval_dimers<MBPOL_INST_ENGINE>(int, const double*, double*, double*, double*, const double*);
#else
template __global__ void k_trimers<MBPOL_INST_ENGINE>(const long long*, const int*, int, const double*, CellInfo, const double*, double*);
template __global__ void k_eval_trimers<MBPOL_INST_ENGINE>(int, const double*, double*, double*, double*, const double*);
#endif
}
