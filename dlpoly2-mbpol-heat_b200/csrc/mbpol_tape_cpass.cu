// mbpol_tape_cpass.cu -- engine 2 row loops as stand-alone device functions (own ptxas unit).
//
// One constant-window pass of a polynomial for ONE cluster per thread:  rec[0..NV-1] = x (in), rec[NV] = E (in/out),
// rec[NV+1..2NV] = g (in/out); `sm` = the CTA's TapeCSmem (per-thread columns).  Deliberately NOT inlined into the big
// kernels and with the pass bounds read from fixed constant addresses: only then does ptxas keep the coefficient loads
// on the uniform datapath (LDCU + UR operands), which is the whole point of this engine.
#include "mbpol_tape.cuh"

template <int WHICH, int NV>
__device__ __forceinline__ void cpass_body(double* rec, TapeCSmem<NV>* sm)
{
    const TapeCPassHeader* hdr = reinterpret_cast<const TapeCPassHeader*>(mbpol_cwin);
    double x[NV], g[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) {
        x[c] = rec[c];
        g[c] = rec[NV + 1 + c];
    }
    double E = rec[NV];
    tape_eval_const_pass<WHICH, NV>(hdr->sg0, hdr->sg1, hdr->row0, hdr->row_last, x, g, E, *sm);
#pragma unroll
    for (int c = 0; c < NV; ++c) rec[NV + 1 + c] = g[c];
    rec[NV] = E;
}

extern "C" __device__ __noinline__ void mbpol_tape3b_cpass(double* rec, TapeCSmem<36>* sm) { cpass_body<3, 36>(rec, sm); }
extern "C" __device__ __noinline__ void mbpol_tape2b_cpass(double* rec, TapeCSmem<31>* sm) { cpass_body<2, 31>(rec, sm); }
