/* mbnrg_b200.h -- C ABI of the MB-nrg ion-water / ion-ion 2-body path in libmbpol_b200.so.
 *
 * Drop-in boundary for the reference's mbnrg.f90 module (plain pointers and sizes, Fortran array conventions: 1-based atom
 * indices in the lists, forces/stress ACCUMULATED into the caller's arrays):
 *
 *   mbnrg_b200_init    replaces  mbnrg_init    (mbnrg.f90:40-157;  call site dlpoly_cmd.f:997)
 *   mbnrg_b200_forces  replaces  mbnrg_forces  (mbnrg.f90:170-485; call site dlpoly_cmd.f:2522-2524)
 *   mbnrg_b200_fini    replaces  mbnrg_fini    (mbnrg.f90:161-166)
 *
 * What is evaluated: for every ion j and every water i whose imaged ion-oxygen distance is <= 10 A (mbnrg.f90:130, :525) the
 * 2B ion-water polynomial of that ion (mbnrg-2b-h2o-<ion>-poly.cpp, 8 variables, 429 coefficients, own switch over the last 1 A below a per-ion cutoff of 6.0 - 7.0 A),
 * and for every ion pair j < k the 1-variable ion-ion polynomial (mbnrg-2b-<ion>-<ion>-poly.cpp) where the reference has
 * one.  The 3-body ion-water-water term, compiled out in the reference build (mbnrg.f90:4), is the separate call
 * mbnrg_b200_forces_3b below.  The reference
 * has no heat-current hooks in this module (no per-atom energies / virials), so neither has this interface.
 *
 * Ion codes (the `key` array of mbnrg_init): 1 f, 2 cl, 3 br, 4 i, 5 li, 6 na, 7 k, 8 rb, 9 cs (mbnrg.f90:541-565).
 * Like the reference, ion POSITIONS are read from the first nions atoms of xxx/yyy/zzz (mbnrg.f90:226-233); listions(j)
 * only names the force slot of ion j.
 *
 * Status codes are those of mbpol_b200.h.  No CPU fallback: without a CUDA device init fails with MBPOL_B200_ENODEVICE. */
#ifndef MBNRG_B200_H
#define MBNRG_B200_H
#include "mbpol_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mbnrg_b200_handle mbnrg_b200_handle;

/* nttm2 = 3 * number of waters; ion_code[nions] = code of every ion in list order; device < 0: current device */
int mbnrg_b200_init(int nttm2, int nions, const int* ion_code, int device, mbnrg_b200_handle** out);
int mbnrg_b200_fini(mbnrg_b200_handle* h);
const char* mbnrg_b200_last_error(const mbnrg_b200_handle* h);

/* this rank evaluates the waters [nw*rank/nranks, nw*(rank+1)/nranks) against all ions; ion-ion pairs on rank 0 only
 * (the reference lets every bead rank add them, mbnrg.f90:319-337, and over-counts after its gdsum when bead_size > 1) */
int mbnrg_b200_set_shard(mbnrg_b200_handle* h, int rank, int nranks);

/* mbnrg_forces: fxx/fyy/fzz[mxatms], *eng2, *vir2 and stress[9] are accumulated into (partial sums of this rank's shard) */
int mbnrg_b200_forces(mbnrg_b200_handle* h, int imcon, const double* cell, int nttm2, const int* listttm2, int nions,
                      const int* listions, int mxatms, const double* xxx, const double* yyy, const double* zzz, double* fxx,
                      double* fyy, double* fzz, double* eng2, double* vir2, double* stress);

/* counts[0] = water-ion pairs inside the 10 A list cutoff, [1] = pairs inside the evaluator's own cutoff (evaluated on the
 * GPU), [2] = ion-ion pairs evaluated; of the last mbnrg_b200_forces call */
int mbnrg_b200_counts(const mbnrg_b200_handle* h, long long counts[3]);
/* device time of the last call: ms[0] pair list, ms[1] ion-water kernel, ms[2] ion-ion kernel, ms[3] total; *launches */
int mbnrg_b200_kernel_times(const mbnrg_b200_handle* h, float ms[4], int* launches);

/* The 3-body ion-water-water term (mbnrg.f90:354-427 + do_trimer :781-993; mbnrg-3b-h2o-h2o-<ion>-poly.cpp, 21 variables, 1831
 * coefficients, degree 5).  The reference build has it compiled out (`!#define ENABLE_X3B`, mbnrg.f90:4), so it is a separate,
 * optional call with the argument conventions of mbnrg_b200_forces: for every ion k and every water pair i < j whose two imaged
 * oxygen-ion distances are both < 7 A (strict; mbnrg.f90:135, :840-842) water i and the ion are placed next to water j and the
 * evaluator of that ion is accumulated into fxx/fyy/fzz, *eng3, *vir3 and stress[9].  Rank r of nranks (mbnrg_b200_set_shard)
 * evaluates the contiguous range [T r/N, T (r+1)/N) of the accepted trimers. */
int mbnrg_b200_forces_3b(mbnrg_b200_handle* h, int imcon, const double* cell, int nttm2, const int* listttm2, int nions,
                         const int* listions, int mxatms, const double* xxx, const double* yyy, const double* zzz, double* fxx,
                         double* fyy, double* fzz, double* eng3, double* vir3, double* stress);
/* counts[0] = accepted trimers (all ranks), counts[1] = evaluated by this rank; ms (may be NULL): list, evaluation, total */
int mbnrg_b200_counts_3b(const mbnrg_b200_handle* h, long long counts[2], float ms[3]);
/* n clusters w[n][21] = (w1[9], w2[9], x[3]) -> E[n] (kcal/mol), g[n][21] */
int mbnrg_b200_eval_h2o_h2o_ion(mbnrg_b200_handle* h, int ion, int n, const double* w, double* E, double* g);

/* batch evaluators for tests: n clusters, w[n][9], x[n][3] -> E[n] (kcal/mol), g1[n][9], g2[n][3]; ion-ion: w[n][3] */
int mbnrg_b200_eval_h2o_ion(mbnrg_b200_handle* h, int ion, int n, const double* w, const double* x, double* E, double* g1, double* g2);
int mbnrg_b200_eval_ion_ion(mbnrg_b200_handle* h, int ion_a, int ion_b, int n, const double* w, const double* x, double* E, double* g1,
                            double* g2);

#ifdef __cplusplus
}
#endif
#endif
