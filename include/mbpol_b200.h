/* mbpol_b200.h -- C ABI of libmbpol_b200.so: the B200 (sm_100a, FP64) drop-in for the MB-pol 2B+3B hot path
 * of ppegolo/dlpoly2-mbpol-heat.  Plain pointers and sizes only; every entry point returns an int status
 * (0 = ok) and never aborts -- the Fortran shim maps non-zero to mb_abort() (multibead.f90:111-123).
 *
 * Reference interfaces replaced (file:line relative to /root/reference/SOURCE):
 *
 *   mbpol_b200_init / _fini        mbpol::mbpol_init / mbpol_fini            mbpol.f90:52-94   (called dlpoly_cmd.f:989-991, 3648)
 *   mbpol_b200_cutoffs             mbpol_2b_cutoff_ / mbpol_3b_cutoff_       mbpol-2b-poly.cpp:1900-1904, mbpol-3b-poly.cpp:1538-1542
 *   mbpol_b200_forces              mbpol::mbpol_forces                       mbpol.f90:98-449  (called dlpoly_cmd.f:2504)
 *                                  + the heatcurrent side effects            heatcurrent.f90:221-231 (update_stress/energy_mbpol),
 *                                                                            :529-533 (update_forces), consumed by :464-483, :703-728
 *   mbpol_b200_eval_dimers/trimers batched form of the per-cluster ABI       mbpol-2b-poly.cpp:1888-1893, mbpol-3b-poly.cpp:1526-1531
 *   mbpol_2b_poly_, mbpol_3b_poly_, mbpol_2b_cutoff_, mbpol_3b_cutoff_       the legacy per-cluster symbols themselves (same
 *                                  signatures; one cluster per call, evaluated on the GPU -- for parity tests, not for speed)
 *
 * Units and conventions are the reference's: coordinates in Angstrom, per-cluster ABI in kcal/mol,
 * mbpol_b200_forces in DL_POLY internal units (x engunit = 418.4, mbpol.f90:37).  All Fortran-visible arrays
 * keep their Fortran layout: xxx/yyy/zzz/fxx/... are SoA over atoms (1-based listttm2 indices), per-atom
 * tensors are atomic_stress(mxatms,3,3) column-major, i.e. element (i,a,b) at  (i-1) + mxatms*((a-1) + 3*(b-1)).
 * Outputs of mbpol_b200_forces are ACCUMULATED into the caller's arrays exactly like the reference
 * (fxx -= g, engsrp += ..., stress(1:9) += ...); the caller owns every host array, the handle owns all
 * device memory, streams and events.  One handle = one GPU = one host thread at a time; calls are synchronous.
 */
#ifndef MBPOL_B200_H
#define MBPOL_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* flags for mbpol_b200_init */
#define MBPOL_B200_WANT_ATOMIC_ENERGY 1   /* -DHEAT_CURRENT: atomic_energy_mbpol                      */
#define MBPOL_B200_WANT_FDECOMP       2   /* -DHEAT_CURRENT: compressed force_matrix tensor W (below) */
#define MBPOL_B200_WANT_ATOMIC_STRESS 4   /* -DHEAT_STRESS : atomic_stress_mbpol                      */
#define MBPOL_B200_DEFAULT_FLAGS      3   /* the reference's default build (Makefile:33-43)           */

/* status codes */
#define MBPOL_B200_OK            0
#define MBPOL_B200_EINVAL        1   /* bad argument (imcon, sizes, NULL)             */
#define MBPOL_B200_ECUDA         2   /* CUDA runtime error, see mbpol_b200_last_error */
#define MBPOL_B200_ECAPACITY     3   /* neighbour capacity exceeded (unphysical density) */
#define MBPOL_B200_ENODEVICE     4   /* no usable sm_100 GPU                          */

typedef struct mbpol_b200_handle mbpol_b200_handle;

/* nttm2 = 3 * (number of waters), as passed to mbpol_init (mbpol.f90:52).  device < 0 = current device. */
int mbpol_b200_init(int nttm2, int device, int flags, mbpol_b200_handle** handle);
int mbpol_b200_fini(mbpol_b200_handle* handle);
const char* mbpol_b200_last_error(const mbpol_b200_handle* handle);
void mbpol_b200_cutoffs(double* r2f, double* r3f);

/* Bead-batched variant (PIMD/CMD, multibead.f90:25-93: every bead is an independent replica of the same topology with its
 * own coordinates): ONE list build and ONE pair of evaluation launches for all beads.  Arrays are [nbeads][mxatms]
 * (bead-major, each bead laid out like the single-system arrays); eng2/vir2/eng3/vir3 are [nbeads], stress and jflux
 * [nbeads][9]; cell and listttm2 are shared by all beads.  mbpol_b200_init == mbpol_b200_init_batched(nbeads = 1). */
int mbpol_b200_init_batched(int nttm2, int nbeads, int device, int flags, mbpol_b200_handle** handle);
int mbpol_b200_forces_batched(mbpol_b200_handle* handle, int nbeads, int imcon, const double* cell, int nttm2,
                              const int* listttm2, int mxatms, const double* xxx, const double* yyy, const double* zzz,
                              double* fxx, double* fyy, double* fzz,
                              double* eng2, double* vir2, double* eng3, double* vir3, double* stress,
                              double* atomic_energy, double* atomic_stress, double* atomic_fdecomp,
                              const double* vxx, const double* vyy, const double* vzz, double* jflux);

/* Work split in the spirit of `do t = bead_rank, n-1, bead_size` (mbpol.f90:218-226, 311-314): with (rank, nranks) != (0,1) the
 * handle BUILDS AND EVALUATES only its shard of the lists -- the dimers whose first water and the trimers whose centre water lie
 * in this rank's contiguous water range -- and the caller's existing gdsum over comm_bead (dlpoly_cmd.f:2551, 2568;
 * heatcurrent.f90:1129-1259) completes the sums.  Default (0,1); uniform water ranges. */
int mbpol_b200_set_shard(mbpol_b200_handle* handle, int rank, int nranks);

/* NCCL inside the boundary (the in-library analogue of those gdsum calls, over NVLink instead of PCIe + MPI).  One process per GPU:
 * rank 0 calls _comm_unique_id and ships the 128 bytes to the others by the host's own means (MPI_Bcast in a Fortran host);
 * every rank then calls _comm_init, which also fixes the handle's shard to (rank, nranks).  From then on every step ends with a
 * reduction on the handle's stream:
 *   mode 2 (default)  ncclReduceScatter of the per-atom part: rank r holds, downloads and ADDS only slice r of the packed buffer
 *                     (1/nranks of the PCIe traffic per rank); the caller's gdsum still completes the arrays, as for the reference.
 *   mode 1            ncclAllReduce: every rank holds everything; the per-atom arrays are returned by every rank (divide or skip the
 *                     caller's gdsum accordingly).
 *   mode 0            no reduction (shards only).
 * In modes 1 and 2 the scalars (eng2, vir2, eng3, vir3, stress, jflux) are all-reduced and returned by RANK 0 ONLY (zero from the
 * others), so that the caller's gdsum of them stays correct.  The shard boundaries follow the per-water work of the previous step
 * (weights = 4 x trimers + dimers, all-reduced with the results), so the ranks stay balanced as the liquid moves. */
#define MBPOL_B200_UNIQUE_ID_BYTES 128
int mbpol_b200_comm_unique_id(void* id /* MBPOL_B200_UNIQUE_ID_BYTES */);
int mbpol_b200_comm_init(mbpol_b200_handle* handle, int rank, int nranks, const void* id);
int mbpol_b200_comm_mode(mbpol_b200_handle* handle, int mode);
/* Host arithmetic behind the shard boundaries (every rank computes the same cut): bounds[r] .. bounds[r+1] = waters of rank r. */
void mbpol_b200_shard_bounds(int nw, int nranks, const int* weight /* nw entries or NULL = uniform */, int* bounds /* nranks + 1 */);
/* cudaHostRegister / cudaHostUnregister for the caller's coordinate and force arrays (pin them ONCE: the per-step copies then run
 * at full PCIe speed and asynchronously). */
int mbpol_b200_host_register(void* p, size_t bytes);
int mbpol_b200_host_unregister(void* p);

/* The drop-in for mbpol_forces.  Argument meaning as in mbpol.f90:98-116; wat_monomer_ener is unused by the
 * reference's live code path and therefore absent.  eng2/vir2/eng3/vir3 are the four scalars the reference
 * gdsum's (mbpol.f90:395-431): this call ADDS its shard's values to them; the shim then does
 * engsrp += eng2+eng3, virsrp += vir2+vir3 (mbpol.f90:434-435).
 *   atomic_energy  [mxatms]      += E_cluster*418.4/n per atom, n = 6 or 9             (may be NULL)
 *   atomic_stress  [mxatms*9]    += S_i(a,b) = 1/2*418.4*sum_j (r_i-r_j)_a FM[i][j]_b  (may be NULL; needs WANT_ATOMIC_STRESS)
 *   atomic_fdecomp [mxatms*9]    += W_i(a,b) = sum_{clusters, j} (x_i-x_j)_a^raw * (-g_i,b*418.4/n)   (may be NULL)
 *       W is the MB-pol slice of sum_j (x_i-x_j) (x) force_matrix(i,j,:) -- everything compute_heat_flux2 and
 *       compute_stress (heatcurrent.f90:703-779) need from the N x N x 3 force_matrix, which is never materialised.
 *   vxx,vyy,vzz (may be NULL) -> jflux[9] is SET to this shard's partial heat-flux contractions:
 *       jflux[0..2] = sum_i e_i v_i, jflux[3..5] = sum_i S_i^{upper-sym} v_i (compute_heat_flux, :464-483),
 *       jflux[6..8] = sum_i W_i v_i (the MB-pol part of jpot in compute_heat_flux2, :703-728).
 * Output arrays in page-locked memory (mbpol_b200_host_register, cudaHostAlloc) take a fused end: their current values are uploaded
 * while the step runs, added on the device and copied straight back (same "+=" result, no host pass); pageable arrays are staged
 * and added by host threads.  After a non-zero return the contents of the output arrays are unspecified (the reference aborts). */
int mbpol_b200_forces(mbpol_b200_handle* handle, int imcon, const double* cell, int nttm2, const int* listttm2,
                      int mxatms, const double* xxx, const double* yyy, const double* zzz,
                      double* fxx, double* fyy, double* fzz,
                      double* eng2, double* vir2, double* eng3, double* vir3, double* stress,
                      double* atomic_energy, double* atomic_stress, double* atomic_fdecomp,
                      const double* vxx, const double* vyy, const double* vzz, double* jflux);

/* Split form of the same call for callers that keep data on the device (bench `value`, NCCL reduction):
 *   _upload  : host -> device copy of coordinates (and velocities)
 *   _compute : list build + 2B + 3B + scatter (+ flux); results stay in the packed device buffer
 *   _download: device -> host copy and accumulation into the caller's arrays (same semantics as _forces)
 * mbpol_b200_device_outputs exposes the packed FP64 device buffer
 *   [ fxx | fyy | fzz | atomic_energy | atomic_stress(9) | atomic_fdecomp(9) ](each mxatms) [ scalars(32) ]
 *   scalars (32 per bead, at the offset mbpol_b200_shard reports): 0 eng2, 1 vir2, 2 eng3, 3 vir3, 4..12 stress, 13..21 jflux,
 *   22 / 23 dimers / trimers of the shard, rest reserved
 * so that one all-reduce(sum, double) over `count` elements finishes a multi-GPU step when the caller reduces itself. */
int mbpol_b200_upload(mbpol_b200_handle* handle, int imcon, const double* cell, int nttm2, const int* listttm2,
                      int mxatms, const double* xxx, const double* yyy, const double* zzz,
                      const double* vxx, const double* vyy, const double* vzz);
int mbpol_b200_compute(mbpol_b200_handle* handle);
int mbpol_b200_download(mbpol_b200_handle* handle, double* fxx, double* fyy, double* fzz,
                        double* eng2, double* vir2, double* eng3, double* vir3, double* stress,
                        double* atomic_energy, double* atomic_stress, double* atomic_fdecomp, double* jflux);
int mbpol_b200_device_outputs(mbpol_b200_handle* handle, void** dev_ptr, long long* count);
/* this rank's water range and the offset of the scalar block inside the packed buffer (= 22*mxatms rounded up to nranks equal,
 * even chunks: the unit of the reduce-scatter) */
int mbpol_b200_shard(mbpol_b200_handle* handle, int* first_water, int* n_waters, long long* scalar_offset);
void* mbpol_b200_stream(mbpol_b200_handle* handle);   /* the cudaStream_t every kernel of the handle runs on */

/* Introspection for tests / benchmarks (valid after _compute or _forces).  The *_total counts are box totals once a reduction
 * has run (in-library communicator) and equal the shard counts otherwise; _get_lists returns THIS RANK's shard of the lists. */
int mbpol_b200_counts(mbpol_b200_handle* handle, long long* n_dimers_total, long long* n_trimers_total,
                      long long* n_dimers_shard, long long* n_trimers_shard);
int mbpol_b200_get_lists(mbpol_b200_handle* handle, int* dimers /*2*n*/, long long dimer_cap,
                         int* trimers /*3*n*/, long long trimer_cap);   /* 0-based water indices, i<j(<k), full lists */
int mbpol_b200_kernel_times(mbpol_b200_handle* handle, float* ms /*[4]: lists, 2B, 3B, total*/, long long* launches);
/* finer split, n <= 11: 0 lists, 1 2B, 2 3B (incidence lists + k_trimers + gather), 3 whole step incl. reduction, 4 k_trimers alone
 * (the roofline kernel), 5 k_trimers_gather, 6 reduction over the ranks, 7 incidence lists (device time, CUDA events);
 * 8 upload, 9 compute, 10 download: host wall time of the three parts of the last mbpol_b200_forces call */
int mbpol_b200_phase_times(mbpol_b200_handle* handle, float* ms, int n);

/* FP64 roofline probe: best-of-3 throughput of independent register-resident DFMA chains on the current device
 * (2 flop per DFMA), the measured denominator for the "fraction of FP64 FMA peak" figures. */
int mbpol_b200_measure_fp64_peak(double* tflops, double* ms);

/* Batched per-cluster evaluation: n clusters, coordinates already imaged as the reference's do_dimer/do_trimer
 * pass them (mbpol.f90:505-517, 875-894).  w: [n][18] or [n][27] (O,H,H per water); outputs in kcal/mol:
 * E[n], g: [n][18|27], fm: [n][108|243] or NULL (diagonal blocks are written as 0; the reference leaves them
 * uninitialised, mbpol-2b-poly.cpp:454). */
int mbpol_b200_eval_dimers(int n, const double* w, double* E, double* g, double* fm);
int mbpol_b200_eval_trimers(int n, const double* w, double* E, double* g, double* fm);

/* The reference's own per-cluster symbols (Fortran-77 calling convention: everything by reference). */
void mbpol_2b_poly_(const double* w1, const double* w2, double* E, double* g1, double* g2, double* force_matrix);
void mbpol_3b_poly_(const double* w1, const double* w2, const double* w3, double* E, double* g1, double* g2,
                    double* g3, double* force_matrix);
void mbpol_2b_cutoff_(double* r);
void mbpol_3b_cutoff_(double* r);

#ifdef __cplusplus
}
#endif
#endif /* MBPOL_B200_H */
