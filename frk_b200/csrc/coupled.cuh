// Internal interface of coupled.cu (observations coupled through shared BAUs: non-diagonal Vfs); used by model.cu.
#pragma once
#include "common.cuh"

#define FRK_COUPLED_NCMAX 64

struct frk_coupling;
struct frk_rowbuckets;

int frk_coupling_create(frk_ctx *ctx, int64_t m, int r, int p, const int *h_vrowptr, const int *h_vcol, const double *h_vval,
                        const int *d_rowptr, const int *d_col, frk_coupling **out);
int frk_coupling_destroy(frk_coupling *C);
int frk_coupling_ncomp(const frk_coupling *C);
int frk_coupling_whiten(frk_coupling *C, frk_ctx *ctx, double sigma2fs, const double *d_val, const double *d_Z, const double *d_X,
                        const double *d_Ve, double *d_logdet_slot);
int frk_coupling_cross(frk_coupling *C, frk_ctx *ctx, const int *d_rowptr, const int *d_col, const double *d_val, const double *d_M);
int frk_coupling_msums(frk_coupling *C, frk_ctx *ctx, const double *d_X, const double *d_Z, const double *d_t, const double *d_Ve,
                       double sigma2fs, double *h_out);
int frk_coupling_case2(frk_coupling *C, frk_ctx *ctx, double sigma2fs, const double *d_Ve, int64_t nbau, const int *d_cz_rowptr,
                       const int *d_cz_members, const double *d_cz_weights, const int *d_rowptr, const int *d_col, const double *d_val,
                       const int *d_rowptr0, const int *d_col0, const double *d_val0, const double *d_Sig, const double *d_res,
                       const double *d_smuS, const double *d_q, const double *d_fs, double *d_mu, double *d_var, double *d_sd);
const int *frk_coupling_rowptr(const frk_coupling *C);
const int *frk_coupling_col(const frk_coupling *C);
const double *frk_coupling_val(const frk_coupling *C);
const double *frk_coupling_Z(const frk_coupling *C);
const double *frk_coupling_X(const frk_coupling *C);
const double *frk_coupling_ones(const frk_coupling *C);
const double *frk_coupling_zeros(const frk_coupling *C);
const frk_rowbuckets *frk_coupling_buckets(const frk_coupling *C);

// bucket.cu: the values of the matrix have changed (same pattern): refresh the bucket-ordered copy
int frk_rowbuckets_refresh(frk_rowbuckets *B, frk_ctx *ctx, const int *d_rowptr, const double *d_val);
