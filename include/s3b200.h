/*
 * s3b200 -- C-ABI of the B200-native hot path of Sparse Spatial Sampling (S^3).
 *
 * The reference (JanisGeise/sparseSpatialSampling) has no FFI: its boundary is the Python class
 * surface.  This header is the boundary a maintainer binds with ctypes; every
 * entry point names the reference code it replaces (paths relative to the reference repository).
 *
 * Conventions
 *   - every function returns 0 on success, a negative S3_ERR_* code otherwise; the message of the
 *     last failure on the calling thread is returned by s3_last_error(); nothing throws;
 *   - all `d_*` pointers are DEVICE pointers to caller-owned, contiguous buffers; the library only
 *     owns the opaque handles it hands out (s3_knn_t, s3_geom_t);
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no hidden device sync
 *     unless stated;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef S3B200_H
#define S3B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S3_F32 0
#define S3_F64 1

typedef struct s3_knn s3_knn_t;
typedef struct s3_geom s3_geom_t;

/* ---- library ---------------------------------------------------------------------------------- */
const char* s3_last_error(void);
int s3_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
int64_t s3_launch_count(void);

/* ---- k-nearest-neighbour index over the original point cloud ---------------------------------
 * replaces sklearn KNeighborsRegressor / NearestNeighbors as used in
 *   sparseSpatialSampling/s_cube.py:161-163 (fit), :224, :328, :372 (predict)
 *   sparseSpatialSampling/export.py:120, :423-441 (fit, kneighbors, inverse-distance weights)      */

/* d_coords: fp64 [n, dim] row-major (dim = 2|3); d_values: fp64 [n] regression targets or NULL.
 * Synchronises `stream` once before returning. */
int s3_knn_build(const double* d_coords, int64_t n, int dim, const double* d_values, void* stream,
                 s3_knn_t** out);
int s3_knn_free(s3_knn_t* h);
/* kneighbors(): d_idx int64 [nq, k], d_dist fp64 [nq, k], ascending distance (ties: smaller index) */
int s3_knn_query(const s3_knn_t* h, const double* d_query, int64_t nq, int k, int64_t* d_idx,
                 double* d_dist, void* stream);
/* KNeighborsRegressor(weights="distance").predict(): d_pred fp64 [nq] */
int s3_knn_predict(const s3_knn_t* h, const double* d_query, int64_t nq, int k, double* d_pred,
                   void* stream);
/* ExportData._build_knn_cache (export.py:403-444): idx int32 [nq,k], normalised inverse-distance
 * weights as fp32 [nq,k] and (optional, may be NULL) fp64 [nq,k] */
int s3_knn_tables(const s3_knn_t* h, const double* d_query, int64_t nq, int k, int32_t* d_idx,
                  float* d_w32, double* d_w64, void* stream);

/* ---- export-stage interpolation ---------------------------------------------------------------
 * replaces interpolate_data (sparseSpatialSampling/export.py:446-468):
 *   out[c, :] = sum_j w[c, j] * data[idx[c, j], :],  data [n_src, row_len], out [n_cells, row_len]
 * dtypes: (data F32, out F32): w is fp32, fp32 FMA accumulation;
 *         (data F32|F64, out F64): w is fp64, products and sequential adds in fp64 (reference order).
 * d_out_row: optional int32 [n_cells] -- row of `out` that receives cell c (cells may be passed in
 * any processing order, e.g. Morton order); NULL = identity.                                       */
int s3_interp_gather(const void* d_data, int data_dtype, int64_t n_src, int64_t row_len,
                     const int32_t* d_idx, const void* d_w, int64_t n_cells, int k,
                     const int32_t* d_out_row, void* d_out, int out_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* S3B200_H */
