/*
 * similarity_transform.h -- C ABI of libsimilarity_transform.so (B200 / sm_100a build).
 *
 * Part 1 is the drop-in boundary: the two symbols the reference's
 * wrapper/similarity_transform.cpp exports and its Python wrapper binds
 * (wrapper/python/similarity_transform.py:19,35-37,66-76).  Same names, same argument
 * meaning, same return value; plain pointers and sizes only.
 *
 * Part 2 is additive (st_*): device-resident inputs, on-device input generation,
 * per-kernel entry points mirroring include/similarity_transform.hpp:55-100, the row-block
 * sharded multi-GPU solve (one process per GPU, or every GPU of the box behind one handle), and
 * the streamed solve of host / file-backed matrices larger than the device.  None of it changes
 * Part 1.
 *
 * Error convention (the reference has none: wrapper/similarity_transform.cpp never checks):
 * no exception ever crosses this boundary; st_* return 0 on success and a negative code on
 * failure, max_eigen_value returns a negative value on failure (the reference only ever
 * returns >= 0), and st_last_error() describes the most recent failure on this thread.
 */
#ifndef SIMILARITY_TRANSFORM_H
#define SIMILARITY_TRANSFORM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef ST_UINT_DEFINED
#define ST_UINT_DEFINED
typedef unsigned int st_uint; /* the reference spells it `uint` */
#endif

/* ------------------------------------------------------------------------------------------
 * Part 1 -- drop-in boundary
 * ---------------------------------------------------------------------------------------- */

/* Replaces make_queue()  (reference wrapper/similarity_transform.cpp:3-12).
 * Writes an opaque heap-allocated solver handle bound to CUDA device 0 (the "default
 * device") to *wq; leaves *wq == NULL when no usable GPU exists, which the reference's
 * Python wrapper already treats as failure (similarity_transform.py:39-40).  Like the
 * reference there is no matching destroy call in Part 1 (st_destroy exists in Part 2). */
void make_queue(void** wq);

/* Replaces max_eigen_value()  (reference wrapper/similarity_transform.cpp:14-37), which
 * forwards to similarity_transform() (reference similarity_transform.cpp:5-75).
 *   wq         handle from make_queue
 *   mat        HOST pointer, dim x dim fp32 row-major, not modified
 *   eigen_val  HOST float[1]   <- s[0] of the last row-sum vector      (:60-65)
 *   eigen_vec  HOST float[dim] <- raw accumulated eigenvector           (:42-43)
 *   dim        matrix dimension, any dim >= 1
 *   iter_cnt   HOST, exactly 4 bytes written: the break index, or 1000 if the stop test
 *              never held                                                (:54)
 * Returns the round loop's duration in whole milliseconds (:36,:56-58), measured on the
 * device; negative on error.  EPS = 1e-3f, MAX_ITR = 1000 (include/similarity_transform.hpp:4-5). */
int64_t max_eigen_value(void* wq, float* mat, float* eigen_val, float* eigen_vec, st_uint dim,
                        st_uint* iter_cnt);

/* ------------------------------------------------------------------------------------------
 * Part 2 -- additive extensions
 * ---------------------------------------------------------------------------------------- */

#define ST_EPS 1e-3f     /* include/similarity_transform.hpp:4 */
#define ST_MAX_ITR 1000u /* include/similarity_transform.hpp:5 */

enum
{
  ST_OK = 0,
  ST_ERR_CUDA = -1,      /* a CUDA runtime call failed; see st_last_error() */
  ST_ERR_ARG = -2,       /* bad argument */
  ST_ERR_NO_DEVICE = -3, /* no usable GPU */
  ST_ERR_TIMEOUT = -4,   /* a device-side barrier timed out (peer rank missing) */
  ST_ERR_NOMEM = -5
};

enum
{
  ST_FORM_READONLY = 0, /* s = (A.e)/e, A read once per round, never written (default) */
  ST_FORM_INPLACE = 1   /* literal W <- D^-1 W D on a working copy (reference :52)       */
};

enum
{
  ST_STOP_ABSOLUTE = 0, /* every circular adjacent pair |s[r] - s[r+1]| < eps: the reference's
                           test (similarity_transform.cpp:413-421), the default                */
  ST_STOP_RELATIVE = 1  /* max_r |s[r] - s[r+1]| < eps * max(0, max_r s[r]): scale-free extension.
                           The absolute test cannot hold once one ulp of lambda exceeds eps
                           (e.g. uniform (0,1] matrices from N = 16384 up run the full max_iter
                           rounds); the relative test stops those after a handful of rounds.
                           Changes WHEN the loop stops, never the arithmetic of a round.        */
};

enum
{
  ST_ACC_F32 = 0, /* row sums accumulated in fp32, like the reference (default)                    */
  ST_ACC_F64 = 1  /* opt-in: the same evaluation order with fp64 accumulators, each row sum rounded
                     to fp32 once at the end (fp32 storage, read-only form; kernels 0, 1, 10, 12, 13).
                     Lowers the rounding noise of the row sums; changes results within the tolerance. */
};

typedef struct st_options
{
  float eps;         /* stop threshold, reference EPS                               */
  uint32_t max_iter; /* round cap, reference MAX_ITR; 1 .. 2^24                     */
  int32_t form;      /* ST_FORM_*                                                   */
  int32_t sweep;     /* bit 0: alternate the row order every round so the tail of one pass is
                        the L2-resident head of the next (default on); bit 1 / bit 2: force
                        static / dynamic work-unit scheduling in the unit-scheduled kernels; bit 3:
                        no in-kernel time stamps                                                 */
  int32_t threads;   /* CTA size of the round kernel, 0 = default                   */
  int32_t ctas;      /* grid size of the round kernel, 0 = one per SM               */
  int32_t kernel;    /* 0 = automatic: on-chip cluster kernel (N <= 512, dim % 4 == 0, one GPU),
                        resident-e kernel (N <= 32768; dim % 4 != 0, bf16 and fp8 storage on
                        configuration 11), wide kernel above (general chunked loop for what the wide
                        kernel is not built for); 1 = general loop, 2 = wide kernel, 10-13 = resident-e
                        configurations (prefetch slot of 2 / 0 / 3 / 1 x 4 KB per warp), 20 = on-chip
                        cluster kernel; any other value is refused (the TMA-ring, 256- / 1024-thread
                        and L2-prefetch variants of round 1 lost on hardware and were removed)         */
  int32_t l2_keep_pct; /* 0..100: share of each CTA's rows loaded with an L2 evict_last policy
                          (the rest evict_first) so that part of A stays L2-resident across
                          rounds; 0 = no cache hints                                       */
  int32_t stop;        /* ST_STOP_*                                                               */
  int32_t accumulate;  /* ST_ACC_*                                                                */
} st_options;

typedef struct st_result
{
  float eigen_val;        /* s[0] of the last pass                                  */
  uint32_t iter_count;    /* break index (reference iter_count)                     */
  uint32_t passes;        /* row passes executed = min(iter_count + 1, max_iter)    */
  uint32_t launches;      /* kernels launched by this call                          */
  float loop_ms;          /* round loop, CUDA events on the solver stream           */
  float total_ms;         /* whole call incl. host<->device copies (host clock)     */
  float round_us_median;  /* per-round time from in-kernel globaltimer stamps       */
  float round_us_min;
  uint64_t bytes_per_round; /* algorithmic bytes one round moves on this GPU        */
  int32_t status;
  uint32_t grid;          /* CTAs the round kernel ran with                          */
  uint32_t kernel_id;     /* 1 general loop, 2-9 TMA ring, 10-19 and 21-26 resident-e, 20 on-chip,
                             30 streamed (host-driven rounds) */
  uint32_t threads;       /* CTA size the round kernel ran with                      */
} st_result;

const char* st_last_error(void);
int st_device_count(void);
void st_default_options(st_options* opt);

/* Solver context bound to one CUDA device (what make_queue creates for device 0). */
int st_create(int device, void** ctx);
void st_destroy(void* ctx);
int st_device_info(void* ctx, int* sm_count, size_t* l2_bytes, size_t* hbm_bytes, char* name,
                   size_t name_len);

/* Device memory owned by the context's device. */
int st_malloc(void* ctx, size_t bytes, void** dptr);
int st_free(void* ctx, void* dptr);
int st_memcpy_h2d(void* ctx, void* dptr, const void* hptr, size_t bytes);
int st_memcpy_d2h(void* ctx, void* hptr, const void* dptr, size_t bytes);
int st_synchronize(void* ctx);

/* Page-lock a caller-owned host buffer (cudaHostRegister) so that max_eigen_value / st_solve_host /
 * st_memcpy_* move it at full PCIe rate: a pageable 8192^2 matrix is staged by the driver at a fraction
 * of the ~55 GB/s a pinned one reaches.  Pin once, solve many times, unpin before freeing the buffer. */
int st_pin_host(void* ctx, void* hptr, size_t bytes);
int st_unpin_host(void* ctx, void* hptr);
/* Matrices the caller cannot pin (the reference's wrapper passes a plain numpy array): pageable host matrices of
 * 32 MiB and up are copied by T host threads through pinned double buffers (4 MiB chunks; the threads bind to the
 * CPUs next to the GPU) instead of through the driver's single staging path -- in max_eigen_value, st_solve_host,
 * st_memcpy_h2d and the block uploads of st_solve_streamed / st_solve_file, where the threads also spread the page
 * faults of a file mapping.  T = 4 by default (measured, Hilbert 8192 end to end: driver staging 24.6 ms, T = 4
 * 7.6 ms, pinned 5.6 ms); ST_UPLOAD_THREADS=T in the environment when the context is created overrides it (0..16;
 * 0 = the driver's staging).  Pinned / registered sources keep the direct copy.  st_staged_upload_bytes reports how
 * many bytes took the threaded path on this context. */
uint64_t st_staged_upload_bytes(void* ctx);

/* Input generation on the device, rows [row0, row0+rows) of the dim x dim matrix written
 * to d_rows (rows x dim, row-major).
 * Hilbert: reference utils.cpp:136-154, A[r][c] = 1.f / (float)(r + c + 1).
 * Uniform: replaces the reference's unseeded host fill utils.cpp:124-134 by a seeded
 * Philox-4x32-10 uniform (0,1] fill that does not depend on the sharding. */
int st_generate_hilbert(void* ctx, float* d_rows, uint32_t dim, uint32_t row0, uint32_t rows);
int st_generate_uniform(void* ctx, float* d_rows, uint32_t dim, uint32_t row0, uint32_t rows,
                        uint64_t seed);

/* similarity_transform() on a matrix already resident in device memory; d_eigen_vec is a
 * device float[dim].  Same semantics as max_eigen_value otherwise. */
int st_solve_device(void* ctx, const float* d_mat, uint32_t dim, const st_options* opt,
                    float* d_eigen_vec, st_result* res);
/* Same with host buffers (what max_eigen_value calls with default options). */
int st_solve_host(void* ctx, const float* h_mat, uint32_t dim, const st_options* opt,
                  float* h_eigen_val, float* h_eigen_vec, st_result* res);
/* bf16 STORAGE of the matrix (opt-in; changes results, so outside reference parity): the matrix is
 * held as bfloat16 -- half the HBM bytes per round -- while the eigenvector, the row sums and every
 * accumulation stay fp32.  bf16 -> fp32 is exact, so the result is bit-identical to an fp32 solve of
 * the bf16-rounded matrix in the fp32 kernels' own evaluation order.  Read-only form only; dim % 4 == 0; kernels 0
 * (automatic), 1 and 11.  st_convert_f32_to_bf16 rounds to nearest even on the device; `count`
 * elements, d_dst 2-byte elements. */
int st_convert_f32_to_bf16(void* ctx, const float* d_src, uint16_t* d_dst, size_t count);
int st_solve_device_bf16(void* ctx, const uint16_t* d_mat, uint32_t dim, const st_options* opt,
                         float* d_eigen_vec, st_result* res);
/* fp8 STORAGE of the matrix (opt-in; changes results, so outside reference parity): one byte per element (e4m3)
 * plus ONE power-of-two fp32 scale per row, A[r][c] ~= d_row_scale[r] * q[r][c] -- a quarter of the HBM bytes per
 * round -- while the eigenvector, the row sums and every accumulation stay fp32.  st_convert_f32_to_fp8 picks each
 * row's scale so that its largest magnitude lands in (224, 448] and rounds to the nearest code, ties to even.
 * e4m3 -> fp32 is exact and the scales are powers of two, so the result is bit-identical to an fp32 solve of the
 * dequantised matrix in the fp32 kernels' own evaluation order.  Read-only form, fp32 accumulation; dim % 4 == 0; kernels 0
 * (automatic), 1 and 11.  Entries below 2^-18 of their row's largest one round to zero. */
int st_convert_f32_to_fp8(void* ctx, const float* d_src, uint8_t* d_dst, float* d_row_scale, uint32_t rows,
                          uint32_t dim);
int st_solve_device_fp8(void* ctx, const uint8_t* d_mat, const float* d_row_scale, uint32_t dim,
                        const st_options* opt, float* d_eigen_vec, st_result* res);
/* Streamed solve: the same round loop for a HOST matrix that does not fit the device (or the share of
 * it the caller grants).  The step before the path in the reference is the host copy-in of the whole
 * matrix (similarity_transform.cpp:14-19); here the device holds a direct-mapped cache of `slots` row
 * blocks (block b lives in slot b % slots) and every round sweeps the blocks in alternating direction,
 * so the blocks a round ends on are the ones the next round starts on: per round only
 * (blocks - slots) blocks cross PCIe, the cached ones are read from HBM.  Copies run on their own
 * stream ahead of the row passes; the round loop is host-driven (one 8-byte read-back per round),
 * which costs microseconds against rounds that take milliseconds to seconds.  Same arithmetic and
 * evaluation order as every other solve, so the result is bit-identical to st_solve_host; read-only
 * form, fp32 accumulation, both stop tests.
 *   h_mat           dim x dim fp32 row-major, pageable, pinned (st_pin_host: full PCIe rate) or a
 *                   read-only file mapping; never modified
 *   device_budget   bytes of device memory to use for the block cache; 0 = what is free now minus 1 GiB,
 *                   and then a matrix that fits is simply handed to st_solve_host
 *   block_rows      rows per block; 0 = about 64 MiB worth of rows
 *   plan            optional: what was done (blocks, slots, bytes over PCIe in round 0 and per later round) */
typedef struct st_stream_plan
{
  uint32_t block_rows;          /* rows per block (the last block may be shorter)          */
  uint32_t blocks;              /* row blocks of the matrix                                */
  uint32_t slots;               /* blocks the device cache holds                           */
  uint32_t streamed;            /* 0: the matrix fitted and st_solve_host ran instead      */
  uint64_t cache_bytes;         /* device memory taken by the cache                        */
  uint64_t h2d_bytes_first;     /* bytes copied host -> device in round 0                  */
  uint64_t h2d_bytes_per_round; /* bytes copied per later round                            */
  uint64_t h2d_bytes_total;     /* bytes copied by this call                               */
} st_stream_plan;
int st_solve_streamed(void* ctx, const float* h_mat, uint32_t dim, const st_options* opt, size_t device_budget,
                      uint32_t block_rows, float* h_eigen_val, float* h_eigen_vec, st_result* res,
                      st_stream_plan* plan);
/* The same on a file: `path` holds dim x dim fp32 row-major starting `offset` bytes into the file (a raw
 * dump, or a .npy file with offset = its header length).  The file is mapped read-only and streamed;
 * nothing but the cache and the vectors is ever resident. */
int st_solve_file(void* ctx, const char* path, uint64_t offset, uint32_t dim, const st_options* opt,
                  size_t device_budget, uint32_t block_rows, float* h_eigen_val, float* h_eigen_vec,
                  st_result* res, st_stream_plan* plan);

/* Device group: several GPUs behind ONE handle, so that the reference's unmodified wrapper uses the whole
 * box.  st_group_attach binds helper contexts on `devices` (NULL / 0 = every other visible GPU; the list
 * must not name the context's own device) to `ctx`.  From then on max_eigen_value / st_solve_host on `ctx`
 * run the row-block sharded solve for matrices with dim >= min_dim (0 = 8192): one host thread per GPU copies
 * its own row block from the caller's host matrix -- every GPU's PCIe link carries 1/G of the matrix -- and
 * joins the collective round kernel; the shards are linked by peer access (st_shard_link_local).  Same bits
 * as on one GPU.  Smaller matrices stay on the context's own GPU.  st_group_detach (also done by
 * st_destroy) frees the helpers.  Setting ST_DEVICES=all or ST_DEVICES=0,1,2,3 in the environment makes
 * make_queue attach the group itself: the handle lives on the first device listed (ST_GROUP_MIN_DIM
 * optionally overrides min_dim). */
int st_group_attach(void* ctx, const int* devices, uint32_t count, uint32_t min_dim);
int st_group_detach(void* ctx);
int st_group_size(void* ctx); /* devices behind this handle, 1 without a group */

/* Per-round device timestamps (ns, globaltimer) of the last solve on this context. */
int st_round_timestamps(void* ctx, uint64_t* out, uint32_t capacity, uint32_t* count);
/* Three stamps per round of the last solve, taken by CTA 0: matrix pass done, round barrier
 * (incl. the cross-GPU exchange) passed, vector tail done -- the per-phase split the
 * reference's per-kernel benchmarks give (benchmarks/similarity_transform.md). */
int st_phase_timestamps(void* ctx, uint64_t* out, uint32_t capacity, uint32_t* count);

/* CUDA-event stopwatch on the context's stream: st_timer_start records an event, st_timer_stop
 * records a second one, waits for it and returns the milliseconds between the two.  For timing the
 * per-kernel entry points below the way the reference's per-kernel benchmarks do
 * (benchmarks/benchmark_similarity_transform.cpp:24-433, main.cpp:37-159). */
int st_timer_start(void* ctx);
int st_timer_stop(void* ctx, float* ms);

/* Per-kernel entry points on device buffers, one per reference L1 function
 * (include/similarity_transform.hpp:55-100; similarity_transform.cpp:77-460). */
int st_sum_across_rows(void* ctx, const float* d_mat, float* d_vec, uint32_t dim);      /* :77-152  */
/* Unfused building block of the collective variant of the sharded loop: one read-only
 * round's row pass on rows [row0,row0+rows): d_vec[row0+r] = (sum_c A[r][c] e[c]) / e[row0+r]. */
int st_row_pass_readonly(void* ctx, const float* d_rows, const float* d_e, float* d_vec,
                         uint32_t dim, uint32_t row0, uint32_t rows);
int st_find_max(void* ctx, const float* d_vec, float* d_max, uint32_t dim);             /* :154-227 */
int st_compute_eigen_vector(void* ctx, const float* d_vec, const float* d_max,
                            float* d_eigen_vec, uint32_t dim);                          /* :229-265 */
int st_initialise_eigen_vector(void* ctx, float* d_eigen_vec, uint32_t dim);            /* :267-284 */
int st_compute_next_matrix(void* ctx, float* d_mat, const float* d_vec, uint32_t dim);  /* :286-330 */
int st_stop(void* ctx, const float* d_vec, uint32_t* d_ret, uint32_t dim, float eps);   /* :332-460 */

/* Row-block sharded solve: one context (one process) per GPU, rank g owns rows
 * [dim*g/world, dim*(g+1)/world).  The per-round exchange of the row-sum slices is done by
 * the round kernel itself with stores into peer memory and a flag barrier over NVLink.
 *   st_shard_create  allocates this rank's exchange block
 *   st_shard_export  writes its 64-byte CUDA IPC handle (ship it to the peers with any
 *                    host-side transport, e.g. torch.distributed all_gather)
 *   st_shard_import  opens the world x 64-byte handle table, in rank order
 *   st_shard_solve   collective: every rank calls it with its own rows.  If a rank does not show up, the others
 *                    return ST_ERR_TIMEOUT after the device-side timeout (10 s) instead of hanging; the group's
 *                    solve counters have then diverged, so destroy the shards of every rank and create new ones. */
#define ST_IPC_HANDLE_BYTES 64
#define ST_MAX_WORLD 8
int st_shard_create(void* ctx, uint32_t dim, uint32_t rank, uint32_t world, void** shard);
int st_shard_export(void* shard, void* handle_out);
int st_shard_import(void* shard, const void* handles);
/* Same-process alternative to export/import (one host thread per GPU): links `world` shards
 * created in this process, in rank order, through plain CUDA peer access. */
int st_shard_link_local(void** shards, uint32_t world);
/* Reserves every device allocation st_shard_solve needs with these options (NULL = defaults).  A sharded solve
 * never allocates: a cudaMalloc / cudaFree on a device with peer mappings may wait for peers that already spin in
 * the collective kernel.  st_shard_create prepares for the default options; before the first solve with options
 * that need more (the in-place form, max_iter above ST_MAX_ITR) call this on EVERY rank and synchronise the ranks
 * on the host.  st_shard_solve returns ST_ERR_ARG when it would have to allocate. */
int st_shard_prepare(void* shard, const st_options* opt);
int st_shard_rows(void* shard, uint32_t* row0, uint32_t* rows);
int st_shard_solve(void* shard, const float* d_rows, const st_options* opt, float* d_eigen_vec,
                   st_result* res);
/* st_shard_solve on bf16 storage of this rank's rows (see st_solve_device_bf16). */
int st_shard_solve_bf16(void* shard, const uint16_t* d_rows, const st_options* opt, float* d_eigen_vec,
                        st_result* res);
/* st_shard_solve on fp8 storage of this rank's rows and their row scales (see st_solve_device_fp8). */
int st_shard_solve_fp8(void* shard, const uint8_t* d_rows, const float* d_row_scale, const st_options* opt,
                       float* d_eigen_vec, st_result* res);
void st_shard_destroy(void* shard);

#ifdef __cplusplus
}
#endif
#endif /* SIMILARITY_TRANSFORM_H */
