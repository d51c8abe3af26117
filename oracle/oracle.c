/*
 * oracle.c -- CPU restatement of itzmeanjan/eigen_value's similarity_transform() path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under eigen_value_b200/ may import, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and there only as the checker / the timed CPU baseline.
 *
 * Parity status: PINNED, two ways.
 * (1) Against the reference's own known answers (tests/test_oracle_golden.py):
 *   - 3x3 golden eigenpair            (reference tests/test.cpp:84-102, main.py:52-58)
 *   - Hilbert round counts 9..17      (reference README.md:70-76, all six devices)
 *   - per-kernel unit fixtures         (reference tests/test.cpp:22-73, utils.cpp:5-122)
 *   - A.v ~= lambda.v acceptance test  (reference wrapper/python/test.py:15-16)
 * (2) Against outputs of the reference ITSELF run here (tests/test_reference_golden.py): the
 *   unmodified reference sources are compiled against a single-threaded CPU SYCL shim
 *   (oracle/sycl_shim, `make -C oracle ref` -> oracle/_ref/libreference_cpu.so); with the
 *   reference's own summation order (ORACLE_SUM_WORKGROUP) this file reproduces lambda, the
 *   eigenvector and iter_count BIT FOR BIT on every case, and the committed fixture
 *   tests/golden/reference_sycl.json carries those outputs to boxes without the reference tree.
 * The reference's real SYCL toolchain (dpcpp) is not available in this image.
 * (3) ORACLE_SUM_CUDA additionally restates the summation order of the CUDA round kernels, so
 *   that the GPU can be held to the oracle bit for bit (tests/test_zz_gpu_bitexact.py); that
 *   order is pinned on the CPU by tests/test_oracle_cuda_order.py (an independent numpy
 *   restatement, and eigenvalues recorded on B200s: tests/golden/gpu_recorded.json).
 *
 * Every function cites the reference lines it restates (paths relative to the reference
 * repository root).  Arithmetic is strict fp32: build with -ffp-contract=off and without
 * -ffast-math (oracle/Makefile does).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_EPS 1e-3f    /* include/similarity_transform.hpp:4 */
#define ORACLE_MAX_ITR 1000 /* include/similarity_transform.hpp:5 */

enum
{
  ORACLE_FORM_INPLACE = 0, /* literal: W <- D^-1 W D every round  (similarity_transform.cpp:52) */
  ORACLE_FORM_READONLY = 1 /* algebraically equal: s = (A.e)/e, A never written              */
};

enum
{
  ORACLE_SUM_SEQUENTIAL = 0, /* left-to-right fp32 */
  ORACLE_SUM_LANES16 = 1,    /* 16 strided partial sums, combined pairwise (vectorisable) */
  ORACLE_SUM_SUBGROUP32 = 2, /* 32-wide group sums added left to right: the shape of the
                                reference's reduce_over_group + atomic adds
                                (similarity_transform.cpp:119-146), one of its legal orders */
  ORACLE_SUM_WORKGROUP = 3,  /* | (wg_size << 8): the reference's literal two-level order for a
                                given work-group size -- per work-group of wg_size columns the
                                32-lane butterfly sums are added into local memory in sub-group
                                order (:119-132), the work-group totals into the row's global
                                cell in work-group order (:139-146).  This is the order the
                                reference executes on oracle/sycl_shim, so the oracle can be
                                compared with the real sources bit for bit. */
  ORACLE_SUM_CUDA = 4,       /* the evaluation order of the CUDA round kernels (lane / accumulator /
                                fold / shuffle tree, FMA): with ORACLE_FORM_READONLY the oracle
                                then matches the GPU bit for bit (row_dot_cuda_order below) */
  ORACLE_SUM_CUDA_F64 = 6    /* ORACLE_SUM_CUDA's order with fp64 accumulators (ST_ACC_F64): products are
                                exact in double, every add rounds once in double, the chunk sum is
                                rounded to fp32 once; chunk sums of a row are added in fp32 */
  /* bf16 and fp8 STORAGE of the matrix need no mode of their own: their kernels reduce 4-element words (one
     64-bit / one 32-bit load) in ORACLE_SUM_CUDA's order, and bf16 -> fp32 / e4m3 -> fp32 are exact: feed the
     oracle the rounded (dequantised) matrix as fp32 and the bits must match the GPU's solve. */
};

int
oracle_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void
oracle_set_threads(int n)
{
#ifdef _OPENMP
  if (n > 0)
    omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static double
now_ms(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec * 1e3 + (double)ts.tv_nsec * 1e-6;
}

/* ---- one row reduction, three legal fp32 orders -------------------------------------- */

static inline float
reduce16(const float* acc)
{
  float a8[8], a4[4];
  for (int i = 0; i < 8; i++)
    a8[i] = acc[i] + acc[i + 8];
  for (int i = 0; i < 4; i++)
    a4[i] = a8[i] + a8[i + 4];
  return (a4[0] + a4[2]) + (a4[1] + a4[3]);
}

/* The evaluation order of the CUDA round kernels (eigen_value_b200/csrc: row_dot_readonly,
 * chunk_dot_prefetched, the cluster kernel), restated so that the GPU can be compared with the
 * oracle BIT FOR BIT, not only within a tolerance:
 *   - the row is cut into 8192-column chunks; chunk sums are added left to right;
 *   - inside a chunk the unit is a float4 when n % 4 == 0 (a single float otherwise); unit j
 *     belongs to lane j % 32 and accumulator (j / 32) % 8 of that lane; a float4 is folded into
 *     its accumulator with four sequential FMAs (x, y, z, w);
 *   - per lane the 8 accumulators are folded pairwise (u += u+4, u += u+2, u += u+1), then the
 *     32 lanes by an xor-shuffle tree (offsets 16, 8, 4, 2, 1).
 * scale == NULL behaves like a vector of ones (fmaf(a, 1, acc) == a + acc exactly). */
#define ORACLE_CUDA_CHUNK 8192
static float
row_dot_cuda_order(const float* row, const float* scale, size_t n)
{
  const size_t vec = (n % 4 == 0) ? 4 : 1;
  float total = 0.f;
  for (size_t c0 = 0; c0 < n; c0 += ORACLE_CUDA_CHUNK) {
    const size_t clen = n - c0 < ORACLE_CUDA_CHUNK ? n - c0 : ORACLE_CUDA_CHUNK;
    const size_t nv = clen / vec;
    float acc[32][8];
    for (int l = 0; l < 32; l++)
      for (int u = 0; u < 8; u++)
        acc[l][u] = 0.f;
    for (size_t j = 0; j < nv; j++) {
      const int l = (int)(j % 32), u = (int)((j / 32) % 8);
      for (size_t k = 0; k < vec; k++) {
        const size_t c = c0 + j * vec + k;
        acc[l][u] = fmaf(row[c], scale ? scale[c] : 1.f, acc[l][u]);
      }
    }
    float lane[32];
    for (int l = 0; l < 32; l++) {
      for (int s = 4; s >= 1; s >>= 1)
        for (int u = 0; u < s; u++)
          acc[l][u] += acc[l][u + s];
      lane[l] = acc[l][0];
    }
    for (int o = 16; o >= 1; o >>= 1) {
      float next[32];
      for (int l = 0; l < 32; l++)
        next[l] = lane[l] + lane[l ^ o];
      for (int l = 0; l < 32; l++)
        lane[l] = next[l];
    }
    total = c0 == 0 ? lane[0] : total + lane[0];
  }
  return total;
}

/* ORACLE_SUM_CUDA_F64: row_dot_cuda_order with double accumulators (fp32 storage: float4 / scalar units) */
static float
row_dot_cuda_order_f64(const float* row, const float* scale, size_t n)
{
  const size_t vec = (n % 4 == 0) ? 4 : 1;
  float total = 0.f;
  for (size_t c0 = 0; c0 < n; c0 += ORACLE_CUDA_CHUNK) {
    const size_t clen = n - c0 < ORACLE_CUDA_CHUNK ? n - c0 : ORACLE_CUDA_CHUNK;
    const size_t nv = clen / vec;
    double acc[32][8];
    for (int l = 0; l < 32; l++)
      for (int u = 0; u < 8; u++)
        acc[l][u] = 0.0;
    for (size_t j = 0; j < nv; j++) {
      const int l = (int)(j % 32), u = (int)((j / 32) % 8);
      for (size_t k = 0; k < vec; k++) {
        const size_t c = c0 + j * vec + k;
        acc[l][u] = fma((double)row[c], scale ? (double)scale[c] : 1.0, acc[l][u]);
      }
    }
    double lane[32];
    for (int l = 0; l < 32; l++) {
      for (int s = 4; s >= 1; s >>= 1)
        for (int u = 0; u < s; u++)
          acc[l][u] += acc[l][u + s];
      lane[l] = acc[l][0];
    }
    for (int o = 16; o >= 1; o >>= 1) {
      double next[32];
      for (int l = 0; l < 32; l++)
        next[l] = lane[l] + lane[l ^ o];
      for (int l = 0; l < 32; l++)
        lane[l] = next[l];
    }
    const float chunk_sum = (float)lane[0];
    total = c0 == 0 ? chunk_sum : total + chunk_sum;
  }
  return total;
}

/* sum_c row[c] * (scale ? scale[c] : 1)   -- scale == NULL is the plain row sum */
static float
row_dot(const float* row, const float* scale, size_t n, int sum_mode)
{
  if (sum_mode == ORACLE_SUM_SEQUENTIAL) {
    float acc = 0.f;
    if (scale)
      for (size_t c = 0; c < n; c++)
        acc += row[c] * scale[c];
    else
      for (size_t c = 0; c < n; c++)
        acc += row[c];
    return acc;
  }
  if (sum_mode == ORACLE_SUM_CUDA)
    return row_dot_cuda_order(row, scale, n);
  if (sum_mode == ORACLE_SUM_CUDA_F64)
    return row_dot_cuda_order_f64(row, scale, n);
  if ((sum_mode & 0xff) == ORACLE_SUM_WORKGROUP) {
    const size_t wg = (size_t)(sum_mode >> 8);
    float cell = 0.f; /* the zero-filled global cell (:85-93) */
    for (size_t g0 = 0; g0 < n; g0 += wg) {
      float lds = 0.f; /* work-group leader resets local memory (:108-110) */
      for (size_t s0 = g0; s0 < g0 + wg && s0 < n; s0 += 32) {
        float v[32];
        for (int l = 0; l < 32; l++) {
          size_t c = s0 + (size_t)l;
          v[l] = (c < g0 + wg && c < n) ? (scale ? row[c] * scale[c] : row[c]) : 0.f;
        }
        for (int w = 16; w >= 1; w >>= 1)
          for (int l = 0; l < w; l++)
            v[l] = v[l] + v[l + w];
        lds += v[0];
      }
      cell += lds;
    }
    return cell;
  }
  if (sum_mode == ORACLE_SUM_SUBGROUP32) {
    float acc = 0.f;
    for (size_t c0 = 0; c0 < n; c0 += 32) {
      /* butterfly order of a 32-lane reduce */
      float v[32];
      for (int l = 0; l < 32; l++) {
        size_t c = c0 + (size_t)l;
        v[l] = c < n ? (scale ? row[c] * scale[c] : row[c]) : 0.f;
      }
      for (int w = 16; w >= 1; w >>= 1)
        for (int l = 0; l < w; l++)
          v[l] = v[l] + v[l + w];
      acc += v[0];
    }
    return acc;
  }
  float acc[16];
  for (int i = 0; i < 16; i++)
    acc[i] = 0.f;
  size_t nb = n & ~(size_t)15;
  if (scale) {
    for (size_t c = 0; c < nb; c += 16)
      for (int i = 0; i < 16; i++)
        acc[i] += row[c + i] * scale[c + i];
    for (size_t c = nb; c < n; c++)
      acc[c - nb] += row[c] * scale[c];
  } else {
    for (size_t c = 0; c < nb; c += 16)
      for (int i = 0; i < 16; i++)
        acc[i] += row[c + i];
    for (size_t c = nb; c < n; c++)
      acc[c - nb] += row[c];
  }
  return reduce16(acc);
}

/* ---- the six device kernels of the reference, one function each ----------------------- */

/* sum_across_rows(): s[r] = sum_c W[r][c]        similarity_transform.cpp:77-152 (:119-146) */
void
oracle_sum_across_rows(const float* mat, float* vec, uint32_t dim, int sum_mode)
{
  const size_t n = dim;
#pragma omp parallel for schedule(static)
  for (size_t r = 0; r < n; r++)
    vec[r] = row_dot(mat + r * n, NULL, n, sum_mode);
}

/* find_max(): m = max(0, max_r s[r])             similarity_transform.cpp:154-227 (:169 zero
 * fill, :195-221 max reductions).  Exact (max is order independent). */
float
oracle_find_max(const float* vec, uint32_t dim)
{
  float m = 0.f;
  for (uint32_t r = 0; r < dim; r++)
    m = vec[r] > m ? vec[r] : m;
  return m;
}

/* compute_eigen_vector(): e[r] *= s[r] / m       similarity_transform.cpp:229-265 (:260) */
void
oracle_compute_eigen_vector(const float* vec, float max, float* eigen_vec, uint32_t dim)
{
  for (uint32_t r = 0; r < dim; r++)
    eigen_vec[r] *= (vec[r] / max);
}

/* initialise_eigen_vector(): e[r] = 1            similarity_transform.cpp:267-284 (:280) */
void
oracle_initialise_eigen_vector(float* eigen_vec, uint32_t dim)
{
  for (uint32_t r = 0; r < dim; r++)
    eigen_vec[r] = 1.f;
}

/* stop(): 1 iff |s[r] - s[(r+1) % dim]| < eps for every r, wrap pair included, strict <
 *                                                 similarity_transform.cpp:332-460 (:413-421) */
uint32_t
oracle_stop(const float* vec, uint32_t dim, float eps)
{
  uint32_t ok = 1;
  for (uint32_t r = 0; r < dim; r++) {
    float self = vec[r];
    float next = vec[(r + 1) % dim];
    float diff = fabsf(self - next);
    if (!(diff < eps))
      ok = 0;
  }
  return ok;
}

/* Relative stop test -- an EXTENSION, not reference behaviour (the CUDA side is
 * ST_STOP_RELATIVE / st::kStopRelative): 1 iff |s[r] - s[(r+1) % dim]| < eps * m for every r, with
 * m = max(0, max_r s[r]) the value find_max() returns.  Same strict <, same wrap pair, same NaN
 * behaviour (any NaN difference fails) as stop(); only the threshold scales with the row sums. */
uint32_t
oracle_stop_relative(const float* vec, uint32_t dim, float eps, float m)
{
  const float thr = eps * m;
  uint32_t ok = 1;
  for (uint32_t r = 0; r < dim; r++) {
    float diff = fabsf(vec[r] - vec[(r + 1) % dim]);
    if (!(diff < thr))
      ok = 0;
  }
  return ok;
}

/* compute_next_matrix(): W[r][c] *= (1.f / s[r]) * s[c]
 *                                                 similarity_transform.cpp:286-330 (:324-325) */
void
oracle_compute_next_matrix(float* mat, const float* vec, uint32_t dim)
{
  const size_t n = dim;
#pragma omp parallel for schedule(static)
  for (size_t r = 0; r < n; r++) {
    float* row = mat + r * n;
    const float inv = 1.f / vec[r];
    for (size_t c = 0; c < n; c++)
      row[c] *= inv * vec[c];
  }
}

/* ---- input generation ----------------------------------------------------------------- */

/* generate_hilbert_matrix(): A[r][c] = 1.f / (float)(r + c + 1)      utils.cpp:136-154 (:150)
 * Writes rows [row0, row0+rows) of the dim x dim matrix into out (rows x dim, row-major). */
void
oracle_generate_hilbert(float* out, uint32_t dim, uint32_t row0, uint32_t rows)
{
  const size_t n = dim;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < (size_t)rows; i++) {
    size_t r = (size_t)row0 + i;
    for (size_t c = 0; c < n; c++)
      out[i * n + c] = 1.f / (float)(r + c + 1);
  }
}

/* Philox-4x32-10 (Salmon et al., SC'11), the published algorithm; counter = (idx/4, 0),
 * key = seed.  Replaces the reference's unseeded host mt19937 uniform [0,1) fill
 * (utils.cpp:124-134) with a seeded, shard-independent uniform (0,1] fill:
 * u = ((x >> 8) + 1) * 2^-24, element idx takes word idx % 4 of block idx / 4. */
static inline void
philox4x32_10(uint64_t ctr, uint64_t key, uint32_t out[4])
{
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0u, c3 = 0u;
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
  for (int i = 0; i < 10; i++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

void
oracle_philox_block(uint64_t ctr, uint64_t key, uint32_t* out4)
{
  philox4x32_10(ctr, key, out4);
}

void
oracle_generate_uniform(float* out, uint32_t dim, uint32_t row0, uint32_t rows, uint64_t seed)
{
  const size_t n = dim;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < (size_t)rows; i++) {
    size_t r = (size_t)row0 + i;
    uint64_t cached_blk = ~(uint64_t)0;
    uint32_t w[4] = { 0, 0, 0, 0 };
    for (size_t c = 0; c < n; c++) {
      uint64_t idx = (uint64_t)r * n + c;
      uint64_t blk = idx >> 2;
      if (blk != cached_blk) {
        philox4x32_10(blk, seed, w);
        cached_blk = blk;
      }
      out[i * n + c] = (float)((w[idx & 3] >> 8) + 1u) * 5.9604644775390625e-08f; /* 2^-24 */
    }
  }
}

/* ---- the host round loop ---------------------------------------------------------------- */

/* Read-only row pass: s[r] = (sum_c A[r][c] * e[c]) / e[r].  D^-1 A D telescopes, so the row
 * sums of the reference's working matrix in round k are exactly this expression with e the
 * eigenvector accumulated before this round's update (similarity_transform.cpp:40,42,52). */
static void
readonly_row_pass(const float* mat, const float* e, float* s, uint32_t dim, uint32_t row0,
                  uint32_t rows, int sum_mode)
{
  const size_t n = dim;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < (size_t)rows; i++) {
    size_t r = (size_t)row0 + i;
    float t = row_dot(mat + i * n, e, n, sum_mode);
    s[r] = t / e[r];
  }
}

/*
 * similarity_transform()                         similarity_transform.cpp:5-75
 *
 *   mat        row-major dim x dim fp32, caller-owned, never modified (:14,:19 copy it)
 *   eigen_val  <- s[0] of the last row-sum vector computed (:60-65)
 *   eigen_vec  <- raw accumulated e, NOT normalised (:42-43, buffer write-back at :67)
 *   iter_count <- break index i; max_itr if the stop test never held (:54)
 *   returns      floor(loop wall time) in ms (:36,:56-58,:74); *elapsed_ms gets it unrounded
 *
 * form / sum_mode / ranks choose among arithmetically legal evaluation orders; ranks > 1
 * executes the row-block sharded algorithm (rank g owns rows [g*dim/ranks, (g+1)*dim/ranks)),
 * each "rank" writing only its slice of s before the gather point.
 */
enum
{
  ORACLE_STOP_ABSOLUTE = 0, /* the reference's test, stop()            similarity_transform.cpp:44 */
  ORACLE_STOP_RELATIVE = 1  /* oracle_stop_relative(): extension, threshold eps * max(s)          */
};

int64_t
oracle_similarity_transform_ex2(const float* mat, float* eigen_val, float* eigen_vec, uint32_t dim,
                                uint32_t* iter_count, float eps, uint32_t max_itr, int form,
                                int sum_mode, uint32_t ranks, int stop_mode, double* elapsed_ms)
{
  const size_t n = dim;
  if (dim == 0 || ranks == 0)
    return -1;
  float* work = NULL;
  float* sum_vec = (float*)malloc(sizeof(float) * n);
  if (!sum_vec)
    return -1;
  if (form == ORACLE_FORM_INPLACE) {
    work = (float*)malloc(sizeof(float) * n * n); /* :14 */
    if (!work) {
      free(sum_vec);
      return -1;
    }
    memcpy(work, mat, sizeof(float) * n * n); /* :19 */
  }

  oracle_initialise_eigen_vector(eigen_vec, dim); /* :34 */

  double t0 = now_ms(); /* :36 */
  uint32_t i = 0;
  for (; i < max_itr; i++) { /* :39 */
    if (form == ORACLE_FORM_INPLACE) {
      for (uint32_t g = 0; g < ranks; g++) { /* :40 */
        uint32_t r0 = (uint32_t)((uint64_t)dim * g / ranks);
        uint32_t r1 = (uint32_t)((uint64_t)dim * (g + 1) / ranks);
        const size_t rows = r1 - r0;
#pragma omp parallel for schedule(static)
        for (size_t k = 0; k < rows; k++)
          sum_vec[r0 + k] = row_dot(work + (r0 + k) * n, NULL, n, sum_mode);
      }
    } else {
      for (uint32_t g = 0; g < ranks; g++) {
        uint32_t r0 = (uint32_t)((uint64_t)dim * g / ranks);
        uint32_t r1 = (uint32_t)((uint64_t)dim * (g + 1) / ranks);
        readonly_row_pass(mat + (size_t)r0 * n, eigen_vec, sum_vec, dim, r0, r1 - r0, sum_mode);
      }
    }
    /* ---- gather point: every rank now holds the full s ---- */
    float m = oracle_find_max(sum_vec, dim);                  /* :41 */
    oracle_compute_eigen_vector(sum_vec, m, eigen_vec, dim);  /* :42-43 */
    const uint32_t converged = stop_mode == ORACLE_STOP_RELATIVE
                                 ? oracle_stop_relative(sum_vec, dim, eps, m)
                                 : oracle_stop(sum_vec, dim, eps);
    if (converged == 1) /* :44-50 */
      break;
    if (form == ORACLE_FORM_INPLACE)
      oracle_compute_next_matrix(work, sum_vec, dim); /* :52 */
  }
  double t1 = now_ms(); /* :56 */

  *iter_count = i;          /* :54 */
  *eigen_val = sum_vec[0];  /* :60-65 */
  if (elapsed_ms)
    *elapsed_ms = t1 - t0;
  free(sum_vec);
  free(work);
  return (int64_t)(t1 - t0); /* :57-58 duration_cast<milliseconds> truncates */
}

/*
 * The same loop on a GENERATED matrix that is never materialised (read-only form only): every row is produced
 * into a per-thread buffer by the generator of SURVEY 8(d) -- kind 0: Hilbert (utils.cpp:150), kind 1: the
 * seeded Philox uniform (0,1] fill -- and reduced right away.  Bit for bit what oracle_similarity_transform_ex2
 * returns on the stored matrix (tests/test_oracle_cuda_order.py), but it needs O(N) memory, so the sizes no host
 * can hold (N = 65536: 16 GiB, N = 131072: 64 GiB -- BASELINE configs 3 to 5) get CPU-computed expected bits too.
 */
int64_t
oracle_similarity_transform_generated(int kind, uint64_t seed, float* eigen_val, float* eigen_vec, uint32_t dim,
                                      uint32_t* iter_count, float eps, uint32_t max_itr, int sum_mode,
                                      int stop_mode, double* elapsed_ms)
{
  const size_t n = dim;
  if (dim == 0 || (kind != 0 && kind != 1))
    return -1;
  float* sum_vec = (float*)malloc(sizeof(float) * n);
  if (!sum_vec)
    return -1;
  oracle_initialise_eigen_vector(eigen_vec, dim); /* :34 */
  double t0 = now_ms();
  uint32_t i = 0;
  int failed = 0;
  for (; i < max_itr && !failed; i++) {
#pragma omp parallel
    {
      float* row = (float*)malloc(sizeof(float) * n);
      if (!row) {
#pragma omp atomic write
        failed = 1;
      } else {
#pragma omp for schedule(dynamic, 16)
        for (size_t r = 0; r < n; r++) {
          if (kind == 0)
            oracle_generate_hilbert(row, dim, (uint32_t)r, 1);
          else
            oracle_generate_uniform(row, dim, (uint32_t)r, 1, seed);
          sum_vec[r] = row_dot(row, eigen_vec, n, sum_mode) / eigen_vec[r];
        }
        free(row);
      }
    }
    if (failed)
      break;
    float m = oracle_find_max(sum_vec, dim);                 /* :41 */
    oracle_compute_eigen_vector(sum_vec, m, eigen_vec, dim); /* :42-43 */
    const uint32_t converged = stop_mode == ORACLE_STOP_RELATIVE ? oracle_stop_relative(sum_vec, dim, eps, m)
                                                                 : oracle_stop(sum_vec, dim, eps);
    if (converged == 1) /* :44-50 */
      break;
  }
  double t1 = now_ms();
  *iter_count = i;         /* :54 */
  *eigen_val = sum_vec[0]; /* :60-65 */
  if (elapsed_ms)
    *elapsed_ms = t1 - t0;
  free(sum_vec);
  return failed ? -1 : (int64_t)(t1 - t0);
}

/* The reference's stop test (the only one parity is defined on). */
int64_t
oracle_similarity_transform_ex(const float* mat, float* eigen_val, float* eigen_vec, uint32_t dim,
                               uint32_t* iter_count, float eps, uint32_t max_itr, int form,
                               int sum_mode, uint32_t ranks, double* elapsed_ms)
{
  return oracle_similarity_transform_ex2(mat, eigen_val, eigen_vec, dim, iter_count, eps, max_itr, form,
                                         sum_mode, ranks, ORACLE_STOP_ABSOLUTE, elapsed_ms);
}

/* Reference defaults: EPS, MAX_ITR, the literal in-place form. */
int64_t
oracle_similarity_transform(const float* mat, float* eigen_val, float* eigen_vec, uint32_t dim,
                            uint32_t* iter_count)
{
  return oracle_similarity_transform_ex(mat, eigen_val, eigen_vec, dim, iter_count, ORACLE_EPS,
                                        ORACLE_MAX_ITR, ORACLE_FORM_INPLACE, ORACLE_SUM_LANES16, 1,
                                        NULL);
}

/* Timed CPU baseline for bench.py: `rounds` full rounds of the reference's in-place round
 * (row sums, max, eigenvector update, stop test, D^-1 W D rescale) with no early exit, on a
 * matrix the caller generated.  Returns elapsed ms of the loop (matrix copy excluded, like
 * similarity_transform.cpp:36). */
double
oracle_time_rounds(const float* mat, uint32_t dim, uint32_t rounds, int form)
{
  const size_t n = dim;
  float* sum_vec = (float*)malloc(sizeof(float) * n);
  float* e = (float*)malloc(sizeof(float) * n);
  float* work = NULL;
  if (!sum_vec || !e)
    return -1.0;
  if (form == ORACLE_FORM_INPLACE) {
    work = (float*)malloc(sizeof(float) * n * n);
    if (!work)
      return -1.0;
    memcpy(work, mat, sizeof(float) * n * n);
  }
  oracle_initialise_eigen_vector(e, dim);
  volatile uint32_t sink = 0;
  double t0 = now_ms();
  for (uint32_t i = 0; i < rounds; i++) {
    if (form == ORACLE_FORM_INPLACE)
      oracle_sum_across_rows(work, sum_vec, dim, ORACLE_SUM_LANES16);
    else
      readonly_row_pass(mat, e, sum_vec, dim, 0, dim, ORACLE_SUM_LANES16);
    float m = oracle_find_max(sum_vec, dim);
    oracle_compute_eigen_vector(sum_vec, m, e, dim);
    sink += oracle_stop(sum_vec, dim, ORACLE_EPS);
    if (form == ORACLE_FORM_INPLACE)
      oracle_compute_next_matrix(work, sum_vec, dim);
  }
  double t1 = now_ms();
  free(sum_vec);
  free(e);
  free(work);
  return t1 - t0;
}
