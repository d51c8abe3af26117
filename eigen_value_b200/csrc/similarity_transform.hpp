// similarity_transform.hpp -- C++ host interface of the B200 build.
//
// Mirrors the reference's include/similarity_transform.hpp:46-100 with the sycl::queue&
// replaced by an st::Context& (a CUDA device + stream + cached scratch).  The reference's
// `wg_size` argument is accepted and ignored: it is a launch-shape knob without numeric
// effect there (wrapper/similarity_transform.cpp:33 vs main.cpp:28 pick different values).
// Vector/matrix arguments of the per-kernel functions are DEVICE pointers (the reference
// passes sycl::buffer objects).
#pragma once

#include <cstddef>
#include <cstdint>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/similarity_transform.h"

typedef unsigned int uint;

inline constexpr float EPS = ST_EPS;        // reference include/similarity_transform.hpp:4
inline constexpr uint MAX_ITR = ST_MAX_ITR; // reference include/similarity_transform.hpp:5

struct CUstream_st;
struct CUevent_st;

namespace st {

struct BarrierState;
struct Shard;

// What make_queue() hands out: one CUDA device, one stream, scratch that is reused across
// solves (the reference allocates and frees its scratch per call, similarity_transform.cpp:14-17).
class Context
{
public:
  explicit Context(int device);
  ~Context();
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;

  int device() const { return device_; }
  int sm_count() const { return sm_count_; }
  size_t l2_bytes() const { return l2_bytes_; }
  size_t hbm_bytes() const { return hbm_bytes_; }
  const std::string& name() const { return name_; }
  CUstream_st* stream() const { return stream_; }

  // Whole round loop on device-resident rows.  shard == nullptr: single GPU (rows == dim).
  // bf16 == true: d_rows points to bfloat16 storage (read-only form, dim % 4 == 0); d_row_scale != nullptr: d_rows
  // points to fp8 (e4m3) storage with one power-of-two scale per owned row (read-only form, dim % 4 == 0); fp32 otherwise.
  int solve(const float* d_rows, uint32_t dim, const st_options& opt, Shard* shard,
            float* d_eigen_vec, st_result* res, bool bf16 = false, const float* d_row_scale = nullptr);
  // shard != nullptr (device group): h_mat is still the whole matrix; this context uploads and solves
  // its own row block, collectively with the other ranks of the shard group.
  int solve_host(const float* h_mat, uint32_t dim, const st_options& opt, float* h_eigen_val,
                 float* h_eigen_vec, st_result* res, Shard* shard = nullptr);
  // Every device allocation a solve() of `rows` rows of a dim-column matrix with these options needs (vectors,
  // stamps, working copy / chunk sums).  solve() calls it itself on one GPU; a SHARDED solve never allocates --
  // a cudaMalloc / cudaFree on a device with peer mappings may wait for peers that already spin in the collective
  // round kernel -- so st_shard_create / st_shard_prepare / upload_rows call it before any rank launches.
  void prepare(uint32_t dim, uint32_t rows, const st_options& opt);
  bool prepared(uint32_t dim, uint32_t rows, const st_options& opt) const;
  // the two halves of solve_host: every allocation + the (asynchronous) upload | the solve + read-back
  void upload_rows(const float* h_mat, uint32_t dim, const st_options& opt, Shard* shard);
  int solve_uploaded(uint32_t dim, const st_options& opt, float* h_eigen_val, float* h_eigen_vec, st_result* res,
                     Shard* shard);
  // Host matrix larger than the device (or than `device_budget`): block cache + alternating sweep,
  // host-driven rounds (include/similarity_transform.h, st_solve_streamed).
  int solve_streamed(const float* h_mat, uint32_t dim, const st_options& opt, size_t device_budget,
                     uint32_t block_rows, float* h_eigen_val, float* h_eigen_vec, st_result* res,
                     st_stream_plan* plan);
  const std::vector<uint64_t>& round_timestamps() const { return last_ts_; }
  const std::vector<uint64_t>& phase_timestamps() const { return last_phase_ts_; }

  // CUDA-event stopwatch on the context's stream (st_timer_start / st_timer_stop)
  void timer_start();
  float timer_stop();

  // host -> device copy on the context's stream, synchronous; pageable sources take the multi-threaded staging path
  void upload(void* d_dst, const void* h_src, size_t bytes);
  uint64_t staged_upload_bytes() const { return staged_bytes_; } // bytes that took the multi-threaded upload path
  std::mutex& mutex() { return mu_; }
  void activate() const; // cudaSetDevice

private:
  friend struct Shard;
  void preload_kernels();
  struct ScratchNeed
  {
    uint32_t vec = 0, stamps = 0;
    size_t work = 0;
  };
  ScratchNeed scratch_need(uint32_t dim, uint32_t rows, const st_options& opt) const;
  void reserve_vectors(uint32_t dim, uint32_t max_iter);
  void reserve_matrix(size_t elems);
  void reserve_work(size_t elems);
  void copy_h2d(float* d_dst, const float* h_src, size_t bytes, CUstream_st* stream = nullptr);
  void load_local_cpus();
  void bind_this_thread_near_device() const;

  int device_ = 0;
  int sm_count_ = 0;
  size_t l2_bytes_ = 0;
  size_t hbm_bytes_ = 0;
  std::string name_;
  CUstream_st* stream_ = nullptr;
  CUevent_st* ev0_ = nullptr;
  CUevent_st* ev1_ = nullptr;
  CUevent_st* ev_timer_[2] = { nullptr, nullptr };
  std::mutex mu_;

  // scratch
  float* d_vec_ = nullptr; // S0 S1 E0 E1 OUT, each vec_cap_ floats
  uint32_t vec_cap_ = 0;
  unsigned long long* d_ts_ = nullptr;
  uint32_t ts_cap_ = 0;
  BarrierState* d_bar_ = nullptr;
  float* d_scalars_ = nullptr; // [0] eigen_val, [1..2] iter/passes (as uint32), [3] spare
  float* d_mat_ = nullptr;     // staging for host-pointer solves
  size_t mat_cap_ = 0;
  float* d_work_ = nullptr; // in-place working copy
  size_t work_cap_ = 0;
  void* h_pinned_ = nullptr; // small pinned read-back block
  // streamed solve: copy stream and two events per cache slot (block landed / block consumed)
  CUstream_st* copy_stream_ = nullptr;
  std::vector<CUevent_st*> slot_ready_, slot_free_;
  // multi-threaded staging of pageable host matrices (ST_UPLOAD_THREADS, default 4; 0 = the driver's own staging)
  int upload_threads_ = 0;
  std::vector<int> local_cpus_; // CPUs on this GPU's NUMA node that the process may use (empty: unknown)
  uint64_t staged_bytes_ = 0;
  void* bounce_ = nullptr;
  std::vector<CUstream_st*> up_streams_;
  std::vector<CUevent_st*> up_events_;
  std::vector<char> up_used_; // per thread x buffer: has its event ever been recorded
  std::vector<uint64_t> last_ts_;
  std::vector<uint64_t> last_phase_ts_;
};

// bytes of the exchange block's head (kernels.cuh: ExchangeHeader -- arrival counter and max slots of the round barrier)
constexpr size_t kExchangeHeaderBytes = 256;

// Row-block shard: this rank's exchange block (barrier words + two N-float row-sum buffers) and the
// mapped blocks of the peers.
struct Shard
{
  Context* ctx = nullptr;
  uint32_t dim = 0, rank = 0, world = 1;
  uint32_t row0 = 0, rows = 0;
  void* block = nullptr; // local exchange block (cudaMalloc)
  size_t block_bytes = 0;
  size_t s_offset[2] = { 0, 0 };
  void* peer_block[ST_MAX_WORLD] = {};
  bool peer_is_ipc[ST_MAX_WORLD] = {};
  bool linked = false;
  uint64_t solves = 0;
  uint32_t flip = 0; // parity offset of the next solve's exchange buffers (RoundParams::flip)
  uint64_t arrive_total = 0; // what every rank's arrival counter reads between solves (never reset)
  uint64_t rounds_total = 0; // rounds run so far (the barrier's max slots rotate by round)
};

} // namespace st

// ---- the reference's entry point, same name and argument order ---------------------------
// reference include/similarity_transform.hpp:46-53 / similarity_transform.cpp:5-75
int64_t
similarity_transform(st::Context& q, const float* mat, float* const eigen_val,
                     float* const eigen_vec, const uint dim, const uint wg_size,
                     uint* const iter_count);

// ---- per-kernel functions (device pointers) ----------------------------------------------
// reference include/similarity_transform.hpp:55-100
int sum_across_rows(st::Context& q, const float* d_mat, float* d_vec, const uint dim, const uint wg_size);
// one read-only round's row pass on a row block: d_vec[row0+r] = (A_r . e) / e[row0+r]
int row_pass_readonly(st::Context& q, const float* d_rows, const float* d_e, float* d_vec,
                      const uint dim, const uint row0, const uint rows);
int find_max(st::Context& q, const float* d_vec, float* d_max, const uint dim, const uint wg_size);
int compute_eigen_vector(st::Context& q, const float* d_vec, const float* d_max, float* d_eigen_vec,
                         const uint dim, const uint wg_size);
int initialise_eigen_vector(st::Context& q, float* d_eigen_vec, const uint dim);
int compute_next_matrix(st::Context& q, float* d_mat, const float* d_vec, const uint dim,
                        const uint wg_size);
int stop(st::Context& q, const float* d_vec, uint* d_ret, const uint dim, const uint wg_size,
         float eps = EPS);

// ---- input generation (reference utils.cpp:136-154, :124-134), on the device -------------
int generate_hilbert_matrix(st::Context& q, float* d_rows, const uint dim, const uint row0 = 0,
                            const uint rows = 0);
int generate_uniform_matrix(st::Context& q, float* d_rows, const uint dim, uint64_t seed,
                            const uint row0 = 0, const uint rows = 0);
// fp32 -> bf16 (round to nearest even) on the device, for the bf16-storage solves
int convert_to_bf16(st::Context& q, const float* d_src, uint16_t* d_dst, size_t count);
// fp32 -> fp8 (e4m3) storage with one power-of-two scale per row, on the device (dim % 4 == 0, 16-byte aligned rows)
int convert_to_fp8(st::Context& q, const float* d_src, uint8_t* d_dst, float* d_row_scale, uint32_t rows, uint32_t dim);

namespace st {
void set_last_error(const std::string& msg);
const char* last_error();
// thrown when a device allocation fails; the C ABI turns it into ST_ERR_NOMEM
struct OutOfDeviceMemory : std::runtime_error
{
  using std::runtime_error::runtime_error;
};
} // namespace st
