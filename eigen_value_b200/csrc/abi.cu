// abi.cu -- the C ABI of libsimilarity_transform.so (include/similarity_transform.h).
//
// Part 1 re-exports the two symbols of reference wrapper/similarity_transform.cpp:3-37 on top
// of the CUDA solver; Part 2 is the additive st_* surface.  Nothing throws across this file.
#include "similarity_transform.hpp"

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#include <cstdio>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cuda_runtime.h>

using st::Context;
using st::Shard;

namespace {

template<typename F>
int
guarded(F&& f) noexcept
{
  try {
    return f();
  } catch (const std::invalid_argument& e) {
    st::set_last_error(e.what());
    return ST_ERR_ARG;
  } catch (const st::OutOfDeviceMemory& e) {
    st::set_last_error(e.what());
    return ST_ERR_NOMEM;
  } catch (const std::bad_alloc&) {
    st::set_last_error("host allocation failed");
    return ST_ERR_NOMEM;
  } catch (const std::exception& e) {
    st::set_last_error(e.what());
    return ST_ERR_CUDA;
  } catch (...) {
    st::set_last_error("unknown failure");
    return ST_ERR_CUDA;
  }
}

#define ABI_CUDA(call)                                                                             \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char buf_[512];                                                                              \
      snprintf(buf_, sizeof buf_, "%s failed: %s", #call, cudaGetErrorString(e_));                 \
      if (e_ == cudaErrorMemoryAllocation) {                                                       \
        (void)cudaGetLastError();                                                                  \
        throw st::OutOfDeviceMemory(buf_);                                                         \
      }                                                                                            \
      throw std::runtime_error(buf_);                                                              \
    }                                                                                              \
  } while (0)

Context*
as_ctx(void* p)
{
  if (!p)
    throw std::invalid_argument("null context");
  return static_cast<Context*>(p);
}

Shard*
as_shard(void* p)
{
  if (!p)
    throw std::invalid_argument("null shard");
  return static_cast<Shard*>(p);
}

constexpr size_t kFlagBytes = st::kExchangeHeaderBytes; // arrival counter, max slots: one 128-byte line each

// ---- device group: several GPUs behind ONE handle --------------------------------------------------
// st_group_attach (or ST_DEVICES in make_queue's environment) binds helper contexts on other GPUs to a
// context.  max_eigen_value / st_solve_host on that context then run the row-block sharded solve for
// matrices of min_dim and up: one host thread per GPU uploads its own row block from the caller's host
// matrix (so the copy runs over every GPU's PCIe link at once) and enters the collective round kernel;
// the shards are linked through plain peer access (st_shard_link_local).  Results are bit-identical to
// the one-GPU solve; the reference's wrapper needs no change to use the whole box.
struct Group
{
  std::vector<Context*> ctx; // [0] is the context the group is attached to (not owned), the others are owned
  std::vector<void*> shards; // built for one dimension at a time
  uint32_t dim = 0, world = 0, min_dim = 0;
  std::mutex mu;
  void drop_shards()
  {
    for (void* sh : shards)
      st_shard_destroy(sh);
    shards.clear();
    dim = world = 0;
  }
  ~Group()
  {
    drop_shards();
    for (size_t i = 1; i < ctx.size(); i++)
      delete ctx[i];
  }
};
constexpr uint32_t kGroupDefaultMinDim = 8192;

// a caller that looked its group up keeps it alive until its solve returns, whatever st_group_detach does meanwhile
// Heap-allocated and never freed on purpose: the reference's wrapper never destroys its queue, so with ST_DEVICES set a
// namespace-scope map would tear its groups down during static destruction -- cudaFree / cudaIpcCloseMemHandle /
// cudaStreamSynchronize while the CUDA runtime unloads, which is undefined.  Groups are released by st_group_detach /
// st_destroy only; whatever is still attached at exit goes with the process.
std::mutex& g_groups_mu = *new std::mutex;
std::map<Context*, std::shared_ptr<Group>>& g_groups = *new std::map<Context*, std::shared_ptr<Group>>;

std::shared_ptr<Group>
group_of(Context* c)
{
  std::lock_guard<std::mutex> lock(g_groups_mu);
  auto it = g_groups.find(c);
  return it == g_groups.end() ? nullptr : it->second;
}

void
group_drop_shards(Group& g)
{
  g.drop_shards();
}

// throws; the caller holds no context mutex
int
group_solve_host(Group& g, const float* h_mat, uint32_t dim, const st_options& opt, float* h_eigen_val, float* h_eigen_vec,
                 st_result* res)
{
  std::lock_guard<std::mutex> lock(g.mu);
  const auto t0 = std::chrono::steady_clock::now();
  const uint32_t world = (uint32_t)std::min<size_t>(g.ctx.size(), dim);
  if (g.dim != dim || g.world != world) { // exchange blocks are sized for one dimension
    group_drop_shards(g);
    for (uint32_t r = 0; r < world; r++) {
      void* sh = nullptr;
      const int rc = st_shard_create(g.ctx[r], dim, r, world, &sh);
      if (rc != ST_OK) {
        group_drop_shards(g);
        return rc; // st_last_error is set on this thread
      }
      g.shards.push_back(sh);
    }
    const int rc = st_shard_link_local(g.shards.data(), world);
    if (rc != ST_OK) {
      group_drop_shards(g);
      return rc;
    }
    g.dim = dim;
    g.world = world;
  }
  std::vector<int> rcs(world, ST_OK);
  std::vector<std::string> errs(world);
  std::vector<st_result> results(world);
  // Two phases with a host barrier between them: (1) every rank allocates and starts its upload, (2) every
  // rank enters the collective kernel.  No allocation may happen on any GPU while a peer already spins in
  // the round barrier, and if a rank fails in phase 1 nobody enters phase 2 (no 10 s timeout).
  std::mutex bar_mu;
  std::condition_variable bar_cv;
  uint32_t arrived = 0, failed = 0;
  auto work = [&](uint32_t r) {
    Context* c = g.ctx[r];
    Shard* sh = static_cast<Shard*>(g.shards[r]);
    std::lock_guard<std::mutex> ctx_lock(c->mutex());
    rcs[r] = guarded([&] {
      c->upload_rows(h_mat, dim, opt, sh);
      return ST_OK;
    });
    if (rcs[r] != ST_OK)
      errs[r] = st::last_error(); // thread-local: carry it to the calling thread
    {
      std::unique_lock<std::mutex> lk(bar_mu);
      failed += rcs[r] != ST_OK;
      if (++arrived == world)
        bar_cv.notify_all();
      else
        bar_cv.wait(lk, [&] { return arrived == world; });
      if (failed)
        return;
    }
    rcs[r] = guarded([&] {
      return c->solve_uploaded(dim, opt, r == 0 ? h_eigen_val : nullptr, r == 0 ? h_eigen_vec : nullptr, &results[r], sh);
    });
    if (rcs[r] != ST_OK)
      errs[r] = st::last_error();
  };
  std::vector<std::thread> threads;
  for (uint32_t r = 1; r < world; r++)
    threads.emplace_back(work, r);
  work(0);
  for (auto& t : threads)
    t.join();
  for (uint32_t r = 0; r < world; r++)
    if (rcs[r] != ST_OK) {
      // a rank that failed leaves the group's flag epochs diverged: rebuild the shards next time
      group_drop_shards(g);
      st::set_last_error("device group, rank " + std::to_string(r) + " (device " + std::to_string(g.ctx[r]->device()) +
                         "): " + errs[r]);
      return rcs[r];
    }
  if (res) {
    *res = results[0];
    res->launches = 0;
    res->bytes_per_round = 0;
    for (uint32_t r = 0; r < world; r++) {
      res->loop_ms = std::max(res->loop_ms, results[r].loop_ms); // max over ranks
      res->launches += results[r].launches;
      res->bytes_per_round += results[r].bytes_per_round;
    }
    res->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  return ST_OK;
}

// ST_DEVICES=all | "0,1,2,3": devices make_queue binds (the first one is the handle's own device)
std::vector<int>
devices_from_env()
{
  std::vector<int> out;
  const char* v = getenv("ST_DEVICES");
  if (!v || !*v)
    return out;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess)
    count = 0;
  if (strcmp(v, "all") == 0) {
    for (int d = 0; d < count; d++)
      out.push_back(d);
    return out;
  }
  for (const char* p = v; *p;) {
    char* end = nullptr;
    const long d = strtol(p, &end, 10);
    if (end == p || d < 0 || d >= count || std::find(out.begin(), out.end(), (int)d) != out.end())
      throw std::invalid_argument(std::string("ST_DEVICES: bad device list '") + v + "'");
    out.push_back((int)d);
    p = *end == ',' ? end + 1 : end;
    if (*end && *end != ',')
      throw std::invalid_argument(std::string("ST_DEVICES: bad device list '") + v + "'");
  }
  return out;
}

} // namespace

extern "C" {

// =============================================================================================
// Part 1 -- drop-in boundary
// =============================================================================================

void
make_queue(void** wq)
{
  if (!wq)
    return;
  *wq = nullptr;
  guarded([&]() -> int {
    // ST_DEVICES (extension): "all" or a list like "0,1,2,3" -- the handle lives on the first device and the
    // others help with matrices of 8192 rows and up; unset = device 0 alone, like the reference's default device
    const std::vector<int> devs = devices_from_env();
    Context* c = new Context(devs.empty() ? 0 : devs[0]);
    if (devs.size() > 1) {
      const char* md = getenv("ST_GROUP_MIN_DIM"); // optional: smallest dim that is sharded (default 8192)
      const long min_dim = md && *md ? strtol(md, nullptr, 10) : 0;
      const int rc = st_group_attach(c, devs.data() + 1, (uint32_t)devs.size() - 1, min_dim > 0 ? (uint32_t)min_dim : 0u);
      if (rc != ST_OK) {
        delete c;
        return rc;
      }
    }
    *wq = c;
    return ST_OK;
  });
}

int64_t
max_eigen_value(void* wq, float* mat, float* eigen_val, float* eigen_vec, st_uint dim,
                st_uint* iter_cnt)
{
  int64_t ms = -1;
  const int rc = guarded([&]() -> int {
    if (!mat || !eigen_val || !eigen_vec || !iter_cnt || dim == 0)
      throw std::invalid_argument("max_eigen_value: null pointer or dim == 0");
    Context* c = as_ctx(wq);
    if (auto g = group_of(c); g && dim >= g->min_dim) { // several GPUs behind this handle
      st_options o;
      st_default_options(&o);
      st_result r{};
      const int rc = group_solve_host(*g, mat, dim, o, eigen_val, eigen_vec, &r);
      if (rc != ST_OK)
        return rc;
      *iter_cnt = r.iter_count;  // exactly 4 bytes
      ms = (int64_t)r.loop_ms;   // whole milliseconds of the loop, like reference :56-58
      return ST_OK;
    }
    ms = similarity_transform(*c, mat, eigen_val, eigen_vec, dim, dim >> 1, iter_cnt);
    return ms < 0 ? (int)ms : ST_OK;
  });
  return rc == ST_OK ? ms : (int64_t)rc;
}

// =============================================================================================
// Part 2 -- extensions
// =============================================================================================

const char*
st_last_error(void)
{
  return st::last_error();
}

int
st_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
    return 0;
  return n;
}

void
st_default_options(st_options* opt)
{
  if (!opt)
    return;
  memset(opt, 0, sizeof *opt);
  opt->eps = ST_EPS;
  opt->max_iter = ST_MAX_ITR;
  opt->form = ST_FORM_READONLY;
  opt->sweep = 1; /* alternate the row order every round: +10 % at N=8192 from L2 reuse */
}

int
st_create(int device, void** ctx)
{
  if (!ctx)
    return ST_ERR_ARG;
  *ctx = nullptr;
  return guarded([&] {
    *ctx = new Context(device);
    return ST_OK;
  });
}

void
st_destroy(void* ctx)
{
  guarded([&] {
    if (ctx)
      st_group_detach(ctx);
    delete static_cast<Context*>(ctx);
    return ST_OK;
  });
}

uint64_t
st_staged_upload_bytes(void* ctx)
{
  uint64_t n = 0;
  guarded([&] {
    n = as_ctx(ctx)->staged_upload_bytes();
    return ST_OK;
  });
  return n;
}

int
st_group_attach(void* ctx, const int* devices, uint32_t count, uint32_t min_dim)
{
  return guarded([&] {
    Context* c = as_ctx(ctx);
    if (group_of(c))
      throw std::invalid_argument("st_group_attach: a group is already attached to this context");
    std::vector<int> devs;
    if (!devices || count == 0) { // every other visible device
      for (int d = 0; d < st_device_count(); d++)
        if (d != c->device())
          devs.push_back(d);
    } else {
      for (uint32_t i = 0; i < count; i++) {
        if (devices[i] == c->device() || std::find(devs.begin(), devs.end(), devices[i]) != devs.end())
          throw std::invalid_argument("st_group_attach: the list repeats a device or names the context's own");
        devs.push_back(devices[i]);
      }
    }
    if (devs.size() + 1 > ST_MAX_WORLD)
      throw std::invalid_argument("st_group_attach: more than ST_MAX_WORLD devices");
    auto g = std::make_shared<Group>();
    g->ctx.push_back(c);
    g->min_dim = min_dim ? min_dim : kGroupDefaultMinDim;
    for (int d : devs)
      g->ctx.push_back(new Context(d)); // a throw here frees the helpers made so far (~Group)
    c->activate();
    std::lock_guard<std::mutex> lock(g_groups_mu);
    g_groups[c] = std::move(g);
    return ST_OK;
  });
}

int
st_group_detach(void* ctx)
{
  return guarded([&] {
    Context* c = as_ctx(ctx);
    std::shared_ptr<Group> g;
    {
      std::lock_guard<std::mutex> lock(g_groups_mu);
      auto it = g_groups.find(c);
      if (it == g_groups.end())
        return ST_OK;
      g = std::move(it->second);
      g_groups.erase(it);
    }
    {
      std::lock_guard<std::mutex> busy(g->mu); // a solve in flight finishes first
    }
    g.reset(); // frees the helpers unless a caller still holds the group; then its last reference does
    c->activate();
    return ST_OK;
  });
}

int
st_group_size(void* ctx)
{
  int n = 0;
  guarded([&] {
    auto g = group_of(as_ctx(ctx));
    n = g ? (int)g->ctx.size() : 1;
    return ST_OK;
  });
  return n;
}

int
st_device_info(void* ctx, int* sm_count, size_t* l2_bytes, size_t* hbm_bytes, char* name,
               size_t name_len)
{
  return guarded([&] {
    Context* c = as_ctx(ctx);
    if (sm_count)
      *sm_count = c->sm_count();
    if (l2_bytes)
      *l2_bytes = c->l2_bytes();
    if (hbm_bytes)
      *hbm_bytes = c->hbm_bytes();
    if (name && name_len) {
      strncpy(name, c->name().c_str(), name_len - 1);
      name[name_len - 1] = 0;
    }
    return ST_OK;
  });
}

int
st_malloc(void* ctx, size_t bytes, void** dptr)
{
  return guarded([&] {
    if (!dptr)
      throw std::invalid_argument("st_malloc: null out pointer");
    as_ctx(ctx)->activate();
    ABI_CUDA(cudaMalloc(dptr, bytes));
    return ST_OK;
  });
}

int
st_free(void* ctx, void* dptr)
{
  return guarded([&] {
    as_ctx(ctx)->activate();
    ABI_CUDA(cudaFree(dptr));
    return ST_OK;
  });
}

int
st_memcpy_h2d(void* ctx, void* dptr, const void* hptr, size_t bytes)
{
  return guarded([&] {
    Context* c = as_ctx(ctx);
    if (!dptr || !hptr)
      throw std::invalid_argument("st_memcpy_h2d: null pointer");
    std::lock_guard<std::mutex> lock(c->mutex());
    c->upload(dptr, hptr, bytes); // pageable sources: multi-threaded staging, like max_eigen_value's own copy
    return ST_OK;
  });
}

int
st_memcpy_d2h(void* ctx, void* hptr, const void* dptr, size_t bytes)
{
  return guarded([&] {
    Context* c = as_ctx(ctx);
    c->activate();
    ABI_CUDA(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, c->stream()));
    ABI_CUDA(cudaStreamSynchronize(c->stream()));
    return ST_OK;
  });
}

int
st_pin_host(void* ctx, void* hptr, size_t bytes)
{
  return guarded([&] {
    if (!hptr || bytes == 0)
      throw std::invalid_argument("st_pin_host: bad argument");
    as_ctx(ctx)->activate();
    ABI_CUDA(cudaHostRegister(hptr, bytes, cudaHostRegisterDefault));
    return ST_OK;
  });
}

int
st_unpin_host(void* ctx, void* hptr)
{
  return guarded([&] {
    if (!hptr)
      throw std::invalid_argument("st_unpin_host: bad argument");
    as_ctx(ctx)->activate();
    ABI_CUDA(cudaHostUnregister(hptr));
    return ST_OK;
  });
}

int
st_synchronize(void* ctx)
{
  return guarded([&] {
    Context* c = as_ctx(ctx);
    c->activate();
    ABI_CUDA(cudaStreamSynchronize(c->stream()));
    return ST_OK;
  });
}

int
st_generate_hilbert(void* ctx, float* d_rows, uint32_t dim, uint32_t row0, uint32_t rows)
{
  return guarded([&] {
    if (!d_rows || dim == 0 || rows == 0 || (uint64_t)row0 + rows > dim)
      throw std::invalid_argument("st_generate_hilbert: bad row range");
    return generate_hilbert_matrix(*as_ctx(ctx), d_rows, dim, row0, rows);
  });
}

int
st_generate_uniform(void* ctx, float* d_rows, uint32_t dim, uint32_t row0, uint32_t rows,
                    uint64_t seed)
{
  return guarded([&] {
    if (!d_rows || dim == 0 || rows == 0 || (uint64_t)row0 + rows > dim)
      throw std::invalid_argument("st_generate_uniform: bad row range");
    return generate_uniform_matrix(*as_ctx(ctx), d_rows, dim, seed, row0, rows);
  });
}

int
st_solve_device(void* ctx, const float* d_mat, uint32_t dim, const st_options* opt,
                float* d_eigen_vec, st_result* res)
{
  return guarded([&] {
    Context* c = as_ctx(ctx);
    st_options o;
    if (opt)
      o = *opt;
    else
      st_default_options(&o);
    std::lock_guard<std::mutex> lock(c->mutex());
    return c->solve(d_mat, dim, o, nullptr, d_eigen_vec, res);
  });
}

int
st_solve_host(void* ctx, const float* h_mat, uint32_t dim, const st_options* opt,
              float* h_eigen_val, float* h_eigen_vec, st_result* res)
{
  return guarded([&] {
    Context* c = as_ctx(ctx);
    st_options o;
    if (opt)
      o = *opt;
    else
      st_default_options(&o);
    if (auto g = group_of(c); g && dim >= g->min_dim && h_mat)
      return group_solve_host(*g, h_mat, dim, o, h_eigen_val, h_eigen_vec, res);
    std::lock_guard<std::mutex> lock(c->mutex());
    return c->solve_host(h_mat, dim, o, h_eigen_val, h_eigen_vec, res);
  });
}

int
st_solve_streamed(void* ctx, const float* h_mat, uint32_t dim, const st_options* opt, size_t device_budget,
                  uint32_t block_rows, float* h_eigen_val, float* h_eigen_vec, st_result* res, st_stream_plan* plan)
{
  return guarded([&] {
    Context* c = as_ctx(ctx);
    st_options o;
    if (opt)
      o = *opt;
    else
      st_default_options(&o);
    std::lock_guard<std::mutex> lock(c->mutex());
    return c->solve_streamed(h_mat, dim, o, device_budget, block_rows, h_eigen_val, h_eigen_vec, res, plan);
  });
}

namespace {
// read-only mapping of a file range, released on every path out of st_solve_file
struct FileMapping
{
  int fd = -1;
  void* base = MAP_FAILED;
  size_t length = 0;
  ~FileMapping()
  {
    if (base != MAP_FAILED)
      munmap(base, length);
    if (fd >= 0)
      close(fd);
  }
};
} // namespace

int
st_solve_file(void* ctx, const char* path, uint64_t offset, uint32_t dim, const st_options* opt, size_t device_budget,
              uint32_t block_rows, float* h_eigen_val, float* h_eigen_vec, st_result* res, st_stream_plan* plan)
{
  return guarded([&] {
    Context* c = as_ctx(ctx);
    if (!path || dim == 0 || offset % sizeof(float) != 0)
      throw std::invalid_argument("st_solve_file: null path, dim == 0 or an offset that is not a multiple of 4");
    const uint64_t need = (uint64_t)dim * dim * sizeof(float);
    FileMapping m;
    m.fd = open(path, O_RDONLY | O_CLOEXEC);
    if (m.fd < 0)
      throw std::invalid_argument(std::string("st_solve_file: cannot open ") + path + ": " + strerror(errno));
    struct stat sb{};
    if (fstat(m.fd, &sb) != 0 || (uint64_t)sb.st_size < offset + need)
      throw std::invalid_argument(std::string("st_solve_file: ") + path + " is shorter than offset + 4 * dim * dim bytes");
    const uint64_t page = (uint64_t)sysconf(_SC_PAGESIZE);
    const uint64_t start = offset / page * page;
    m.length = (size_t)(offset - start + need);
    m.base = mmap(nullptr, m.length, PROT_READ, MAP_PRIVATE, m.fd, (off_t)start);
    if (m.base == MAP_FAILED)
      throw std::runtime_error(std::string("st_solve_file: mmap failed: ") + strerror(errno));
    const float* mat = reinterpret_cast<const float*>(static_cast<const char*>(m.base) + (offset - start));
    st_options o;
    if (opt)
      o = *opt;
    else
      st_default_options(&o);
    std::lock_guard<std::mutex> lock(c->mutex());
    return c->solve_streamed(mat, dim, o, device_budget, block_rows, h_eigen_val, h_eigen_vec, res, plan);
  });
}

int
st_convert_f32_to_bf16(void* ctx, const float* d_src, uint16_t* d_dst, size_t count)
{
  return guarded([&] {
    if (!d_src || !d_dst || count == 0)
      throw std::invalid_argument("st_convert_f32_to_bf16: bad argument");
    return convert_to_bf16(*as_ctx(ctx), d_src, d_dst, count);
  });
}

int
st_solve_device_bf16(void* ctx, const uint16_t* d_mat, uint32_t dim, const st_options* opt,
                     float* d_eigen_vec, st_result* res)
{
  return guarded([&] {
    Context* c = as_ctx(ctx);
    st_options o;
    if (opt)
      o = *opt;
    else
      st_default_options(&o);
    std::lock_guard<std::mutex> lock(c->mutex());
    return c->solve(reinterpret_cast<const float*>(d_mat), dim, o, nullptr, d_eigen_vec, res, true);
  });
}

int
st_convert_f32_to_fp8(void* ctx, const float* d_src, uint8_t* d_dst, float* d_row_scale, uint32_t rows, uint32_t dim)
{
  return guarded([&] {
    if (!d_src || !d_dst || !d_row_scale || rows == 0 || dim == 0 || dim % 4u != 0u ||
        (reinterpret_cast<uintptr_t>(d_src) & 15u) != 0 || (reinterpret_cast<uintptr_t>(d_dst) & 15u) != 0)
      throw std::invalid_argument("st_convert_f32_to_fp8: needs dim % 4 == 0 and 16-byte aligned buffers");
    return convert_to_fp8(*as_ctx(ctx), d_src, d_dst, d_row_scale, rows, dim);
  });
}

int
st_solve_device_fp8(void* ctx, const uint8_t* d_mat, const float* d_row_scale, uint32_t dim, const st_options* opt,
                    float* d_eigen_vec, st_result* res)
{
  return guarded([&] {
    Context* c = as_ctx(ctx);
    if (!d_mat || !d_row_scale)
      throw std::invalid_argument("st_solve_device_fp8: null matrix or row scales");
    st_options o;
    if (opt)
      o = *opt;
    else
      st_default_options(&o);
    std::lock_guard<std::mutex> lock(c->mutex());
    return c->solve(reinterpret_cast<const float*>(d_mat), dim, o, nullptr, d_eigen_vec, res, false, d_row_scale);
  });
}

int
st_round_timestamps(void* ctx, uint64_t* out, uint32_t capacity, uint32_t* count)
{
  return guarded([&] {
    Context* c = as_ctx(ctx);
    const auto& ts = c->round_timestamps();
    const uint32_t n = (uint32_t)std::min<size_t>(ts.size(), capacity);
    if (out)
      memcpy(out, ts.data(), sizeof(uint64_t) * n);
    if (count)
      *count = (uint32_t)ts.size();
    return ST_OK;
  });
}

int
st_phase_timestamps(void* ctx, uint64_t* out, uint32_t capacity, uint32_t* count)
{
  return guarded([&] {
    Context* c = as_ctx(ctx);
    const auto& ts = c->phase_timestamps();
    const uint32_t n = (uint32_t)std::min<size_t>(ts.size(), capacity);
    if (out)
      memcpy(out, ts.data(), sizeof(uint64_t) * n);
    if (count)
      *count = (uint32_t)ts.size();
    return ST_OK;
  });
}

int
st_timer_start(void* ctx)
{
  return guarded([&] {
    as_ctx(ctx)->timer_start();
    return ST_OK;
  });
}

int
st_timer_stop(void* ctx, float* ms)
{
  return guarded([&] {
    if (!ms)
      throw std::invalid_argument("st_timer_stop: null out pointer");
    *ms = as_ctx(ctx)->timer_stop();
    return ST_OK;
  });
}

// ---- per-kernel entry points ------------------------------------------------------------------
int
st_sum_across_rows(void* ctx, const float* d_mat, float* d_vec, uint32_t dim)
{
  return guarded([&] {
    if (!d_mat || !d_vec || dim == 0)
      throw std::invalid_argument("st_sum_across_rows: bad argument");
    return sum_across_rows(*as_ctx(ctx), d_mat, d_vec, dim, 0);
  });
}
int
st_row_pass_readonly(void* ctx, const float* d_rows, const float* d_e, float* d_vec, uint32_t dim,
                     uint32_t row0, uint32_t rows)
{
  return guarded([&] {
    if (!d_rows || !d_e || !d_vec || rows == 0 || (uint64_t)row0 + rows > dim)
      throw std::invalid_argument("st_row_pass_readonly: bad argument");
    return row_pass_readonly(*as_ctx(ctx), d_rows, d_e, d_vec, dim, row0, rows);
  });
}
int
st_find_max(void* ctx, const float* d_vec, float* d_max, uint32_t dim)
{
  return guarded([&] {
    if (!d_vec || !d_max || dim == 0)
      throw std::invalid_argument("st_find_max: bad argument");
    return find_max(*as_ctx(ctx), d_vec, d_max, dim, 0);
  });
}
int
st_compute_eigen_vector(void* ctx, const float* d_vec, const float* d_max, float* d_eigen_vec,
                        uint32_t dim)
{
  return guarded([&] {
    if (!d_vec || !d_max || !d_eigen_vec || dim == 0)
      throw std::invalid_argument("st_compute_eigen_vector: bad argument");
    return compute_eigen_vector(*as_ctx(ctx), d_vec, d_max, d_eigen_vec, dim, 0);
  });
}
int
st_initialise_eigen_vector(void* ctx, float* d_eigen_vec, uint32_t dim)
{
  return guarded([&] {
    if (!d_eigen_vec || dim == 0)
      throw std::invalid_argument("st_initialise_eigen_vector: bad argument");
    return initialise_eigen_vector(*as_ctx(ctx), d_eigen_vec, dim);
  });
}
int
st_compute_next_matrix(void* ctx, float* d_mat, const float* d_vec, uint32_t dim)
{
  return guarded([&] {
    if (!d_mat || !d_vec || dim == 0)
      throw std::invalid_argument("st_compute_next_matrix: bad argument");
    return compute_next_matrix(*as_ctx(ctx), d_mat, d_vec, dim, 0);
  });
}
int
st_stop(void* ctx, const float* d_vec, uint32_t* d_ret, uint32_t dim, float eps)
{
  return guarded([&] {
    if (!d_vec || !d_ret || dim == 0)
      throw std::invalid_argument("st_stop: bad argument");
    return stop(*as_ctx(ctx), d_vec, d_ret, dim, 0, eps);
  });
}

// ---- row-block sharding -------------------------------------------------------------------------
int
st_shard_create(void* ctx, uint32_t dim, uint32_t rank, uint32_t world, void** shard)
{
  if (!shard)
    return ST_ERR_ARG;
  *shard = nullptr;
  return guarded([&] {
    Context* c = as_ctx(ctx);
    if (dim == 0 || world == 0 || world > ST_MAX_WORLD || rank >= world || world > dim)
      throw std::invalid_argument("st_shard_create: bad rank/world/dim");
    c->activate();
    Shard* s = new Shard();
    s->ctx = c;
    s->dim = dim;
    s->rank = rank;
    s->world = world;
    s->row0 = (uint32_t)((uint64_t)dim * rank / world);
    s->rows = (uint32_t)((uint64_t)dim * (rank + 1) / world) - s->row0;
    const size_t vec_bytes = (((size_t)dim * sizeof(float)) + 127) & ~(size_t)127;
    s->s_offset[0] = kFlagBytes;
    s->s_offset[1] = kFlagBytes + vec_bytes;
    s->block_bytes = kFlagBytes + 2 * vec_bytes;
    cudaError_t e = cudaMalloc(&s->block, s->block_bytes);
    if (e != cudaSuccess) {
      delete s;
      (void)cudaGetLastError();
      if (e == cudaErrorMemoryAllocation)
        throw st::OutOfDeviceMemory(std::string("cudaMalloc(exchange block): ") + cudaGetErrorString(e));
      throw std::runtime_error(std::string("cudaMalloc(exchange block): ") + cudaGetErrorString(e));
    }
    ABI_CUDA(cudaMemset(s->block, 0, s->block_bytes));
    s->peer_block[rank] = s->block;
    s->linked = (world == 1);
    // scratch for solves with the default options, now -- a sharded solve never allocates (Context::solve)
    st_options o;
    st_default_options(&o);
    try {
      std::lock_guard<std::mutex> lock(c->mutex());
      c->prepare(dim, s->rows, o);
    } catch (...) {
      cudaFree(s->block);
      delete s;
      throw;
    }
    *shard = s;
    return ST_OK;
  });
}

int
st_shard_prepare(void* shard, const st_options* opt)
{
  return guarded([&] {
    Shard* s = as_shard(shard);
    st_options o;
    if (opt)
      o = *opt;
    else
      st_default_options(&o);
    std::lock_guard<std::mutex> lock(s->ctx->mutex());
    s->ctx->prepare(s->dim, s->rows, o);
    return ST_OK;
  });
}

int
st_shard_export(void* shard, void* handle_out)
{
  return guarded([&] {
    Shard* s = as_shard(shard);
    if (!handle_out)
      throw std::invalid_argument("st_shard_export: null handle");
    static_assert(sizeof(cudaIpcMemHandle_t) == ST_IPC_HANDLE_BYTES, "IPC handle size");
    s->ctx->activate();
    cudaIpcMemHandle_t h;
    ABI_CUDA(cudaIpcGetMemHandle(&h, s->block));
    memcpy(handle_out, &h, sizeof h);
    return ST_OK;
  });
}

int
st_shard_import(void* shard, const void* handles)
{
  return guarded([&] {
    Shard* s = as_shard(shard);
    if (!handles)
      throw std::invalid_argument("st_shard_import: null handle table");
    s->ctx->activate();
    const char* tab = static_cast<const char*>(handles);
    for (uint32_t g = 0; g < s->world; g++) {
      if (g == s->rank)
        continue;
      cudaIpcMemHandle_t h;
      memcpy(&h, tab + (size_t)g * ST_IPC_HANDLE_BYTES, sizeof h);
      void* p = nullptr;
      ABI_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      s->peer_block[g] = p;
      s->peer_is_ipc[g] = true;
    }
    s->linked = true;
    return ST_OK;
  });
}

int
st_shard_link_local(void** shards, uint32_t world)
{
  return guarded([&] {
    if (!shards || world == 0 || world > ST_MAX_WORLD)
      throw std::invalid_argument("st_shard_link_local: bad world");
    for (uint32_t a = 0; a < world; a++) {
      Shard* sa = as_shard(shards[a]);
      if (sa->rank != a || sa->world != world)
        throw std::invalid_argument("st_shard_link_local: shards must be passed in rank order");
      sa->ctx->activate();
      for (uint32_t b = 0; b < world; b++) {
        Shard* sb = as_shard(shards[b]);
        if (a != b && sa->ctx->device() != sb->ctx->device()) {
          int can = 0;
          ABI_CUDA(cudaDeviceCanAccessPeer(&can, sa->ctx->device(), sb->ctx->device()));
          if (!can)
            throw std::runtime_error("st_shard_link_local: no peer access between the devices");
          cudaError_t e = cudaDeviceEnablePeerAccess(sb->ctx->device(), 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            throw std::runtime_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
          (void)cudaGetLastError();
        }
        sa->peer_block[b] = sb->block;
      }
      sa->linked = true;
    }
    return ST_OK;
  });
}

int
st_shard_rows(void* shard, uint32_t* row0, uint32_t* rows)
{
  return guarded([&] {
    Shard* s = as_shard(shard);
    if (row0)
      *row0 = s->row0;
    if (rows)
      *rows = s->rows;
    return ST_OK;
  });
}

int
st_shard_solve(void* shard, const float* d_rows, const st_options* opt, float* d_eigen_vec,
               st_result* res)
{
  return guarded([&] {
    Shard* s = as_shard(shard);
    st_options o;
    if (opt)
      o = *opt;
    else
      st_default_options(&o);
    std::lock_guard<std::mutex> lock(s->ctx->mutex());
    return s->ctx->solve(d_rows, s->dim, o, s, d_eigen_vec, res);
  });
}

int
st_shard_solve_bf16(void* shard, const uint16_t* d_rows, const st_options* opt, float* d_eigen_vec,
                    st_result* res)
{
  return guarded([&] {
    Shard* s = as_shard(shard);
    st_options o;
    if (opt)
      o = *opt;
    else
      st_default_options(&o);
    std::lock_guard<std::mutex> lock(s->ctx->mutex());
    return s->ctx->solve(reinterpret_cast<const float*>(d_rows), s->dim, o, s, d_eigen_vec, res, true);
  });
}

int
st_shard_solve_fp8(void* shard, const uint8_t* d_rows, const float* d_row_scale, const st_options* opt, float* d_eigen_vec,
                   st_result* res)
{
  return guarded([&] {
    Shard* s = as_shard(shard);
    if (!d_rows || !d_row_scale)
      throw std::invalid_argument("st_shard_solve_fp8: null rows or row scales");
    st_options o;
    if (opt)
      o = *opt;
    else
      st_default_options(&o);
    std::lock_guard<std::mutex> lock(s->ctx->mutex());
    return s->ctx->solve(reinterpret_cast<const float*>(d_rows), s->dim, o, s, d_eigen_vec, res, false, d_row_scale);
  });
}

void
st_shard_destroy(void* shard)
{
  guarded([&] {
    Shard* s = static_cast<Shard*>(shard);
    if (!s)
      return ST_OK;
    s->ctx->activate();
    for (uint32_t g = 0; g < s->world; g++)
      if (s->peer_is_ipc[g] && s->peer_block[g])
        cudaIpcCloseMemHandle(s->peer_block[g]);
    cudaFree(s->block);
    delete s;
    return ST_OK;
  });
}

} // extern "C"
