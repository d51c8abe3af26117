// C++ acceptance test of the B200 build, scenario by scenario the reference's tests/test.cpp:
// per-kernel checks on N = 1024 (identity row sums :22-30, max of 1..N :32-41, first
// eigenvector update :43-54, stop success :56-64, stop failure through the wrap pair :66-73) and
// the 3x3 golden eigenpair through the reference-named entry point similarity_transform()
// (:79-104), with st::Context standing where sycl::queue stood.
//
// Built by tests/test_gpu_cpp.py:  g++ -std=c++17 test_similarity_transform.cpp -lsimilarity_transform
// Exit code 0 = all checks passed.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../eigen_value_b200/csrc/similarity_transform.hpp"

static const uint N = 1 << 10; // reference tests/test.cpp:7
static const uint B = 1 << 7;  // reference tests/test.cpp:8 (work-group size; ignored here)

#define CHECK(cond)                                                                                \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      std::printf("FAILED %s:%d: %s   [%s]\n", __FILE__, __LINE__, #cond, st_last_error());        \
      return 1;                                                                                    \
    }                                                                                              \
  } while (0)

struct DeviceVec
{
  void* ctx;
  void* p = nullptr;
  size_t bytes;
  DeviceVec(void* c, size_t n_bytes)
    : ctx(c)
    , bytes(n_bytes)
  {
    if (st_malloc(ctx, bytes, &p) != ST_OK)
      p = nullptr;
  }
  ~DeviceVec() { st_free(ctx, p); }
  template<typename T>
  T* as()
  {
    return static_cast<T*>(p);
  }
  int up(const void* h) { return st_memcpy_h2d(ctx, p, h, bytes); }
  int down(void* h) { return st_memcpy_d2h(ctx, h, p, bytes); }
};

int
main()
{
  if (st_device_count() < 1) {
    std::printf("no CUDA device\n");
    return 2;
  }
  st::Context q(0);
  void* ctx = &q; // the C ABI's context handle is the st::Context
  std::printf("running on %s\n\n", q.name().c_str());

  // ---- sum across rows of the identity matrix (reference :22-30, utils.cpp:5-35) ----
  std::vector<float> mat((size_t)N * N, 0.f), vec(N, 0.f), eigen_vec(N, 0.f);
  for (uint r = 0; r < N; r++)
    mat[(size_t)r * N + r] = 1.f;
  DeviceVec d_mat(ctx, sizeof(float) * N * N), d_vec(ctx, sizeof(float) * N), d_e(ctx, sizeof(float) * N),
    d_max(ctx, sizeof(float)), d_ret(ctx, sizeof(uint));
  CHECK(d_mat.p && d_vec.p && d_e.p && d_max.p && d_ret.p);
  CHECK(d_mat.up(mat.data()) == ST_OK);
  CHECK(sum_across_rows(q, d_mat.as<float>(), d_vec.as<float>(), N, B) == ST_OK);
  CHECK(st_synchronize(ctx) == ST_OK && d_vec.down(vec.data()) == ST_OK);
  for (uint i = 0; i < N; i++)
    CHECK(vec[i] == 1.f);
  std::printf("sum across row works !\n");

  // ---- max of v[r] = r + 1 (reference :32-41, utils.cpp:37-59) ----
  for (uint r = 0; r < N; r++)
    vec[r] = (float)(r + 1);
  CHECK(d_vec.up(vec.data()) == ST_OK);
  CHECK(find_max(q, d_vec.as<float>(), d_max.as<float>(), N, B) == ST_OK);
  float max = 0.f;
  CHECK(st_synchronize(ctx) == ST_OK && d_max.down(&max) == ST_OK);
  CHECK(max == (float)N);
  std::printf("max from vector works !\n");

  // ---- first eigenvector update (reference :43-54, utils.cpp:61-72) ----
  CHECK(initialise_eigen_vector(q, d_e.as<float>(), N) == ST_OK);
  CHECK(compute_eigen_vector(q, d_vec.as<float>(), d_max.as<float>(), d_e.as<float>(), N, B) == ST_OK);
  CHECK(st_synchronize(ctx) == ST_OK && d_e.down(eigen_vec.data()) == ST_OK);
  float max_dev = 0.f;
  for (uint i = 0; i < N; i++)
    max_dev = std::fmax(max_dev, std::fabs(vec[i] / max - eigen_vec[i]));
  std::printf("maximum deviation in computing eigen vector %g\n", max_dev);
  CHECK(max_dev == 0.f);

  // ---- stop criterion (reference :56-73, utils.cpp:74-122) ----
  uint ret = 7;
  for (uint r = 0; r < N; r++)
    vec[r] = 1.f + 1e-4f;
  CHECK(d_vec.up(vec.data()) == ST_OK);
  CHECK(stop(q, d_vec.as<float>(), d_ret.as<uint>(), N, B) == ST_OK);
  CHECK(st_synchronize(ctx) == ST_OK && d_ret.down(&ret) == ST_OK);
  std::printf("stopping criteria test result [success]: %u\n", ret);
  CHECK(ret == 1);
  for (uint r = 0; r < N; r++)
    vec[r] = (float)(r + 1) * 1e-4f;
  CHECK(d_vec.up(vec.data()) == ST_OK);
  CHECK(stop(q, d_vec.as<float>(), d_ret.as<uint>(), N, B) == ST_OK);
  CHECK(st_synchronize(ctx) == ST_OK && d_ret.down(&ret) == ST_OK);
  std::printf("stopping criteria test result [fail]: %u\n", ret);
  CHECK(ret == 0); // only the wrap pair |v[N-1] - v[0]| breaks it

  // ---- the 3x3 golden (reference :79-104) ----
  float m3[9] = { 1, 1, 2, 2, 1, 3, 2, 3, 5 }, keep[9];
  for (int i = 0; i < 9; i++)
    keep[i] = m3[i];
  float eigen_val = 0.f, e3[3] = { 0, 0, 0 };
  uint iter_count = 0;
  int64_t ts = similarity_transform(q, m3, &eigen_val, e3, 3, 3, &iter_count);
  CHECK(ts >= 0);
  CHECK(std::fabs(eigen_val - 7.53114f) < EPS);
  CHECK(std::fabs(e3[0] - 0.394074f) < EPS);
  CHECK(std::fabs(e3[1] - 0.578844f) < EPS);
  CHECK(std::fabs(e3[2] - 0.997451f) < EPS);
  CHECK(iter_count == 4);
  for (int i = 0; i < 9; i++)
    CHECK(m3[i] == keep[i]); // caller's matrix untouched (similarity_transform.cpp:14,19)
  std::printf("similarity transform worked !\t\t[ %u iterations ]\t\t%ld ms\n", iter_count, (long)ts);

  // ---- main.cpp:23-35, one row of the README table: Hilbert 1024 -> 13 rounds ----
  {
    const uint dim = 1024;
    DeviceVec d_h(ctx, sizeof(float) * dim * dim);
    std::vector<float> h((size_t)dim * dim), ev(dim);
    CHECK(generate_hilbert_matrix(q, d_h.as<float>(), dim) == ST_OK);
    CHECK(st_synchronize(ctx) == ST_OK && d_h.down(h.data()) == ST_OK);
    CHECK(h[0] == 1.f && h[(size_t)dim * dim - 1] == 1.f / (float)(2 * dim - 1));
    uint itr = 0;
    float lam = 0.f;
    int64_t tm = similarity_transform(q, h.data(), &lam, ev.data(), dim, dim >> 1, &itr);
    CHECK(tm >= 0 && itr == 13); // README.md:73
    std::printf("%-5ux%5u\t\t\t%10ld ms\t\t\t%6u round(s)   lambda = %.7f\n", dim, dim, (long)tm, itr, lam);

    // ---- the same matrix streamed through a device cache of a quarter of its rows (st::Context::solve_streamed):
    //      the step before the path in the reference is the whole-matrix copy-in (similarity_transform.cpp:14-19) ----
    st_options opt;
    st_default_options(&opt);
    st_result res{};
    st_stream_plan plan{};
    std::vector<float> ev_streamed(dim);
    float lam_streamed = 0.f;
    CHECK(q.solve_streamed(h.data(), dim, opt, sizeof(float) * dim * (dim / 4), 64, &lam_streamed, ev_streamed.data(), &res,
                           &plan) == ST_OK);
    CHECK(res.iter_count == 13 && plan.streamed == 1 && plan.blocks == 16 && plan.slots == 4);
    CHECK(plan.h2d_bytes_first == sizeof(float) * dim * dim && plan.h2d_bytes_per_round == sizeof(float) * dim * 64 * 12);
    CHECK(lam_streamed == lam);
    for (uint i = 0; i < dim; i++)
      CHECK(ev_streamed[i] == ev[i]);
    std::printf("streamed solve: same bits, %u blocks through %u slots, %.1f MiB over PCIe per later round\n", plan.blocks,
                plan.slots, plan.h2d_bytes_per_round / 1048576.0);

    // ---- every GPU of the box behind the one handle (st_group_attach), when there is more than one ----
    if (st_device_count() > 1) {
      CHECK(st_group_attach(ctx, nullptr, 0, 256) == ST_OK && st_group_size(ctx) == st_device_count());
      std::vector<float> ev_group(dim);
      float lam_group = 0.f;
      uint itr_group = 0;
      CHECK(max_eigen_value(ctx, h.data(), &lam_group, ev_group.data(), dim, &itr_group) >= 0);
      CHECK(itr_group == 13 && lam_group == lam);
      for (uint i = 0; i < dim; i++)
        CHECK(ev_group[i] == ev[i]);
      CHECK(st_group_detach(ctx) == ST_OK && st_group_size(ctx) == 1);
      std::printf("device group of %d GPUs behind max_eigen_value: same bits\n", st_device_count());
    }
  }
  std::printf("\nall checks passed\n");
  return 0;
}
