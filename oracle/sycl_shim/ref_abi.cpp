// ref_abi.cpp -- runtime half of the CPU SYCL shim plus a plain C surface over the reference's
// own functions, linked together with the UNMODIFIED reference sources into
// oracle/_ref/libreference_cpu.so (recipe: oracle/Makefile, target `ref`).
// TEST INFRASTRUCTURE ONLY: used to validate oracle/oracle.c and to generate tests/golden/.
#include "similarity_transform.hpp" // the reference's header (-I$(REFERENCE)/include)
#include "utils.hpp"

// ---- context switch: save callee-saved registers, swap stacks ------------------------------
asm(R"(
.text
.globl shim_switch
.type shim_switch,@function
shim_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size shim_switch,.-shim_switch
)");

namespace sycl {
namespace detail {
GroupRun&
run()
{
  static GroupRun r;
  return r;
}
} // namespace detail
} // namespace sycl

extern "C" void
shim_fiber_entry()
{
  sycl::detail::GroupRun& r = sycl::detail::run();
  sycl::detail::Fiber* self = r.current;
  r.invoke(r.kernel_ctx, self->local_linear);
  self->done = true;
  shim_switch(&self->sp, r.sched_sp); // park; never resumed
  abort();
}

// ---- C surface ----------------------------------------------------------------------------------
extern "C" void make_queue(void** wq); // reference wrapper/similarity_transform.cpp:3-12
extern "C" int64_t max_eigen_value(void* wq, float* mat, float* eigen_val, float* eigen_vec, uint dim,
                                   uint* iter_cnt); // reference wrapper/similarity_transform.cpp:14-37

static sycl::queue&
the_queue()
{
  static sycl::queue q{ sycl::device{ sycl::default_selector{} } };
  return q;
}

extern "C" {

// the Python wrapper's path: make_queue + max_eigen_value, wg_size = min(dim >> 1, max_wg)
int64_t
ref_max_eigen_value(float* mat, float* eigen_val, float* eigen_vec, uint dim, uint* iter_cnt)
{
  static void* wq = nullptr;
  if (!wq)
    make_queue(&wq);
  try {
    return max_eigen_value(wq, mat, eigen_val, eigen_vec, dim, iter_cnt);
  } catch (const std::exception&) {
    return -1;
  }
}

// similarity_transform() with an explicit work-group size (tests/test.cpp:96-97 uses wg_size = 3)
int64_t
ref_similarity_transform(const float* mat, float* eigen_val, float* eigen_vec, uint dim, uint wg_size,
                         uint* iter_count)
{
  try {
    return similarity_transform(the_queue(), mat, eigen_val, eigen_vec, dim, wg_size, iter_count);
  } catch (const std::exception&) {
    return -1;
  }
}

int
ref_sum_across_rows(float* mat, float* vec, uint dim, uint wg_size)
{
  try {
    buffer_2d b_mat{ mat, sycl::range<2>{ dim, dim } };
    buffer_1d b_vec{ vec, sycl::range<1>{ dim } };
    sum_across_rows(the_queue(), b_mat, b_vec, dim, wg_size, {}).wait();
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}

int
ref_find_max(float* vec, float* max, uint dim, uint wg_size)
{
  try {
    buffer_1d b_vec{ vec, sycl::range<1>{ dim } };
    buffer_1d b_max{ max, sycl::range<1>{ 1 } };
    find_max(the_queue(), b_vec, b_max, dim, wg_size, {}).wait();
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}

int
ref_compute_eigen_vector(float* vec, float* max, float* eigen_vec, uint dim, uint wg_size)
{
  try {
    buffer_1d b_vec{ vec, sycl::range<1>{ dim } };
    buffer_1d b_max{ max, sycl::range<1>{ 1 } };
    buffer_1d b_e{ eigen_vec, sycl::range<1>{ dim } };
    compute_eigen_vector(the_queue(), b_vec, b_max, b_e, dim, wg_size, {}).wait();
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}

int
ref_initialise_eigen_vector(float* eigen_vec, uint dim)
{
  buffer_1d b_e{ eigen_vec, sycl::range<1>{ dim } };
  initialise_eigen_vector(the_queue(), b_e, dim, {}).wait();
  return 0;
}

int
ref_compute_next_matrix(float* mat, float* vec, uint dim, uint wg_size)
{
  try {
    buffer_2d b_mat{ mat, sycl::range<2>{ dim, dim } };
    buffer_1d b_vec{ vec, sycl::range<1>{ dim } };
    compute_next_matrix(the_queue(), b_mat, b_vec, dim, wg_size, {}).wait();
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}

int
ref_stop(float* vec, uint* ret, uint dim, uint wg_size)
{
  try {
    buffer_1d b_vec{ vec, sycl::range<1>{ dim } };
    sycl::buffer<uint, 1> b_ret{ ret, sycl::range<1>{ 1 } };
    stop(the_queue(), b_vec, b_ret, dim, wg_size, {}).wait();
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}

// reference utils.cpp fixtures and generator
int
ref_generate_hilbert_matrix(float* mat, uint dim)
{
  try {
    generate_hilbert_matrix(the_queue(), mat, dim);
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}
int
ref_identity_matrix(float* mat, uint dim, uint wg_size)
{
  identity_matrix(the_queue(), mat, dim, wg_size, {}).wait();
  return 0;
}
int
ref_generate_vector(float* vec, uint dim, uint wg_size)
{
  generate_vector(the_queue(), vec, dim, wg_size, {}).wait();
  return 0;
}
int
ref_stop_criteria_test_success_data(float* vec, uint dim, uint wg_size)
{
  stop_criteria_test_success_data(the_queue(), vec, dim, wg_size, {}).wait();
  return 0;
}
int
ref_stop_criteria_test_fail_data(float* vec, uint dim, uint wg_size)
{
  stop_criteria_test_fail_data(the_queue(), vec, dim, wg_size, {}).wait();
  return 0;
}
uint
ref_max_work_group_size(void)
{
  return (uint)sycl::kShimMaxWorkGroup;
}

} // extern "C"
