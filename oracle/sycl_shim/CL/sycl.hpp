// CL/sycl.hpp -- a single-threaded CPU stand-in for the slice of SYCL 2020 / oneAPI DPC++ that
// itzmeanjan/eigen_value uses, so that the reference's UNMODIFIED sources
// (similarity_transform.cpp, utils.cpp, wrapper/similarity_transform.cpp) compile with g++ and run
// here.  TEST INFRASTRUCTURE ONLY (oracle/_ref); nothing in the product links against it.
//
// Execution model: queue::submit runs the command group synchronously.  parallel_for executes the
// work-groups one after another; inside a group every work-item is a fiber (own stack, hand-made
// context switch), scheduled round-robin.  group_barrier, sub_group::barrier and every sub-group
// collective are rendezvous points: an item deposits its value, yields, and continues once the
// whole (sub-)group has arrived.  Sub-groups are 32 consecutive local ids (the reference pins
// reqd_sub_group_size(32)).  Atomics degenerate to plain read-modify-write; float atomic adds thus
// happen in one fixed, legal order (sub-group 0, 1, 2 ... then work-group 0, 1, 2 ...).
#pragma once

#include <sys/types.h>

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

extern "C" void shim_switch(void** save_sp, void* load_sp); // oracle/sycl_shim/ref_abi.cpp

namespace sycl {

// ---------------------------------------------------------------------------------------------
// index types
// ---------------------------------------------------------------------------------------------
template<int D>
struct range
{
  size_t v[D > 0 ? D : 1] = {};
  range() = default;
  range(size_t a) { v[0] = a; }
  range(size_t a, size_t b)
  {
    static_assert(D >= 2, "two extents need two dimensions");
    v[0] = a;
    v[1] = b;
  }
  size_t operator[](int i) const { return v[i]; }
  size_t& operator[](int i) { return v[i]; }
  size_t get(int i) const { return v[i]; }
  size_t size() const
  {
    size_t s = 1;
    for (int i = 0; i < D; i++)
      s *= v[i];
    return s;
  }
};
template<int D>
using id = range<D>;

template<int D>
struct nd_range
{
  range<D> global, local;
  nd_range(range<D> g, range<D> l)
    : global(g)
    , local(l)
  {}
};

struct event
{
  void wait() {}
};

// ---------------------------------------------------------------------------------------------
// enums, tags
// ---------------------------------------------------------------------------------------------
namespace access {
enum class mode
{
  read,
  write,
  read_write
};
enum class target
{
  global_buffer,
  local,
  device
};
enum class address_space
{
  global_space,
  local_space
};
} // namespace access
using access_mode = access::mode;

enum class memory_scope
{
  work_item,
  sub_group,
  work_group,
  device,
  system
};
enum class memory_order
{
  relaxed,
  acquire,
  release,
  acq_rel,
  seq_cst
};

struct no_init_t
{};
inline constexpr no_init_t no_init{};

template<typename T = void>
struct plus
{
  T operator()(const T& a, const T& b) const { return a + b; }
  static T identity() { return T(0); }
};
template<typename T = void>
struct maximum
{
  T operator()(const T& a, const T& b) const { return a < b ? b : a; }
  static T identity() { return std::numeric_limits<T>::lowest(); }
};
template<typename T = void>
struct minimum
{
  T operator()(const T& a, const T& b) const { return b < a ? b : a; }
  static T identity() { return std::numeric_limits<T>::max(); }
};

inline float
abs(float x)
{
  return std::fabs(x);
}
inline double
abs(double x) // tests/test.cpp:99-102 calls abs(float - double) under `using namespace sycl`
{
  return std::fabs(x);
}
inline int
abs(int x)
{
  return x < 0 ? -x : x;
}

// ---------------------------------------------------------------------------------------------
// device / context / queue
// ---------------------------------------------------------------------------------------------
struct default_selector
{};

namespace info {
namespace device {
struct max_work_group_size
{
  using return_type = size_t;
};
struct name
{
  using return_type = std::string;
};
} // namespace device
} // namespace info

constexpr size_t kShimMaxWorkGroup = 256;

struct device
{
  device() = default;
  device(const default_selector&) {}
  template<typename I>
  typename I::return_type get_info() const
  {
    if constexpr (std::is_same_v<I, info::device::max_work_group_size>)
      return kShimMaxWorkGroup;
    else
      return std::string("sycl-shim single-thread CPU (oracle/sycl_shim)");
  }
};

struct context
{
  context() = default;
  context(const device&) {}
};

// ---------------------------------------------------------------------------------------------
// fibers and the per-work-group rendezvous state
// ---------------------------------------------------------------------------------------------
namespace detail {

constexpr size_t kSubGroup = 32;
constexpr size_t kStackBytes = 32 * 1024;

struct Fiber
{
  void* sp = nullptr;
  bool done = false;
  size_t local_linear = 0;
  std::unique_ptr<unsigned char[]> stack;
};

struct Rendezvous
{
  size_t size = 0, arrived = 0, gen = 0;
  unsigned char vals[kSubGroup][8];
  unsigned char res[kSubGroup][8];
};

struct GroupRun
{
  void (*invoke)(void* kernel_ctx, size_t local_linear) = nullptr;
  void* kernel_ctx = nullptr;
  size_t n_items = 0;
  Rendezvous group;                // group_barrier
  std::vector<Rendezvous> sub;     // one per sub-group
  std::vector<Fiber> fibers;
  void* sched_sp = nullptr;
  Fiber* current = nullptr;
};

GroupRun& run(); // the one active work-group (ref_abi.cpp)

inline void
yield()
{
  GroupRun& r = run();
  shim_switch(&r.current->sp, r.sched_sp);
}

extern "C" void shim_fiber_entry(); // ref_abi.cpp: runs run().invoke for run().current, then parks

inline void
prepare(Fiber& f)
{
  if (!f.stack)
    f.stack.reset(new unsigned char[kStackBytes]);
  uintptr_t top = reinterpret_cast<uintptr_t>(f.stack.get()) + kStackBytes;
  top &= ~uintptr_t(15);
  void** sp = reinterpret_cast<void**>(top);
  *--sp = nullptr;                                   // keeps (rsp % 16 == 8) at entry
  *--sp = reinterpret_cast<void*>(&shim_fiber_entry); // return address taken by shim_switch's ret
  for (int i = 0; i < 6; i++)
    *--sp = nullptr;                                 // r15 r14 r13 r12 rbx rbp
  f.sp = sp;
  f.done = false;
}

// run one work-group of n items to completion
inline void
run_group(size_t n, void (*invoke)(void*, size_t), void* ctx)
{
  GroupRun& r = run();
  r.invoke = invoke;
  r.kernel_ctx = ctx;
  r.n_items = n;
  if (r.fibers.size() < n)
    r.fibers.resize(n);
  const size_t nsub = (n + kSubGroup - 1) / kSubGroup;
  r.sub.assign(nsub, Rendezvous{});
  for (size_t s = 0; s < nsub; s++)
    r.sub[s].size = std::min(kSubGroup, n - s * kSubGroup);
  r.group = Rendezvous{};
  r.group.size = n;
  for (size_t i = 0; i < n; i++) {
    r.fibers[i].local_linear = i;
    prepare(r.fibers[i]);
  }
  size_t live = n;
  while (live) {
    for (size_t i = 0; i < n; i++) {
      Fiber& f = r.fibers[i];
      if (f.done)
        continue;
      r.current = &f;
      shim_switch(&r.sched_sp, f.sp);
      if (f.done)
        live--;
    }
  }
  r.current = nullptr;
}

// generic rendezvous: deposit `v`, wait for everybody, `finish` turns vals[] into res[] once
template<typename T, typename Finish>
inline T
rendezvous(Rendezvous& z, size_t lane, const T& v, Finish&& finish)
{
  static_assert(sizeof(T) <= 8, "shim collectives carry at most 8 bytes");
  const size_t gen = z.gen;
  std::memcpy(z.vals[lane % kSubGroup], &v, sizeof(T));
  if (++z.arrived == z.size) {
    finish(z);
    z.arrived = 0;
    z.gen++;
  }
  while (z.gen == gen)
    yield();
  T out;
  std::memcpy(&out, z.res[lane % kSubGroup], sizeof(T));
  return out;
}

// barrier over more than 32 items: no payload
inline void
barrier(Rendezvous& z)
{
  const size_t gen = z.gen;
  if (++z.arrived == z.size) {
    z.arrived = 0;
    z.gen++;
  }
  while (z.gen == gen)
    yield();
}

} // namespace detail

// ---------------------------------------------------------------------------------------------
// groups and items
// ---------------------------------------------------------------------------------------------
struct sub_group
{
  size_t sg_id = 0, lane = 0, sz = 0;
  range<1> get_local_id() const { return range<1>(lane); }
  range<1> get_local_range() const { return range<1>(sz); }
  size_t get_local_linear_id() const { return lane; }
  bool leader() const { return lane == 0; }
  detail::Rendezvous& state() const { return detail::run().sub[sg_id]; }
  void barrier() const
  {
    int dummy = 0;
    detail::rendezvous<int>(state(), lane, dummy, [](detail::Rendezvous&) {});
  }
  template<typename T>
  T shuffle_down(const T& x, size_t delta) const
  {
    const size_t n = sz;
    return detail::rendezvous<T>(state(), lane, x, [n, delta](detail::Rendezvous& z) {
      for (size_t l = 0; l < n; l++)
        std::memcpy(z.res[l], z.vals[l + delta < n ? l + delta : l], 8);
    });
  }
};

template<int D>
struct group
{
  range<D> gid, lrange;
  size_t local_linear = 0;
  size_t get_id(int d) const { return gid[d]; }
  size_t get_group_id(int d) const { return gid[d]; }
  range<D> get_local_range() const { return lrange; }
  bool leader() const { return local_linear == 0; }
};

template<int D>
struct nd_item
{
  range<D> gid, lid, grange, lrange, group_id;
  size_t local_linear = 0;

  size_t get_global_id(int d) const { return gid[d]; }
  range<D> get_global_id() const { return gid; }
  size_t get_local_id(int d) const { return lid[d]; }
  size_t get_global_range(int d) const { return grange[d]; }
  size_t get_local_range(int d) const { return lrange[d]; }
  size_t get_local_linear_id() const { return local_linear; }
  size_t get_global_linear_id() const
  {
    size_t lin = 0;
    for (int d = 0; d < D; d++)
      lin = lin * grange[d] + gid[d];
    return lin;
  }
  group<D> get_group() const { return group<D>{ group_id, lrange, local_linear }; }
  sub_group get_sub_group() const
  {
    const size_t n = lrange.size();
    const size_t s = local_linear / detail::kSubGroup;
    return sub_group{ s, local_linear % detail::kSubGroup, std::min(detail::kSubGroup, n - s * detail::kSubGroup) };
  }
  void barrier() const { detail::barrier(detail::run().group); }
};

template<int D>
inline void
group_barrier(const group<D>&, memory_scope = memory_scope::work_group)
{
  detail::barrier(detail::run().group);
}
inline void
group_barrier(const sub_group& sg, memory_scope = memory_scope::sub_group)
{
  sg.barrier();
}

// butterfly order: slot l += slot l+w for w = 16, 8, 4, 2, 1 (slots beyond the sub-group hold the
// operator's identity) -- the order oracle.c's ORACLE_SUM_WORKGROUP mode restates
template<typename T, typename Op>
inline T
reduce_over_group(const sub_group& sg, const T& x, Op op)
{
  const size_t n = sg.sz;
  return detail::rendezvous<T>(sg.state(), sg.lane, x, [n, op](detail::Rendezvous& z) {
    T v[detail::kSubGroup];
    for (size_t l = 0; l < detail::kSubGroup; l++) {
      if (l < n)
        std::memcpy(&v[l], z.vals[l], sizeof(T));
      else
        v[l] = Op::identity();
    }
    for (size_t w = detail::kSubGroup / 2; w >= 1; w >>= 1)
      for (size_t l = 0; l < w; l++)
        v[l] = op(v[l], v[l + w]);
    for (size_t l = 0; l < n; l++)
      std::memcpy(z.res[l], &v[0], sizeof(T));
  });
}

template<typename T>
inline T
group_broadcast(const sub_group& sg, const T& x, size_t src = 0)
{
  const size_t n = sg.sz;
  return detail::rendezvous<T>(sg.state(), sg.lane, x, [n, src](detail::Rendezvous& z) {
    for (size_t l = 0; l < n; l++)
      std::memcpy(z.res[l], z.vals[src], 8);
  });
}

inline bool
all_of_group(const sub_group& sg, bool pred)
{
  const size_t n = sg.sz;
  const int v = pred ? 1 : 0;
  return detail::rendezvous<int>(sg.state(), sg.lane, v, [n](detail::Rendezvous& z) {
           int all = 1;
           for (size_t l = 0; l < n; l++) {
             int p;
             std::memcpy(&p, z.vals[l], sizeof p);
             all &= p;
           }
           for (size_t l = 0; l < n; l++)
             std::memcpy(z.res[l], &all, sizeof all);
         }) != 0;
}

// ---------------------------------------------------------------------------------------------
// buffers, accessors, handler, queue
// ---------------------------------------------------------------------------------------------
template<typename T, int D = 1>
struct buffer
{
  T* ptr = nullptr; // aliases the host allocation: the write-back at destruction is a no-op
  range<D> r;
  buffer(T* host, range<D> rg)
    : ptr(host)
    , r(rg)
  {}
  range<D> get_range() const { return r; }
};

struct handler;

template<typename T, int D, access::mode M = access::mode::read_write,
         access::target Tg = access::target::global_buffer>
struct accessor;

template<typename T, access::mode M>
struct accessor<T, 1, M, access::target::global_buffer>
{
  T* ptr = nullptr;
  range<1> r;
  accessor(buffer<T, 1>& b, handler&)
    : ptr(b.ptr)
    , r(b.r)
  {}
  accessor(buffer<T, 1>& b, handler&, no_init_t)
    : ptr(b.ptr)
    , r(b.r)
  {}
  accessor(buffer<T, 1>& b, handler&, range<1> sub)
    : ptr(b.ptr)
    , r(sub)
  {}
  T& operator[](size_t i) const { return ptr[i]; }
  T& operator[](const id<1>& i) const { return ptr[i[0]]; }
  size_t size() const { return r[0]; }
};

template<typename T, access::mode M>
struct accessor<T, 2, M, access::target::global_buffer>
{
  T* ptr = nullptr;
  range<2> r;
  accessor(buffer<T, 2>& b, handler&)
    : ptr(b.ptr)
    , r(b.r)
  {}
  accessor(buffer<T, 2>& b, handler&, no_init_t)
    : ptr(b.ptr)
    , r(b.r)
  {}
  struct row
  {
    T* p;
    T& operator[](size_t c) const { return p[c]; }
  };
  row operator[](size_t i) const { return row{ ptr + i * r[1] }; }
  size_t size() const { return r.size(); }
};

// local memory: one allocation shared by all copies of the accessor; work-groups run one after
// another, and every reference kernel re-initialises its local memory behind a barrier
template<typename T, access::mode M>
struct accessor<T, 1, M, access::target::local>
{
  std::shared_ptr<std::vector<T>> mem;
  accessor(range<1> n, handler&)
    : mem(std::make_shared<std::vector<T>>(n[0]))
  {}
  T& operator[](size_t i) const { return (*mem)[i]; }
};

template<typename T, int D = 1, access::mode M = access::mode::read_write>
struct host_accessor
{
  T* ptr;
  host_accessor(buffer<T, D>& b)
    : ptr(b.ptr)
  {}
  T& operator[](size_t i) const { return ptr[i]; }
};

struct handler
{
  template<typename V>
  void depends_on(const V&)
  {}
  template<typename Acc, typename T>
  void fill(Acc acc, const T& value)
  {
    for (size_t i = 0; i < acc.size(); i++)
      acc.ptr[i] = value;
  }
  template<typename Src, typename Dst>
  void copy(Src src, Dst dst)
  {
    for (size_t i = 0; i < src.size(); i++)
      dst.ptr[i] = src.ptr[i];
  }

  template<int D, typename K>
  struct Launch
  {
    const K* kernel;
    range<D> grange, lrange, group_id;
    static void invoke(void* ctx, size_t local_linear)
    {
      Launch* self = static_cast<Launch*>(ctx);
      nd_item<D> it;
      it.grange = self->grange;
      it.lrange = self->lrange;
      it.group_id = self->group_id;
      it.local_linear = local_linear;
      size_t rem = local_linear;
      for (int d = D - 1; d >= 0; d--) {
        it.lid[d] = rem % self->lrange[d];
        rem /= self->lrange[d];
        it.gid[d] = self->group_id[d] * self->lrange[d] + it.lid[d];
      }
      (*self->kernel)(it);
    }
  };

  template<typename Name = void, int D, typename K>
  void parallel_for(nd_range<D> ndr, const K& kernel)
  {
    for (int d = 0; d < D; d++)
      if (ndr.local[d] == 0 || ndr.global[d] % ndr.local[d] != 0)
        throw std::runtime_error("sycl-shim: global range is not a multiple of the work-group size");
    if (ndr.local.size() > 4096)
      throw std::runtime_error("sycl-shim: work-group too large");
    Launch<D, K> launch{ &kernel, ndr.global, ndr.local, range<D>() };
    range<D> groups;
    for (int d = 0; d < D; d++)
      groups[d] = ndr.global[d] / ndr.local[d];
    const size_t total = groups.size();
    for (size_t g = 0; g < total; g++) {
      size_t rem = g;
      for (int d = D - 1; d >= 0; d--) {
        launch.group_id[d] = rem % groups[d];
        rem /= groups[d];
      }
      detail::run_group(ndr.local.size(), &Launch<D, K>::invoke, &launch);
    }
  }
};

struct queue
{
  device dev;
  queue() = default;
  queue(const device& d)
    : dev(d)
  {}
  queue(const context&, const device& d)
    : dev(d)
  {}
  device get_device() const { return dev; }
  template<typename F>
  event submit(F&& f)
  {
    handler h;
    f(h);
    return event{};
  }
  void wait() {}
};

// ---------------------------------------------------------------------------------------------
// oneAPI extension namespace as DPC++ 2021.4 spelled it
// ---------------------------------------------------------------------------------------------
namespace ext {
namespace oneapi {
using sub_group = ::sycl::sub_group;
enum class memory_order
{
  relaxed,
  acq_rel,
  seq_cst
};
enum class memory_scope
{
  work_item,
  sub_group,
  work_group,
  device,
  system
};
template<int D>
inline bool
leader(const group<D>& g)
{
  return g.leader();
}
inline bool
leader(const ::sycl::sub_group& sg)
{
  return sg.leader();
}

template<typename T, memory_order, memory_scope, access::address_space>
struct atomic_ref
{
  T& ref;
  explicit atomic_ref(T& r)
    : ref(r)
  {}
  T fetch_add(T v)
  {
    T old = ref;
    ref = old + v;
    return old;
  }
  T fetch_max(T v)
  {
    T old = ref;
    ref = old < v ? v : old;
    return old;
  }
  T fetch_min(T v)
  {
    T old = ref;
    ref = v < old ? v : old;
    return old;
  }
};
} // namespace oneapi
} // namespace ext

} // namespace sycl
