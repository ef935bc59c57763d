// Minimal thrust::device_vector / thrust::device_ptr for the sort front door of this repo (include/thrust/sort.h).
//
// Only what the radix-sort path of the reference touches is provided: contiguous device storage, begin()/end()
// iterators that unwrap to raw device pointers (the reference's sort unwraps with thrust::raw_pointer_cast /
// try_unwrap_contiguous_iterator before it reaches cub::DeviceRadixSort, thrust/system/cuda/detail/sort.h:288-339),
// and host <-> device assignment.  This is NOT a general Thrust container.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <iterator>
#include <stdexcept>
#include <string>
#include <vector>

namespace thrust
{

struct system_error : std::runtime_error
{
  cudaError_t code;
  system_error(cudaError_t e, const std::string& what)
      : std::runtime_error(what + ": " + cudaGetErrorString(e))
      , code(e)
  {}
};

namespace detail
{
inline void throw_on_error(cudaError_t e, const char* what)
{
  if (e != cudaSuccess)
  {
    cudaGetLastError();
    throw system_error(e, what);
  }
}
} // namespace detail

/// Typed device pointer; random-access iterator over device memory (not dereferenceable on the host).
template <class T>
struct device_ptr
{
  using iterator_category = std::random_access_iterator_tag;
  using value_type        = T;
  using difference_type   = std::ptrdiff_t;
  using pointer           = T*;
  using reference         = T&;

  T* p = nullptr;
  device_ptr() = default;
  explicit device_ptr(T* raw)
      : p(raw)
  {}
  T* get() const
  {
    return p;
  }
  device_ptr operator+(difference_type n) const
  {
    return device_ptr(p + n);
  }
  device_ptr operator-(difference_type n) const
  {
    return device_ptr(p - n);
  }
  difference_type operator-(const device_ptr& o) const
  {
    return p - o.p;
  }
  device_ptr& operator+=(difference_type n)
  {
    p += n;
    return *this;
  }
  device_ptr& operator++()
  {
    ++p;
    return *this;
  }
  bool operator==(const device_ptr& o) const
  {
    return p == o.p;
  }
  bool operator!=(const device_ptr& o) const
  {
    return p != o.p;
  }
  bool operator<(const device_ptr& o) const
  {
    return p < o.p;
  }
};

template <class T>
device_ptr<T> device_pointer_cast(T* raw)
{
  return device_ptr<T>(raw);
}
template <class T>
T* raw_pointer_cast(const device_ptr<T>& d)
{
  return d.get();
}
template <class T>
T* raw_pointer_cast(T* raw)
{
  return raw;
}

template <class T>
class device_vector
{
public:
  using value_type = T;
  using iterator   = device_ptr<T>;

  device_vector() = default;
  explicit device_vector(size_t n)
  {
    resize_uninitialised(n);
    if (n)
    {
      detail::throw_on_error(cudaMemset(d_, 0, n * sizeof(T)), "device_vector: memset");
    }
  }
  device_vector(const std::vector<T>& h)
  {
    *this = h;
  }
  device_vector(const device_vector& o)
  {
    resize_uninitialised(o.n_);
    if (n_)
    {
      detail::throw_on_error(cudaMemcpy(d_, o.d_, n_ * sizeof(T), cudaMemcpyDeviceToDevice), "device_vector: copy");
    }
  }
  device_vector(device_vector&& o) noexcept
      : d_(o.d_)
      , n_(o.n_)
  {
    o.d_ = nullptr;
    o.n_ = 0;
  }
  device_vector& operator=(const std::vector<T>& h)
  {
    resize_uninitialised(h.size());
    if (n_)
    {
      detail::throw_on_error(cudaMemcpy(d_, h.data(), n_ * sizeof(T), cudaMemcpyHostToDevice), "device_vector: H2D");
    }
    return *this;
  }
  device_vector& operator=(device_vector o)
  {
    std::swap(d_, o.d_);
    std::swap(n_, o.n_);
    return *this;
  }
  ~device_vector()
  {
    if (d_)
    {
      cudaFree(d_);
    }
  }

  size_t size() const
  {
    return n_;
  }
  bool empty() const
  {
    return n_ == 0;
  }
  iterator begin() const
  {
    return iterator(d_);
  }
  iterator end() const
  {
    return iterator(d_ + n_);
  }
  device_ptr<T> data() const
  {
    return iterator(d_);
  }
  /// copy to the host (thrust::host_vector<T> h = d; in the reference)
  std::vector<T> to_host() const
  {
    std::vector<T> h(n_);
    if (n_)
    {
      detail::throw_on_error(cudaMemcpy(h.data(), d_, n_ * sizeof(T), cudaMemcpyDeviceToHost), "device_vector: D2H");
    }
    return h;
  }
  operator std::vector<T>() const
  {
    return to_host();
  }

private:
  void resize_uninitialised(size_t n)
  {
    if (d_)
    {
      cudaFree(d_);
      d_ = nullptr;
    }
    n_ = n;
    if (n)
    {
      detail::throw_on_error(cudaMalloc(reinterpret_cast<void**>(&d_), n * sizeof(T)), "device_vector: cudaMalloc");
    }
  }
  T* d_     = nullptr;
  size_t n_ = 0;
};

} // namespace thrust
