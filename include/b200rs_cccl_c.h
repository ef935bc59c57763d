/*
 * b200rs_cccl_c.h -- the `cccl.c.parallel` radix-sort entry points, exported by libb200rs.so with the reference's
 * names and binary layout, so that a host built against the reference's own headers
 * (/root/reference/c/parallel/include/cccl/c/radix_sort.h:63-143, types.h:31-168; implementation being replaced:
 * c/parallel/src/radix_sort.cu:232-700) can link -lb200rs instead of libcccl.c.parallel.so -- this is what
 * cuda.compute's `make_radix_sort` binds (python/cuda_cccl/cuda/compute/_bindings_impl.pyx).
 *
 * This header only restates the layouts those calls take (same field order, same enumerator values); use it, or the
 * reference's headers, interchangeably.  Differences in behaviour, all on the build side:
 *   - cccl_device_radix_sort_build / _build_ex / _compile compile nothing (the kernels are ahead-of-time sm_100a code):
 *     they record key type, value type and order; the path / config arguments are ignored; _load is a no-op;
 *   - keys must be arithmetic (CCCL_INT8 .. CCCL_FLOAT64, CCCL_BOOLEAN); CCCL_STORAGE keys (user-defined types with a
 *     decomposer) return CUDA_ERROR_NOT_SUPPORTED; the decomposer argument is ignored for arithmetic keys, as in
 *     the reference (identity decomposer);
 *   - iterators must be pointers (the reference has the same restriction, radix_sort.cu:598-606);
 *   - serialize / deserialize round-trip the recorded fields; link_ltoir returns CUDA_ERROR_NOT_SUPPORTED.
 * cccl_device_radix_sort itself is b200rs_sort: same temp-storage protocol, selector and stream semantics.
 */
#ifndef B200RS_CCCL_C_H_
#define B200RS_CCCL_C_H_

#include <cuda.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef CCCL_C_API
#  define CCCL_C_API __attribute__((__visibility__("default")))
#endif

/* ---- cccl/c/types.h, the part these calls use (skipped when the reference's header was included first) */
#ifndef B200RS_HAVE_CCCL_C_TYPES
typedef enum cccl_type_enum { CCCL_INT8 = 0, CCCL_INT16 = 1, CCCL_INT32 = 2, CCCL_INT64 = 3, CCCL_UINT8 = 4,
  CCCL_UINT16 = 5, CCCL_UINT32 = 6, CCCL_UINT64 = 7, CCCL_FLOAT16 = 8, CCCL_FLOAT32 = 9, CCCL_FLOAT64 = 10,
  CCCL_STORAGE = 11, CCCL_BOOLEAN = 12 } cccl_type_enum;
typedef struct cccl_type_info { size_t size; size_t alignment; cccl_type_enum type; } cccl_type_info;
typedef enum cccl_op_kind_t { CCCL_STATELESS = 0, CCCL_STATEFUL = 1, CCCL_IDENTITY = 20 /* others unused here */ }
  cccl_op_kind_t;
typedef enum cccl_op_code_type { CCCL_OP_LTOIR = 0, CCCL_OP_CPP_SOURCE = 1 } cccl_op_code_type;
typedef struct cccl_op_t { cccl_op_kind_t type; const char* name; const char* code; size_t code_size;
  cccl_op_code_type code_type; size_t size; size_t alignment; void* state; const char** extra_ltoirs;
  size_t* extra_ltoir_sizes; size_t num_extra_ltoirs; cccl_op_code_type* extra_code_types; } cccl_op_t;
typedef struct cccl_build_config { const char** extra_compile_flags; size_t num_extra_compile_flags;
  const char** extra_include_dirs; size_t num_extra_include_dirs; } cccl_build_config;
typedef enum cccl_iterator_kind_t { CCCL_POINTER = 0, CCCL_ITERATOR = 1 } cccl_iterator_kind_t;
typedef union { int64_t signed_offset; uint64_t unsigned_offset; } cccl_increment_t;
typedef void (*cccl_host_op_fn_ptr_t)(void*, cccl_increment_t);
/* type == CCCL_POINTER: `state` IS the device pointer */
typedef struct cccl_iterator_t { size_t size; size_t alignment; cccl_iterator_kind_t type; cccl_op_t advance;
  cccl_op_t dereference; cccl_type_info value_type; void* state; cccl_host_op_fn_ptr_t host_advance; } cccl_iterator_t;
typedef enum cccl_sort_order_t { CCCL_ASCENDING = 0, CCCL_DESCENDING = 1 } cccl_sort_order_t;
typedef enum cccl_payload_kind_t { CCCL_PAYLOAD_LTOIR = 0, CCCL_PAYLOAD_CUBIN = 1 } cccl_payload_kind_t;
#endif

/* ---- cccl/c/radix_sort.h:27-61: the kernel handles / lowered names stay null here (nothing is JIT-compiled) */
typedef struct cccl_device_radix_sort_build_result_t
{
  int cc; void* payload; size_t payload_size; cccl_payload_kind_t payload_kind; CUlibrary library;
  cccl_type_info key_type; cccl_type_info value_type;
  CUkernel single_tile_kernel, upsweep_kernel, alt_upsweep_kernel, scan_bins_kernel, downsweep_kernel,
    alt_downsweep_kernel, histogram_kernel, exclusive_sum_kernel, init_bins_and_counters_kernel, init_lookback_kernel,
    onesweep_kernel;
  cccl_sort_order_t order; void* runtime_policy; size_t runtime_policy_size;
  char *single_tile_kernel_lowered_name, *upsweep_kernel_lowered_name, *alt_upsweep_kernel_lowered_name,
    *scan_bins_kernel_lowered_name, *downsweep_kernel_lowered_name, *alt_downsweep_kernel_lowered_name,
    *histogram_kernel_lowered_name, *exclusive_sum_kernel_lowered_name, *init_bins_and_counters_kernel_lowered_name,
    *init_lookback_kernel_lowered_name, *onesweep_kernel_lowered_name;
} cccl_device_radix_sort_build_result_t;

CCCL_C_API CUresult cccl_device_radix_sort_build(cccl_device_radix_sort_build_result_t* build, cccl_sort_order_t sort_order,
  cccl_iterator_t input_keys_it, cccl_iterator_t input_values_it, cccl_op_t decomposer, const char* decomposer_return_type,
  int cc_major, int cc_minor, const char* cub_path, const char* thrust_path, const char* libcudacxx_path,
  const char* ctk_path);
CCCL_C_API CUresult cccl_device_radix_sort_build_ex(cccl_device_radix_sort_build_result_t* build,
  cccl_sort_order_t sort_order, cccl_iterator_t input_keys_it, cccl_iterator_t input_values_it, cccl_op_t decomposer,
  const char* decomposer_return_type, int cc_major, int cc_minor, const char* cub_path, const char* thrust_path,
  const char* libcudacxx_path, const char* ctk_path, cccl_build_config* config);
CCCL_C_API CUresult cccl_device_radix_sort_compile(cccl_device_radix_sort_build_result_t* build,
  cccl_sort_order_t sort_order, cccl_iterator_t input_keys_it, cccl_iterator_t input_values_it, cccl_op_t decomposer,
  const char* decomposer_return_type, int cc_major, int cc_minor, const char* cub_path, const char* thrust_path,
  const char* libcudacxx_path, const char* ctk_path, cccl_build_config* config);
CCCL_C_API CUresult cccl_device_radix_sort_load(cccl_device_radix_sort_build_result_t* build);
CCCL_C_API CUresult cccl_device_radix_sort(cccl_device_radix_sort_build_result_t build, void* d_temp_storage,
  size_t* temp_storage_bytes, cccl_iterator_t d_keys_in, cccl_iterator_t d_keys_out, cccl_iterator_t d_values_in,
  cccl_iterator_t d_values_out, cccl_op_t decomposer, uint64_t num_items, int begin_bit, int end_bit,
  bool is_overwrite_okay, int* selector, CUstream stream);
CCCL_C_API CUresult cccl_device_radix_sort_link_ltoir(cccl_device_radix_sort_build_result_t* build,
  const void** input_blobs, const size_t* input_sizes, size_t num_inputs);
CCCL_C_API CUresult cccl_device_radix_sort_serialize(const cccl_device_radix_sort_build_result_t* build, void** out_buf,
  size_t* out_size);
CCCL_C_API CUresult cccl_device_radix_sort_deserialize(cccl_device_radix_sort_build_result_t* build, const void* buf,
  size_t size);
CCCL_C_API CUresult cccl_device_radix_sort_cleanup(cccl_device_radix_sort_build_result_t* bld_ptr);
/* frees a buffer returned by cccl_device_radix_sort_serialize (reference: cccl/c/serialization.h) */
CCCL_C_API void cccl_serialization_buffer_free(void* buf);

#ifdef __cplusplus
}
#endif
#endif /* B200RS_CCCL_C_H_ */
