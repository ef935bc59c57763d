// `cccl.c.parallel` radix-sort entry points over b200rs_sort (include/b200rs_cccl_c.h).
//
// Replaces /root/reference/c/parallel/src/radix_sort.cu:232-700 (cccl_device_radix_sort_build* / _compile / _load /
// cccl_device_radix_sort / _serialize / _deserialize / _cleanup).  The reference JIT-compiles CUB's kernels for the
// (key, value, order) triple with NVRTC and keeps CUkernel handles in the build result; here the kernels already exist
// in this library, so "build" only records the triple and "run" is one b200rs_sort call.
#include <cstdlib>
#include <cstring>

#include "../../include/b200rs.h"
#include "../../include/b200rs_cccl_c.h"

namespace
{
// arithmetic key -> (key kind, ok)
bool key_kind_of(cccl_type_enum t, int* kind)
{
  switch (t)
  {
    case CCCL_INT8:
    case CCCL_INT16:
    case CCCL_INT32:
    case CCCL_INT64: *kind = B200RS_KEY_INT; return true;
    case CCCL_UINT8:
    case CCCL_UINT16:
    case CCCL_UINT32:
    case CCCL_UINT64:
    case CCCL_BOOLEAN: *kind = B200RS_KEY_UINT; return true;
    case CCCL_FLOAT16:
    case CCCL_FLOAT32:
    case CCCL_FLOAT64: *kind = B200RS_KEY_FLOAT; return true;
    default: return false; // CCCL_STORAGE: user-defined key + decomposer
  }
}

bool keys_only_of(const cccl_iterator_t& values)
{
  // the reference's own convention (radix_sort.cu:240): a pointer iterator without a pointer
  return values.type == CCCL_POINTER && values.state == nullptr;
}

struct Wire // serialized form of a build result
{
  char magic[8];
  int cc;
  int order;
  cccl_type_info key_type, value_type;
};
const char MAGIC[8] = {'b', '2', '0', '0', 'r', 's', '0', '1'};

CUresult record(cccl_device_radix_sort_build_result_t* build, cccl_sort_order_t order, const cccl_iterator_t& keys,
                const cccl_iterator_t& values, int cc_major, int cc_minor)
{
  if (build == nullptr)
  {
    return CUDA_ERROR_INVALID_VALUE;
  }
  int kind = 0;
  if (keys.type != CCCL_POINTER || values.type != CCCL_POINTER || !key_kind_of(keys.value_type.type, &kind))
  {
    return CUDA_ERROR_NOT_SUPPORTED;
  }
  const size_t kb = keys.value_type.size;
  const size_t vb = keys_only_of(values) ? 0 : values.value_type.size;
  if ((kb != 1 && kb != 2 && kb != 4 && kb != 8) || (kind == B200RS_KEY_FLOAT && kb < 2)
      || (vb != 0 && vb != 1 && vb != 2 && vb != 4 && vb != 8 && vb != 16))
  {
    return CUDA_ERROR_NOT_SUPPORTED;
  }
  memset(build, 0, sizeof(*build));
  build->cc           = cc_major * 10 + cc_minor;
  build->payload_kind = CCCL_PAYLOAD_CUBIN;
  build->key_type     = keys.value_type;
  build->value_type   = values.value_type;
  if (vb == 0)
  {
    build->value_type.size = 0; // keys only (the reference keeps cub::NullType here)
  }
  build->order = order;
  return CUDA_SUCCESS;
}
} // namespace

extern "C" {

CUresult cccl_device_radix_sort_compile(
  cccl_device_radix_sort_build_result_t* build, cccl_sort_order_t sort_order, cccl_iterator_t input_keys_it,
  cccl_iterator_t input_values_it, cccl_op_t, const char*, int cc_major, int cc_minor, const char*, const char*,
  const char*, const char*, cccl_build_config*)
{
  return record(build, sort_order, input_keys_it, input_values_it, cc_major, cc_minor);
}

CUresult cccl_device_radix_sort_load(cccl_device_radix_sort_build_result_t* build)
{
  return build != nullptr ? CUDA_SUCCESS : CUDA_ERROR_INVALID_VALUE; // nothing to load: the kernels are in this library
}

CUresult cccl_device_radix_sort_build_ex(
  cccl_device_radix_sort_build_result_t* build, cccl_sort_order_t sort_order, cccl_iterator_t input_keys_it,
  cccl_iterator_t input_values_it, cccl_op_t decomposer, const char* decomposer_return_type, int cc_major, int cc_minor,
  const char* cub_path, const char* thrust_path, const char* libcudacxx_path, const char* ctk_path,
  cccl_build_config* config)
{
  const CUresult r =
    cccl_device_radix_sort_compile(build, sort_order, input_keys_it, input_values_it, decomposer, decomposer_return_type,
                                   cc_major, cc_minor, cub_path, thrust_path, libcudacxx_path, ctk_path, config);
  return r != CUDA_SUCCESS ? r : cccl_device_radix_sort_load(build);
}

CUresult cccl_device_radix_sort_build(
  cccl_device_radix_sort_build_result_t* build, cccl_sort_order_t sort_order, cccl_iterator_t input_keys_it,
  cccl_iterator_t input_values_it, cccl_op_t decomposer, const char* decomposer_return_type, int cc_major, int cc_minor,
  const char* cub_path, const char* thrust_path, const char* libcudacxx_path, const char* ctk_path)
{
  return cccl_device_radix_sort_build_ex(build, sort_order, input_keys_it, input_values_it, decomposer,
                                         decomposer_return_type, cc_major, cc_minor, cub_path, thrust_path,
                                         libcudacxx_path, ctk_path, nullptr);
}

CUresult cccl_device_radix_sort(
  cccl_device_radix_sort_build_result_t build, void* d_temp_storage, size_t* temp_storage_bytes, cccl_iterator_t d_keys_in,
  cccl_iterator_t d_keys_out, cccl_iterator_t d_values_in, cccl_iterator_t d_values_out, cccl_op_t, uint64_t num_items,
  int begin_bit, int end_bit, bool is_overwrite_okay, int* selector, CUstream stream)
{
  if (d_keys_in.type != CCCL_POINTER || d_values_in.type != CCCL_POINTER || d_keys_out.type != CCCL_POINTER
      || d_values_out.type != CCCL_POINTER)
  {
    return CUDA_ERROR_UNKNOWN; // as the reference (radix_sort.cu:598-606)
  }
  int kind = 0;
  if (!key_kind_of(build.key_type.type, &kind) || temp_storage_bytes == nullptr)
  {
    return CUDA_ERROR_INVALID_VALUE;
  }
  int sel      = 0;
  const int rc = b200rs_sort(
    d_temp_storage, temp_storage_bytes, d_keys_in.state, d_keys_out.state, d_values_in.state, d_values_out.state, num_items,
    kind, int(build.key_type.size), int(build.value_type.size), begin_bit, end_bit, build.order == CCCL_DESCENDING ? 1 : 0,
    is_overwrite_okay ? 1 : 0, &sel, reinterpret_cast<b200rs_stream_t>(stream));
  if (selector != nullptr && rc == 0 && d_temp_storage != nullptr)
  {
    *selector = sel;
  }
  return static_cast<CUresult>(rc); // cudaError_t and CUresult share their numeric values
}

CUresult cccl_device_radix_sort_link_ltoir(cccl_device_radix_sort_build_result_t*, const void**, const size_t*, size_t)
{
  return CUDA_ERROR_NOT_SUPPORTED; // there is no LTO-IR to link: keys are arithmetic, the kernels ahead-of-time
}

CUresult cccl_device_radix_sort_serialize(const cccl_device_radix_sort_build_result_t* build, void** out_buf,
                                          size_t* out_size)
{
  if (build == nullptr || out_buf == nullptr || out_size == nullptr)
  {
    return CUDA_ERROR_INVALID_VALUE;
  }
  Wire* w = static_cast<Wire*>(malloc(sizeof(Wire)));
  if (w == nullptr)
  {
    return CUDA_ERROR_OUT_OF_MEMORY;
  }
  memset(w, 0, sizeof(Wire));
  memcpy(w->magic, MAGIC, sizeof(MAGIC));
  w->cc         = build->cc;
  w->order      = int(build->order);
  w->key_type   = build->key_type;
  w->value_type = build->value_type;
  *out_buf      = w;
  *out_size     = sizeof(Wire);
  return CUDA_SUCCESS;
}

CUresult cccl_device_radix_sort_deserialize(cccl_device_radix_sort_build_result_t* build, const void* buf, size_t size)
{
  if (build == nullptr || buf == nullptr || size != sizeof(Wire) || memcmp(buf, MAGIC, sizeof(MAGIC)) != 0)
  {
    return CUDA_ERROR_INVALID_VALUE; // build is left unchanged
  }
  const Wire* w = static_cast<const Wire*>(buf);
  memset(build, 0, sizeof(*build));
  build->cc           = w->cc;
  build->payload_kind = CCCL_PAYLOAD_CUBIN;
  build->order        = static_cast<cccl_sort_order_t>(w->order);
  build->key_type     = w->key_type;
  build->value_type   = w->value_type;
  return CUDA_SUCCESS;
}

CUresult cccl_device_radix_sort_cleanup(cccl_device_radix_sort_build_result_t* bld_ptr)
{
  if (bld_ptr == nullptr)
  {
    return CUDA_ERROR_INVALID_VALUE; // as the reference
  }
  memset(bld_ptr, 0, sizeof(*bld_ptr)); // nothing was allocated
  return CUDA_SUCCESS;
}

void cccl_serialization_buffer_free(void* buf)
{
  free(buf);
}

} // extern "C"
