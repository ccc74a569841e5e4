// abi_check.cpp -- compiled ONLY where the reference tree exists (oracle/Makefile target `ref`): static_asserts
// that ggllm.cpp_b200/csrc/ggml_abi_mirror.h matches the reference's ggml.h / ggml-cuda.h byte for byte.
// TEST INFRASTRUCTURE; produces no code the product links.
#include "ggml.h"
#include "ggml-cuda.h"
#include "../ggllm.cpp_b200/csrc/ggml_abi_mirror.h"
#include <cstddef>

#define SAME_OFF(f) static_assert(offsetof(ggml_tensor, f) == offsetof(abi::tensor, f), "offset of " #f)
static_assert(sizeof(ggml_tensor) == sizeof(abi::tensor), "sizeof(ggml_tensor)");
SAME_OFF(type); SAME_OFF(backend); SAME_OFF(n_dims); SAME_OFF(ne); SAME_OFF(nb); SAME_OFF(op); SAME_OFF(is_param); SAME_OFF(grad);
SAME_OFF(src0); SAME_OFF(src1); SAME_OFF(opt); SAME_OFF(n_tasks); SAME_OFF(perf_runs); SAME_OFF(perf_cycles); SAME_OFF(perf_time_us);
SAME_OFF(data); SAME_OFF(name); SAME_OFF(extra); SAME_OFF(meta); SAME_OFF(padding);
static_assert(sizeof(tensor_meta) == sizeof(abi::tensor_meta), "sizeof(tensor_meta)");
#define SAME_MOFF(f) static_assert(offsetof(tensor_meta, f) == offsetof(abi::tensor_meta, f), "meta offset of " #f)
SAME_MOFF(layer_id); SAME_MOFF(short_name); SAME_MOFF(cuda_op_directive); SAME_MOFF(cuda_info_op_on_device); SAME_MOFF(cuda_perf_mal_mul_type);
SAME_MOFF(f_custom); SAME_MOFF(i_custom); SAME_MOFF(debug_flag);
static_assert(sizeof(ggml_compute_params) == sizeof(abi::compute_params), "sizeof(ggml_compute_params)");
static_assert(offsetof(ggml_compute_params, ith) == offsetof(abi::compute_params, ith) && offsetof(ggml_compute_params, wdata) == offsetof(abi::compute_params, wdata), "params");
static_assert(GGML_BACKEND_CPU == abi::BACKEND_CPU && GGML_BACKEND_GPU == abi::BACKEND_GPU && GGML_BACKEND_GPU_SPLIT == abi::BACKEND_GPU_SPLIT, "backends");
static_assert(GGML_TASK_INIT == abi::TASK_INIT && GGML_TASK_COMPUTE == abi::TASK_COMPUTE && GGML_TASK_FINALIZE == abi::TASK_FINALIZE, "tasks");
static_assert(GGML_OP_NONE == abi::OP_NONE && GGML_OP_ADD == abi::OP_ADD && GGML_OP_MUL == abi::OP_MUL && GGML_OP_REPEAT == abi::OP_REPEAT &&
              GGML_OP_GELU == abi::OP_GELU && GGML_OP_SILU == abi::OP_SILU && GGML_OP_NORM == abi::OP_NORM && GGML_OP_RMS_NORM == abi::OP_RMS_NORM &&
              GGML_OP_MUL_MAT == abi::OP_MUL_MAT && GGML_OP_SCALE == abi::OP_SCALE && GGML_OP_SET == abi::OP_SET && GGML_OP_CPY == abi::OP_CPY &&
              GGML_OP_CONT == abi::OP_CONT && GGML_OP_RESHAPE == abi::OP_RESHAPE && GGML_OP_VIEW == abi::OP_VIEW && GGML_OP_PERMUTE == abi::OP_PERMUTE &&
              GGML_OP_TRANSPOSE == abi::OP_TRANSPOSE && GGML_OP_GET_ROWS == abi::OP_GET_ROWS && GGML_OP_DIAG_MASK_INF == abi::OP_DIAG_MASK_INF &&
              GGML_OP_SOFT_MAX == abi::OP_SOFT_MAX && GGML_OP_ROPE == abi::OP_ROPE, "op values");
static_assert(GGML_TYPE_F32 == 0 && GGML_TYPE_F16 == 1 && GGML_TYPE_Q4_0 == 2 && GGML_TYPE_Q4_1 == 3 && GGML_TYPE_Q5_0 == 6 && GGML_TYPE_Q5_1 == 7 &&
              GGML_TYPE_Q8_0 == 8 && GGML_TYPE_Q2_K == 10 && GGML_TYPE_Q3_K == 11 && GGML_TYPE_Q4_K == 12 && GGML_TYPE_Q5_K == 13 && GGML_TYPE_Q6_K == 14, "type ids");
// GPUStatus / extra: the surface header re-declares them; compare with the reference's
namespace ours {
#undef GGML_CUDA_MAX_DEVICES
#define ggml_tensor_extra_gpu ggml_tensor_extra_gpu_ours
#define GPUStatus GPUStatus_ours
#define GGML_B200_SURFACE_TYPES_ONLY
#include "../include/ggml_b200_cuda_surface.h"
#undef GPUStatus
#undef ggml_tensor_extra_gpu
}
static_assert(sizeof(ours::GPUStatus_ours) == sizeof(GPUStatus) && offsetof(ours::GPUStatus_ours, device_props) == offsetof(GPUStatus, device_props) &&
              offsetof(ours::GPUStatus_ours, total_free_vram) == offsetof(GPUStatus, total_free_vram), "GPUStatus");
static_assert(sizeof(ours::ggml_tensor_extra_gpu_ours) == sizeof(ggml_tensor_extra_gpu), "extra");
int abi_check_ok(void) { return 1; }
