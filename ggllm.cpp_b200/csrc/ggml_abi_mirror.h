// ggml_abi_mirror.h -- layout mirror of the reference structs the operator hook receives.
//
// The surface functions take `struct ggml_tensor *` / `struct ggml_compute_params *` owned by the reference's ggml.c.
// We cannot include the reference's ggml.h in this repository, so the fields we read are mirrored here with the
// same types, order and padding (ggml.h:385-403 tensor_meta, 421-459 ggml_tensor, 501-516 compute params,
// 269-273 backends, 296-370 ops).  oracle/abi_check.cpp is compiled against the real headers in the build container
// and static_asserts every size / offset / enum value below, so a drift is a build error, not a silent corruption.
#pragma once
#include <cstddef>
#include <cstdint>

namespace abi {

constexpr int MAX_DIMS = 4, MAX_OPT = 4, MAX_NAME = 64;

enum backend : int { BACKEND_CPU = 0, BACKEND_GPU = 10, BACKEND_GPU_SPLIT = 20 };
enum task : int { TASK_INIT = 0, TASK_COMPUTE = 1, TASK_FINALIZE = 2 };
// the ops this backend looks at (numeric values of enum ggml_op)
enum op : int {
    OP_NONE = 0, OP_ADD = 2, OP_MUL = 6, OP_REPEAT = 14, OP_GELU = 23, OP_SILU = 25, OP_NORM = 27, OP_RMS_NORM = 28, OP_MUL_MAT = 30,
    OP_SCALE = 32, OP_SET = 33, OP_CPY = 34, OP_CONT = 35, OP_RESHAPE = 36, OP_VIEW = 37, OP_PERMUTE = 38, OP_TRANSPOSE = 39,
    OP_GET_ROWS = 40, OP_DIAG_MASK_INF = 43, OP_SOFT_MAX = 45, OP_ROPE = 47,
};

struct tensor_meta {
    int8_t  layer_id;
    char    short_name[MAX_NAME];
    int8_t  cuda_op_directive;       // -1 default, 0 = CUDA forbidden, 1 = forced
    int8_t  cuda_info_op_on_device;
    uint8_t cuda_perf_mal_mul_type;  // 1 = quantised kernel, 16 / 32 = cuBLAS 16/32-bit (ggml.c:18266-18358 prints it)
    float   f_custom[4];
    int     i_custom[4];
    uint8_t debug_flag;
    char    padding[15];
};

struct tensor {
    int      type;
    int      backend;
    int      n_dims;
    int64_t  ne[MAX_DIMS];
    size_t   nb[MAX_DIMS];
    int      op;
    bool     is_param;
    tensor * grad;
    tensor * src0;
    tensor * src1;
    tensor * opt[MAX_OPT];
    int      n_tasks;
    int      perf_runs;
    int64_t  perf_cycles;
    int64_t  perf_time_us;
    void *   data;
    char     name[MAX_NAME];
    void *   extra;
    tensor_meta meta;
    char     padding[4];
};

struct compute_params {
    int    type;          // enum ggml_task_type
    int    ith, nth;
    size_t wsize;
    void * wdata;
};

} // namespace abi
