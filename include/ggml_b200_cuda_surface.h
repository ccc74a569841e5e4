/*
 * ggml_b200_cuda_surface.h -- the reference's GPU operator surface, as exported by libggml_b200.so.
 *
 * These are exactly the extern "C" symbols of the reference's ggml-cuda.h (file:line cited per symbol; paths are
 * relative to the ggllm.cpp tree).  A ggllm.cpp build compiled with -DGGML_USE_CUBLAS needs NO source change: it
 * keeps including its own ggml-cuda.h and links libggml_b200.so instead of compiling ggml-cuda.cu.
 * This header exists so that the contract is written down next to the library; `struct ggml_tensor`,
 * `struct ggml_compute_params` are the reference's types (ggml.h:421-459, 507-516) and stay opaque here -- the
 * library reads them through a layout mirror that oracle/abi_check.cpp static_asserts against the reference headers.
 *
 * Differences from the reference backend that a caller can observe (all documented in INTEGRATION.md):
 *   - one process drives ONE device: GPUStatus.num_devices is 1 (the main device); multi-GPU runs use the
 *     layer-range pipeline of ggml_b200.h part B, one process per GPU, instead of GGML_BACKEND_GPU_SPLIT row splits
 *   - MUL_MAT numerics follow the CPU path (int8-quantised activations, integer block dots), not the reference CUDA
 *     kernels' fp32 activations
 *   - weights resident on the CPU are never uploaded per call: ggml_cuda_can_mul_mat() is false for them
 */
#ifndef GGML_B200_CUDA_SURFACE_H
#define GGML_B200_CUDA_SURFACE_H
#include <cuda_runtime.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct ggml_tensor;
struct ggml_compute_params;

#ifndef GGML_CUDA_MAX_DEVICES
#define GGML_CUDA_MAX_DEVICES 16                       /* ggml-cuda.h:10 */
struct ggml_tensor_extra_gpu {                         /* ggml-cuda.h:13-15; owned by the backend, opaque to callers */
    void * data_device[GGML_CUDA_MAX_DEVICES];
};
typedef struct {                                       /* ggml-cuda.h:16-27 */
    int max_gpus;
    int num_devices;
    int main_device_id;
    size_t total_vram;
    size_t total_free_vram;
    size_t device_vram_free[GGML_CUDA_MAX_DEVICES];
    size_t device_vram_total[GGML_CUDA_MAX_DEVICES];
    int64_t device_vram_reserved[GGML_CUDA_MAX_DEVICES];
    struct cudaDeviceProp device_props[GGML_CUDA_MAX_DEVICES];
} GPUStatus;
#endif

#ifndef GGML_B200_SURFACE_TYPES_ONLY   /* oracle/abi_check.cpp compares just the types with the reference's */
/* ggml-cuda.h:31  pointer to a static struct, never freed by the caller */
const GPUStatus * ggml_cuda_get_system_gpu_status(void);
/* ggml-cuda.h:33  check_only: has init finished? (non-blocking); otherwise initialise (idempotent).  Called from a
 * helper pthread (ggml.c:4319-4326) and polled by libfalcon.cpp:901,909,1947 */
bool   ggml_init_cublas(bool check_only);
/* ggml-cuda.h:34-35  -1 = refresh all; works before init (ggml.c:4322) */
void   ggml_cuda_update_gpu_status(int device_id);
void   ggml_cuda_print_gpu_status(const GPUStatus * status, bool print_summary);
/* ggml-cuda.h:36-39, 57-59  plain setters, legal before init (ggml-cuda.cu:3164-3172) */
void   ggml_cuda_set_max_gpus(int max_gpus);
void   ggml_cuda_set_vram_reserved(int64_t vram_reserved);
void   ggml_cuda_set_tensor_split_prepare(const float * tensor_split, int num_devices);
void   ggml_cuda_set_tensor_split(const float * tensor_split);
void   ggml_cuda_set_main_device(int main_device);
void   ggml_cuda_set_scratch_size(size_t scratch_size);
void   ggml_cuda_free_scratch(void);
/* ggml-cuda.h:41-42  (ggml_cuda_mul is only reached through the dispatcher) */
void   ggml_cuda_mul(const struct ggml_tensor * src0, const struct ggml_tensor * src1, struct ggml_tensor * dst);
bool   ggml_cuda_can_mul_mat(const struct ggml_tensor * src0, const struct ggml_tensor * src1, struct ggml_tensor * dst);
/* ggml-cuda.h:47-50  pinned host memory (may return NULL -> caller uses new[], llama-util.h:462-477); pool GC hooks */
void * ggml_cuda_host_malloc(size_t size);
void   ggml_cuda_host_free(void * ptr);
void   ggml_cuda_pool_reset_all_counters(int device_id);
int    ggml_cuda_pool_purge_buffers_with_access_count(int min_access_count, int device_id);
/* ggml-cuda.h:52  upload the tensor's raw bytes, set tensor->extra; tensor->data keeps pointing at the host copy */
void   ggml_cuda_transform_tensor(void * data, struct ggml_tensor * tensor);
/* ggml-cuda.h:54-56 */
void   ggml_cuda_free_data(struct ggml_tensor * tensor);
void   ggml_cuda_assign_buffers(struct ggml_tensor * tensor);
void   ggml_cuda_assign_buffers_no_scratch(struct ggml_tensor * tensor);
/* ggml-cuda.h:60  THE operator hook (called by every worker thread, ggml.c:15779-15790): true = handled */
bool   ggml_cuda_compute_forward(struct ggml_compute_params * params, struct ggml_tensor * tensor);
#endif /* GGML_B200_SURFACE_TYPES_ONLY */

#ifdef __cplusplus
}
#endif
#endif
