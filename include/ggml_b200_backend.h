/* ggml_b200_backend.h — the drop-in boundary of chatllm.cpp_b200: the symbols libggml-cuda.so exports.
 *
 * The reference loads backends as shared modules and binds exactly two C symbols with dlsym
 * (ggml/src/ggml-backend-reg.cpp:211-246; typedefs ggml/src/ggml-backend-impl.h:214-251):
 *
 *     ggml_backend_reg_t ggml_backend_init(void);     required
 *     int                ggml_backend_score(void);    optional, 0 = "not usable on this system"
 *
 * Everything else crosses the boundary through the vtables reachable from the returned registry object
 * (ggml/src/ggml-backend-impl.h:11-210, GGML_BACKEND_API_VERSION == 2):
 *
 *   ggml_backend_reg_i          get_name="CUDA", get_device_count, get_device, get_proc_address
 *   ggml_backend_device_i       get_name="CUDA<i>", get_description, get_memory, get_type=GPU, get_props,
 *                               init_backend, get_buffer_type, get_host_buffer_type (pinned), supports_op,
 *                               supports_buft                       (replaces ggml-cuda.cu:5057-5073)
 *   ggml_backend_buffer_type_i  get_name, alloc_buffer (NULL on OOM), get_alignment=256, is_host=false
 *                                                                    (replaces ggml-cuda.cu:685-751)
 *   ggml_backend_buffer_i       free_buffer, get_base, memset_tensor, set_tensor, get_tensor, cpy_tensor, clear
 *                               — synchronous; Q4_0/Q8_0 tensors are converted to/from the device row layout here
 *                                                                    (replaces ggml-cuda.cu:568-683)
 *   ggml_backend_i              get_name, free, synchronize, graph_compute -> enum ggml_status
 *                                                                    (replaces ggml-cuda.cu:4375-4390)
 *
 * The structs themselves are the host application's SDK (ggml/include/ggml-backend.h, ggml/src/ggml-backend-impl.h);
 * the module is compiled against those headers in place and resolves libggml-base.so symbols
 * (ggml_backend_buffer_init, ggml_nbytes, ...) from the host process at dlopen time.
 *
 * Error convention: no exceptions; NULL from alloc_buffer on OOM, false from supports_* /
 * cpy_tensor, GGML_STATUS_FAILED from graph_compute, GGML_ABORT on unrecoverable CUDA errors.
 * No CPU fallback exists inside the module: unsupported ops are reported through supports_op and nothing else.
 */
#ifndef GGML_B200_BACKEND_H
#define GGML_B200_BACKEND_H

#ifdef __cplusplus
extern "C" {
#endif

struct ggml_backend_reg;
struct ggml_backend;

/* ggml/src/ggml-backend-impl.h:216  typedef ggml_backend_reg_t (*ggml_backend_init_t)(void); */
struct ggml_backend_reg * ggml_backend_init(void);
/* ggml/src/ggml-backend-impl.h:219  typedef int (*ggml_backend_score_t)(void); */
int ggml_backend_score(void);
/* introspection for tests / bench: kernels launched so far by a backend instance of this module (-1 if not ours) */
long long ggml_backend_b200_launch_count(struct ggml_backend * backend);

#ifdef __cplusplus
}
#endif
#endif
