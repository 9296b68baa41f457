/*
 * cesm_b200.h -- C ABI of libcesm_b200.so, the B200 (sm_100a) kernel library behind the
 * cesm_emulator_b200 PyTorch modules.
 *
 * The reference (kallenordling/cesm_emulator) has no FFI of its own: its hot path is the PyTorch
 * nn.Module tree in video_net.py / model.py, which lowers to aten/cuDNN/cuBLAS calls.  Each entry
 * point below replaces one of those lowered call groups and cites the reference lines it stands in
 * for.  Conventions shared by every function:
 *
 *   - plain C types only: raw DEVICE pointers, sizes, and an opaque `stream` (a cudaStream_t);
 *   - activations are bf16, channels-last: [N, H, W, C] with N = batch*frames (NDHWC flattened);
 *   - parameters / parameter gradients / statistics are fp32;
 *   - the library owns no tensor memory: callers (PyTorch's caching allocator) own every buffer;
 *   - no implicit synchronisation, nothing on the default stream, CUDA-graph capturable;
 *   - return 0 on success, a negative cesm_status otherwise; cesm_last_error() gives the
 *     thread-local message.  There is no CPU fallback: without a CUDA device every compute call
 *     fails with CESM_ERR_CUDA.
 */
#ifndef CESM_B200_H
#define CESM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum cesm_status {
    CESM_OK = 0,
    CESM_ERR_INVALID = -1, /* bad argument (shape not supported, misaligned pointer, ...) */
    CESM_ERR_CUDA = -2,    /* a CUDA runtime / driver call failed */
    CESM_ERR_INTERNAL = -3
} cesm_status;

/* Thread-local description of the last failing call on this thread ("" if none). */
const char* cesm_last_error(void);
/* Library version string, e.g. "cesm_b200 0.1 sm_100a". */
const char* cesm_version(void);

/* ------------------------------------------------------------------------------------------------
 * Dense contractions on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 * ---------------------------------------------------------------------------------------------- */

#define CESM_MAX_TAPS 16

/*
 * Implicit GEMM over pixels:
 *     out[n, oh, ow, co] = sum_t sum_ci  A[n, oh*stride + tap_dh[t], ow*stride + tap_dw[t], ci]
 *                                        * wt[co, t, ci]          (+ bias[co]) (+ residual[...])
 * A is the channel concatenation of a0 (c0 channels) and a1 (c1 channels, may be NULL/0); reads
 * outside the image are zero (conv padding).  The output pixel (n, oh, ow) is written to row
 * ((n*out_h + oh*o_sh + o_h0)*out_w + ow*o_sw + o_w0) of an [*, ldo] matrix, which lets the four
 * sub-pixel phases of a transposed conv interleave into one tensor.
 *
 * Replaces: Conv3d(1,3,3) video_net.py:215; Conv3d 1x1x1 :246; Downsample Conv3d(1,4,4)/s2 :62;
 * Upsample ConvTranspose3d(1,4,4)/s2 :66 (as 4 phase GEMMs); Conv2d 1x1 to_qkv/to_out :322-323;
 * nn.Linear to_qkv/to_out :380-381; and the data-gradient of each (same op, transformed weights).
 *
 * Constraints: c0, c1, cout multiples of 64; stride in {1, 2}; stride 2 needs c1 == 0 and even h, w.
 */
typedef struct cesm_igemm_args {
    const void* a0; /* bf16 [n, h, w, c0] */
    const void* a1; /* bf16 [n, h, w, c1] or NULL */
    int32_t c0, c1;
    int32_t n, h, w;
    int32_t stride;
    int32_t num_taps;
    int32_t tap_dh[CESM_MAX_TAPS];
    int32_t tap_dw[CESM_MAX_TAPS];
    const void* wt; /* bf16 [cout, num_taps*(c0+c1)] */
    int32_t cout;
    int32_t oh, ow; /* output pixels iterated per image */
    void* out;      /* bf16 (or fp32 if out_fp32) */
    int32_t out_fp32;
    int32_t ldo;
    int32_t out_h, out_w, o_sh, o_sw, o_h0, o_w0;
    const float* bias;    /* fp32 [cout] or NULL */
    const void* residual; /* bf16, addressed like out with pitch ldr, or NULL */
    int32_t ldr;
} cesm_igemm_args;

int cesm_igemm(const cesm_igemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CESM_B200_H */
