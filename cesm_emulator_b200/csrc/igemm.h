// Internal launch interface of the tcgen05 implicit-GEMM kernels (igemm.cu, wgrad.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cesm {

struct IgemmParams {
    // K loop: num_taps taps x (c0 + c1)/64 channel blocks.  Tap t reads TMA map tap_map[t] (+1 for
    // the second concat source) at pixel offset (tap_dh[t], tap_dw[t]) from the output pixel.
    int32_t c0, c1;
    int32_t num_taps;
    int32_t tap_map[16];
    int32_t tap_dh[16];
    int32_t tap_dw[16];
    // pixel space iterated by the M tiles and the tile box (bw*bh*bn <= 128)
    int32_t n, oh, ow;
    int32_t bw, bh, bn;
    uint32_t a_box_bytes;
    // output: pixel (n, oh, ow) -> row ((n*out_h + oh*o_sh + o_h0)*out_w + ow*o_sw + o_w0)
    int32_t cout;
    void* out;
    int32_t out_fp32;
    int32_t ldo;
    int32_t out_h, out_w, o_sh, o_sw, o_h0, o_w0;
    const float* bias;     // [cout] or null
    const void* residual;  // bf16, same pixel addressing as out, row pitch ldr; or null
    int32_t ldr;
};

struct WgradParams {
    int32_t c0, c1;
    int32_t num_taps;
    int32_t tap_map[16];
    int32_t tap_dh[16];
    int32_t tap_dw[16];
    int32_t n, oh, ow;      // pixel grid contracted over
    int32_t bw, bh, bn;     // 64-pixel K tile box (bw*bh*bn == 64)
    uint32_t box_bytes;     // bytes of one 64-channel x 64-pixel TMA box
    int32_t cout;
    float* dw;              // fp32 [cout][num_taps][c0+c1], accumulated with red.add
};

cudaError_t wgrad_launch(const CUtensorMap* xmaps, int n_xmaps, const CUtensorMap& ymap, const WgradParams& p,
                         int block_n, int ksplit, cudaStream_t stream);

cudaError_t igemm_launch(const CUtensorMap* amaps, int n_amaps, const CUtensorMap& bmap, const IgemmParams& p,
                         int block_n, cudaStream_t stream);

}  // namespace cesm
