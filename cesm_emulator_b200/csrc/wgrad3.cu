// tcgen05 weight gradient of the 3x3 stride-1 convolutions with halo reuse (round 2).
//
//     dW[co][tap][ci] += sum_{pixels p}  dY[p][co] * X[p + tap][ci]          (video_net.py:215 backward)
//
// wgrad.cu fetches a separate, tap-shifted activation tile for every (tap, 64-channel block): at 64 output
// channels a CTA needs 48 KB from L2 per eight 128x64x16 MMAs and the L2 -> shared-memory feed (~42 B/clk per
// SM with every SM pulling), not the tensor pipe, sets the time.  Here a CTA owns ONE 64-channel input block
// and ALL NINE taps: per bh x bw pixel tile it fetches the (bh+2) x (bw+2) activation halo once (the nine taps
// are row-shifted MN-major views of it; tools/probe_umma_desc.cu: exact on B200 for any 128-byte row shift)
// and the dY tile once -- 42 KB per forty MMAs, 5.6x less L2 traffic.  Five accumulators (tap pairs stacked
// along M through the descriptor's leading-byte offset, the ninth tap alone) live in TMEM (5 x 64 columns).
//
// Grid: (input 64-channel blocks, cout / 64, split-K over pixel tiles); fp32 partials are added into dW with
// red.global.add like wgrad.cu.  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocation,
// 4..7 = epilogue.
#include "api_common.h"
#include "common.cuh"
#include "igemm.h"

namespace cesm {

static constexpr int kW3Threads = 256;
static constexpr int kW3Stages = 4;
static constexpr int kW3DyBytes = 128 * 128;   // 128 pixels x 64 output channels, fp16

__global__ void __launch_bounds__(kW3Threads, 1)
wgrad3_kernel(const __grid_constant__ Wgrad3Maps maps, const Wgrad3Params p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t stage_bytes = p.halo_stage_bytes + kW3DyBytes;
    const uint32_t bar_base = smem_base + kW3Stages * stage_bytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kW3Stages + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * kW3Stages);
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * kW3Stages + 1);
    volatile uint32_t* tmem_ptr_gen =
        reinterpret_cast<volatile uint32_t*>(smem_gen + kW3Stages * stage_bytes + 8 * (2 * kW3Stages + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cb = blockIdx.x;                 // 64-channel input block of this CTA
    const int col0 = blockIdx.y * 64;          // output channels of this CTA
    const int pw = p.bw + 2;

    const int tiles = p.tiles_w * p.tiles_h * p.n;
    const int per = (tiles + gridDim.z - 1) / gridDim.z;
    const int t_begin = blockIdx.z * per;
    const int t_end = min(tiles, t_begin + per);
    const int num_kb = t_end - t_begin;
    if (num_kb <= 0) return;  // uniform across the CTA

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.x[0]);
        tma_prefetch_desc(&maps.y);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kW3Stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr_addr, 512);   // 5 x 64 accumulator columns
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer: one halo box + one dY box per pixel tile =====
        pdl_wait();
        const int cblk0 = p.c0 >> 6;
        const int midx = cb >= cblk0 ? 1 : 0;
        const int c = (cb >= cblk0 ? cb - cblk0 : cb) << 6;
        const uint32_t tx_bytes = p.halo_box_bytes + kW3DyBytes;
        int stage = 0;
        uint32_t phase = 0;
        for (int t = t_begin; t < t_end; ++t) {
            mbar_wait(empty_bar(stage), phase ^ 1u, 41);
            const int tw = t % p.tiles_w, th = (t / p.tiles_w) % p.tiles_h, tn = t / (p.tiles_w * p.tiles_h);
            const int ow0 = tw * p.bw, oh0 = th * p.bh;
            const uint32_t sa = smem_base + stage * stage_bytes;
            mbar_arrive_expect_tx(full_bar(stage), tx_bytes);
            tma_load_4d(sa, &maps.x[midx], full_bar(stage), c, ow0 - 1, oh0 - 1, tn);       // zero padding = TMA OOB fill
            tma_load_4d(sa + p.halo_stage_bytes, &maps.y, full_bar(stage), col0, ow0, oh0, tn);
            if (++stage == kW3Stages) {
                stage = 0;
                phase ^= 1u;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected thread) =====
        // M tile m stacks taps (2m, 2m+1) along M: two MN-major 64-channel atoms `lbo[m]` bytes apart inside the SAME
        // halo; K runs over the 16-pixel pieces of the tile's image rows (pixels of different rows are not
        // contiguous in the halo, so every (row, piece) has its own start address).
        constexpr uint32_t idesc = make_idesc_f16(128, 64, 1, 1);
        if (elect_one()) {
            uint64_t a_hi[5];
#pragma unroll
            for (int m = 0; m < 5; ++m)
                a_hi[m] = make_smem_desc_sw128(0, (uint32_t)(p.tap_row[2 * m + 1] - p.tap_row[2 * m]) * 128u, 1024);
            const uint64_t b_hi = make_smem_desc_sw128(0, 0, 1024);
            uint32_t a_row8[5];
#pragma unroll
            for (int m = 0; m < 5; ++m) a_row8[m] = (uint32_t)p.tap_row[2 * m] * 8u;   // 16-byte units
            const int bh = p.bh, pieces = p.bw >> 4;
            const uint32_t pw8 = (uint32_t)pw * 8u, bw8 = (uint32_t)p.bw * 8u;
            const uint32_t halo16 = p.halo_stage_bytes >> 4, stage16 = stage_bytes >> 4, base16 = smem_base >> 4;
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(full_bar(stage), phase, 42);
                tc_fence_after();
                const uint32_t sa16 = base16 + stage * stage16;
                const uint32_t sb16 = sa16 + halo16;
                for (int r = 0; r < bh; ++r)
                    for (int s = 0; s < pieces; ++s) {
                        const uint32_t xa = sa16 + r * pw8 + s * 128u;   // 16 pixels = 2048 B = 128 units
                        const uint64_t db = b_hi | (uint64_t)(sb16 + r * bw8 + s * 128u);
                        const uint32_t accum = (kb | r | s) != 0;
#pragma unroll
                        for (int m = 0; m < 5; ++m)
                            umma_f16(tmem_base + m * 64, a_hi[m] | (uint64_t)(xa + a_row8[m]), db, idesc, accum);
                    }
                umma_commit(empty_bar(stage));
                if (kb == num_kb - 1) umma_commit(tmem_full_bar);
                if (++stage == kW3Stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===== epilogue: fp32 partials -> red.global.add into dW[co][tap][ci] =====
        const int q = warp & 3;
        const int r = q * 32 + lane;               // accumulator row: unit r >> 6 of the M tile, channel r & 63
        const int ci = (cb << 6) + (r & 63);
        mbar_wait(tmem_full_bar, 0, 43);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const size_t co_stride = (size_t)p.so;
#pragma unroll 1
        for (int m = 0; m < 5; ++m) {
            const int t = 2 * m + (r >> 6);
            const bool valid = t < 9;
            float* base = p.dw + (size_t)p.tap_off[valid ? t : 0] + (size_t)ci * p.si;
#pragma unroll 1
            for (int cc = 0; cc < 64; cc += 32) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + m * 64 + cc, v);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        atomicAdd(base + (size_t)(col0 + cc + i) * co_stride, __uint_as_float(v[i]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

cudaError_t wgrad3_launch(const Wgrad3Maps& maps, const Wgrad3Params& p, int cblk, int ksplit, cudaStream_t stream) {
    const size_t smem = 1024 + (size_t)kW3Stages * (p.halo_stage_bytes + kW3DyBytes) + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kIgemm2MaxSmem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (smem > kIgemm2MaxSmem) return cudaErrorInvalidValue;
    dim3 grid(cblk, p.cout / 64, ksplit);
    launch_pdl(wgrad3_kernel, grid, kW3Threads, smem, stream, maps, p);
    return cudaGetLastError();
}

}  // namespace cesm
