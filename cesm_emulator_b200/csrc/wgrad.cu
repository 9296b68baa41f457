// tcgen05 weight-gradient kernel for sm_100a: the contraction runs over PIXELS, read straight
// from the channels-last activations with no transposed copies.
//
//     dW[co][tap][ci] += sum_{pixels p}  dY[p][co] * X[p + tap][ci]
//
// Both operands are "MN-major" for the tensor core (the contracted pixel axis is the strided one,
// channels are contiguous), which tcgen05 consumes directly: SWIZZLE_128B atoms of 64 channels x 8
// pixels, LBO = stride between 64-channel atoms, SBO = 1024 B between 8-pixel groups (verified on
// B200 by tools/probe_umma_desc.cu).  A CTA owns a 128-row block of (tap, ci) = two 64-channel
// "units" and BLOCK_N output channels, walks its share of the pixel tiles (split-K over grid.z),
// and adds its fp32 partial into dW with red.global.add.
//
// Replaces the weight-gradient halves of cuDNN/cuBLAS backward for video_net.py:215, :246, :62,
// :66, :322-323, :380-381.
#include <cstdlib>
#include "api_common.h"
#include "common.cuh"
#include "igemm.h"

namespace cesm {

static constexpr int kPix = 64;  // pixels per K block

struct WgradMaps {
    CUtensorMap x[4];
    CUtensorMap y;
};

template <int BLOCK_N, int STAGES>
struct WgradSmem {
    static constexpr int kABytes = 2 * kPix * 128;
    static constexpr int kBBytes = (BLOCK_N / 64) * kPix * 128;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = STAGES * kStageBytes;
    static constexpr int kTotal = kBarOffset + 1024 + 1024;
};

// PAIR (BLOCK_N >= 128): the two CTAs of a cluster take two neighbouring 128-row (tap, ci) blocks of the SAME output
// channels and pixel range and work as one cta_group::2 unit (M = 256): each loads its own activation tiles and HALF
// of the dY channel atoms, so the L2 -> shared-memory feed that bounds these launches drops by a third at 256
// columns (48 -> 32 KB per 64 pixels).  The even CTA issues; loads of both report to its barriers; commits multicast.
template <int BLOCK_N, int STAGES, bool PAIR>
__global__ void __launch_bounds__(256)
wgrad_kernel(const __grid_constant__ WgradMaps maps, const WgradParams p) {
    static_assert(!PAIR || BLOCK_N >= 128, "a CTA of a pair holds whole 64-channel atoms of dY");
    pdl_trigger();  // pdl_wait() sits in the TMA producer: MMA and epilogue depend on its data
    using S = WgradSmem<BLOCK_N, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + S::kBarOffset;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 1);
    volatile uint32_t* tmem_ptr_gen =
        reinterpret_cast<volatile uint32_t*>(smem_gen + S::kBarOffset + 8 * (2 * STAGES + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cblk0 = p.c0 >> 6;
    const int cblk = (p.c0 + p.c1) >> 6;
    const int units = p.num_taps * cblk;
    const int u0 = blockIdx.x * 2, u1 = u0 + 1;
    const bool has_u1 = u1 < units;
    const int col0 = blockIdx.y * BLOCK_N;
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    constexpr int kBAtoms = PAIR ? BLOCK_N / 128 : BLOCK_N / 64;   // dY atoms THIS CTA loads

    const int tiles_w = (p.ow + p.bw - 1) / p.bw;
    const int tiles_h = (p.oh + p.bh - 1) / p.bh;
    const int tiles_n = (p.n + p.bn - 1) / p.bn;
    const int tiles = tiles_w * tiles_h * tiles_n;
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
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR) tmem_alloc_pair(tmem_ptr_addr, BLOCK_N);
        else tmem_alloc(tmem_ptr_addr, BLOCK_N);
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        pdl_wait();
        int um[2], uc[2], udh[2], udw[2];
        for (int i = 0; i < 2; ++i) {
            const int u = u0 + i;
            const int tap = (u < units) ? u / cblk : 0;
            const int cb = (u < units) ? u - tap * cblk : 0;
            um[i] = p.tap_map[tap] + (cb >= cblk0 ? 1 : 0);
            uc[i] = (cb >= cblk0 ? cb - cblk0 : cb) << 6;
            udh[i] = p.tap_dh[tap];
            udw[i] = p.tap_dw[tap];
        }
        // PAIR: every CTA always loads two activation tiles (an out-of-range unit re-loads unit 0; its rows are
        // dropped by the epilogue) so that the leader can expect a fixed byte count for both CTAs
        const uint32_t tx_bytes = PAIR ? 2u * p.box_bytes * (2 + kBAtoms) : p.box_bytes * ((has_u1 ? 2 : 1) + BLOCK_N / 64);
        int stage = 0;
        uint32_t phase = 0;
        for (int t = t_begin; t < t_end; ++t) {
            mbar_wait(empty_bar(stage), phase ^ 1u, 21);
            const int tw = t % tiles_w, th = (t / tiles_w) % tiles_h, tn = t / (tiles_w * tiles_h);
            const int ow0 = tw * p.bw, oh0 = th * p.bh, n0 = tn * p.bn;
            const uint32_t sa = smem_base + stage * S::kStageBytes;
            const uint32_t sb = sa + S::kABytes;
            if (PAIR) {
                if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), tx_bytes);
                tma_load_4d_pair(sa, &maps.x[um[0]], full_bar(stage), uc[0], ow0 + udw[0], oh0 + udh[0], n0);
                tma_load_4d_pair(sa + kPix * 128, &maps.x[um[1]], full_bar(stage), uc[1], ow0 + udw[1], oh0 + udh[1], n0);
#pragma unroll
                for (int a = 0; a < kBAtoms; ++a)
                    tma_load_4d_pair(sb + a * kPix * 128, &maps.y, full_bar(stage),
                                     col0 + ((int)cta_rank * kBAtoms + a) * 64, ow0, oh0, n0);
            } else {
                mbar_arrive_expect_tx(full_bar(stage), tx_bytes);
                tma_load_4d(sa, &maps.x[um[0]], full_bar(stage), uc[0], ow0 + udw[0], oh0 + udh[0], n0);
                if (has_u1)
                    tma_load_4d(sa + kPix * 128, &maps.x[um[1]], full_bar(stage), uc[1], ow0 + udw[1], oh0 + udh[1], n0);
#pragma unroll
                for (int a = 0; a < BLOCK_N / 64; ++a)
                    tma_load_4d(sb + a * kPix * 128, &maps.y, full_bar(stage), col0 + a * 64, ow0, oh0, n0);
            }
            if (++stage == STAGES) {
                stage = 0;
                phase ^= 1u;
            }
        }
        if (PAIR) {
            // producer tail: the leader's multicast commits land on THIS CTA's empty barriers asynchronously; stay
            // until the last one has arrived (the shared memory must not change hands under an in-flight arrive)
            for (int k = 0; k < STAGES; ++k) {
                mbar_wait(empty_bar(stage), phase ^ 1u, 24);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one elected thread runs the loop (its instruction stream is the critical path: no
        // per-MMA predicate, no re-convergence points); the constant descriptor word is hoisted and only the
        // start-address field advances =====
        constexpr uint32_t idesc = make_idesc_f16(PAIR ? 256 : 128, BLOCK_N, 1, 1);
        if (cta_rank == 0 && elect_one()) {
            const uint64_t desc_hi = make_smem_desc_sw128(0, kPix * 128, 1024);
            const uint32_t base16 = smem_base >> 4;
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(full_bar(stage), phase, 22);
                tc_fence_after();
                const uint32_t sa16 = base16 + stage * (S::kStageBytes >> 4);
                const uint32_t sb16 = sa16 + (S::kABytes >> 4);
#pragma unroll
                for (int k = 0; k < kPix / 16; ++k) {   // 16 pixels = two 8-row groups = 2048 B further along K
                    const uint64_t da = desc_hi | (uint64_t)(sa16 + k * 128), db = desc_hi | (uint64_t)(sb16 + k * 128);
                    if (PAIR) umma_f16_pair(tmem_base, da, db, idesc, k == 0 ? (uint32_t)(kb != 0) : 1u);
                    else umma_f16(tmem_base, da, db, idesc, k == 0 ? (uint32_t)(kb != 0) : 1u);
                }
                if (PAIR) {
                    umma_commit_pair(empty_bar(stage));
                    if (kb == num_kb - 1) umma_commit_pair(tmem_full_bar);
                } else {
                    umma_commit(empty_bar(stage));
                    if (kb == num_kb - 1) umma_commit(tmem_full_bar);
                }
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===== epilogue: fp32 partial -> red.global.add into dW[co][tap][ci] =====
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int u = u0 + (r >> 6);
        const bool valid = u < units;
        const int tap = valid ? u / cblk : 0;
        const int ci = valid ? ((u - tap * cblk) << 6) + (r & 63) : 0;
        float* base = p.dw + (size_t)p.tap_off[tap] + (size_t)ci * p.si;
        const size_t co_stride = (size_t)p.so;
        mbar_wait(tmem_full_bar, 0, 23);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
        for (int cc = 0; cc < BLOCK_N; cc += 32) {
            uint32_t v[32];
            tmem_ld_32x32(taddr + cc, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int i = 0; i < 32; ++i) atomicAdd(base + (size_t)(col0 + cc + i) * co_stride, __uint_as_float(v[i]));
            }
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();
    else __syncthreads();
    if (warp == 2) {
        if (PAIR) tmem_dealloc_pair(tmem_base, BLOCK_N);
        else tmem_dealloc(tmem_base, BLOCK_N);
    }
}

template <int BLOCK_N, int STAGES, bool PAIR>
static cudaError_t launch_wgrad(const WgradMaps& maps, const WgradParams& p, dim3 grid, cudaStream_t stream) {
    using S = WgradSmem<BLOCK_N, STAGES>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<BLOCK_N, STAGES, PAIR>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    launch_pdl_cluster(wgrad_kernel<BLOCK_N, STAGES, PAIR>, grid, 256, S::kTotal, stream, PAIR ? 2 : 1, maps, p);
    return cudaGetLastError();
}

cudaError_t wgrad_launch(const CUtensorMap* xmaps, int n_xmaps, const CUtensorMap& ymap, const WgradParams& p,
                         int block_n, int ksplit, bool pair, cudaStream_t stream) {
    WgradMaps maps;
    for (int i = 0; i < 4; ++i) maps.x[i] = xmaps[i < n_xmaps ? i : 0];
    maps.y = ymap;
    const int units = p.num_taps * ((p.c0 + p.c1) >> 6);
    const int m_tiles = (units + 1) / 2;
    dim3 grid(pair ? (m_tiles + 1) / 2 * 2 : m_tiles, p.cout / block_n, ksplit);
    switch (block_n) {
        case 64: return launch_wgrad<64, 4, false>(maps, p, grid, stream);
        case 128: return pair ? launch_wgrad<128, 3, true>(maps, p, grid, stream)
                              : launch_wgrad<128, 3, false>(maps, p, grid, stream);
        // 256 columns: one CTA per SM with a 4-deep 48 KB ring beats two CTAs with 2 stages each
        // (64->768 projection gradient at 192x288: 104 -> 87 us, 6.3 TB/s; tools/bench_wgrad.py)
        case 256: return pair ? launch_wgrad<256, 4, true>(maps, p, grid, stream)
                              : launch_wgrad<256, 4, false>(maps, p, grid, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace cesm
