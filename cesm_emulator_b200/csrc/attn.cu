// Attention cores (fp16 activations, fp32 math).  Head dim is 32 for both attention types
// (video_net.py:314 dim_head=32; model.py:55 attn_dim_head=32), which these kernels require.
//
//  Temporal attention (video_net.py:413-453 + rotary_embedding.py:29-48): per (batch, pixel, head)
//  a softmax attention over the F frames with q scaled then rotary-rotated, k rotated, and the
//  relative-position bias added.  One thread per (b, pixel, head, query frame) streams the F
//  keys with an online softmax: sim never exists in memory, whatever F is.
//
//  Spatial linear attention (video_net.py:338-344): per (frame, head) softmax(q) over d,
//  softmax(k) over the n pixels, ctx = k^T v (32x32), out = ctx^T q.
#include "api_common.h"
#include "common.cuh"

namespace cesm {

static constexpr int D = 32;  // head dim

__device__ __forceinline__ void load32(const h16* p, float (&f)[D]) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 u = __ldg(q + j);
        float2 a = unpack_h2(u.x), b = unpack_h2(u.y), c = unpack_h2(u.z), d = unpack_h2(u.w);
        f[8 * j + 0] = a.x; f[8 * j + 1] = a.y; f[8 * j + 2] = b.x; f[8 * j + 3] = b.y;
        f[8 * j + 4] = c.x; f[8 * j + 5] = c.y; f[8 * j + 6] = d.x; f[8 * j + 7] = d.y;
    }
}
__device__ __forceinline__ void store32(h16* p, const float (&f)[D]) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack_h2(f[8 * j + 0], f[8 * j + 1]);
        u.y = pack_h2(f[8 * j + 2], f[8 * j + 3]);
        u.z = pack_h2(f[8 * j + 4], f[8 * j + 5]);
        u.w = pack_h2(f[8 * j + 6], f[8 * j + 7]);
        q[j] = u;
    }
}
// interleaved-pair rotation by the angles of frame `pos`: (x0,x1) -> (x0 c - x1 s, x1 c + x0 s)
__device__ __forceinline__ void rope(float (&f)[D], const float* __restrict__ cs, const float* __restrict__ sn,
                                     int pos, float scale) {
#pragma unroll
    for (int m = 0; m < D / 2; ++m) {
        const float c = cs[pos * (D / 2) + m], s = sn[pos * (D / 2) + m];
        const float x0 = f[2 * m] * scale, x1 = f[2 * m + 1] * scale;
        f[2 * m] = x0 * c - x1 * s;
        f[2 * m + 1] = x1 * c + x0 * s;
    }
}
// transpose of the rotation (gradient w.r.t. the un-rotated vector), then scale
__device__ __forceinline__ void rope_t(float (&f)[D], const float* __restrict__ cs, const float* __restrict__ sn,
                                       int pos, float scale) {
#pragma unroll
    for (int m = 0; m < D / 2; ++m) {
        const float c = cs[pos * (D / 2) + m], s = sn[pos * (D / 2) + m];
        const float g0 = f[2 * m], g1 = f[2 * m + 1];
        f[2 * m] = (g0 * c + g1 * s) * scale;
        f[2 * m + 1] = (g1 * c - g0 * s) * scale;
    }
}
__device__ __forceinline__ float dot32(const float (&a)[D], const float (&b)[D]) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) s = fmaf(a[i], b[i], s);
    return s;
}

// ------------------------------------------------------------------------------------------------
// temporal attention
//   qkv : [B*F*HW][3*H*D]  (q | k | v, each head-major), row = (b*F + f)*HW + hw
//   bias: [H][F][F] fp32 ; cs/sn: [F][D/2] fp32 rotary cos/sin ; out: [B*F*HW][H*D] ; lse: [rows][H]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
tattn_fwd_kernel(const h16* __restrict__ qkv, const float* __restrict__ bias,
                 const float* __restrict__ cs, const float* __restrict__ sn, h16* __restrict__ out,
                 float* __restrict__ lse, int B, int F, int HW, int H, float scale) {
    pdl_trigger();
    pdl_wait();
    const long long total = (long long)B * F * HW * H;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int h = t % H;
    const int hw = (t / H) % HW;
    const int i = (t / ((long long)H * HW)) % F;
    const int b = t / ((long long)H * HW * F);
    const int ld = 3 * H * D;
    const long long row_i = ((long long)b * F + i) * HW + hw;
    float q[D];
    load32(qkv + row_i * ld + h * D, q);
    rope(q, cs, sn, i, scale);
    float m = -INFINITY, l = 0.f, acc[D];
#pragma unroll
    for (int e = 0; e < D; ++e) acc[e] = 0.f;
    for (int j = 0; j < F; ++j) {
        const long long row_j = ((long long)b * F + j) * HW + hw;
        float k[D];
        load32(qkv + row_j * ld + H * D + h * D, k);
        rope(k, cs, sn, j, 1.f);
        const float s = dot32(q, k) + bias[(h * F + i) * F + j];
        const float mn = fmaxf(m, s);
        const float corr = __expf(m - mn), p = __expf(s - mn);
        float v[D];
        load32(qkv + row_j * ld + 2 * H * D + h * D, v);
        l = l * corr + p;
#pragma unroll
        for (int e = 0; e < D; ++e) acc[e] = acc[e] * corr + p * v[e];
        m = mn;
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int e = 0; e < D; ++e) acc[e] *= inv;
    store32(out + row_i * (H * D) + h * D, acc);
    lse[row_i * H + h] = m + __logf(l);
}

// One thread per (b, hw, h, r): first acts as query r (dq, dbias), then as key/value r (dk, dv).
__global__ void __launch_bounds__(128)
tattn_bwd_kernel(const h16* __restrict__ qkv, const float* __restrict__ bias,
                 const float* __restrict__ cs, const float* __restrict__ sn, const h16* __restrict__ out,
                 const float* __restrict__ lse, const h16* __restrict__ dout,
                 h16* __restrict__ dqkv, float* __restrict__ dbias, int B, int F, int HW, int H,
                 float scale) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sbias[];  // [H*F*F] block-local dbias accumulator when it fits
    const bool use_sh = (H * F * F) <= 2048;
    if (use_sh) {
        for (int x = threadIdx.x; x < H * F * F; x += blockDim.x) sbias[x] = 0.f;
        __syncthreads();
    }
    const long long total = (long long)B * F * HW * H;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total) {
        const int h = t % H;
        const int hw = (t / H) % HW;
        const int r = (t / ((long long)H * HW)) % F;
        const int b = t / ((long long)H * HW * F);
        const int ld = 3 * H * D, lo = H * D;
        const long long row_r = ((long long)b * F + r) * HW + hw;
        // ---- query role ----
        {
            float q[D], go[D], o[D], dq[D];
            load32(qkv + row_r * ld + h * D, q);
            rope(q, cs, sn, r, scale);
            load32(dout + row_r * lo + h * D, go);
            load32(out + row_r * lo + h * D, o);
            const float Di = dot32(go, o);
            const float L = lse[row_r * H + h];
#pragma unroll
            for (int e = 0; e < D; ++e) dq[e] = 0.f;
            for (int j = 0; j < F; ++j) {
                const long long row_j = ((long long)b * F + j) * HW + hw;
                float k[D], v[D];
                load32(qkv + row_j * ld + H * D + h * D, k);
                rope(k, cs, sn, j, 1.f);
                load32(qkv + row_j * ld + 2 * H * D + h * D, v);
                const float p = __expf(dot32(q, k) + bias[(h * F + r) * F + j] - L);
                const float ds = p * (dot32(go, v) - Di);
#pragma unroll
                for (int e = 0; e < D; ++e) dq[e] = fmaf(ds, k[e], dq[e]);
                if (use_sh) atomicAdd(&sbias[(h * F + r) * F + j], ds);
                else atomicAdd(&dbias[(h * F + r) * F + j], ds);
            }
            rope_t(dq, cs, sn, r, scale);
            store32(dqkv + row_r * ld + h * D, dq);
        }
        // ---- key / value role ----
        {
            float k[D], v[D], dk[D], dv[D];
            load32(qkv + row_r * ld + H * D + h * D, k);
            rope(k, cs, sn, r, 1.f);
            load32(qkv + row_r * ld + 2 * H * D + h * D, v);
#pragma unroll
            for (int e = 0; e < D; ++e) dk[e] = dv[e] = 0.f;
            for (int i = 0; i < F; ++i) {
                const long long row_i = ((long long)b * F + i) * HW + hw;
                float q[D], go[D], o[D];
                load32(qkv + row_i * ld + h * D, q);
                rope(q, cs, sn, i, scale);
                load32(dout + row_i * lo + h * D, go);
                load32(out + row_i * lo + h * D, o);
                const float p = __expf(dot32(q, k) + bias[(h * F + i) * F + r] - lse[row_i * H + h]);
                const float ds = p * (dot32(go, v) - dot32(go, o));
#pragma unroll
                for (int e = 0; e < D; ++e) {
                    dk[e] = fmaf(ds, q[e], dk[e]);
                    dv[e] = fmaf(p, go[e], dv[e]);
                }
            }
            rope_t(dk, cs, sn, r, 1.f);
            store32(dqkv + row_r * ld + H * D + h * D, dk);
            store32(dqkv + row_r * ld + 2 * H * D + h * D, dv);
        }
    }
    if (use_sh) {
        __syncthreads();
        for (int x = threadIdx.x; x < H * F * F; x += blockDim.x) {
            const float v = sbias[x];
            if (v != 0.f) atomicAdd(&dbias[x], v);
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Small-window specialisation (F <= 4, the training window K=3 and the F=1 sampling call): four
// lanes share one (pixel column, head), 8 of the 32 features each, so a warp reads one pixel's
// 512 contiguous bytes of q (then k, then v) per frame; the F x F scores live in registers and
// are reduced with two quad shuffles.  Nothing but qkv (and dout) is read: the backward recomputes
// the softmax instead of loading out / lse.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld8(const h16* p, float (&f)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    float2 a = unpack_h2(u.x), b = unpack_h2(u.y), c = unpack_h2(u.z), d = unpack_h2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void st8(h16* p, const float (&f)[8]) {
    uint4 u;
    u.x = pack_h2(f[0], f[1]); u.y = pack_h2(f[2], f[3]);
    u.z = pack_h2(f[4], f[5]); u.w = pack_h2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ float qsum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ float dot8(const float (&a)[8], const float (&b)[8]) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(a[i], b[i], s);
    return s;
}
// rotate the 4 interleaved pairs this lane owns by the angles (c4, s4) of one frame
__device__ __forceinline__ void rope8(float (&f)[8], const float (&c4)[4], const float (&s4)[4], float scale) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const float x0 = f[2 * m] * scale, x1 = f[2 * m + 1] * scale;
        f[2 * m] = x0 * c4[m] - x1 * s4[m];
        f[2 * m + 1] = x1 * c4[m] + x0 * s4[m];
    }
}
__device__ __forceinline__ void rope8_t(float (&f)[8], const float (&c4)[4], const float (&s4)[4], float scale) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const float g0 = f[2 * m], g1 = f[2 * m + 1];
        f[2 * m] = (g0 * c4[m] + g1 * s4[m]) * scale;
        f[2 * m + 1] = (g1 * c4[m] - g0 * s4[m]) * scale;
    }
}

// Both small-window kernels are persistent and feed themselves through a thread-private cp.async
// ring in shared memory: the 16-byte pieces of item i+2 are already in flight (global -> shared,
// no registers involved) while item i is being computed, so a thread keeps 2 items' worth of bytes
// outstanding regardless of its register budget.  A thread only ever reads back what it copied
// itself, so cp.async.wait_group is the only synchronisation.
static constexpr int TS_STAGES = 3;
__device__ __forceinline__ void ts_cp16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ts_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ts_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(TS_STAGES - 1) : "memory"); }
__device__ __forceinline__ void lds8(uint32_t addr, float (&f)[8]) {
    uint4 u;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr));
    float2 a = unpack_h2(u.x), b = unpack_h2(u.y), c = unpack_h2(u.z), d = unpack_h2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
extern __shared__ __align__(16) uint8_t ts_smem[];

template <int F>
__global__ void __launch_bounds__(256, 2)
tattn_small_fwd_kernel(const h16* __restrict__ qkv, const float* __restrict__ bias,
                       const float* __restrict__ cs, const float* __restrict__ sn, h16* __restrict__ out,
                       float* __restrict__ lse, long long npix /* B*HW */, int HW, int H, float scale) {
    pdl_trigger();
    pdl_wait();
    constexpr int NV = 3 * F;  // 16-byte vectors per item: (q, k, v) x F frames
    const int c = threadIdx.x & 3;
    const int HD = H * D, ld = 3 * HD;
    const long long items = npix * H * 4;
    const long long stride = (long long)gridDim.x * blockDim.x;  // multiple of 4*H: (h, c) fixed per thread
    const long long idx0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int h = (idx0 >> 2) % H;
    float cf[F][4], sf[F][4], bs[F][F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            cf[f][m] = __ldg(cs + f * (D / 2) + c * 4 + m);
            sf[f][m] = __ldg(sn + f * (D / 2) + c * 4 + m);
        }
#pragma unroll
        for (int j = 0; j < F; ++j) bs[f][j] = __ldg(bias + (h * F + f) * F + j);
    }
    // ring: [stage][vector][thread] of 16 bytes
    const uint32_t ring = smem_u32(ts_smem) + threadIdx.x * 16u;
    constexpr uint32_t kVec = 256 * 16, kStage = NV * kVec;
    auto issue = [&](long long idx, int stage) {
        if (idx < items) {
            const long long pix = idx / (4 * H);
            const long long b = pix / HW, hw = pix % HW;
            const uint32_t st = ring + stage * kStage;
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const h16* row = qkv + ((b * F + f) * HW + hw) * ld + h * D + c * 8;
                ts_cp16(st + (3 * f + 0) * kVec, row);
                ts_cp16(st + (3 * f + 1) * kVec, row + HD);
                ts_cp16(st + (3 * f + 2) * kVec, row + 2 * HD);
            }
        }
        ts_commit();
    };
#pragma unroll
    for (int s0 = 0; s0 < TS_STAGES - 1; ++s0) issue(idx0 + s0 * stride, s0);
    int stage = 0;
    for (long long idx = idx0; idx - threadIdx.x + (threadIdx.x & ~31) < items; idx += stride) {
        {
            int ns = stage + TS_STAGES - 1;
            if (ns >= TS_STAGES) ns -= TS_STAGES;
            issue(idx + (TS_STAGES - 1) * stride, ns);
        }
        ts_wait();
        const bool valid = idx < items;  // warp-uniform trip count: the quad shuffles use the full mask
        const long long pix = valid ? idx / (4 * H) : 0;
        const long long b = pix / HW, hw = pix % HW;
        const uint32_t st = ring + stage * kStage;
        float q[F][8], k[F][8], v[F][8];
#pragma unroll
        for (int f = 0; f < F; ++f) {
            lds8(st + (3 * f + 0) * kVec, q[f]);
            lds8(st + (3 * f + 1) * kVec, k[f]);
            lds8(st + (3 * f + 2) * kVec, v[f]);
            rope8(q[f], cf[f], sf[f], scale);
            rope8(k[f], cf[f], sf[f], 1.f);
        }
#pragma unroll
        for (int i = 0; i < F; ++i) {
            float s[F], m = -INFINITY;
#pragma unroll
            for (int j = 0; j < F; ++j) {
                s[j] = qsum(dot8(q[i], k[j])) + bs[i][j];
                m = fmaxf(m, s[j]);
            }
            float l = 0.f;
#pragma unroll
            for (int j = 0; j < F; ++j) {
                s[j] = __expf(s[j] - m);
                l += s[j];
            }
            const float inv = 1.f / l;
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = 0.f;
#pragma unroll
            for (int j = 0; j < F; ++j) {
                const float p = s[j] * inv;
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = fmaf(p, v[j][e], o[e]);
            }
            const long long row_i = (b * F + i) * HW + hw;
            if (valid) {
                st8(out + row_i * HD + h * D + c * 8, o);
                if (c == 0 && lse) lse[row_i * H + h] = m + __logf(l);
            }
        }
        if (++stage == TS_STAGES) stage = 0;
    }
}

// 128-thread blocks, three per SM: the kernel needs ~170 registers, so smaller blocks are what buys
// the extra resident warps (12 instead of 8 per SM) that hide the load latency of this pure stream.
template <int F>
__global__ void __launch_bounds__(128, (F <= 3 ? 3 : 2))
tattn_small_bwd_kernel(const h16* __restrict__ qkv, const float* __restrict__ bias,
                       const float* __restrict__ cs, const float* __restrict__ sn,
                       const h16* __restrict__ dout, h16* __restrict__ dqkv,
                       float* __restrict__ dbias, long long npix, int HW, int H, float scale) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sbias[8 * F * F];
    for (int x = threadIdx.x; x < H * F * F; x += blockDim.x) sbias[x] = 0.f;
    __syncthreads();
    const int c = threadIdx.x & 3;
    const int HD = H * D, ld = 3 * HD;
    float cf[F][4], sf[F][4];
#pragma unroll
    for (int f = 0; f < F; ++f)
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            cf[f][m] = __ldg(cs + f * (D / 2) + c * 4 + m);
            sf[f][m] = __ldg(sn + f * (D / 2) + c * 4 + m);
        }
    float dbacc[F][F];
#pragma unroll
    for (int i = 0; i < F; ++i)
#pragma unroll
        for (int j = 0; j < F; ++j) dbacc[i][j] = 0.f;
    const long long items = npix * H * 4;
    const long long stride = (long long)gridDim.x * blockDim.x;  // multiple of 4*H: (h, c) fixed per thread
    const long long idx0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int h = (idx0 >> 2) % H;
    constexpr int NV = 4 * F;  // (q, k, v, dout) x F frames
    const uint32_t ring = smem_u32(ts_smem) + threadIdx.x * 16u;
    constexpr uint32_t kVec = 128 * 16, kStage = NV * kVec;
    auto issue = [&](long long idx, int stage) {
        if (idx < items) {
            const long long pix = idx / (4 * H);
            const long long b = pix / HW, hw = pix % HW;
            const uint32_t st = ring + stage * kStage;
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const long long r = (b * F + f) * HW + hw;
                const h16* row = qkv + r * ld + h * D + c * 8;
                ts_cp16(st + (4 * f + 0) * kVec, row);
                ts_cp16(st + (4 * f + 1) * kVec, row + HD);
                ts_cp16(st + (4 * f + 2) * kVec, row + 2 * HD);
                ts_cp16(st + (4 * f + 3) * kVec, dout + r * HD + h * D + c * 8);
            }
        }
        ts_commit();
    };
#pragma unroll
    for (int s0 = 0; s0 < TS_STAGES - 1; ++s0) issue(idx0 + s0 * stride, s0);
    int stage = 0;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < items; base += stride) {
        const long long idx = base + threadIdx.x;
        {
            int ns = stage + TS_STAGES - 1;
            if (ns >= TS_STAGES) ns -= TS_STAGES;
            issue(idx + (TS_STAGES - 1) * stride, ns);
        }
        ts_wait();
        const bool valid = idx < items;  // warp-uniform trip count: shuffles use the full mask
        const long long pix = valid ? idx / (4 * H) : 0;
        const long long b = pix / HW, hw = pix % HW;
        const uint32_t st = ring + stage * kStage;
        if (++stage == TS_STAGES) stage = 0;
        float q[F][8], k[F][8], v[F][8], go[F][8];
#pragma unroll
        for (int f = 0; f < F; ++f) {
            lds8(st + (4 * f + 0) * kVec, q[f]);
            lds8(st + (4 * f + 1) * kVec, k[f]);
            lds8(st + (4 * f + 2) * kVec, v[f]);
            lds8(st + (4 * f + 3) * kVec, go[f]);
            rope8(q[f], cf[f], sf[f], scale);
            rope8(k[f], cf[f], sf[f], 1.f);
        }
        float p[F][F], ds[F][F];
#pragma unroll
        for (int i = 0; i < F; ++i) {
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < F; ++j) {
                p[i][j] = qsum(dot8(q[i], k[j])) + __ldg(bias + (h * F + i) * F + j);
                m = fmaxf(m, p[i][j]);
            }
            float l = 0.f;
#pragma unroll
            for (int j = 0; j < F; ++j) {
                p[i][j] = __expf(p[i][j] - m);
                l += p[i][j];
            }
            const float inv = 1.f / l;
            float Di = 0.f;
#pragma unroll
            for (int j = 0; j < F; ++j) {
                p[i][j] *= inv;
                ds[i][j] = qsum(dot8(go[i], v[j]));  // dp_ij
                Di = fmaf(p[i][j], ds[i][j], Di);
            }
#pragma unroll
            for (int j = 0; j < F; ++j) {
                ds[i][j] = p[i][j] * (ds[i][j] - Di);
                if (valid) dbacc[i][j] += ds[i][j];
            }
        }
#pragma unroll
        for (int f = 0; f < F; ++f) {
            float dq[8], dk[8], dv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) dq[e] = dk[e] = dv[e] = 0.f;
#pragma unroll
            for (int j = 0; j < F; ++j)
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    dq[e] = fmaf(ds[f][j], k[j][e], dq[e]);   // query role: f = i
                    dk[e] = fmaf(ds[j][f], q[j][e], dk[e]);   // key role:   f = j, sum over queries
                    dv[e] = fmaf(p[j][f], go[j][e], dv[e]);
                }
            rope8_t(dq, cf[f], sf[f], scale);
            rope8_t(dk, cf[f], sf[f], 1.f);
            h16* drow = dqkv + ((b * F + f) * HW + hw) * ld + h * D + c * 8;
            if (valid) {
                st8(drow, dq);
                st8(drow + HD, dk);
                st8(drow + 2 * HD, dv);
            }
        }
    }
    if (c == 0) {
#pragma unroll
        for (int i = 0; i < F; ++i)
#pragma unroll
            for (int j = 0; j < F; ++j) atomicAdd(&sbias[(h * F + i) * F + j], dbacc[i][j]);
    }
    __syncthreads();
    for (int x = threadIdx.x; x < H * F * F; x += blockDim.x) atomicAdd(&dbias[x], sbias[x]);
}

}  // namespace cesm

using namespace cesm;

extern "C" int cesm_tattn_fwd(const void* qkv, const float* bias, const float* cs, const float* sn, void* out,
                              float* lse, int B, int F, int HW, int H, int dim_head, float scale, void* stream) {
    CESM_REQUIRE(dim_head == D, "temporal attention kernel needs dim_head == 32 (got %d)", dim_head);
    CESM_REQUIRE(H >= 1 && H <= 8, "temporal attention kernel supports 1..8 heads (got %d)", H);
    cudaStream_t st = as_stream(stream);
    if (F <= 4) {
        const long long npix = (long long)B * HW, items = npix * H * 4;
        long long want = (items + 255) / 256;
        if (want > 148 * 2) want = 148 * 2;  // persistent: one resident wave (two 110 KB rings per SM at F = 3)
        const int blocks = (int)want;
#define TATTN_FWD(FF)                                                                                               \
    {                                                                                                               \
        constexpr int kSm = TS_STAGES * 3 * FF * 256 * 16;                                                          \
        static bool cfg = false;                                                                                    \
        if (!cfg) {                                                                                                 \
            CESM_CHECK_CUDA(cudaFuncSetAttribute(tattn_small_fwd_kernel<FF>,                                        \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, kSm));                \
            cfg = true;                                                                                             \
        }                                                                                                           \
        launch_pdl(tattn_small_fwd_kernel<FF>, blocks, 256, kSm, st, (const h16*)qkv, bias, cs, sn,               \
                                                             (h16*)out, lse, npix, HW, H, scale);         \
    }
        switch (F) {
            case 1: TATTN_FWD(1) break;
            case 2: TATTN_FWD(2) break;
            case 3: TATTN_FWD(3) break;
            default: TATTN_FWD(4) break;
        }
#undef TATTN_FWD
        CESM_CHECK_LAUNCH();
        return CESM_OK;
    }
    CESM_REQUIRE(lse != nullptr, "lse is required for F > 4");
    const long long total = (long long)B * F * HW * H;
    const int blocks = (int)((total + 127) / 128);
    launch_pdl(tattn_fwd_kernel, blocks, 128, 0, st, (const h16*)qkv, bias, cs, sn, (h16*)out, lse, B, F, HW,
                                             H, scale);
    CESM_CHECK_LAUNCH();
    return CESM_OK;
}

extern "C" int cesm_tattn_bwd(const void* qkv, const float* bias, const float* cs, const float* sn, const void* out,
                              const float* lse, const void* dout, void* dqkv, float* dbias, int B, int F, int HW, int H,
                              int dim_head, float scale, void* stream) {
    CESM_REQUIRE(dim_head == D, "temporal attention kernel needs dim_head == 32 (got %d)", dim_head);
    CESM_REQUIRE(H >= 1 && H <= 8, "temporal attention kernel supports 1..8 heads (got %d)", H);
    cudaStream_t st = as_stream(stream);
    CESM_ZERO_SCRATCH(dbias, sizeof(float) * H * F * F, st);
    if (F <= 4) {
        const long long npix = (long long)B * HW, items = npix * H * 4;
        long long want = (items + 127) / 128;
        const int unit = H;  // blocks of 128 threads: a grid that is a multiple of H keeps (h, c) fixed per thread
        long long cap = 148LL * 3 * 4;
        if (want > cap) want = cap;
        const int blocks = (int)(((want + unit - 1) / unit) * unit);
#define TATTN_BWD(FF)                                                                                          \
    {                                                                                                          \
        constexpr int kSm = TS_STAGES * 4 * FF * 128 * 16;                                                     \
        static bool cfg = false;                                                                               \
        if (!cfg) {                                                                                            \
            CESM_CHECK_CUDA(cudaFuncSetAttribute(tattn_small_bwd_kernel<FF>,                                   \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, kSm));           \
            cfg = true;                                                                                        \
        }                                                                                                      \
        launch_pdl(tattn_small_bwd_kernel<FF>, blocks, 128, kSm, st, (const h16*)qkv, bias, cs, sn,          \
                                                             (const h16*)dout, (h16*)dqkv, \
                                                             dbias, npix, HW, H, scale);                       \
    }
        switch (F) {
            case 1: TATTN_BWD(1) break;
            case 2: TATTN_BWD(2) break;
            case 3: TATTN_BWD(3) break;
            default: TATTN_BWD(4) break;
        }
#undef TATTN_BWD
        CESM_CHECK_LAUNCH();
        return CESM_OK;
    }
    CESM_REQUIRE(out != nullptr && lse != nullptr, "out and lse are required for F > 4");
    const long long total = (long long)B * F * HW * H;
    const int blocks = (int)((total + 127) / 128);
    const size_t sh = (H * F * F <= 2048) ? sizeof(float) * H * F * F : 0;
    launch_pdl(tattn_bwd_kernel, blocks, 128, sh, st, (const h16*)qkv, bias, cs, sn, (const h16*)out, lse,
                                              (const h16*)dout, (h16*)dqkv, dbias, B, F, HW, H,
                                              scale);
    CESM_CHECK_LAUNCH();
    return CESM_OK;
}
