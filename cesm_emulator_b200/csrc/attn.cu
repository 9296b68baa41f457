// Attention cores (bf16 activations, fp32 math).  Head dim is 32 for both attention types
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

__device__ __forceinline__ void load32(const __nv_bfloat16* p, float (&f)[D]) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 u = __ldg(q + j);
        float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        f[8 * j + 0] = a.x; f[8 * j + 1] = a.y; f[8 * j + 2] = b.x; f[8 * j + 3] = b.y;
        f[8 * j + 4] = c.x; f[8 * j + 5] = c.y; f[8 * j + 6] = d.x; f[8 * j + 7] = d.y;
    }
}
__device__ __forceinline__ void store32(__nv_bfloat16* p, const float (&f)[D]) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
        u.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
        u.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
        u.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
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
tattn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ bias,
                 const float* __restrict__ cs, const float* __restrict__ sn, __nv_bfloat16* __restrict__ out,
                 float* __restrict__ lse, int B, int F, int HW, int H, float scale) {
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
tattn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ bias,
                 const float* __restrict__ cs, const float* __restrict__ sn, const __nv_bfloat16* __restrict__ out,
                 const float* __restrict__ lse, const __nv_bfloat16* __restrict__ dout,
                 __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dbias, int B, int F, int HW, int H,
                 float scale) {
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
// spatial linear attention
//   qkv: [NI*n][3*H*D], image ni = rows [ni*n, (ni+1)*n)
// ------------------------------------------------------------------------------------------------
// per (image, k column) partial (max, sum exp) over a strip of pixels; thread = one column
__global__ void __launch_bounds__(256)
la_kstats_partial_kernel(const __nv_bfloat16* __restrict__ qkv, float* __restrict__ part, int n, int HD, int nstrips) {
    const int ni = blockIdx.y, strip = blockIdx.x;
    const int col = threadIdx.x;  // HD == 256 threads (checked on host) or loop
    const int per = (n + nstrips - 1) / nstrips;
    const int p0 = strip * per, p1 = min(n, p0 + per);
    for (int c = col; c < HD; c += blockDim.x) {
        float m = -INFINITY, z = 0.f;
        const __nv_bfloat16* base = qkv + ((size_t)ni * n) * (3 * HD) + HD + c;
        for (int p = p0; p < p1; ++p) {
            const float v = __bfloat162float(base[(size_t)p * 3 * HD]);
            const float mn = fmaxf(m, v);
            z = z * __expf(m - mn) + __expf(v - mn);
            m = mn;
        }
        float* o = part + (((size_t)ni * nstrips + strip) * HD + c) * 2;
        o[0] = m;
        o[1] = z;
    }
}
__global__ void la_kstats_combine_kernel(const float* __restrict__ part, float* __restrict__ kstat, int HD,
                                         int nstrips, int total) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // ni*HD + c
    if (idx >= total) return;
    const int ni = idx / HD, c = idx % HD;
    float m = -INFINITY, z = 0.f;
    for (int s = 0; s < nstrips; ++s) {
        const float* p = part + (((size_t)ni * nstrips + s) * HD + c) * 2;
        if (p[1] == 0.f) continue;
        const float mn = fmaxf(m, p[0]);
        z = z * __expf(m - mn) + p[1] * __expf(p[0] - mn);
        m = mn;
    }
    kstat[(size_t)idx * 2] = m;
    kstat[(size_t)idx * 2 + 1] = z;
}

// qk[row][0:HD] = scale*softmax_d(q), qk[row][HD:2HD] = exp(k-m)/Z ; thread per (row, head)
__global__ void __launch_bounds__(128)
la_prep_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ kstat, __nv_bfloat16* __restrict__ qk,
               long long rows, int n, int H, float scale) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * H) return;
    const int h = t % H;
    const long long row = t / H;
    const int ni = row / n;
    const int HD = H * D;
    float q[D], k[D];
    load32(qkv + row * 3 * HD + h * D, q);
    load32(qkv + row * 3 * HD + HD + h * D, k);
    float m = q[0];
#pragma unroll
    for (int i = 1; i < D; ++i) m = fmaxf(m, q[i]);
    float z = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        q[i] = __expf(q[i] - m);
        z += q[i];
    }
    const float inv = scale / z;
#pragma unroll
    for (int i = 0; i < D; ++i) q[i] *= inv;
    const float* ks = kstat + ((size_t)ni * HD + h * D) * 2;
#pragma unroll
    for (int i = 0; i < D; ++i) k[i] = __expf(k[i] - ks[2 * i]) / ks[2 * i + 1];
    store32(qk + row * 2 * HD + h * D, q);
    store32(qk + row * 2 * HD + HD + h * D, k);
}

// ctx[ni][h][d][e] += sum_pixels a[p][h*D+d] * b[p][h*D+e]   (a, b: bf16 with row pitches lda, ldb)
// block = (pixel chunk, head, image); each warp owns a slice of the chunk, lane = column e.
__global__ void __launch_bounds__(256)
la_context_kernel(const __nv_bfloat16* __restrict__ a, int lda, const __nv_bfloat16* __restrict__ b, int ldb,
                  float* __restrict__ ctx, int n, int H, int chunk) {
    const int ni = blockIdx.z, h = blockIdx.y;
    const int p0 = blockIdx.x * chunk, p1 = min(n, p0 + chunk);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = 0.f;
    const __nv_bfloat16* ab = a + ((size_t)ni * n) * lda + h * D;
    const __nv_bfloat16* bb = b + ((size_t)ni * n) * ldb + h * D;
    for (int p = p0 + warp; p < p1; p += 8) {
        const float av = __bfloat162float(ab[(size_t)p * lda + lane]);  // a[p][d = lane]
        const float bv = __bfloat162float(bb[(size_t)p * ldb + lane]);  // b[p][e = lane]
#pragma unroll
        for (int d = 0; d < D; ++d) acc[d] = fmaf(__shfl_sync(0xffffffffu, av, d), bv, acc[d]);
    }
    __shared__ float red[8][D][D + 1];
#pragma unroll
    for (int d = 0; d < D; ++d) red[warp][d][lane] = acc[d];
    __syncthreads();
    for (int x = threadIdx.x; x < D * D; x += 256) {
        const int d = x / D, e = x % D;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][d][e];
        atomicAdd(&ctx[(((size_t)ni * H + h) * D + d) * D + e], s);
    }
}

// out[p][h*D+e] = sum_d ctx[ni][h][d][e] * q[p][h*D+d]   ; thread per (pixel, head)
__global__ void __launch_bounds__(128)
la_apply_kernel(const __nv_bfloat16* __restrict__ qk, const float* __restrict__ ctx, __nv_bfloat16* __restrict__ out,
                int n, int H, int chunk) {
    __shared__ float sc[D][D];
    const int ni = blockIdx.z, h = blockIdx.y;
    const int HD = H * D;
    for (int x = threadIdx.x; x < D * D; x += blockDim.x) sc[x / D][x % D] = ctx[((size_t)ni * H + h) * D * D + x];
    __syncthreads();
    const int p1 = min(n, (int)(blockIdx.x + 1) * chunk);
    for (int p = blockIdx.x * chunk + threadIdx.x; p < p1; p += blockDim.x) {
        const size_t row = (size_t)ni * n + p;
        float q[D], o[D];
        load32(qk + row * 2 * HD + h * D, q);
#pragma unroll
        for (int e = 0; e < D; ++e) o[e] = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
#pragma unroll
            for (int e = 0; e < D; ++e) o[e] = fmaf(sc[d][e], q[d], o[e]);
        }
        store32(out + row * HD + h * D, o);
    }
}

// delta[ni][h*D+d] = sum_e dctx[d][e] * ctx[d][e]
__global__ void la_delta_kernel(const float* __restrict__ ctx, const float* __restrict__ dctx, float* __restrict__ delta,
                                int total) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (ni*H + h)*D + d
    if (idx >= total) return;
    float s = 0.f;
    for (int e = 0; e < D; ++e) s = fmaf(dctx[(size_t)idx * D + e], ctx[(size_t)idx * D + e], s);
    delta[idx] = s;
}

// Backward of prep+apply for one (pixel, head):
//   dqh = ctx dout ; dkh = dctx v ; dv = dctx^T kh ; dq = qsm*(g - sum qsm*g), g = scale*dqh ;
//   dk = kh*(dkh - delta)
__global__ void __launch_bounds__(128)
la_bwd_apply_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ qk,
                    const __nv_bfloat16* __restrict__ dout, const float* __restrict__ ctx,
                    const float* __restrict__ dctx, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv,
                    int n, int H, int chunk, float scale) {
    __shared__ float sc[D][D], sd[D][D], sdel[D];
    const int ni = blockIdx.z, h = blockIdx.y;
    const int HD = H * D;
    for (int x = threadIdx.x; x < D * D; x += blockDim.x) {
        sc[x / D][x % D] = ctx[((size_t)ni * H + h) * D * D + x];
        sd[x / D][x % D] = dctx[((size_t)ni * H + h) * D * D + x];
    }
    if (threadIdx.x < D) sdel[threadIdx.x] = delta[((size_t)ni * H + h) * D + threadIdx.x];
    __syncthreads();
    const int p1 = min(n, (int)(blockIdx.x + 1) * chunk);
    for (int p = blockIdx.x * chunk + threadIdx.x; p < p1; p += blockDim.x) {
        const size_t row = (size_t)ni * n + p;
        float go[D], t[D], r[D];
        load32(dout + row * HD + h * D, go);
        // dq
        load32(qk + row * 2 * HD + h * D, t);  // scale*softmax(q)
        float dot = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            float g = 0.f;
#pragma unroll
            for (int e = 0; e < D; ++e) g = fmaf(sc[d][e], go[e], g);
            r[d] = g;              // dqh[d]
            dot = fmaf(t[d], g, dot);  // sum_d (scale*sm_d) * dqh_d
        }
        // t = scale*sm ; dq_d = sm_d*(scale*dqh_d - sum_j sm_j*scale*dqh_j) = t_d*(dqh_d - dot/scale)
        const float dots = dot / scale;
#pragma unroll
        for (int d = 0; d < D; ++d) r[d] = t[d] * (r[d] - dots);
        store32(dqkv + row * 3 * HD + h * D, r);
        // dk, dv
        float v[D];
        load32(qkv + row * 3 * HD + 2 * HD + h * D, v);
        load32(qk + row * 2 * HD + HD + h * D, t);  // kh
        float dv[D];
#pragma unroll
        for (int e = 0; e < D; ++e) dv[e] = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            float g = 0.f;
#pragma unroll
            for (int e = 0; e < D; ++e) {
                g = fmaf(sd[d][e], v[e], g);
                dv[e] = fmaf(sd[d][e], t[d], dv[e]);
            }
            r[d] = t[d] * (g - sdel[d]);
        }
        store32(dqkv + row * 3 * HD + HD + h * D, r);
        store32(dqkv + row * 3 * HD + 2 * HD + h * D, dv);
    }
}

}  // namespace cesm

using namespace cesm;

extern "C" int cesm_tattn_fwd(const void* qkv, const float* bias, const float* cs, const float* sn, void* out,
                              float* lse, int B, int F, int HW, int H, int dim_head, float scale, void* stream) {
    CESM_REQUIRE(dim_head == D, "temporal attention kernel needs dim_head == 32 (got %d)", dim_head);
    const long long total = (long long)B * F * HW * H;
    const int blocks = (int)((total + 127) / 128);
    tattn_fwd_kernel<<<blocks, 128, 0, as_stream(stream)>>>((const __nv_bfloat16*)qkv, bias, cs, sn, (__nv_bfloat16*)out,
                                                            lse, B, F, HW, H, scale);
    CESM_CHECK_LAUNCH();
    return CESM_OK;
}

extern "C" int cesm_tattn_bwd(const void* qkv, const float* bias, const float* cs, const float* sn, const void* out,
                              const float* lse, const void* dout, void* dqkv, float* dbias, int B, int F, int HW, int H,
                              int dim_head, float scale, void* stream) {
    CESM_REQUIRE(dim_head == D, "temporal attention kernel needs dim_head == 32 (got %d)", dim_head);
    cudaStream_t st = as_stream(stream);
    CESM_CHECK_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * H * F * F, st));
    const long long total = (long long)B * F * HW * H;
    const int blocks = (int)((total + 127) / 128);
    const size_t sh = (H * F * F <= 2048) ? sizeof(float) * H * F * F : 0;
    tattn_bwd_kernel<<<blocks, 128, sh, st>>>((const __nv_bfloat16*)qkv, bias, cs, sn, (const __nv_bfloat16*)out, lse,
                                              (const __nv_bfloat16*)dout, (__nv_bfloat16*)dqkv, dbias, B, F, HW, H,
                                              scale);
    CESM_CHECK_LAUNCH();
    return CESM_OK;
}

static int la_chunk(int n) { return n >= 4096 ? 1024 : (n >= 512 ? 256 : 64); }

// kstat: [NI][H*D][2] (max, sumexp); part: scratch [NI][nstrips<=64][H*D][2]; qk: [NI*n][2*H*D];
// ctx: [NI][H][D][D] fp32; out: [NI*n][H*D]
extern "C" int cesm_linattn_fwd(const void* qkv, float* part, float* kstat, void* qk, float* ctx, void* out, int NI,
                                int n, int H, int dim_head, float scale, void* stream) {
    CESM_REQUIRE(dim_head == D, "linear attention kernel needs dim_head == 32 (got %d)", dim_head);
    cudaStream_t st = as_stream(stream);
    const int HD = H * D;
    const int nstrips = n >= 64 * 64 ? 64 : (n >= 256 ? 16 : 1);
    la_kstats_partial_kernel<<<dim3(nstrips, NI), 256, 0, st>>>((const __nv_bfloat16*)qkv, part, n, HD, nstrips);
    CESM_CHECK_LAUNCH();
    la_kstats_combine_kernel<<<ceil_div(NI * HD, 256), 256, 0, st>>>(part, kstat, HD, nstrips, NI * HD);
    CESM_CHECK_LAUNCH();
    const long long rows = (long long)NI * n;
    la_prep_kernel<<<(int)((rows * H + 127) / 128), 128, 0, st>>>((const __nv_bfloat16*)qkv, kstat, (__nv_bfloat16*)qk,
                                                                  rows, n, H, scale);
    CESM_CHECK_LAUNCH();
    CESM_CHECK_CUDA(cudaMemsetAsync(ctx, 0, sizeof(float) * NI * H * D * D, st));
    const int chunk = la_chunk(n);
    dim3 grid(ceil_div(n, chunk), H, NI);
    const __nv_bfloat16* qkp = (const __nv_bfloat16*)qk;
    la_context_kernel<<<grid, 256, 0, st>>>(qkp + HD, 2 * HD, (const __nv_bfloat16*)qkv + 2 * HD, 3 * HD, ctx, n, H, chunk);
    CESM_CHECK_LAUNCH();
    la_apply_kernel<<<grid, 128, 0, st>>>(qkp, ctx, (__nv_bfloat16*)out, n, H, chunk);
    CESM_CHECK_LAUNCH();
    return CESM_OK;
}

// dctx, delta: scratch [NI][H][D][D], [NI][H][D]
extern "C" int cesm_linattn_bwd(const void* qkv, const void* qk, const float* ctx, const void* dout, float* dctx,
                                float* delta, void* dqkv, int NI, int n, int H, int dim_head, float scale,
                                void* stream) {
    CESM_REQUIRE(dim_head == D, "linear attention kernel needs dim_head == 32 (got %d)", dim_head);
    cudaStream_t st = as_stream(stream);
    const int HD = H * D;
    CESM_CHECK_CUDA(cudaMemsetAsync(dctx, 0, sizeof(float) * NI * H * D * D, st));
    const int chunk = la_chunk(n);
    dim3 grid(ceil_div(n, chunk), H, NI);
    // dctx[d][e] = sum_p (scale*softmax(q))[p][d] * dout[p][e]
    la_context_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)qk, 2 * HD, (const __nv_bfloat16*)dout, HD, dctx, n, H,
                                            chunk);
    CESM_CHECK_LAUNCH();
    la_delta_kernel<<<ceil_div(NI * HD, 128), 128, 0, st>>>(ctx, dctx, delta, NI * HD);
    CESM_CHECK_LAUNCH();
    la_bwd_apply_kernel<<<grid, 128, 0, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)qk,
                                              (const __nv_bfloat16*)dout, ctx, dctx, delta, (__nv_bfloat16*)dqkv, n, H,
                                              chunk, scale);
    CESM_CHECK_LAUNCH();
    return CESM_OK;
}
