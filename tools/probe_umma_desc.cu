// Hardware probe (developer tool, not part of the product library): which shared-memory matrix
// descriptors does tcgen05.mma accept for row-shifted views of a SWIZZLE_128B tile?
//
// Answers two design questions for the conv kernels:
//  (1) K-major A tile whose start address is shifted by s rows (s*128 B, not 1024-B aligned):
//      needed for halo reuse (one smem patch serves all 9 taps of a 3x3 conv).
//  (2) MN-major operands (K = pixels) and K-shifted views of them: needed for weight gradients
//      straight from NHWC activations.
// Integer-valued bf16 inputs make every result exact in fp32, so the check is bitwise.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_umma_desc probe_umma_desc.cu
#include <cstdlib>
#include <vector>

#include "../cesm_emulator_b200/csrc/common.cuh"

using namespace cesm;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct ProbeArgs {
    int a_mn_major, b_mn_major;
    int a_rows, b_rows;        // rows per TMA box (K-major: M/N rows + slack; MN-major: K rows + slack)
    int a_atoms, b_atoms;      // MN-major: number of 64-wide MN atoms (separate boxes)
    uint32_t a_shift_bytes, b_shift_bytes;
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    uint32_t a_kstep_bytes, b_kstep_bytes;  // start-address advance per UMMA_K=16
    uint32_t a_base_off, b_base_off;
    int n;  // UMMA N
    float* d;
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap map_a,
                                                    const __grid_constant__ CUtensorMap map_b, ProbeArgs p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sa = base;                 // up to 48 KB
    const uint32_t sb = base + 48 * 1024;     // up to 48 KB
    const uint32_t bar = base + 96 * 1024;
    const uint32_t bar2 = bar + 8;
    const uint32_t tptr = bar + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(bar2, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tptr, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + 96 * 1024 + 16);
    if (threadIdx.x == 0) {
        const uint32_t a_bytes = p.a_rows * 128 * p.a_atoms, b_bytes = p.b_rows * 128 * p.b_atoms;
        mbar_arrive_expect_tx(bar, a_bytes + b_bytes);
        for (int i = 0; i < p.a_atoms; ++i) tma_load_2d(sa + i * p.a_rows * 128, &map_a, bar, i * 64, 0);
        for (int i = 0; i < p.b_atoms; ++i) tma_load_2d(sb + i * p.b_rows * 128, &map_b, bar, i * 64, 0);
        mbar_wait(bar, 0, 10);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, p.n, p.a_mn_major, p.b_mn_major);
        for (int k = 0; k < 4; ++k) {
            uint64_t da = make_smem_desc_sw128(sa + p.a_shift_bytes + k * p.a_kstep_bytes, p.a_lbo, p.a_sbo, p.a_base_off);
            uint64_t db = make_smem_desc_sw128(sb + p.b_shift_bytes + k * p.b_kstep_bytes, p.b_lbo, p.b_sbo, p.b_base_off);
            umma_bf16(tmem, da, db, idesc, k != 0);
        }
        umma_commit(bar2);
    }
    mbar_wait(bar2, 0, 11);
    tc_fence_after();
    for (int cc = 0; cc < p.n; cc += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + cc, v);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) p.d[(warp * 32 + lane) * p.n + cc + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 64);
}

static EncodeTiledFn g_enc;
static CUtensorMap make_map(void* ptr, uint64_t inner, uint64_t rows, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t str[1] = {inner * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = g_enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("encode failed %d\n", (int)r);
        exit(2);
    }
    return m;
}

static float bf(int v) { return (float)v; }

int main() {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) {
        printf("no driver entry point\n");
        return 2;
    }
    g_enc = (EncodeTiledFn)fp;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int N = 64;
    srand(7);
    // ---------------- K-major: A [144 rows][64 k], B [64 n][64 k] ----------------
    const int AR = 144;
    std::vector<__nv_bfloat16> hA(AR * 64), hB(N * 64);
    std::vector<int> iA(AR * 64), iB(N * 64);
    for (int i = 0; i < AR * 64; ++i) { iA[i] = rand() % 5 - 2; hA[i] = __float2bfloat16(bf(iA[i])); }
    for (int i = 0; i < N * 64; ++i) { iB[i] = rand() % 5 - 2; hB[i] = __float2bfloat16(bf(iB[i])); }
    __nv_bfloat16 *dA, *dB;
    float* dD;
    cudaMalloc(&dA, hA.size() * 2);
    cudaMalloc(&dB, hB.size() * 2);
    cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap mA = make_map(dA, 64, AR, AR), mB = make_map(dB, 64, N, N);
    std::vector<float> hD(128 * N);
    printf("== K-major A shifted by s rows (start += s*128 B) ==\n");
    for (int s : {0, 1, 2, 3, 4, 7, 8, 9, 16}) {
        for (int mode = 0; mode < 2; ++mode) {
            ProbeArgs p{};
            p.a_rows = AR; p.b_rows = N; p.a_atoms = 1; p.b_atoms = 1;
            p.a_shift_bytes = s * 128; p.a_lbo = 0; p.a_sbo = 1024; p.b_lbo = 0; p.b_sbo = 1024;
            p.a_kstep_bytes = 32; p.b_kstep_bytes = 32;
            p.a_base_off = mode ? (s & 7) : 0;
            p.n = N; p.d = dD;
            cudaMemset(dD, 0xff, 128 * N * 4);
            probe_kernel<<<1, 128, 100 * 1024>>>(mA, mB, p);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("s=%d mode=%d CUDA error %s\n", s, mode, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hD.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int r = 0; r < 128; ++r)
                for (int n = 0; n < N; ++n) {
                    int acc = 0;
                    for (int k = 0; k < 64; ++k) acc += iA[(r + s) * 64 + k] * iB[n * 64 + k];
                    if (hD[r * N + n] != (float)acc) ++bad;
                }
            printf("kmajor shift=%2d base_offset=%d : %s (%d mismatches)\n", s, p.a_base_off, bad ? "WRONG" : "exact", bad);
        }
    }
    // ---------------- MN-major: At [K=80 rows][128 m], Bt [80][64 n] ----------------
    const int KR = 80;
    std::vector<__nv_bfloat16> hAt(KR * 128), hBt(KR * N);
    std::vector<int> iAt(KR * 128), iBt(KR * N);
    for (int i = 0; i < KR * 128; ++i) { iAt[i] = rand() % 5 - 2; hAt[i] = __float2bfloat16(bf(iAt[i])); }
    for (int i = 0; i < KR * N; ++i) { iBt[i] = rand() % 5 - 2; hBt[i] = __float2bfloat16(bf(iBt[i])); }
    __nv_bfloat16 *dAt, *dBt;
    cudaMalloc(&dAt, hAt.size() * 2);
    cudaMalloc(&dBt, hBt.size() * 2);
    cudaMemcpy(dAt, hAt.data(), hAt.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dBt, hBt.data(), hBt.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap mAt = make_map(dAt, 128, KR, KR), mBt = make_map(dBt, N, KR, KR);
    printf("== MN-major A (2 atoms) and B (1 atom), K shifted by s rows ==\n");
    for (int s : {0, 1, 2, 8, 9}) {
        for (int mode = 0; mode < 4; ++mode) {
            // mode bit0: base_offset = s&7 ; mode bit1: swap LBO/SBO roles
            ProbeArgs p{};
            p.a_mn_major = 1; p.b_mn_major = 1;
            p.a_rows = KR; p.b_rows = KR; p.a_atoms = 2; p.b_atoms = 1;
            p.a_shift_bytes = s * 128; p.b_shift_bytes = s * 128;
            uint32_t atom_stride = KR * 128;  // bytes between the two 64-wide MN atoms of A
            if (!(mode & 2)) { p.a_lbo = atom_stride; p.a_sbo = 1024; p.b_lbo = atom_stride; p.b_sbo = 1024; }
            else { p.a_lbo = 1024; p.a_sbo = atom_stride; p.b_lbo = 1024; p.b_sbo = atom_stride; }
            p.a_kstep_bytes = 16 * 128; p.b_kstep_bytes = 16 * 128;
            p.a_base_off = (mode & 1) ? (s & 7) : 0; p.b_base_off = p.a_base_off;
            p.n = N; p.d = dD;
            cudaMemset(dD, 0xff, 128 * N * 4);
            probe_kernel<<<1, 128, 100 * 1024>>>(mAt, mBt, p);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mn s=%d mode=%d CUDA error %s\n", s, mode, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hD.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    int acc = 0;
                    for (int k = 0; k < 64; ++k) acc += iAt[(k + s) * 128 + m] * iBt[(k + s) * N + n];
                    if (hD[m * N + n] != (float)acc) ++bad;
                }
            printf("mnmajor shift=%d base_offset=%d lbo/sbo=%s : %s (%d mismatches)\n", s, p.a_base_off,
                   (mode & 2) ? "swapped" : "lbo=atom,sbo=1024", bad ? "WRONG" : "exact", bad);
        }
    }
    return 0;
}
