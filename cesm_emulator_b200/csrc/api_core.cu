// C-ABI core: error reporting, TMA descriptor cache, and the cesm_igemm entry point
// (tile-shape selection + tensor-map construction for igemm.cu).
#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>

#include "api_common.h"
#include "igemm.h"

namespace cesm {

static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int set_error(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

// ------------------------------------------------------------------------------------------------
// TMA descriptors
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn resolve_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct MapKey {
    uint64_t v[14];
    bool operator==(const MapKey& o) const {
        for (int i = 0; i < 14; ++i)
            if (v[i] != o.v[i]) return false;
        return true;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        uint64_t h = 1469598103934665603ull;
        for (int i = 0; i < 14; ++i) {
            h ^= k.v[i];
            h *= 1099511628211ull;
        }
        return static_cast<size_t>(h);
    }
};
static std::mutex g_map_mutex;  // autograd runs backward on its own threads
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

int get_tensor_map_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box) {
    CESM_REQUIRE(rank >= 2 && rank <= 4, "tensor map rank %d unsupported", rank);
    CESM_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15u) == 0, "tensor base %p not 16-byte aligned", base);
    MapKey key{};
    key.v[0] = reinterpret_cast<uint64_t>(base);
    key.v[1] = static_cast<uint64_t>(rank);
    for (int i = 0; i < rank; ++i) {
        key.v[2 + i] = dims[i];
        key.v[6 + i] = (i == 0) ? 0 : strides_bytes[i - 1];
        key.v[10 + i] = box[i];
    }
    {
        std::lock_guard<std::mutex> lk(g_map_mutex);
        auto it = g_map_cache.find(key);
        if (it != g_map_cache.end()) {
            *out = it->second;
            return CESM_OK;
        }
    }
    EncodeTiledFn enc = resolve_encode();
    if (!enc) return set_error(CESM_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    cuuint64_t gdim[4], gstr[3];
    cuuint32_t bx[4], es[4];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) {
            gstr[i - 1] = strides_bytes[i - 1];
            CESM_REQUIRE(gstr[i - 1] % 16 == 0, "tensor map stride %llu not a multiple of 16 B",
                         (unsigned long long)gstr[i - 1]);
        }
        CESM_REQUIRE(bx[i] >= 1 && bx[i] <= 256, "tensor map box[%d]=%u out of range", i, bx[i]);
    }
    alignas(64) CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(CESM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    {
        std::lock_guard<std::mutex> lk(g_map_mutex);
        if (g_map_cache.size() > 8192) g_map_cache.clear();
        g_map_cache.emplace(key, m);
    }
    *out = m;
    return CESM_OK;
}

// ------------------------------------------------------------------------------------------------
// tile-shape selection: split the 128 rows of an M tile over (w, h, n) to waste the fewest rows
// ------------------------------------------------------------------------------------------------
static void choose_tile(int n, int oh, int ow, int* bw, int* bh, int* bn) {
    long long best_tiles = -1;
    int b_w = 1, b_h = 1, b_n = 1;
    for (int w = 1; w <= 128 && w <= ow; ++w) {
        // only widths that are the full row, or powers of two, keep epilogue stores well formed
        if (!(w == ow || (w & (w - 1)) == 0)) continue;
        for (int h = 1; h * w <= 128 && h <= oh; ++h) {
            if (!(h == oh || (h & (h - 1)) == 0)) continue;
            int nn = 128 / (w * h);
            if (nn > n) nn = n;
            if (nn < 1) nn = 1;
            if (w != ow || h != oh) nn = 1;  // only batch whole images into one tile
            long long tiles = 1LL * ceil_div(ow, w) * ceil_div(oh, h) * ceil_div(n, nn);
            if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && w > b_w)) {
                best_tiles = tiles;
                b_w = w;
                b_h = h;
                b_n = nn;
            }
        }
    }
    *bw = b_w;
    *bh = b_h;
    *bn = b_n;
}

}  // namespace cesm

using namespace cesm;

extern "C" const char* cesm_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* cesm_version(void) { return "cesm_b200 0.1 sm_100a"; }
extern "C" long long cesm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int cesm_igemm(const cesm_igemm_args* a, void* stream) {
    CESM_REQUIRE(a != nullptr, "args is NULL");
    CESM_REQUIRE(a->c0 > 0 && a->c0 % 64 == 0 && a->c1 >= 0 && a->c1 % 64 == 0,
                 "channel counts must be multiples of 64 (c0=%d c1=%d)", a->c0, a->c1);
    CESM_REQUIRE(a->cout > 0 && a->cout % 64 == 0, "cout=%d must be a multiple of 64", a->cout);
    CESM_REQUIRE(a->num_taps >= 1 && a->num_taps <= CESM_MAX_TAPS, "num_taps=%d out of range", a->num_taps);
    CESM_REQUIRE(a->stride == 1 || a->stride == 2, "stride=%d unsupported", a->stride);
    CESM_REQUIRE(a->stride == 1 || (a->c1 == 0 && a->h % 2 == 0 && a->w % 2 == 0),
                 "stride 2 needs a single source with even h, w");
    CESM_REQUIRE(a->n > 0 && a->h > 0 && a->w > 0 && a->oh > 0 && a->ow > 0, "empty geometry");
    CESM_REQUIRE((a->c1 == 0) == (a->a1 == nullptr), "a1 / c1 mismatch");
    CESM_REQUIRE(a->ldo % 8 == 0 && (a->residual == nullptr || a->ldr % 8 == 0), "row pitches must be multiples of 8");

    IgemmParams p{};
    p.c0 = a->c0;
    p.c1 = a->c1;
    p.num_taps = a->num_taps;
    p.n = a->n;
    p.oh = a->oh;
    p.ow = a->ow;
    choose_tile(a->n, a->oh, a->ow, &p.bw, &p.bh, &p.bn);
    p.a_box_bytes = 64u * 2u * p.bw * p.bh * p.bn;
    p.cout = a->cout;
    p.out = a->out;
    p.out_fp32 = a->out_fp32;
    p.ldo = a->ldo;
    p.out_h = a->out_h;
    p.out_w = a->out_w;
    p.o_sh = a->o_sh;
    p.o_sw = a->o_sw;
    p.o_h0 = a->o_h0;
    p.o_w0 = a->o_w0;
    p.bias = a->bias;
    p.residual = a->residual;
    p.ldr = a->ldr;

    CUtensorMap amaps[4];
    int n_amaps = 0;
    const uint32_t box[4] = {64u, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    if (a->stride == 1) {
        const void* src[2] = {a->a0, a->a1};
        const int cs[2] = {a->c0, a->c1};
        for (int s = 0; s < 2; ++s) {
            if (!src[s]) continue;
            const uint64_t dims[4] = {(uint64_t)cs[s], (uint64_t)a->w, (uint64_t)a->h, (uint64_t)a->n};
            const uint64_t str[3] = {(uint64_t)cs[s] * 2, (uint64_t)a->w * cs[s] * 2, (uint64_t)a->h * a->w * cs[s] * 2};
            int rc = get_tensor_map_bf16(&amaps[s], src[s], 4, dims, str, box);
            if (rc) return rc;
            n_amaps = s + 1;
        }
        for (int t = 0; t < a->num_taps; ++t) {
            p.tap_map[t] = 0;
            p.tap_dh[t] = a->tap_dh[t];
            p.tap_dw[t] = a->tap_dw[t];
        }
    } else {
        // stride 2: four parity sub-images of the source, each a unit-stride [n, h/2, w/2, c] view
        const int c = a->c0;
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                const char* base = static_cast<const char*>(a->a0) + (size_t)(ph * a->w + pw) * c * 2;
                const uint64_t dims[4] = {(uint64_t)c, (uint64_t)a->w / 2, (uint64_t)a->h / 2, (uint64_t)a->n};
                const uint64_t str[3] = {(uint64_t)c * 4, (uint64_t)a->w * c * 4, (uint64_t)a->h * a->w * c * 2};
                int rc = get_tensor_map_bf16(&amaps[ph * 2 + pw], base, 4, dims, str, box);
                if (rc) return rc;
            }
        n_amaps = 4;
        for (int t = 0; t < a->num_taps; ++t) {
            const int ph = a->tap_dh[t] & 1, pw = a->tap_dw[t] & 1;
            p.tap_map[t] = ph * 2 + pw;
            p.tap_dh[t] = (a->tap_dh[t] - ph) / 2;
            p.tap_dw[t] = (a->tap_dw[t] - pw) / 2;
        }
    }

    const int ktot = a->num_taps * (a->c0 + a->c1);
    const int block_n = (a->cout % 256 == 0) ? 256 : (a->cout % 128 == 0 ? 128 : 64);
    CUtensorMap bmap;
    {
        const uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)a->cout};
        const uint64_t str[1] = {(uint64_t)ktot * 2};
        const uint32_t bbox[2] = {64u, (uint32_t)block_n};
        int rc = get_tensor_map_bf16(&bmap, a->wt, 2, dims, str, bbox);
        if (rc) return rc;
    }
    note_launch();
    CESM_CHECK_CUDA(igemm_launch(amaps, n_amaps, bmap, p, block_n, as_stream(stream)));
    return CESM_OK;
}

// 64-pixel K tile for the weight gradient: powers of two with bw*bh*bn == 64, fewest tiles
static void choose_tile64(int n, int oh, int ow, int* bw, int* bh, int* bn) {
    long long best = -1;
    for (int w = 1; w <= 64; w <<= 1)
        for (int h = 1; h * w <= 64; h <<= 1) {
            const int nn = 64 / (w * h);
            const long long tiles = 1LL * ceil_div(ow, w) * ceil_div(oh, h) * ceil_div(n, nn);
            if (best < 0 || tiles < best || (tiles == best && w > *bw)) {
                best = tiles;
                *bw = w;
                *bh = h;
                *bn = nn;
            }
        }
}

extern "C" int cesm_wgrad(const cesm_wgrad_args* a, void* stream) {
    CESM_REQUIRE(a != nullptr, "args is NULL");
    CESM_REQUIRE(a->c0 > 0 && a->c0 % 64 == 0 && a->c1 >= 0 && a->c1 % 64 == 0,
                 "channel counts must be multiples of 64 (c0=%d c1=%d)", a->c0, a->c1);
    CESM_REQUIRE(a->cout > 0 && a->cout % 64 == 0, "cout=%d must be a multiple of 64", a->cout);
    CESM_REQUIRE(a->num_taps >= 1 && a->num_taps <= CESM_MAX_TAPS, "num_taps=%d out of range", a->num_taps);
    CESM_REQUIRE(a->stride == 1 || a->stride == 2, "stride=%d unsupported", a->stride);
    CESM_REQUIRE(a->stride == 1 || (a->c1 == 0 && a->h % 2 == 0 && a->w % 2 == 0),
                 "stride 2 needs a single source with even h, w");
    CESM_REQUIRE((a->c1 == 0) == (a->x1 == nullptr), "x1 / c1 mismatch");
    cudaStream_t st = as_stream(stream);

    WgradParams p{};
    p.c0 = a->c0;
    p.c1 = a->c1;
    p.num_taps = a->num_taps;
    p.n = a->n;
    p.oh = a->oh;
    p.ow = a->ow;
    choose_tile64(a->n, a->oh, a->ow, &p.bw, &p.bh, &p.bn);
    p.box_bytes = 64u * 2u * 64u;
    p.cout = a->cout;
    p.dw = a->dw;
    const uint32_t box[4] = {64u, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};

    CUtensorMap xmaps[4];
    int n_xmaps = 0;
    if (a->stride == 1) {
        const void* src[2] = {a->x0, a->x1};
        const int cs[2] = {a->c0, a->c1};
        for (int s = 0; s < 2; ++s) {
            if (!src[s]) continue;
            const uint64_t dims[4] = {(uint64_t)cs[s], (uint64_t)a->w, (uint64_t)a->h, (uint64_t)a->n};
            const uint64_t str[3] = {(uint64_t)cs[s] * 2, (uint64_t)a->w * cs[s] * 2, (uint64_t)a->h * a->w * cs[s] * 2};
            int rc = get_tensor_map_bf16(&xmaps[s], src[s], 4, dims, str, box);
            if (rc) return rc;
            n_xmaps = s + 1;
        }
        for (int t = 0; t < a->num_taps; ++t) {
            p.tap_map[t] = 0;
            p.tap_dh[t] = a->tap_dh[t];
            p.tap_dw[t] = a->tap_dw[t];
        }
    } else {
        const int c = a->c0;
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                const char* base = static_cast<const char*>(a->x0) + (size_t)(ph * a->w + pw) * c * 2;
                const uint64_t dims[4] = {(uint64_t)c, (uint64_t)a->w / 2, (uint64_t)a->h / 2, (uint64_t)a->n};
                const uint64_t str[3] = {(uint64_t)c * 4, (uint64_t)a->w * c * 4, (uint64_t)a->h * a->w * c * 2};
                int rc = get_tensor_map_bf16(&xmaps[ph * 2 + pw], base, 4, dims, str, box);
                if (rc) return rc;
            }
        n_xmaps = 4;
        for (int t = 0; t < a->num_taps; ++t) {
            const int ph = a->tap_dh[t] & 1, pw = a->tap_dw[t] & 1;
            p.tap_map[t] = ph * 2 + pw;
            p.tap_dh[t] = (a->tap_dh[t] - ph) / 2;
            p.tap_dw[t] = (a->tap_dw[t] - pw) / 2;
        }
    }
    CUtensorMap ymap;
    {
        const int co = a->cout;
        const char* base = static_cast<const char*>(a->dy) + ((size_t)a->y_h0 * a->y_w + a->y_w0) * co * 2;
        const uint64_t dims[4] = {(uint64_t)co, (uint64_t)a->ow, (uint64_t)a->oh, (uint64_t)a->n};
        const uint64_t str[3] = {(uint64_t)a->y_sw * co * 2, (uint64_t)a->y_sh * a->y_w * co * 2,
                                 (uint64_t)a->y_h * a->y_w * co * 2};
        int rc = get_tensor_map_bf16(&ymap, base, 4, dims, str, box);
        if (rc) return rc;
    }
    const int ctot = a->c0 + a->c1;
    CESM_CHECK_CUDA(cudaMemsetAsync(a->dw, 0, sizeof(float) * (size_t)a->cout * a->num_taps * ctot, st));
    const int block_n = (a->cout % 256 == 0) ? 256 : (a->cout % 128 == 0 ? 128 : 64);
    const int units = a->num_taps * (ctot / 64);
    const int base_ctas = ((units + 1) / 2) * (a->cout / block_n);
    const int tiles = ceil_div(a->ow, p.bw) * ceil_div(a->oh, p.bh) * ceil_div(a->n, p.bn);
    int ksplit = ceil_div(148 * 2, base_ctas);
    if (ksplit > tiles) ksplit = tiles;
    if (ksplit < 1) ksplit = 1;
    note_launch();
    CESM_CHECK_CUDA(wgrad_launch(xmaps, n_xmaps, ymap, p, block_n, ksplit, st));
    return CESM_OK;
}
