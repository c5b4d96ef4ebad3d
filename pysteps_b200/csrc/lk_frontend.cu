// lk_frontend.cu -- the per-frame front end of dense Lucas-Kanade as three passes over the frame
// instead of seven (sm_100a): what dense_lucaskanade does to a frame before any feature is
// looked for --
//   np.ma.masked_invalid + fill value          pysteps/motion/lucaskanade.py:207-219
//   utils.images.morph_opening (3x3 cross)     pysteps/utils/images.py:60-86
//   min / max of the opened image (4 row sets) pysteps/tracking/lucaskanade.py:144-160,
//   scaling to uint8 for the tracker ...       pysteps/feature/shitomasi.py:131-151
//   ... and for the detector, buffered mask
// Pass A  mask + min/max of the raw frame (the opening's threshold is the frame minimum);
// Pass B  opening recomputed from the raw frame in shared memory + min/max/count of the opened
//         image -- the opened float64 image is never written to HBM;
// Pass C  opening again + both uint8 scalings + the detector's validity map.
// (A global minimum separates A from B and B from C, hence three kernels; the 32 MB frame stays
// in the 126 MB L2 between them.)
//
// Passes B and C are stencils with a 2-pixel halo (opening = erode o dilate with the 3x3 cross:
// 5x5 footprint; 5x5 mask buffer).  Their float64 tiles are staged in shared memory by TMA:
// one elected thread issues cp.async.bulk.tensor.2d for a (64+4) x (16+4) box -- halo included,
// out-of-image parts filled by the copy engine -- into a two-deep ring, an mbarrier per slot
// counts the bytes in, and the CTA computes tile i while the copy of tile i+1 is in flight.
// Persistent CTAs (a multiple of the SM count) walk the tiles.  Results are those of the
// stand-alone stage kernels of lk_dense.cu bit for bit (tests/test_lk_gpu.py).
#include <cuda.h>

#include "common.cuh"
#include "lk_common.cuh"
#include "quantise_body.cuh"

namespace {

constexpr int FW = 64, FH = 16;          // pixel tile of one step
constexpr int HALO = 2;
constexpr int BW = FW + 2 * HALO, BH = FH + 2 * HALO;   // TMA box: 68 x 20 float64 (544-byte rows)
constexpr int FTHREADS = 256;
constexpr int MW = 80;                   // mask tile row pitch (bytes)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// box (x0 .. x0+BW, y0 .. y0+BH) of the tensor map -> shared memory; coordinates may lie outside
// the image (halo of border tiles): those elements are filled by the copy engine
__device__ __forceinline__ void tma_load_box(void *dst, const CUtensorMap *map, int x0, int y0, unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x0), "r"(y0), "r"(smem_u32(bar)) : "memory");
}

struct FrontParams {
    const uint8_t *mask;     // (m, n) from pass A
    const double *stats0;    // [min, max, count] of the raw frame
    const double *stats;     // 12 doubles of pass B (pass C only)
    int m, n, opening, dil, f32;
    MM *part;                // pass B: [4][nparts]
    int nparts;
    uint8_t *q_track, *q_det, *valid;   // pass C (q_det / valid may be null)
};

// Per tile the stencil is evaluated in three sweeps over shared memory instead of per output pixel
// (25 threshold tests per pixel otherwise): the threshold image of the box, its erosion, then the
// output pixels (dilation of the erosion = the opening).
//   s_bin: 0 background / masked, 1 foreground, 2 outside the image
//   s_ero: eroded foreground (out-of-image neighbours do not erode), 0 outside the image
struct TileGeom {
    int x0, y0, m, n;
    __device__ __forceinline__ bool inside(int ly, int lx) const {
        const int y = y0 + ly - HALO, x = x0 + lx - HALO;
        return y >= 0 && y < m && x >= 0 && x < n;
    }
};

// k x k buffered mask (feature/shitomasi.py:131-137), out-of-image taps ignored (their mask byte is 0)
__device__ __forceinline__ bool buffered(const uint8_t *msk, int ly, int lx, int dil) {
    bool d = false;
    const int r = dil / 2;
    for (int dy = -r; dy <= dil - 1 - r; dy++)
        for (int dx = -r; dx <= dil - 1 - r; dx++) d |= msk[(ly + dy) * MW + lx + dx] != 0;
    return d;
}

// PASS = 1: statistics of the opened image.  PASS = 2: uint8 images.
template <int PASS>
__global__ void __launch_bounds__(FTHREADS)
front_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ FrontParams p) {
    __shared__ __align__(128) double s_img[2][BH * BW];
    __shared__ __align__(16) uint8_t s_msk[2][BH * MW];
    __shared__ uint8_t s_bin[BH * MW], s_ero[BH * MW];
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ MM s_mm[32];
    const int tid = threadIdx.x;
    const int m = p.m, n = p.n;
    const int tiles_x = (n + FW - 1) / FW, tiles = tiles_x * ((m + FH - 1) / FH);
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    constexpr unsigned BOX_BYTES = BW * BH * sizeof(double);
    auto issue = [&](int t, int slot) {
        const int x0 = (t % tiles_x) * FW, y0 = (t / tiles_x) * FH;
        mbar_expect_tx(&s_bar[slot], BOX_BYTES);
        tma_load_box(s_img[slot], &tmap, x0 - HALO, y0 - HALO, &s_bar[slot]);
    };
    int t = blockIdx.x;
    if (t < tiles && tid == 0) issue(t, 0);
    // mask bytes of a tile (1 byte per pixel, 2-pixel halo, zero outside the image): fetched into
    // registers one tile ahead, so their L2 latency hides behind the current tile's arithmetic
    constexpr int MPT = (BH * BW + FTHREADS - 1) / FTHREADS;
    uint8_t mreg[MPT];
    auto fetch_mask = [&](int tt) {
        const int x0 = (tt % tiles_x) * FW, y0 = (tt / tiles_x) * FH;
#pragma unroll
        for (int k = 0; k < MPT; k++) {
            const int e = tid + k * FTHREADS;
            const int ly = e / BW, lx = e - ly * BW;
            const int y = y0 + ly - HALO, x = x0 + lx - HALO;
            mreg[k] = (e < BH * BW && y >= 0 && y < m && x >= 0 && x < n) ? p.mask[(size_t)y * n + x] : 0;
        }
    };
    if (t < tiles) fetch_mask(t);

    const double minval = p.stats0[0];
    const bool opening = p.opening != 0;
    const bool any_masked = p.stats0[2] < (double)m * (double)n;
    MM acc[4];
    if (PASS == 1) {
#pragma unroll
        for (int k = 0; k < 4; k++) { acc[k].mn = CUDART_INF; acc[k].mx = -CUDART_INF; acc[k].cnt = 0; }
    }
    for (int it = 0; t < tiles; t += gridDim.x, it++) {
        const int slot = it & 1;
        const int tn = t + gridDim.x;
        // the other slot was last read in iteration it-1, which ended with a block barrier
        if (tn < tiles && tid == 0) issue(tn, slot ^ 1);
        TileGeom G;
        G.x0 = (t % tiles_x) * FW; G.y0 = (t / tiles_x) * FH; G.m = m; G.n = n;
        const double *img = s_img[slot];
        uint8_t *msk = s_msk[slot];
#pragma unroll
        for (int k = 0; k < MPT; k++) {
            const int e = tid + k * FTHREADS;
            if (e < BH * BW) msk[(e / BW) * MW + e % BW] = mreg[k];
        }
        if (tn < tiles) fetch_mask(tn);
        mbar_wait(&s_bar[slot], (unsigned)((it >> 1) & 1));
        __syncthreads();
        if (opening) {
            // sweep 1: utils/images.py:66-70 filled > thr (masked pixels count as background)
            for (int e = tid; e < BH * BW; e += FTHREADS) {
                const int ly = e / BW, lx = e - ly * BW;
                s_bin[ly * MW + lx] = !G.inside(ly, lx) ? 2 : ((!msk[ly * MW + lx] && img[ly * BW + lx] > minval) ? 1 : 0);
            }
            __syncthreads();
            // sweep 2: erosion with the 3x3 cross on the box minus its outermost ring
            for (int e = tid; e < (BH - 2) * (BW - 2); e += FTHREADS) {
                const int ly = 1 + e / (BW - 2), lx = 1 + e % (BW - 2);
                const int c = s_bin[ly * MW + lx];
                s_ero[ly * MW + lx] = (c == 1) & (s_bin[(ly - 1) * MW + lx] != 0) & (s_bin[(ly + 1) * MW + lx] != 0) &
                                      (s_bin[ly * MW + lx - 1] != 0) & (s_bin[ly * MW + lx + 1] != 0);
            }
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < FW * FH / FTHREADS; k++) {
            const int lx = tid % FW, ly = tid / FW + k * (FTHREADS / FW);
            const int x = G.x0 + lx, y = G.y0 + ly;
            if (x >= n || y >= m) continue;
            const int hy = ly + HALO, hx = lx + HALO;
            double v = img[hy * BW + hx];
            if (opening && s_bin[hy * MW + hx] == 1) {
                // sweep 3: a foreground pixel survives the opening if the cross around it holds an
                // eroded pixel; otherwise it takes the frame minimum (:72-81)
                const bool keep = s_ero[hy * MW + hx] | s_ero[(hy - 1) * MW + hx] | s_ero[(hy + 1) * MW + hx] |
                                  s_ero[hy * MW + hx - 1] | s_ero[hy * MW + hx + 1];
                if (!keep) v = minval;
            }
            const bool mk0 = msk[hy * MW + hx] != 0;
            if (PASS == 1) {
                if (!mk0) {
                    acc[0].mn = fmin(acc[0].mn, v); acc[0].mx = fmax(acc[0].mx, v); acc[0].cnt++;
                    if (y >= 1) { acc[1].mn = fmin(acc[1].mn, v); acc[1].mx = fmax(acc[1].mx, v); acc[1].cnt++; }
                    if (y >= 2) { acc[2].mn = fmin(acc[2].mn, v); acc[2].mx = fmax(acc[2].mx, v); acc[2].cnt++; }
                }
                bool d = mk0;
                if (p.dil > 0 && any_masked) d = buffered(msk, hy, hx, p.dil);
                if (!d) acc[3].cnt++;
            } else {
                const size_t i = (size_t)y * n + x;
                const double *st = p.stats;
                auto scale = [&](double val, int set) -> uint8_t {
                    const double im_min = st[3 * set + 0], im_max = st[3 * set + 1];
                    double q;
                    if (p.f32) q = qz::scale_f32(val, im_min, im_max);
                    else if (__dsub_rn(im_max, im_min) > 1e-8)
                        q = __dmul_rn(__ddiv_rn(__dsub_rn(val, im_min), __dsub_rn(im_max, im_min)), 255.0);
                    else q = __dsub_rn(val, im_min);
                    return cast_u8(q);
                };
                const double fill = st[0];
                // tracking/lucaskanade.py:144-160
                p.q_track[i] = scale(mk0 ? fill : v, 0);
                if (p.q_det) {
                    // feature/shitomasi.py:131-151 (see quantise_kernel of lk_dense.cu, mode 1)
                    bool dmask = mk0;
                    if (p.dil > 0 && any_masked) dmask = buffered(msk, hy, hx, p.dil);
                    if (p.valid) p.valid[i] = dmask ? 0 : 1;
                    const bool any_clear = st[11] > 0.0;
                    bool mk = mk0;
                    int set = 0;
                    if (p.dil > 0) {
                        set = (any_clear ? 1 : 0) + (any_masked ? 1 : 0);
                        if ((y == 0 && any_clear) || (y == 1 && any_masked)) mk = true;
                        if (!any_clear && any_masked) { set = 2; if (y == 0) mk = true; }
                    }
                    p.q_det[i] = scale(mk ? fill : v, set);
                }
            }
        }
        __syncthreads();  // everyone is done with this slot before it is refilled
    }
    if (PASS == 1) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const MM r = mm_block(acc[k], s_mm);
            if (tid == 0) p.part[(size_t)k * p.nparts + blockIdx.x] = r;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

}  // namespace

extern "C" int b200_lk_frontend(const double *img, const uint8_t *user_mask, int m, int n, int size_opening,
                                int buffer_mask, int flags, uint8_t *mask, double *stats0, double *stats,
                                uint8_t *q_track, uint8_t *q_det, uint8_t *valid, void *stream) {
    B200_REQUIRE(img && mask && stats0 && stats && q_track && m >= 1 && n >= 1, "bad arguments");
    B200_REQUIRE(size_opening == 0 || size_opening == 3, "only the 3x3 structuring element is implemented");
    B200_REQUIRE(buffer_mask >= 0 && buffer_mask <= 5, "buffer_mask must be 0..5 for the fused front end");
    // TMA needs 16-byte aligned rows: an even number of float64 columns
    B200_REQUIRE(n % 2 == 0 && ((uintptr_t)img & 15) == 0, "the fused front end needs an even width and a 16-byte aligned frame");
    cudaStream_t s = (cudaStream_t)stream;
    // ---- pass A: mask + min/max of the raw frame (kernels of lk_dense.cu) ---------------------
    if (int rc = b200_mask_invalid(img, user_mask, m, n, mask, stats0, stream)) return rc;
    // ---- tensor map of the frame: 2-D float64, box 68 x 20, zero fill outside ------------------
    EncodeTiledFn enc = encode_tiled();
    if (!enc) {
        b200::set_error("cuTensorMapEncodeTiled is not available from this driver");
        return B200_ENOTSUP;
    }
    CUtensorMap tmap;
    const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)m};
    const cuuint64_t strides[1] = {(cuuint64_t)n * sizeof(double)};
    const cuuint32_t box[2] = {BW, BH};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)img, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        b200::set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr);
        return B200_EINVAL;
    }
    const int tiles = b200::ceil_div(n, FW) * b200::ceil_div(m, FH);
    const int nparts = std::min(tiles, b200::num_sms() * 4);  // persistent CTAs: a multiple of the SM count
    b200::Scratch part;
    B200_CUDA(part.alloc(sizeof(MM) * nparts * 4, s));
    FrontParams p;
    memset(&p, 0, sizeof(p));
    p.mask = mask; p.stats0 = stats0; p.stats = stats; p.m = m; p.n = n; p.opening = size_opening != 0;
    p.dil = buffer_mask; p.f32 = (flags & B200_QUANTISE_F32) != 0; p.part = (MM *)part.p; p.nparts = nparts;
    p.q_track = q_track; p.q_det = q_det; p.valid = valid;
    front_kernel<1><<<nparts, FTHREADS, 0, s>>>(tmap, p);
    B200_LAUNCH_CHECK();
    mm_final_kernel<<<1, 256, 0, s>>>((const MM *)part.p, nparts, 4, stats);
    B200_LAUNCH_CHECK();
    front_kernel<2><<<nparts, FTHREADS, 0, s>>>(tmap, p);
    B200_LAUNCH_CHECK();
    return 0;
}
