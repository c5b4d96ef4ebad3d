// lk_common.cuh -- pieces shared by the dense Lucas-Kanade stage kernels (lk_dense.cu) and their
// fused TMA-tiled front end (lk_frontend.cu): deterministic min / max / count reductions and the
// NumPy float -> uint8 cast.
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace {

// ---------------------------------------------------------------- block min/max reduce
struct MM {
    double mn, mx;
    unsigned long long cnt;
};

__device__ __forceinline__ MM mm_merge(MM a, const MM &b) {
    a.mn = fmin(a.mn, b.mn);
    a.mx = fmax(a.mx, b.mx);
    a.cnt += b.cnt;
    return a;
}

__device__ __forceinline__ MM mm_warp(MM v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MM t;
        t.mn = __shfl_xor_sync(0xffffffffu, v.mn, o);
        t.mx = __shfl_xor_sync(0xffffffffu, v.mx, o);
        t.cnt = __shfl_xor_sync(0xffffffffu, v.cnt, o);
        v = mm_merge(v, t);
    }
    return v;
}

__device__ __forceinline__ MM mm_block(MM v, MM *sm) {
    v = mm_warp(v);
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int w = tid >> 5, l = tid & 31;
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
    __syncthreads();
    if (l == 0) sm[w] = v;
    __syncthreads();
    if (w == 0) {
        MM t;
        t.mn = CUDART_INF; t.mx = -CUDART_INF; t.cnt = 0;
        if (l < nw) t = sm[l];
        v = mm_warp(t);
    }
    return v;
}

// stats layout: [min, max, count] (+3 per set)
__device__ __forceinline__ void mm_write_stats(double *__restrict__ stats, int s, const MM &v) {
    // numpy's masked min()/max() of an all-masked array is `masked`; report NaN
    stats[3 * s + 0] = v.cnt ? v.mn : CUDART_NAN;
    stats[3 * s + 1] = v.cnt ? v.mx : CUDART_NAN;
    stats[3 * s + 2] = (double)v.cnt;
}

__global__ void __launch_bounds__(256) mm_final_kernel(const MM *__restrict__ part, int nparts, int nsets,
                                                       double *__restrict__ stats) {
    __shared__ MM sm[32];
    for (int s = 0; s < nsets; s++) {
        MM v;
        v.mn = CUDART_INF; v.mx = -CUDART_INF; v.cnt = 0;
        for (int i = threadIdx.x; i < nparts; i += blockDim.x) v = mm_merge(v, part[(size_t)s * nparts + i]);
        v = mm_block(v, sm);
        if (threadIdx.x == 0) mm_write_stats(stats, s, v);
        __syncthreads();
    }
}

// The same final reduction WITHOUT a second launch: every CTA publishes its partials, takes a
// ticket, and the CTA that draws the last one reduces them all (`ticket` is zeroed by the caller on
// the stream before the launch).  Order of the merges: min / max / integer counts are associative,
// so the statistics are those of mm_final_kernel bit for bit.  Call from ALL threads of the CTA,
// after thread 0 has stored this CTA's partials (flat thread index `tid`, 1-D or 2-D blocks).
__device__ __forceinline__ void mm_finish(const MM *part, int nparts, int nsets, double *__restrict__ stats,
                                          unsigned *ticket, MM *sm, int tid, int nthreads) {
    __shared__ int s_last;
    if (tid == 0) {
        __threadfence();  // the partials before the ticket
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int s = 0; s < nsets; s++) {
        MM v;
        v.mn = CUDART_INF; v.mx = -CUDART_INF; v.cnt = 0;
        for (int i = tid; i < nparts; i += nthreads) {
            const MM *q = part + (size_t)s * nparts + i;
            MM t;
            t.mn = __ldcg(&q->mn); t.mx = __ldcg(&q->mx); t.cnt = __ldcg(&q->cnt);  // other CTAs' stores: from L2
            v = mm_merge(v, t);
        }
        v = mm_block(v, sm);
        if (tid == 0) mm_write_stats(stats, s, v);
        __syncthreads();
    }
}

// (x - im_min) / (im_max - im_min) * 255 -> astype(uint8): truncation toward zero, and the
// x86-64 behaviour of NumPy for out-of-range values (through int32, low byte kept)
__device__ __forceinline__ uint8_t cast_u8(double v) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return 0;  // cvttsd2si -> INT_MIN -> low byte 0
    return (uint8_t)((int)v & 0xff);
}

}  // namespace
