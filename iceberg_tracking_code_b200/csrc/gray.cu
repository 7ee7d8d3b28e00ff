// K0: RGB(A) u8 -> gray u8 in OpenCV's fixed point.
// Replaces cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY), s1_lucaskanade_tracking.py:283,311.
// HBM-bound: 3 B read + 1 B write per pixel.  Vector path: one thread turns 16 pixels
// (3 x 16-byte loads) into one 16-byte store; a warp covers 1536 contiguous input bytes.
#include "common.cuh"

namespace ibt {

template <int SH>
__device__ __forceinline__ uint32_t gray1(uint32_t c0, uint32_t c1, uint32_t c2, int k0, int k1, int k2)
{
    return (c0 * k0 + c1 * k1 + c2 * k2 + (1u << (SH - 1))) >> SH;
}

// 16 pixels from 12 packed words (cn == 3).
template <int SH>
__device__ __forceinline__ uint4 gray16_c3(const uint4 a, const uint4 b, const uint4 c, int k0, int k1, int k2)
{
    const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {          // 4 pixels = 3 words
        const uint32_t w0 = w[3 * q], w1 = w[3 * q + 1], w2 = w[3 * q + 2];
        uint32_t g0 = gray1<SH>(w0 & 0xff, (w0 >> 8) & 0xff, (w0 >> 16) & 0xff, k0, k1, k2);
        uint32_t g1 = gray1<SH>(w0 >> 24, w1 & 0xff, (w1 >> 8) & 0xff, k0, k1, k2);
        uint32_t g2 = gray1<SH>((w1 >> 16) & 0xff, w1 >> 24, w2 & 0xff, k0, k1, k2);
        uint32_t g3 = gray1<SH>((w2 >> 8) & 0xff, (w2 >> 16) & 0xff, w2 >> 24, k0, k1, k2);
        o[q] = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

template <int SH>
__global__ void __launch_bounds__(256)
gray_c3_vec_kernel(const uint8_t *__restrict__ src, int64_t src_pitch, uint8_t *__restrict__ dst,
                   int64_t dst_pitch, int H, int units_per_row, int64_t total_units, int k0, int k1, int k2)
{
    pdl_launch_dependents();                              // (the pyramid launch that follows waits for this grid before it reads)
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < total_units;
         u += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(u / units_per_row);
        const int ux = (int)(u - (int64_t)y * units_per_row);
        const uint4 *p = reinterpret_cast<const uint4 *>(src + y * src_pitch + (int64_t)ux * 48);
        const uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        *reinterpret_cast<uint4 *>(dst + y * dst_pitch + (int64_t)ux * 16) = gray16_c3<SH>(a, b, c, k0, k1, k2);
    }
}

// Generic path: any cn (3/4), any pitch/alignment, and the row tails of the vector path.
template <int SH>
__global__ void __launch_bounds__(256)
gray_scalar_kernel(const uint8_t *__restrict__ src, int64_t src_pitch, int cn, uint8_t *__restrict__ dst,
                   int64_t dst_pitch, int H, int x_begin, int W, int k0, int k1, int k2)
{
    const int wspan = W - x_begin;
    const int64_t total = (int64_t)H * wspan;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / wspan);
        const int x = x_begin + (int)(i - (int64_t)y * wspan);
        const uint8_t *p = src + y * src_pitch + (int64_t)x * cn;
        dst[y * dst_pitch + x] = (uint8_t)gray1<SH>(p[0], p[1], p[2], k0, k1, k2);
    }
}

template <int SH>
static int launch_gray(const uint8_t *src, int H, int W, int cn, int64_t src_pitch, uint8_t *dst, int64_t dst_pitch,
                       int k0, int k1, int k2, cudaStream_t st)
{
    int x_done = 0;
    const bool vec_ok = cn == 3 && (reinterpret_cast<uintptr_t>(src) % 16 == 0) && (src_pitch % 16 == 0) &&
                        (reinterpret_cast<uintptr_t>(dst) % 16 == 0) && (dst_pitch % 16 == 0) && W >= 16;
    if (vec_ok) {
        const int upr = W / 16;
        const int64_t total = (int64_t)H * upr;
        int64_t blocks = (total + 255) / 256;
        const int64_t maxb = (int64_t)kNumSMs * 8 * 4;     // 8 resident CTAs/SM x 4 waves, grid-stride beyond
        if (blocks > maxb) blocks = maxb;
        gray_c3_vec_kernel<SH><<<(unsigned)blocks, 256, 0, st>>>(src, src_pitch, dst, dst_pitch, H, upr, total, k0, k1, k2);
        x_done = upr * 16;
    }
    if (x_done < W) {
        const int64_t total = (int64_t)H * (W - x_done);
        int64_t blocks = (total + 255) / 256;
        const int64_t maxb = (int64_t)kNumSMs * 8 * 4;
        if (blocks > maxb) blocks = maxb;
        gray_scalar_kernel<SH><<<(unsigned)blocks, 256, 0, st>>>(src, src_pitch, cn, dst, dst_pitch, H, x_done, W, k0, k1, k2);
    }
    return check_launch("ibt_gray_u8");
}

} // namespace ibt

IBT_API int ibt_gray_u8(const uint8_t *src, int H, int W, int cn, int64_t src_pitch, uint8_t *dst,
                        int64_t dst_pitch, int coeffset, void *stream)
{
    if (!src || !dst || H <= 0 || W <= 0 || (cn != 3 && cn != 4) || src_pitch < (int64_t)W * cn || dst_pitch < W)
        return IBT_E_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (coeffset == IBT_GRAY_CV4_15BIT)
        return ibt::launch_gray<15>(src, H, W, cn, src_pitch, dst, dst_pitch, 3735, 19235, 9798, st);
    if (coeffset == IBT_GRAY_CV3_14BIT)
        return ibt::launch_gray<14>(src, H, W, cn, src_pitch, dst, dst_pitch, 1868, 9617, 4899, st);
    return IBT_E_INVALID;
}
