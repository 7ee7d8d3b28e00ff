// K1: fused pyrDown (5x5 binomial /256) + Scharr derivative for one pyramid level
// (SURVEY.md A.2, A.4).  Replaces the buildOpticalFlowPyramid / calcScharrDeriv work that
// cv2.calcOpticalFlowPyrLK repeats on every call (s1_lucaskanade_tracking.py:323,326).
//
// HBM-bound: per level-l pixel 1 B is read once, 4 B (int16 dx,dy interleaved) and 1/4 B
// (level l+1) are written.  A CTA stages a (TH+3) x (TW+8) u8 halo tile in shared memory
// (REFLECT_101 applied while staging), then every thread slides a 4-pixel-wide column strip down
// 16 rows: the horizontal halves of both separable filters are computed once per staged row and
// kept in registers, the vertical halves finish them; each derivative row leaves as one 16-byte
// store per thread (512 contiguous bytes per warp).
#include "common.cuh"

namespace ibt {

constexpr int TW = 128;                 // tile width  (input pixels)
constexpr int TH = 32;                  // tile height (input pixels), 2 warps x 16 rows
constexpr int RPW = 16;                 // rows per warp
constexpr int HX = 4;                   // staged columns left/right of the tile (keeps words aligned)
constexpr int SROWS = TH + 3;           // rows y0-2 .. y0+TH
constexpr int SWORDS = (TW + 2 * HX) / 4;   // 34 words per staged row (conflict-free word reads)

__device__ __forceinline__ uint32_t pack_i16(int lo, int hi)
{
    return (static_cast<uint32_t>(lo) & 0xffffu) | (static_cast<uint32_t>(hi) << 16);
}

template <bool DERIV, bool DOWN>
__global__ void __launch_bounds__(64)
pyr_level_kernel(const uint8_t *__restrict__ src, int h, int w, int64_t pitch,
                 uint8_t *__restrict__ deriv, int64_t dpitch,
                 uint8_t *__restrict__ down, int64_t downpitch, int src_word_ok, int deriv_vec_ok)
{
    __shared__ uint32_t tile[SROWS * SWORDS];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;

    // ---- stage the halo tile -------------------------------------------------------------
    for (int idx = tid; idx < SROWS * SWORDS; idx += 64) {
        const int r = idx / SWORDS, cw = idx - r * SWORDS;
        const int gx = x0 - HX + 4 * cw;
        const uint8_t *rowp = src + (int64_t)r101(y0 - 2 + r, h) * pitch;
        uint32_t v;
        if (src_word_ok && gx >= 0 && gx + 3 < w) {
            v = __ldg(reinterpret_cast<const uint32_t *>(rowp + gx));
        } else {
            v = (uint32_t)rowp[r101(gx, w)] | ((uint32_t)rowp[r101(gx + 1, w)] << 8) |
                ((uint32_t)rowp[r101(gx + 2, w)] << 16) | ((uint32_t)rowp[r101(gx + 3, w)] << 24);
        }
        tile[idx] = v;
    }
    __syncthreads();

    const int tx = tid & 31, wy = tid >> 5;
    const int x = x0 + 4 * tx;
    const int ybase = y0 + RPW * wy;
    if (x >= w || ybase >= h) return;

    // horizontal halves kept for the previous rows
    int hx1[4], hx2[4], hs1[4], hs2[4];     // row j-1, row j-2
    int hp[5][2];                           // pyrDown horizontal sums of rows j-4 .. j
#pragma unroll
    for (int i = 0; i < 4; i++) { hx1[i] = hx2[i] = hs1[i] = hs2[i] = 0; }
#pragma unroll
    for (int i = 0; i < 5; i++) { hp[i][0] = hp[i][1] = 0; }

    const uint32_t *trow = tile + (RPW * wy) * SWORDS + tx;
#pragma unroll
    for (int j = 0; j <= RPW + 2; j++) {            // staged rows ybase-2 .. ybase+16
        const uint32_t w0 = trow[j * SWORDS], w1 = trow[j * SWORDS + 1], w2 = trow[j * SWORDS + 2];
        // p[k] = pixel at column x - 4 + k
        const int p2 = (w0 >> 16) & 0xff, p3 = w0 >> 24;
        const int p4 = w1 & 0xff, p5 = (w1 >> 8) & 0xff, p6 = (w1 >> 16) & 0xff, p7 = w1 >> 24;
        const int p8 = w2 & 0xff;
        const int p[7] = {p2, p3, p4, p5, p6, p7, p8};   // p[k] here = column x - 2 + k

        int hx[4], hs[4];
        if (DERIV) {
#pragma unroll
            for (int i = 0; i < 4; i++) {                // column x+i is p[2+i]
                hx[i] = p[3 + i] - p[1 + i];
                hs[i] = 3 * (p[1 + i] + p[3 + i]) + 10 * p[2 + i];
            }
            if (j >= 3) {                                // emit derivative row of staged row j-1
                const int yo = ybase + j - 3;
                if (yo < h) {
                    uint32_t o[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int dx = 3 * (hx2[i] + hx[i]) + 10 * hx1[i];
                        const int dy = hs[i] - hs2[i];
                        o[i] = pack_i16(dx, dy);
                    }
                    uint8_t *dp = deriv + (int64_t)yo * dpitch + (int64_t)x * 4;
                    if (deriv_vec_ok && x + 3 < w) {
                        *reinterpret_cast<uint4 *>(dp) = make_uint4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            if (x + i < w) reinterpret_cast<uint32_t *>(dp)[i] = o[i];
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) { hx2[i] = hx1[i]; hx1[i] = hx[i]; hs2[i] = hs1[i]; hs1[i] = hs[i]; }
        }
        if (DOWN) {
#pragma unroll
            for (int o = 0; o < 2; o++) {                // output column x/2+o is centred on p[2+2o]
                hp[4][o] = p[2 * o] + p[4 + 2 * o] + 4 * (p[1 + 2 * o] + p[3 + 2 * o]) + 6 * p[2 + 2 * o];
            }
            if (j >= 4 && (j & 1) == 0) {                // centre row = staged row j-2 = image row ybase+j-4
                const int Y = (ybase + j - 4) >> 1;
                const int oh = (h + 1) >> 1, ow = (w + 1) >> 1;
                if (Y < oh) {
#pragma unroll
                    for (int o = 0; o < 2; o++) {
                        const int X = (x >> 1) + o;
                        const int v = (hp[0][o] + hp[4][o] + 4 * (hp[1][o] + hp[3][o]) + 6 * hp[2][o] + 128) >> 8;
                        if (X < ow) down[(int64_t)Y * downpitch + X] = (uint8_t)v;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) { hp[i][0] = hp[i + 1][0]; hp[i][1] = hp[i + 1][1]; }
        }
    }
}

static int launch_level(const uint8_t *src, int h, int w, int64_t pitch, int16_t *deriv, int64_t dpitch,
                        uint8_t *down, int64_t downpitch, cudaStream_t st)
{
    if (!src || h <= 0 || w <= 0 || pitch < w) return IBT_E_INVALID;
    if (!deriv && !down) return IBT_OK;
    if (deriv && (dpitch < (int64_t)w * 4 || dpitch % 4 != 0 || reinterpret_cast<uintptr_t>(deriv) % 4 != 0))
        return IBT_E_INVALID;
    if (down && downpitch < (w + 1) / 2) return IBT_E_INVALID;
    const int src_word_ok = (reinterpret_cast<uintptr_t>(src) % 4 == 0) && (pitch % 4 == 0);
    const int deriv_vec_ok = deriv && (reinterpret_cast<uintptr_t>(deriv) % 16 == 0) && (dpitch % 16 == 0);
    dim3 grid((w + TW - 1) / TW, (h + TH - 1) / TH);
    uint8_t *d8 = reinterpret_cast<uint8_t *>(deriv);
    if (deriv && down)
        pyr_level_kernel<true, true><<<grid, 64, 0, st>>>(src, h, w, pitch, d8, dpitch, down, downpitch, src_word_ok, deriv_vec_ok);
    else if (deriv)
        pyr_level_kernel<true, false><<<grid, 64, 0, st>>>(src, h, w, pitch, d8, dpitch, nullptr, 0, src_word_ok, deriv_vec_ok);
    else
        pyr_level_kernel<false, true><<<grid, 64, 0, st>>>(src, h, w, pitch, nullptr, 0, down, downpitch, src_word_ok, 0);
    return check_launch("ibt_pyr_level_u8");
}

} // namespace ibt

IBT_API int ibt_pyramid_levels(int H, int W, int winW, int winH, int maxLevel, int *sizes_hw)
{
    if (H <= 0 || W <= 0 || maxLevel < 0 || maxLevel >= IBT_MAX_LEVELS || !sizes_hw) return IBT_E_INVALID;
    int l = 0, h = H, w = W;
    sizes_hw[0] = h; sizes_hw[1] = w;
    while (l < maxLevel) {
        const int nh = (h + 1) / 2, nw = (w + 1) / 2;
        if (!(nw > winW && nh > winH)) break;
        h = nh; w = nw; ++l;
        sizes_hw[2 * l] = h; sizes_hw[2 * l + 1] = w;
    }
    return l;
}

IBT_API int ibt_pyr_level_u8(const uint8_t *src, int h, int w, int64_t src_pitch, int16_t *deriv, int64_t deriv_pitch,
                             uint8_t *down, int64_t down_pitch, void *stream)
{
    return ibt::launch_level(src, h, w, src_pitch, deriv, deriv_pitch, down, down_pitch, static_cast<cudaStream_t>(stream));
}

IBT_API int ibt_pyramid_build(const ibt_pyramid_t *pyr, int with_derivs, void *stream)
{
    if (!pyr || pyr->nlevels < 1 || pyr->nlevels > IBT_MAX_LEVELS) return IBT_E_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int l = 0; l < pyr->nlevels; l++) {
        const bool last = (l + 1 == pyr->nlevels);
        if (!last && (pyr->rows[l + 1] != (pyr->rows[l] + 1) / 2 || pyr->cols[l + 1] != (pyr->cols[l] + 1) / 2))
            return IBT_E_INVALID;
        int16_t *d = with_derivs ? const_cast<int16_t *>(pyr->deriv[l]) : nullptr;
        if (with_derivs && !d) return IBT_E_INVALID;
        uint8_t *dn = last ? nullptr : const_cast<uint8_t *>(pyr->img[l + 1]);
        if (!last && !dn) return IBT_E_INVALID;
        int rc = ibt::launch_level(pyr->img[l], pyr->rows[l], pyr->cols[l], pyr->img_pitch[l], d,
                                   with_derivs ? pyr->deriv_pitch[l] : 0, dn, last ? 0 : pyr->img_pitch[l + 1], st);
        if (rc != IBT_OK) return rc;
    }
    return IBT_OK;
}
